// ekgsim_b200/host/ekg_support.h -- host-side plumbing of the B200 EkgSim drop-in: .ini reader,
// the reference's text formats (.matrix, measuring points, .column) and console helpers.
//
// Newly written; the behaviour follows the reference so that the same input files and the same
// console/`.column` consumers keep working (paths relative to synergy-twinning/ekgsim):
//   ini semantics            copyOfLibs/Ini.cpp:209-257 (File ctor), :357-382 (parse), Ini.h:268-300 (arrays)
//   .matrix                  simlib/matrix.h:124-248, simlib/simulator.cpp:71-106 (export), :288-367 (delays)
//   measuring points         simlib/simulator.cpp:369-397
//   .column read / write     simlib/columnFile.h:125-174, simlib/simulator.h:196-240
//   vector printing <a,b>    AMS-DEMO/VectorArithmetics.h:74-83
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
#include <sys/time.h>
#include <vector>

namespace ekg {

// ---- small helpers ---------------------------------------------------------------------------------
inline std::string file_extension(const std::string& f) {
	const size_t p = f.rfind('.');
	return p == std::string::npos ? std::string() : f.substr(p);
}

inline double wall_seconds() {
	timeval tv;
	gettimeofday(&tv, nullptr);
	return tv.tv_sec + 1e-6 * tv.tv_usec;
}

/// "<a,b,c>" with the stream's current precision (what the reference prints for criteria / parameters)
template <class T>
std::string angle_list(const std::vector<T>& v, int precision = 6) {
	std::ostringstream o;
	o.precision(precision);
	o << "<";
	for (size_t i = 0; i < v.size(); ++i) o << (i ? "," : "") << v[i];
	o << ">";
	return o.str();
}

/// stderr line "<message>done (Xs)" when the scope ends, like the reference's loading banners
class LogTimer {
	std::ostream& log_;
	double t0_;

public:
	LogTimer(std::ostream& log, const std::string& message) : log_(log), t0_(wall_seconds()) { log_ << message; }
	~LogTimer() { log_ << "done (" << (wall_seconds() - t0_) << "s)\n"; }
};

// ---- .ini -------------------------------------------------------------------------------------------
// Line oriented: "[section]", "name = value" (name/value trimmed of blanks and tabs, split at the
// FIRST '='), everything else (including ';' comments) is kept as a value-less entry.  When a name
// occurs more than once in a section the LAST occurrence wins.  Names are case and space sensitive.
class IniFile {
	struct Var { std::string name, value; };
	struct Sec { std::string name; std::deque<Var> vars; };
	std::vector<Sec> secs_;
	bool found_ = false;

	static std::string trim(const std::string& s) {
		size_t a = 0, b = s.size();
		while (a < b && (s[a] == ' ' || s[a] == '\t')) ++a;
		while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r')) --b;
		return s.substr(a, b - a);
	}

public:
	explicit IniFile(const std::string& fname) {
		secs_.push_back(Sec{"default", {}});
		std::ifstream f(fname.c_str());
		if (!f.is_open()) return;
		found_ = true;
		std::string line;
		size_t active = 0;
		while (std::getline(f, line)) {
			if (line.empty()) continue;
			std::string name, value;
			const size_t eq = line.find('=');
			if (eq == std::string::npos) name = trim(line);
			else { name = trim(line.substr(0, eq)); value = trim(line.substr(eq + 1)); }
			if (name.size() >= 2 && name.front() == '[' && name.back() == ']') {
				const std::string sn = trim(name.substr(1, name.size() - 2));
				active = secs_.size();
				for (size_t i = 1; i < secs_.size(); ++i) if (secs_[i].name == sn) active = i;
				if (active == secs_.size()) secs_.push_back(Sec{sn, {}});
			} else {
				secs_[active].vars.push_front(Var{name, value});  // newest first -> last duplicate wins
			}
		}
	}

	bool found() const { return found_; }
	bool empty() const { return secs_.size() == 1 && secs_[0].vars.empty(); }

	/// index of a named section, -1 if absent (then every lookup in it fails and defaults stay)
	int section(const std::string& name) const {
		for (size_t i = 1; i < secs_.size(); ++i) if (secs_[i].name == name) return (int)i;
		return -1;
	}

	bool raw(std::string& out, const std::string& name, int sec) const {
		if (sec < 0 || sec >= (int)secs_.size()) return false;
		for (const Var& v : secs_[sec].vars) if (v.name == name) { out = v.value; return true; }
		return false;
	}

	bool load(std::string& var, const std::string& name, int sec) const { return raw(var, name, sec); }

	template <class T>
	bool load(T& var, const std::string& name, int sec) const {
		std::string s;
		if (!raw(s, name, sec)) return false;
		std::istringstream ss(s);
		ss >> var;  // a failed extraction leaves what operator>> leaves (0 since C++11), like the reference
		return true;
	}

	/// list separated by any of " ;,\t"; at most `limit` elements (0 = unlimited)
	template <class T>
	bool load_array(std::vector<T>& out, const std::string& name, int sec, unsigned limit = 0) const {
		std::string s;
		if (!raw(s, name, sec)) return false;
		std::istringstream ss(s);
		while (ss) {
			T v;
			ss >> v;
			out.push_back(v);
			int removed = 0;
			while (ss.good()) {
				const int c = ss.peek();
				if (c == ' ' || c == ';' || c == ',' || c == '\t') { ss.ignore(); ++removed; } else break;
			}
			if (!removed) break;
			if (limit && --limit == 0) break;
		}
		return true;
	}
};

// ---- .matrix ---------------------------------------------------------------------------------------
struct MatrixHeader {
	int dims = 0;
	int64_t size[3] = {1, 1, 1};  // [z][y][x]; header lists X x Y [x Z]
	char separator = 0;           // '\t', ',', ' ' or 0 (one character per voxel)
};

inline MatrixHeader read_matrix_header(std::istream& in, const std::string& fname, int max_dims) {
	MatrixHeader h;
	in.ignore(100000000, '\n');  // free comment line
	std::string dim;
	in >> dim;
	if (dim.size() < 2 || !(dim[1] == 'D' || dim[1] == 'd') || !(dim[0] > '0' && dim[0] <= '0' + max_dims))
		throw std::runtime_error("error while loading matrix, could not read dimensionality");
	h.dims = dim[0] - '0';
	int w = 2;
	in >> h.size[w];
	for (int d = h.dims - 1; d > 0; --d) {
		in.ignore(100, 'x');
		--w;
		in >> h.size[w];
	}
	std::string rest, sep;
	std::getline(in, rest, '\n');
	std::istringstream(rest) >> sep;
	h.separator = sep == "tab" ? '\t' : (sep == "comma" || sep == ",") ? ',' : sep == "space" ? ' ' : 0;
	if (in.fail()) throw std::runtime_error("reading file " + fname + " failed (while reading size)");
	return h;
}

/// Shape file -> uint16 layers, raster z,y,x; negative entries / 'X' mark excitation start voxels
/// (bit 0x1000, matrix.h:90,124-129,206-218).
inline void load_shape_matrix(const std::string& fname, std::vector<uint16_t>& layers, int64_t& Z, int64_t& Y, int64_t& X) {
	if (file_extension(fname) != ".matrix") throw std::runtime_error("can only import from .matrix files for the time being");
	std::ifstream in(fname.c_str(), std::ios::binary);
	if (!in.is_open()) throw std::runtime_error("could not open file " + fname);
	const MatrixHeader h = read_matrix_header(in, fname, 3);
	Z = h.size[0]; Y = h.size[1]; X = h.size[2];
	const int64_t n = Z * Y * X;
	layers.assign((size_t)n, 0);
	// slurp the rest: a 4x-resolution heart is ~200 MB of text, iostream extraction is too slow for that
	std::string body((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
	const char* p = body.c_str();
	const char* end = p + body.size();
	int64_t i = 0;
	if (h.separator) {
		for (; i < n; ++i) {
			while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == ',')) ++p;
			if (p >= end) break;
			char* q;
			const long v = strtol(p, &q, 10);
			if (q == p) break;
			p = q;
			layers[(size_t)i] = (uint16_t)(v < 0 ? 0x1000 - v : v);
		}
	} else {
		for (; i < n; ++i) {
			while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
			if (p >= end) break;
			const char ch = *p++;
			layers[(size_t)i] = (ch == 'x' || ch == 'X') ? (uint16_t)(0x1000 + 1) : ch >= 'a' ? (uint16_t)(ch - 'a' + 10)
			                    : ch >= 'A' ? (uint16_t)(ch - 'A' + 10) : (uint16_t)(ch - '0');
		}
	}
	if (i < n) throw std::runtime_error("reading file " + fname + " failed (while reading data)");
}

/// Binary side-cars of the two big text inputs (SURVEY 8(f) "input pipeline"): `<file>.b200bin` next to a shape file
/// (u16 layers) or an excitation-sequence dump (f64 delays at FULL precision -- the text form keeps 3 decimals,
/// simulator.cpp:92).  A side-car is USED whenever it exists and is fresh (size and mtime of the text file as recorded in
/// its header); it is WRITTEN for shape files only with EKGSIM_B200_CACHE=1 (nothing appears in the user's directory
/// unasked) and always together with an excitation-sequence dump (that file is an output of this program anyway).
/// EKGSIM_B200_CACHE=0 switches both off.
inline bool sidecar_enabled() { const char* e = getenv("EKGSIM_B200_CACHE"); return !(e && e[0] == '0'); }
inline bool sidecar_write_enabled() { const char* e = getenv("EKGSIM_B200_CACHE"); return e && e[0] != '0'; }
struct SidecarHeader { char magic[8]; int64_t z, y, x, src_size, src_mtime; };

inline void load_shape_matrix_cached(const std::string& fname, std::vector<uint16_t>& layers, int64_t& Z, int64_t& Y, int64_t& X) {
	if (!sidecar_enabled()) { load_shape_matrix(fname, layers, Z, Y, X); return; }
	struct stat st;
	const bool have_src = stat(fname.c_str(), &st) == 0;
	const std::string cache = fname + ".b200bin";
	SidecarHeader h;
	if (have_src) {
		std::ifstream in(cache.c_str(), std::ios::binary);
		if (in.read(reinterpret_cast<char*>(&h), sizeof h) && !memcmp(h.magic, "EKGSHP1", 8) && h.src_size == (int64_t)st.st_size &&
		    h.src_mtime == (int64_t)st.st_mtime && h.z > 0 && h.y > 0 && h.x > 0) {
			layers.resize((size_t)(h.z * h.y * h.x));
			if (in.read(reinterpret_cast<char*>(layers.data()), (std::streamsize)(layers.size() * 2))) { Z = h.z; Y = h.y; X = h.x; return; }
		}
	}
	load_shape_matrix(fname, layers, Z, Y, X);
	if (have_src && sidecar_write_enabled()) {
		memcpy(h.magic, "EKGSHP1", 8);
		h.z = Z; h.y = Y; h.x = X; h.src_size = (int64_t)st.st_size; h.src_mtime = (int64_t)st.st_mtime;
		std::ofstream out(cache.c_str(), std::ios::binary);
		out.write(reinterpret_cast<const char*>(&h), sizeof h);
		out.write(reinterpret_cast<const char*>(layers.data()), (std::streamsize)(layers.size() * 2));
	}
}

/// 2-D matrix of doubles (conduction / transfer matrix), row major [Y][X]
inline void load_double_matrix(const std::string& fname, std::vector<double>& m, int64_t& rows, int64_t& cols) {
	if (file_extension(fname) != ".matrix") throw std::runtime_error("can only import from .matrix files for the time being");
	std::ifstream in(fname.c_str());
	if (!in.is_open()) throw std::runtime_error("could not open file " + fname);
	const MatrixHeader h = read_matrix_header(in, fname, 2);
	rows = h.size[1]; cols = h.size[2];
	m.assign((size_t)(rows * cols), 0.0);
	for (int64_t i = 0; i < rows * cols; ++i) {
		in >> m[(size_t)i];
		if (h.separator == ',' && (i + 1) % cols != 0) in.ignore(100, ',');
	}
	if (in.fail()) throw std::runtime_error("reading file " + fname + " failed (while reading data)");
}

/// excitation sequence stored as a .matrix of doubles of the model's size (simulator.cpp:288-367)
inline void load_delay_matrix(const std::string& fname, int64_t Z, int64_t Y, int64_t X, std::vector<double>& delay) {
	if (file_extension(fname) != ".matrix") throw std::runtime_error("can only import from .matrix for the time being");
	struct stat st;
	if (sidecar_enabled() && stat(fname.c_str(), &st) == 0) {   // full-precision side-car written with the dump
		SidecarHeader h;
		std::ifstream bin((fname + ".b200bin").c_str(), std::ios::binary);
		if (bin.read(reinterpret_cast<char*>(&h), sizeof h) && !memcmp(h.magic, "EKGACT1", 8) && h.src_size == (int64_t)st.st_size &&
		    h.src_mtime == (int64_t)st.st_mtime && h.z == Z && h.y == Y && h.x == X) {
			delay.resize((size_t)(Z * Y * X));
			if (bin.read(reinterpret_cast<char*>(delay.data()), (std::streamsize)(delay.size() * 8))) return;
		}
	}
	std::ifstream in(fname.c_str());
	if (!in.is_open()) throw std::runtime_error("could not open file " + fname + " to load excitation sequence");
	const MatrixHeader h = read_matrix_header(in, fname, 3);
	if (h.size[0] != Z || h.size[1] != Y || h.size[2] != X)
		throw std::runtime_error("failed to read excitation sequence due to its incompatible size");
	delay.assign((size_t)(Z * Y * X), 0.0);
	for (size_t i = 0; i < delay.size(); ++i) {
		in >> delay[i];
		if (h.separator == ',') in.ignore(1);
	}
}

/// "Shape file" dump (simulator.cpp:71-106 with excit = false): one character per voxel -- 'X' for a start voxel, '0'..'9',
/// then 'A'.. for layers from 10 on --, one text line per y, an empty line after every z block
inline void export_shape_matrix(const std::string& fname, int64_t Z, int64_t Y, int64_t X, const std::vector<uint16_t>& layers) {
	if (file_extension(fname) != ".matrix") throw std::runtime_error("can only export to .matrix for the time being");
	std::ofstream out(fname.c_str());
	out << "Shape file (next line specifies matrix type (2D) size (X x Y x Z) and elemet separator type (tab) after that comes data"
	       " (x are rows, y are lines, z are blocks - separated with empty line)\n";
	if (Z > 1) out << "3D " << X << " x " << Y << " x " << Z;
	else out << "2D " << X << " x " << Y;
	out << "\n";
	for (int64_t z = 0; z < Z; ++z) {
		for (int64_t y = 0; y < Y; ++y) {
			for (int64_t x = 0; x < X; ++x) {
				const uint16_t l = layers[(size_t)((z * Y + y) * X + x)];
				out << ((l & 0x1000) ? 'X' : (l > 9 ? (char)(l - 10 + 'A') : (char)('0' + (unsigned char)l)));
			}
			out << "\n";
		}
		out << "\n";
	}
}

/// The reference's built-in test shape (InputLoader::generateTestShape, simulator.h:600-644; used by EkgSim::loadShape when
/// `[model] shape` is empty): a 2-D ring of 160 x 120 voxels around (y, x) = (80, 30), layer = floor(floor(d - 50) / 2) for
/// 50 < d < 76 (so 0 = empty up to d = 52, then layers 1..12), start voxel = first occupied voxel at x = 30 from y = 80 upwards;
/// like the reference it also writes the shape to autoGeneratedShape.matrix in the current directory.
inline void generate_test_shape(std::vector<uint16_t>& layers, int64_t& Z, int64_t& Y, int64_t& X, const char* export_as = "autoGeneratedShape.matrix") {
	Z = 1; Y = 160; X = 120;
	const int dLow = 50, dHigh = 76;
	layers.assign((size_t)(Y * X), 0);
	const int64_t mid1 = Y / 2, mid2 = X / 4;
	for (int64_t i = 0; i < Y; ++i)
		for (int64_t j = 0; j < X; ++j) {
			const double d = std::sqrt((double)((i - mid1) * (i - mid1)) + (double)((j - mid2) * (j - mid2)));
			if (d > dLow && d < dHigh) layers[(size_t)(i * X + j)] = (uint16_t)((size_t)(d - dLow) / 2);
		}
	int64_t sy = Y / 2;
	const int64_t sx = X / 4;
	for (; sy < Y; ++sy) if (layers[(size_t)(sy * X + sx)] > 0) break;
	if (sy == Y) throw std::runtime_error("Could not create starting point for excitation sequence");
	layers[(size_t)(sy * X + sx)] = (uint16_t)(layers[(size_t)(sy * X + sx)] + 0x1000);   // ShapeElement::layerStartingPoint
	if (export_as && *export_as) export_shape_matrix(export_as, Z, Y, X, layers);
}

/// "Excitation file" dump, 3 decimals, tab separated (simulator.cpp:71-106 with excit = true)
inline void export_delay_matrix(const std::string& fname, int64_t Z, int64_t Y, int64_t X, const std::vector<double>& delay) {
	if (file_extension(fname) != ".matrix") throw std::runtime_error("can only export to .matrix for the time being");
	std::ofstream out(fname.c_str());
	out << "Excitation file (next line specifies matrix type (2D) size (X x Y x Z) and elemet separator type (tab) after that comes data"
	       " (x are rows, y are lines, z are blocks - separated with empty line)\n";
	if (Z > 1) out << "3D " << X << " x " << Y << " x " << Z;
	else out << "2D " << X << " x " << Y;
	out << " tab\n";
	out << std::fixed << std::setprecision(3);
	for (int64_t z = 0; z < Z; ++z) {
		for (int64_t y = 0; y < Y; ++y) {
			for (int64_t x = 0; x < X; ++x) {
				if (x) out << '\t';
				out << delay[(size_t)((z * Y + y) * X + x)];
			}
			out << "\n";
		}
		out << "\n";
	}
	out.close();
	struct stat st;
	if (sidecar_enabled() && stat(fname.c_str(), &st) == 0) {   // the same map at full precision (see sidecar_enabled)
		SidecarHeader h;
		memcpy(h.magic, "EKGACT1", 8);
		h.z = Z; h.y = Y; h.x = X; h.src_size = (int64_t)st.st_size; h.src_mtime = (int64_t)st.st_mtime;
		std::ofstream bin((fname + ".b200bin").c_str(), std::ios::binary);
		bin.write(reinterpret_cast<const char*>(&h), sizeof h);
		bin.write(reinterpret_cast<const char*>(delay.data()), (std::streamsize)(delay.size() * 8));
	}
}

// ---- measuring points: "x, y, z" per line, stored (z,y,x) ---------------------------------------------
struct Vec3 {
	double v[3];
	Vec3() : v{0, 0, 0} {}
	Vec3(double a, double b, double c) : v{a, b, c} {}
	double& operator[](size_t i) { return v[i]; }
	double operator[](size_t i) const { return v[i]; }
};

inline std::vector<Vec3> load_points(const std::string& fname) {
	std::vector<Vec3> pts;
	std::ifstream in(fname.c_str());
	while (in.is_open() && !in.eof()) {
		Vec3 p;
		in >> p[2]; in.ignore(1000, ',');
		in >> p[1]; in.ignore(1000, ',');
		in >> p[0]; in.ignore(1000, '\n');
		if (in.fail()) break;
		pts.push_back(p);
	}
	if (pts.empty()) throw std::runtime_error("failed to open file containing measuring points: " + fname);
	return pts;
}

// ---- .column ----------------------------------------------------------------------------------------
/// reads whitespace separated columns; leading '#' / '%' lines are skipped; the first data line
/// fixes the column count; reading stops at the first line that does not fill every column
inline std::vector<std::vector<double>> load_column_file(const std::string& fname) {
	if (file_extension(fname) != ".column") throw std::runtime_error("ColumnFile::load only works on .column files");
	std::ifstream in(fname.c_str());
	if (!in.is_open()) throw std::runtime_error("could not open " + fname);
	while (in.peek() == '#' || in.peek() == '%') in.ignore(10000, '\n');
	std::string line;
	std::getline(in, line);
	std::vector<double> row;
	{
		std::istringstream ls(line);
		double d;
		while (ls >> d) row.push_back(d);
	}
	if (row.empty()) throw std::runtime_error("File has 0 columns");
	std::vector<std::vector<double>> cols(row.size());
	bool ok = true;
	while (in && ok) {
		for (size_t i = 0; i < cols.size(); ++i) cols[i].push_back(row[i]);
		std::getline(in, line);
		std::istringstream ls(line);
		for (size_t i = 0; i < row.size(); ++i) ls >> row[i];
		ok = !ls.fail();
	}
	return cols;
}

struct NamedColumn {
	std::string name, comment;
	const std::vector<double>* data = nullptr;
	size_t size() const { return data->size(); }
	double operator[](size_t i) const { return (*data)[i]; }
};

/// the reference's multi-vector .column writer: "#Comment: ...", optional per-vector comments, "# ",
/// header row, then rows "setw(10) fixed(5) time \t setw(10) scientific value ..."; '\n' between
/// rows, none after the last.
template <class Vec>
void export_columns(const Vec& vec, const std::string& fname, double start_time, double time_step = -1.0, const std::string& comment = "") {
	if (file_extension(fname) != ".column") throw std::runtime_error("saving vectors to .column files only for the time being");
	std::ofstream out(fname.c_str());
	out << "#Comment: " << comment << '\n';
	for (size_t i = 0; i < vec.size(); ++i)
		if (vec[i].comment != "") out << "# " << vec[i].name << " : " << vec[i].comment << '\n';
	out << "# \n";
	out << (time_step > 0 ? "#Time[ms]\t" : "# ");
	for (size_t i = 0; i < vec.size(); ++i) out << (i ? "\t" : "") << vec[i].name;
	out << '\n';
	double time = start_time;
	for (size_t i = 0; i < vec[0].size(); ++i) {
		if (i) out << "\n";
		if (time_step > 0) {
			out << std::fixed << std::setprecision(5) << std::setw(10) << time << '\t';
			time += time_step;
		}
		for (size_t j = 0; j < vec.size(); ++j) {
			if (j) out << '\t';
			out << std::scientific << std::setw(10) << vec[j][i];
		}
	}
}

}  // namespace ekg
