// oracle/ref_dump.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Driver around the UNMODIFIED reference (compiled from /root/reference by oracle/Makefile,
// never copied).  It textually includes the reference's sim.cpp so that the evaluation glue
// (class SimImplementation, sim.cpp:322-1043, which the reference hides inside the .cpp) can
// be driven directly and its intermediate values dumped at full double precision -- the
// reference's own outputs (.column files, stdout) only carry 6 significant digits
// (simulator.h:227,233).
//
// Run it in a directory holding simulator.ini + the files that ini names (like the reference).
//
//   ref_dump activation <out.bin>
//        layers (u16) + activation times (f64), raster z,y,x           (simulator.cpp:248-286)
//   ref_dump eval <vectors.txt> <out.bin> [--length N] [--glue-only]
//        for every parameter vector (one per line, comma/space separated, %.17g):
//        24x9 layer coefficients (sim.cpp:825-916), displaced lead positions (sim_lib.h:197-208),
//        ECG[L][T] (simulator.cpp:452-550), criteria, violation (sim.cpp:443-491, :600-702).
//        --glue-only sets "fast approximation limit" to -1 so Simulation::run is skipped
//        (sim.cpp:741-746) -> only the glue outputs are meaningful.
//
// Binary layout (little endian), all records self-describing:
//   activation: "EKGACT1\0" u64 Z,Y,X ; u16 layer[Z*Y*X] ; f64 delay[Z*Y*X]
//   eval:       "EKGEVL1\0" u64 nVec, nLayers, L, T, nCrit ; then per vector:
//               f64 params[16pad: u64 nParams + f64[nParams]] f64 layerK[nLayers*9] f64 leads[L*3] (z,y,x)
//               f64 ecg[L*T] f64 crit[nCrit] f64 violation f64 seconds u64 simulationDone

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <queue>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <stdint.h>
#include <sys/time.h>

// compiled with g++ -fno-access-control (oracle/Makefile) so this driver can read the
// reference's private members (EkgSim::sim, SimImplementation::layerAps, ...).
#include REF_SIM_CPP

static double now() {
	timeval tv; gettimeofday(&tv, 0);
	return tv.tv_sec + 1e-6 * tv.tv_usec;
}

static void wr(FILE* f, const void* p, size_t n) {
	if (fwrite(p, 1, n, f) != n) throw std::runtime_error("short write");
}
static void wr64(FILE* f, uint64_t v) { wr(f, &v, 8); }

static int dumpActivation(const char* out) {
	EkgSim sim;
	sim.loadSettings("simulator.ini");
	sim.loadTransferMatrix();
	sim.loadShape();
	double t0 = now();
	sim.simExcitationSequence();
	double t1 = now();
	const SimLib::ShapeMatrix& sh = sim.getModelShape();
	size_t Z = sh.size()[0], Y = sh.size()[1], X = sh.size()[2];
	FILE* f = fopen(out, "wb");
	if (!f) throw std::runtime_error("cannot open output");
	wr(f, "EKGACT1\0", 8);
	wr64(f, Z); wr64(f, Y); wr64(f, X);
	std::vector<uint16_t> lay(Z * Y * X);
	std::vector<double> del(Z * Y * X);
	size_t n = 0;
	for (size_t z = 0; z < Z; ++z) for (size_t y = 0; y < Y; ++y) for (size_t x = 0; x < X; ++x, ++n) {
		lay[n] = (uint16_t)sh[z][y][x].layer;
		del[n] = sh[z][y][x].excitationDelay;
	}
	wr(f, lay.data(), lay.size() * 2);
	wr(f, del.data(), del.size() * 8);
	fclose(f);
	fprintf(stderr, "ref_dump: automaton %.3f s, %zu x %zu x %zu\n", t1 - t0, Z, Y, X);
	printf("{\"automaton_seconds\": %.6f}\n", t1 - t0);
	return 0;
}

static std::vector<std::vector<double> > readVectors(const char* fname) {
	std::ifstream in(fname);
	if (!in.is_open()) throw std::runtime_error("cannot open vectors file");
	std::vector<std::vector<double> > v;
	std::string line;
	while (std::getline(in, line)) {
		for (size_t i = 0; i < line.size(); ++i) if (line[i] == ',' || line[i] == ';') line[i] = ' ';
		std::istringstream ss(line);
		std::vector<double> row; double d;
		while (ss >> d) row.push_back(d);
		if (!row.empty()) v.push_back(row);
	}
	return v;
}

static int dumpEval(const char* vecFile, const char* out, int length, bool glueOnly) {
	std::vector<std::vector<double> > vecs = readVectors(vecFile);
	SimImplementation impl;                       // sim.cpp:362 (loads model, runs automaton)
	if (length > 0) impl.sim->getSettings().simulationLength = length;
	if (glueOnly) impl.settings.fastApproxLimit = -1;
	size_t nLayers = impl.sim->requiredAps();
	size_t L = impl.sim->numMeasurements();
	size_t T = (size_t)ceil(impl.sim->getSettings().simulationLength / impl.sim->getSettings().simulationTimeStep);
	size_t nCrit = impl.deducedNumOfCriteria;
	FILE* f = fopen(out, "wb");
	if (!f) throw std::runtime_error("cannot open output");
	wr(f, "EKGEVL1\0", 8);
	wr64(f, vecs.size()); wr64(f, nLayers); wr64(f, L); wr64(f, T); wr64(f, nCrit);
	for (size_t v = 0; v < vecs.size(); ++v) {
		std::vector<double> result;
		double t0 = now();
		double violation = impl.eval(vecs[v], result);      // sim.cpp:443
		double t1 = now();
		wr64(f, vecs[v].size());
		wr(f, vecs[v].data(), vecs[v].size() * 8);
		for (size_t l = 0; l < nLayers; ++l) wr(f, impl.layerAps[l].getK(), 9 * 8);
		std::vector<EkgSim::PositionVec> pos;
		impl.sim->sim.getMeasuringPoints(pos);
		for (size_t l = 0; l < L; ++l) { double p[3] = {pos[l][0], pos[l][1], pos[l][2]}; wr(f, p, 24); }
		for (size_t l = 0; l < L; ++l) {
			std::vector<double> m(T, 0.0);
			if (impl.simulationDone) {
				const std::vector<double>& r = impl.sim->getMeasurement(l);
				for (size_t t = 0; t < T && t < r.size(); ++t) m[t] = r[t];
			}
			wr(f, m.data(), T * 8);
		}
		result.resize(nCrit, 0.0);
		wr(f, result.data(), nCrit * 8);
		wr(f, &violation, 8);
		double secs = t1 - t0; wr(f, &secs, 8);
		wr64(f, impl.simulationDone ? 1 : 0);
		fflush(f);
		fprintf(stderr, "\nref_dump: vector %zu/%zu  %.3f s  violation %.17g\n", v + 1, vecs.size(), secs, violation);
	}
	fclose(f);
	return 0;
}

int main(int argc, char** argv) {
	try {
		if (argc >= 3 && !strcmp(argv[1], "activation")) return dumpActivation(argv[2]);
		if (argc >= 4 && !strcmp(argv[1], "eval")) {
			int length = -1; bool glueOnly = false;
			for (int i = 4; i < argc; ++i) {
				if (!strcmp(argv[i], "--length") && i + 1 < argc) length = atoi(argv[++i]);
				else if (!strcmp(argv[i], "--glue-only")) glueOnly = true;
			}
			return dumpEval(argv[2], argv[3], length, glueOnly);
		}
		fprintf(stderr, "usage: ref_dump activation <out.bin> | eval <vectors.txt> <out.bin> [--length N] [--glue-only]\n");
		return 2;
	} catch (std::exception& e) {
		fprintf(stderr, "ref_dump: %s\n", e.what());
		return 1;
	} catch (const char* e) {
		fprintf(stderr, "ref_dump: %s\n", e);
		return 1;
	}
}
