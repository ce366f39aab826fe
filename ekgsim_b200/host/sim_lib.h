// ekgsim_b200/host/sim_lib.h -- the "simlib Simulator interface" of the B200 build.
//
// Source-compatible stand-in for the reference's simlib/sim_lib.h (class SimLib::EkgSim,
// sim_lib.h:93-287) and the types its callers touch (Settings simulator.h:83-147, WohlfartPlus
// Wohlfart.h:167-230, ActionPotential simulator.h:468-480).  Same method names, argument meaning,
// console output and exceptions (std::runtime_error) -- but where the reference walks the voxel
// model on the CPU, this class hands the work to libekgsim_b200.so through the C ABI
// (include/ekgsim_b200.h):
//
//     loadShape + loadTransferMatrix  -> ekg_model_create       (device-resident model)
//     simExcitationSequence           -> ekg_model_activation / ekg_model_set_activation
//     run                             -> ekg_simulate           (fused AP + stencil + lead kernel)
//     runBatch (new)                  -> ekg_simulate with B parameter sets in one launch
//     fitLayers / evaluateBatch (new) -> ekg_fit_layers / ekg_evaluate (layer-AP construction on the device)
//
// Host-only pieces that stay on the CPU exactly like in the reference: file parsing, the lead
// displacement plane (u,v) construction, the 2*T-sample "string model" approximation
// (simulator.cpp:552-559) and the AP formula used by the evaluation glue's layer fitting.
// There is no CPU fallback for the voxel loops: without a CUDA device the calls throw.
#pragma once

#include <algorithm>
#include <cmath>
#include <memory>
#include <sstream>
#include <thread>
#include <unordered_map>

#include "../../include/ekgsim_b200.h"
#include "ekg_support.h"

namespace SimLib {

using ekg::Vec3;

// ---- simulator settings, same ini keys as the reference (simulator.h:107-134) ------------------------
struct Settings {
	std::string inputShapeFilename, inputExcitationSequenceFilename, inputPointsFilename, inputTransferFilename;
	std::string inputActionPotentials, inputActionPotentialsFilter;
	double inputActionPotentialsFilterParam = 0;
	double inputActionPotentialsTimeStep = 1;
	std::string inputActionPotentialsFunction;
	double inputActionPotentialsScale = 0;
	std::string outputExcitationSequence, outputApFilename, outputFilename;
	int simulationLength = 800;
	int simulationStart = 0;
	std::string neighbourhoodType;
	double simulationTimeStep = 10;

	void loadFromIni(const ekg::IniFile& ini) {
		const int model = ini.section("model");
		ini.load(inputShapeFilename, "shape", model);
		ini.load(inputExcitationSequenceFilename, "excitation sequence", model);
		ini.load(inputPointsFilename, "points", model);
		ini.load(inputTransferFilename, "transfer", model);
		const int ap = ini.section("action potentials");
		ini.load(inputActionPotentials, "input file", ap);
		ini.load(inputActionPotentialsFilter, "filter", ap);
		ini.load(inputActionPotentialsTimeStep, "resample time step", ap);
		ini.load(inputActionPotentialsFilterParam, "filter parameter", ap);
		ini.load(inputActionPotentialsFunction, "combination function", ap);
		ini.load(inputActionPotentialsScale, "expected length", ap);
		const int out = ini.section("output files");
		ini.load(outputExcitationSequence, "excitation sequence", out);
		ini.load(outputApFilename, "action potentials filename", out);
		outputFilename = "result.column";
		ini.load(outputFilename, "results filename", out);
		const int sim = ini.section("simulation");
		ini.load(simulationLength, "length", sim);
		ini.load(simulationStart, "start", sim);
		ini.load(neighbourhoodType, "neighbourhood type", sim);
		simulationTimeStep = 1.0 / 3.0;
		ini.load(simulationTimeStep, "time step", sim);
	}
};

// ---- extended Wohlfart AP, 9 coefficients (Wohlfart.h:167-230) ----------------------------------------
class WohlfartPlus {
public:
	static const size_t numParams = 9;

protected:
	double k[numParams];

public:
	WohlfartPlus() {}
	explicit WohlfartPlus(const double newK[numParams]) { setK(newK); }
	void setK(const double newK[numParams]) { std::copy(newK, newK + numParams, k); }
	void setK(const std::vector<double>& newK) { std::copy(newK.begin(), newK.end(), k); }
	const double* getK() const { return k; }
	double* getK() { return k; }

	/// value of the AP at time t
	double operator[](double t) const {
		return (1.0 / (1.0 + std::exp(-k[1] * t)))
		     * (k[2] * ((1.0 - k[3]) * std::exp(-k[4] * t) + k[3]))
		     * (std::exp(-k[5] * t) * (1 - std::pow((1 + std::exp(-k[7] * (t - k[8])
		            + std::log(std::pow(2, (k[7] / k[6])) - 1))), -(k[6] / k[7]))))
		     + k[0];
	}

	/// time at which the AP has repolarised to k0 + 0.1 k2 (last downward crossing on a 1 ms grid
	/// over [0,1000), linearly interpolated); -1 if it never does
	double apd90() const {
		const size_t n = 1000;
		std::vector<double> d(n);
		for (size_t i = 0; i < n; ++i) d[i] = (*this)[(double)i];
		const double target = k[0] + k[2] * 0.1;
		double time = -1.0;
		for (size_t i = 1; i < n; ++i)
			if (d[i - 1] > target && d[i] <= target) {
				const double wa = d[i - 1] - target, wb = target - d[i];
				time = (double(i - 1) * wa + double(i) * wb) / (wa + wb);
			}
		return time;
	}
};

// ---- AP + activation time (simulator.h:468-480, simulator.cpp:154-170) -----------------------------------
struct ActionPotential {
	WohlfartPlus wohl;
	double at;

	void init(const double* wohlK, double newAt) { wohl.setK(wohlK); wohl.getK()[8] -= newAt; at = newAt; }
	void init(const ActionPotential& ap, double newAt) { wohl.setK(ap.getK()); wohl.getK()[8] -= newAt; at = newAt; }
	double operator()(double time) const { return wohl[time - at]; }
	double& operator[](size_t i) { return wohl.getK()[i]; }
	double operator[](size_t i) const { return wohl.getK()[i]; }
	const double* getK() const { return wohl.getK(); }
	size_t size() const { return wohl.numParams; }
};

/// element of the vector handed to exportVectors (simulator.h:488-500); a type of this namespace so
/// that unqualified exportVectors(...) calls in caller code resolve here by ADL like in the reference
struct saveVecElement : ekg::NamedColumn {};

template <class Vec>
void exportVectors(const Vec& vec, const std::string& filename, double startTime, double timeStep = -1.0, const std::string& comment = "") {
	ekg::export_columns(vec, filename, startTime, timeStep, comment);
}

/// read-only view of the voxel model for callers of getModelShape()
struct ShapeElement { size_t layer; double excitationDelay; size_t apIndex; };

class EkgSim {
public:
	typedef Vec3 PositionVec;

private:
	Settings settings;
	// model on the host (as parsed) and on the device
	std::vector<uint16_t> layers_;
	int64_t Z_ = 0, Y_ = 0, X_ = 0;
	std::vector<double> transfer_;
	int64_t tRows_ = 0, tCols_ = 0;
	ekg_model* model_ = nullptr;
	int device_ = 0;
	/// z-slab mode (new; BASELINE config 4 "single sim z-slab sharded"): ONE large model spread over several GPUs from this
	/// process.  Every device holds the model and sums the ECG over its slab of voxel planes (ekg_model_set_slab, slabs
	/// balanced by occupied voxels); run() adds the partial ECGs [leads][T] on the host.  The automaton runs peer-linked
	/// over the same slabs (ekg_model_activation_link: one kernel per device, face planes exchanged through NVLink peer
	/// memory); devices without native peer atomics each compute the whole map instead.  slabModels_[0] == model_.
	std::vector<int> slabDevices_;
	std::vector<ekg_model*> slabModels_;
	std::vector<std::pair<int64_t, int64_t>> slabRanges_;
	bool slabAutomatonLinked_ = false;
	size_t targetNumOfAps_ = 12;   // Simulation::targetNumOfAps default (simulator.h:556)
	int nbhd_ = -1;
	double timeStep_ = 0;
	bool haveActivation_ = false;

	std::vector<ActionPotential> aps_;            // layer APs before run(), per-class APs after (getAps)
	std::vector<PositionVec> originalMeasuringPositions, measuringPositions, vectorU, vectorV;
	std::vector<PositionVec> mps_;                // what Simulation::mps holds
	std::vector<std::vector<double>> measurement_;
	double startTime_ = 0;
	int mode_ = EKG_MODE_DEFAULT;
	double lastRunSeconds_ = 0;

	static void check(int rc) { if (rc != EKG_OK) throw std::runtime_error(ekg_last_error()); }

public:
	/// [0, Z) cut into n contiguous z-slabs of (nearly) equal numbers of occupied voxels (the same cuts as
	/// ekgsim_b200/dist.py::slab_ranges makes for the torchrun layout); slabs may be empty for tiny models
	static std::vector<std::pair<int64_t, int64_t>> balancedSlabs(const std::vector<uint16_t>& layers, int64_t Z, int64_t Y, int64_t X, size_t n) {
		std::vector<int64_t> cum((size_t)Z + 1, 0);
		for (int64_t z = 0; z < Z; ++z) {
			int64_t c = 0;
			const uint16_t* pl = &layers[(size_t)(z * Y * X)];
			for (int64_t i = 0; i < Y * X; ++i) c += (pl[i] & 0x0fff) != 0;
			cum[(size_t)z + 1] = cum[(size_t)z] + c;
		}
		std::vector<std::pair<int64_t, int64_t>> slabs(n, std::make_pair<int64_t, int64_t>(0, 0));
		int64_t prev = 0;
		for (size_t d = 0; d < n; ++d) {
			int64_t cut = Z;
			if (d + 1 < n) {
				// the first plane count whose prefix reaches the target, or the one before it if that prefix is at least as close
				const double target = (double)cum[(size_t)Z] * (double)(d + 1) / (double)n;
				cut = (int64_t)(std::lower_bound(cum.begin(), cum.end(), target, [](int64_t a, double t) { return (double)a < t; }) - cum.begin());
				if (cut > 0 && std::fabs((double)cum[(size_t)cut - 1] - target) <= std::fabs((double)cum[(size_t)std::min<int64_t>(cut, Z)] - target)) --cut;
				cut = std::min(std::max(cut, prev), Z);
			}
			slabs[d] = std::make_pair(prev, cut);
			prev = cut;
		}
		return slabs;
	}

private:
	void ensureModel() {
		if (model_) return;
		if (layers_.empty()) throw std::runtime_error("shape not loaded");
		if (transfer_.empty()) throw std::runtime_error("transfer matrix not loaded");
		if (slabDevices_.size() < 2) {
			check(ekg_model_create(layers_.data(), Z_, Y_, X_, transfer_.data(), tRows_, tCols_, device_, &model_));
			return;
		}
		// one handle per device, created side by side (each upload + device-side set-up is independent)
		const size_t n = slabDevices_.size();
		slabModels_.assign(n, nullptr);
		std::vector<std::string> errors(n);
		forEachSlab([&](size_t d) {
			if (ekg_model_create(layers_.data(), Z_, Y_, X_, transfer_.data(), tRows_, tCols_, slabDevices_[d], &slabModels_[d]) != EKG_OK)
				throw std::runtime_error(ekg_last_error());
		});
		model_ = slabModels_[0];
		slabRanges_ = balancedSlabs(layers_, Z_, Y_, X_, n);
		forEachSlab([&](size_t d) {
			if (ekg_model_set_slab(slabModels_[d], slabRanges_[d].first, slabRanges_[d].second) != EKG_OK) throw std::runtime_error(ekg_last_error());
		});
	}

	/// f(d) for every slab device on its own host thread (the C ABI is one caller per handle; errors are per thread)
	template <class F>
	void forEachSlab(F f) {
		const size_t n = slabDevices_.size();
		std::vector<std::string> errors(n);
		std::vector<std::thread> pool;
		for (size_t d = 0; d < n; ++d)
			pool.emplace_back([&, d]() { try { f(d); } catch (std::exception& e) { errors[d] = e.what(); if (errors[d].empty()) errors[d] = "error"; } });
		for (std::thread& t : pool) t.join();
		for (const std::string& e : errors) if (!e.empty()) throw std::runtime_error(e);
	}

	void destroyModels() {
		if (slabModels_.empty()) { if (model_) ekg_model_destroy(model_); }
		else for (ekg_model* m : slabModels_) if (m) ekg_model_destroy(m);
		slabModels_.clear();
		model_ = nullptr;
		haveActivation_ = false;
	}

	/// the automaton over the slabs: peer-linked when the devices allow it, else every device computes the whole map
	void slabActivation() {
		const size_t n = slabModels_.size();
		std::vector<unsigned char> infos(n * EKG_LINK_INFO_BYTES);
		std::vector<int64_t> slabs(2 * n);
		for (size_t d = 0; d < n; ++d) {
			check(ekg_model_activation_link_info(slabModels_[d], &infos[d * EKG_LINK_INFO_BYTES]));
			slabs[2 * d] = slabRanges_[d].first; slabs[2 * d + 1] = slabRanges_[d].second;
		}
		bool linked = getenv("EKGSIM_B200_SLAB_AUTOMATON") == nullptr || std::string(getenv("EKGSIM_B200_SLAB_AUTOMATON")) != "replicated";
		for (size_t d = 0; d < n && linked; ++d) {
			const int rc = ekg_model_activation_link(slabModels_[d], (int)d, (int)n, infos.data(), slabs.data());
			if (rc == EKG_E_UNSUPPORTED) linked = false;
			else check(rc);
		}
		slabAutomatonLinked_ = linked;
		if (!linked) {
			for (ekg_model* m : slabModels_) ekg_model_activation_unlink(m);
			forEachSlab([&](size_t d) { if (ekg_model_activation(slabModels_[d], nullptr, nullptr) != EKG_OK) throw std::runtime_error(ekg_last_error()); });
			return;
		}
		// begin everywhere (rings ready) -> launch everywhere -> wait everywhere -> pull the other slabs -> publish
		forEachSlab([&](size_t d) { if (ekg_model_activation_begin(slabModels_[d]) != EKG_OK) throw std::runtime_error(ekg_last_error()); });
		for (ekg_model* m : slabModels_) check(ekg_model_activation_linked_launch(m, 0));
		forEachSlab([&](size_t d) { if (ekg_model_activation_linked_wait(slabModels_[d], nullptr, nullptr) != EKG_OK) throw std::runtime_error(ekg_last_error()); });
		forEachSlab([&](size_t d) { if (ekg_model_activation_linked_gather(slabModels_[d]) != EKG_OK) throw std::runtime_error(ekg_last_error()); });
		forEachSlab([&](size_t d) { if (ekg_model_activation_end(slabModels_[d], nullptr) != EKG_OK) throw std::runtime_error(ekg_last_error()); });
		// the mappings are only needed while the automaton runs; with peer access left on, every later allocation on these
		// devices would be mapped for all the peers as well
		for (ekg_model* m : slabModels_) ekg_model_activation_unlink(m);
	}

	static PositionVec cross(const PositionVec& a, const PositionVec& b) {
		return PositionVec(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
	}
	static void normalize(PositionVec& v) {
		double s = v[0] * v[0];
		s += v[1] * v[1];
		s += v[2] * v[2];
		const double l = std::sqrt(s);
		const double f = l != 0 ? 1.0 / l : 0.0;
		for (int i = 0; i < 3; ++i) v[i] *= f;
	}

public:
	EkgSim() {
		if (const char* d = getenv("EKGSIM_B200_DEVICE")) device_ = atoi(d);
		if (const char* sl = getenv("EKGSIM_B200_SLABS")) setSlabDevices(sl);
		if (const char* m = getenv("EKGSIM_B200_MODE")) {
			const std::string s(m);
			mode_ = s == "direct" ? EKG_MODE_DIRECT : s == "hoisted" ? EKG_MODE_HOISTED : s == "separable" ? EKG_MODE_SEPARABLE : EKG_MODE_DEFAULT;
		}
	}
	~EkgSim() { destroyModels(); }

	void setDevice(int device) { device_ = device; }
	/// z-slab mode over these GPUs: "all" | "<count>" | "<id>,<id>,..." (an id may repeat: several slabs on one GPU);
	/// fewer than two devices = off.  Takes effect when the device model is (re)built.
	void setSlabDevices(const std::string& spec) {
		std::vector<int> ids;
		const int have = ekg_device_count();
		if (spec == "all") for (int i = 0; i < have; ++i) ids.push_back(i);
		else if (spec.find(',') == std::string::npos) { for (int i = 0; i < atoi(spec.c_str()); ++i) ids.push_back(i); }
		else {
			std::istringstream in(spec);
			for (std::string tok; std::getline(in, tok, ',');) if (!tok.empty()) ids.push_back(atoi(tok.c_str()));
		}
		for (int id : ids) if (id < 0 || id >= have) throw std::runtime_error("no such CUDA device in slab list: " + spec);
		if (ids.size() > 16) throw std::runtime_error("at most 16 slabs");
		destroyModels();
		slabDevices_ = ids.size() >= 2 ? ids : std::vector<int>();
		if (!ids.empty()) device_ = ids[0];
	}
	size_t numSlabs() const { return slabDevices_.size() >= 2 ? slabDevices_.size() : 1; }
	bool sharded() const { return slabDevices_.size() >= 2; }
	bool slabAutomatonLinked() const { return slabAutomatonLinked_; }
	const std::vector<std::pair<int64_t, int64_t>>& slabRanges() const { return slabRanges_; }
	int device() const { return device_; }

	/// A second simulator with the same inputs and settings on another GPU (new; multi-GPU evaluation in one process,
	/// the in-process counterpart of the reference's "every MPI rank builds its own OptimizationFunction",
	/// main.cpp:301-302): the parsed host state is copied, the device model is created and the activation map
	/// computed on `device` (bit-identical on every device).  Quiet: the console contract belongs to the primary.
	std::unique_ptr<EkgSim> replicate(int device) const {
		if (sharded()) throw std::runtime_error("a z-slab sharded simulator cannot be replicated (slabs and batch devices are alternatives)");
		std::unique_ptr<EkgSim> r(new EkgSim);
		r->slabDevices_.clear();
		r->settings = settings;
		r->layers_ = layers_; r->Z_ = Z_; r->Y_ = Y_; r->X_ = X_;
		r->transfer_ = transfer_; r->tRows_ = tRows_; r->tCols_ = tCols_;
		r->device_ = device;
		r->targetNumOfAps_ = targetNumOfAps_;
		r->nbhd_ = nbhd_; r->timeStep_ = timeStep_; r->mode_ = mode_;
		r->originalMeasuringPositions = originalMeasuringPositions; r->measuringPositions = measuringPositions;
		r->vectorU = vectorU; r->vectorV = vectorV; r->mps_ = mps_;
		r->measurement_.assign(mps_.size(), std::vector<double>());
		if (haveActivation_) {
			r->ensureModel();
			if (settings.inputExcitationSequenceFilename == "") check(ekg_model_activation(r->model_, nullptr, nullptr));
			else {
				std::vector<double> delay((size_t)(Z_ * Y_ * X_));
				check(ekg_model_get_activation(model_, delay.data()));
				check(ekg_model_set_activation(r->model_, delay.data()));
			}
			r->haveActivation_ = true;
		}
		return r;
	}
	void setMode(int mode) { mode_ = mode; }
	double lastRunSeconds() const { return lastRunSeconds_; }
	ekg_model* handle() { ensureModel(); return model_; }

	void loadSettings(const char* fname = "settings.ini") {
		ekg::LogTimer tm(std::cerr, "loading settings                            ");
		ekg::IniFile ini(fname);
		if (!ini.found()) std::cout << "warning, " << fname << " not found. ";
		else if (ini.empty()) std::cout << "warning, " << fname << " is empty. ";
		else settings.loadFromIni(ini);
	}

	Settings& getSettings() { return settings; }
	const Settings& settingsView() const { return settings; }

	/// applies the selected timestep and neighbourhood (sim_lib.h:131-148; note that "3D4" selects the
	/// 8 cube corners and "2D4" the 4 in-plane diagonals, simulator.h:338-364)
	void applySettings() {
		const std::string& t = settings.neighbourhoodType;
		if (t == "2D4") nbhd_ = EKG_NBHD_2D4;
		else if (t == "2D8") nbhd_ = EKG_NBHD_2D8;
		else if (t == "3D4") nbhd_ = EKG_NBHD_3D4;
		else if (t == "3D8" || t == "cube") nbhd_ = EKG_NBHD_3D8;
		else throw std::runtime_error("\n   unknown neighbourhood (only know of these: 2D4, 3D4, 2D8, 3D8 (cube))");
		timeStep_ = settings.simulationTimeStep;
	}

	/// swaps: the caller's vector receives the previous contents (sim_lib.h:151, simulator.cpp:208-210)
	void setApsDestructive(std::vector<ActionPotential>& destructible) { aps_.swap(destructible); cellApsPending_ = false; }
	/// layer APs before run(); after run() one AP per (layer, delay) class in first-seen raster order
	/// (built on first use -- the class table costs a pass over the whole model)
	const std::vector<ActionPotential>& getAps() const {
		if (cellApsPending_) const_cast<EkgSim*>(this)->buildCellAps();
		return aps_;
	}

	void loadTransferMatrix() {
		ekg::LogTimer tm(std::cerr, "loading transfer matrix                     ");
		ekg::load_double_matrix(settings.inputTransferFilename, transfer_, tRows_, tCols_);
		// Simulation::loadTransferMatrix checks against the default / previously loaded layer count
		if ((size_t)tRows_ < targetNumOfAps_ || (size_t)tCols_ < targetNumOfAps_) throw std::runtime_error("loaded transfer matrix too small");
		destroyModels();
	}

	void loadMeasuringPoints() {
		ekg::LogTimer tm(std::cerr, "loading measuring points                    ");
		std::cerr << "measuring points: \n";
		mps_ = ekg::load_points(settings.inputPointsFilename);
		for (const PositionVec& p : mps_) std::cerr << " -> " << p[2] << ", " << p[1] << ", " << p[0] << "\n";
		originalMeasuringPositions = mps_;
		const size_t n = mps_.size();
		measuringPositions = mps_;
		vectorU.resize(n);
		vectorV.resize(n);
		measurement_.assign(n, std::vector<double>());
		// plane of movement of every lead: u = -(x_axis x p), v = (x_axis x p) x p, coordinates are (z,y,x)
		const PositionVec xAxis(0, 0, 1);
		for (size_t i = 0; i < n; ++i) {
			const PositionVec& p = originalMeasuringPositions[i];
			vectorU[i] = cross(xAxis, p);
			vectorV[i] = cross(vectorU[i], p);
			for (int c = 0; c < 3; ++c) vectorU[i][c] *= -1;
			normalize(vectorU[i]);
			normalize(vectorV[i]);
			std::cout << " u" << i << " = " << vectorU[i][2] << "," << vectorU[i][1] << "," << vectorU[i][0] << "\n";
			std::cout << " v" << i << " = " << vectorV[i][2] << "," << vectorV[i][1] << "," << vectorV[i][0] << "\n";
		}
	}

	void moveMeasuringPoints(const std::vector<PositionVec>& uvwDisplacement) {
		if (uvwDisplacement.size() != measuringPositions.size()) {
			std::cerr << uvwDisplacement.size() << " != " << measuringPositions.size() << " " << originalMeasuringPositions.size() << std::endl;
			throw std::runtime_error("invalid vector of measuring point displacements");
		}
		for (size_t i = 0; i < originalMeasuringPositions.size(); ++i)
			measuringPositions[i] = displaced(i, uvwDisplacement[i][0], uvwDisplacement[i][1]);
		mps_ = measuringPositions;
		measurement_.assign(mps_.size(), std::vector<double>());
	}

	/// sets the lead positions directly, (z,y,x) each (Simulation::setMeasuringPoints, simulator.cpp:400-404)
	void moveMeasuringPointsTo(const std::vector<PositionVec>& positions) {
		if (positions.size() != measuringPositions.size()) throw std::runtime_error("invalid vector of measuring point displacements");
		measuringPositions = positions;
		mps_ = positions;
		measurement_.assign(mps_.size(), std::vector<double>());
	}

	/// original position + du * u + dv * v, evaluated left to right like the reference (sim_lib.h:201)
	PositionVec displaced(size_t i, double du, double dv) const {
		PositionVec r;
		for (int c = 0; c < 3; ++c) r[c] = (originalMeasuringPositions[i][c] + du * vectorU[i][c]) + dv * vectorV[i][c];
		return r;
	}

	void loadShape() {
		ekg::LogTimer tm(std::cerr, "loading shape                               ");
		// no file name: the built-in test shape, like InputLoader::loadShape (simulator.h:646-651)
		if (settings.inputShapeFilename == "") ekg::generate_test_shape(layers_, Z_, Y_, X_);
		else ekg::load_shape_matrix_cached(settings.inputShapeFilename, layers_, Z_, Y_, X_);
		size_t maxLayer = 0;
		for (uint16_t l : layers_) maxLayer = std::max<size_t>(maxLayer, l & 0x0fff);
		targetNumOfAps_ = maxLayer;
		destroyModels();
	}

	size_t requiredAps() const { return targetNumOfAps_; }

	void simExcitationSequence() {
		ensureModel();
		if (settings.inputExcitationSequenceFilename == "") {
			ekg::LogTimer tm(std::cerr, "calculating excitation sequence             ");
			if (sharded()) slabActivation();
			else check(ekg_model_activation(model_, nullptr, nullptr));
		} else {
			ekg::LogTimer tm(std::cerr, "loading excitation sequence                 ");
			std::vector<double> delay;
			ekg::load_delay_matrix(settings.inputExcitationSequenceFilename, Z_, Y_, X_, delay);
			if (sharded()) forEachSlab([&](size_t d) { if (ekg_model_set_activation(slabModels_[d], delay.data()) != EKG_OK) throw std::runtime_error(ekg_last_error()); });
			else check(ekg_model_set_activation(model_, delay.data()));
		}
		haveActivation_ = true;
		delayCache_.clear();
		classFirstVoxel_.clear();
		if (settings.outputExcitationSequence != "") {
			std::vector<double> delay((size_t)(Z_ * Y_ * X_));
			check(ekg_model_get_activation(model_, delay.data()));
			ekg::export_delay_matrix(settings.outputExcitationSequence, Z_, Y_, X_, delay);
		}
	}

	void printSettings() {
		std::cout << "model: " << settings.inputShapeFilename << " (" << Z_ << ", " << Y_ << ", " << X_ << ")\n";
		std::cout << "neighbourhood: ";
		const std::string& t = settings.neighbourhoodType;
		if (t == "2D4") std::cout << "2D, 4 neighbours\n";
		else if (t == "2D8") std::cout << "2D, 8 neighbours\n";
		else if (t == "3D4") std::cout << "3D, 6 neighbours\n";
		else if (t == "3D8" || t == "cube") std::cout << "3D, 26 neighbours (full cube)\n";
		std::cout << "simulation start = " << settings.simulationStart << "\n" << "simulation time step = " << settings.simulationTimeStep << "\n"
		          << "simulation length = " << settings.simulationLength << "\n" << "total steps = "
		          << settings.simulationLength / settings.simulationTimeStep << "\n";
	}

	void printMessages(std::ostream&) {}

	/// full simulation of the current layer APs and lead positions (Simulation::run)
	void run() {
		if (aps_.size() != targetNumOfAps_) throw std::runtime_error("run: need exactly one action potential per layer");
		std::vector<double> k(aps_.size() * 9);
		for (size_t l = 0; l < aps_.size(); ++l) {
			// as stored: setApIndices builds every cell AP with init(aps[layer], delay), which keeps the layer AP's k8 as it
			// is (already shifted by the layer AP's own `at`, if a caller set one) and replaces `at` (simulator.cpp:161-165, :600)
			std::copy(aps_[l].getK(), aps_[l].getK() + 9, k.begin() + 9 * l);
		}
		std::vector<double> leads(mps_.size() * 3);
		for (size_t i = 0; i < mps_.size(); ++i) for (int c = 0; c < 3; ++c) leads[3 * i + c] = mps_[i][c];
		std::vector<double> ecg;
		runBatch(k.data(), leads.data(), 1, ecg);
		const size_t T = ecg.size() / std::max<size_t>(mps_.size(), 1);
		measurement_.assign(mps_.size(), std::vector<double>());
		for (size_t i = 0; i < mps_.size(); ++i) measurement_[i].assign(ecg.begin() + i * T, ecg.begin() + (i + 1) * T);
		cellApsPending_ = true;
	}

	/// B parameter sets in one launch: layerK [B][layers][9], leadsZyx [B][leads][3] -> ecg [B][leads][T]
	void runBatch(const double* layerK, const double* leadsZyx, size_t B, std::vector<double>& ecg) {
		ensureModel();
		if (!haveActivation_) throw std::runtime_error("excitation sequence missing: call simExcitationSequence() first");
		if (nbhd_ < 0) throw std::runtime_error("applySettings() must be called before run()");
		startTime_ = settings.simulationStart;
		const size_t T = (size_t)std::ceil(settings.simulationLength / timeStep_);
		ecg.assign(B * mps_.size() * T, 0.0);
		const double t0 = ekg::wall_seconds();
		if (!sharded()) {
			check(ekg_simulate(model_, layerK, leadsZyx, (int64_t)B, (int64_t)mps_.size(), nbhd_, (double)settings.simulationStart, timeStep_,
			                   (double)settings.simulationLength, mode_, ecg.data()));
		} else {
			// every device sums its slab; the partial ECGs are added in slab order (the same sum whatever the thread timing)
			std::vector<std::vector<double>> part(slabModels_.size(), std::vector<double>(ecg.size()));
			forEachSlab([&](size_t d) {
				if (ekg_simulate(slabModels_[d], layerK, leadsZyx, (int64_t)B, (int64_t)mps_.size(), nbhd_, (double)settings.simulationStart, timeStep_,
				                 (double)settings.simulationLength, mode_, part[d].data()) != EKG_OK)
					throw std::runtime_error(ekg_last_error());
			});
			for (const std::vector<double>& p : part) for (size_t i = 0; i < ecg.size(); ++i) ecg[i] += p[i];
		}
		lastRunSeconds_ = ekg::wall_seconds() - t0;
	}

	/// layer-AP construction on the device (new): borderK [B][nBorder][9] -> layerK [B][layers][9]
	void fitLayers(const double* borderK, size_t B, size_t nBorder, size_t mid, const double* d9, double step, double eps, int iterations,
	               std::vector<double>& layerK) {
		ensureModel();
		layerK.assign(B * targetNumOfAps_ * 9, 0.0);
		check(ekg_fit_layers(model_, borderK, (int64_t)B, (int64_t)nBorder, (int64_t)mid, d9, step, eps, iterations, layerK.data()));
	}

	/// fit + simulation in one device pass (new): borderK [B][nBorder][9], leadsZyx [B][leads][3] -> ecg [B][leads][T]
	/// and, if asked for, layerK [B][layers][9]
	void evaluateBatch(const double* borderK, size_t nBorder, size_t mid, const double* d9, double step, double eps, int iterations,
	                   const double* leadsZyx, size_t B, std::vector<double>& ecg, std::vector<double>* layerK) {
		ensureModel();
		if (!haveActivation_) throw std::runtime_error("excitation sequence missing: call simExcitationSequence() first");
		if (nbhd_ < 0) throw std::runtime_error("applySettings() must be called before run()");
		startTime_ = settings.simulationStart;
		if (sharded()) {   // the fit on the first device, the simulation over the slabs
			std::vector<double> k;
			fitLayers(borderK, B, nBorder, mid, d9, step, eps, iterations, k);
			runBatch(k.data(), leadsZyx, B, ecg);
			if (layerK) layerK->swap(k);
			return;
		}
		const size_t T = (size_t)std::ceil(settings.simulationLength / timeStep_);
		ecg.assign(B * mps_.size() * T, 0.0);
		if (layerK) layerK->assign(B * targetNumOfAps_ * 9, 0.0);
		const double t0 = ekg::wall_seconds();
		check(ekg_evaluate(model_, borderK, (int64_t)nBorder, (int64_t)mid, d9, step, eps, iterations, leadsZyx, (int64_t)B,
		                   (int64_t)mps_.size(), nbhd_, (double)settings.simulationStart, timeStep_, (double)settings.simulationLength, mode_,
		                   nullptr, 0, nullptr, 0, nullptr, layerK ? layerK->data() : nullptr, ecg.data()));
		lastRunSeconds_ = ekg::wall_seconds() - t0;
	}

	/// fit + simulation + curve comparison in one device pass (new): only B x leads criteria come back.
	/// targets [leads][nTarget] (normalised / resampled by the caller), comparison = the ini's ECG comparison mode
	void evaluateBatchCriteria(const double* borderK, size_t nBorder, size_t mid, const double* d9, double step, double eps, int iterations,
	                           const double* leadsZyx, size_t B, const double* targets, size_t nTarget, const double* targetOffsets,
	                           int comparison, std::vector<double>& criteria) {
		ensureModel();
		if (!haveActivation_) throw std::runtime_error("excitation sequence missing: call simExcitationSequence() first");
		if (nbhd_ < 0) throw std::runtime_error("applySettings() must be called before run()");
		if (sharded()) throw std::runtime_error("evaluateBatchCriteria: not available on a z-slab sharded simulator (use evaluateBatch)");
		startTime_ = settings.simulationStart;
		criteria.assign(B * mps_.size(), 0.0);
		const double t0 = ekg::wall_seconds();
		check(ekg_evaluate(model_, borderK, (int64_t)nBorder, (int64_t)mid, d9, step, eps, iterations, leadsZyx, (int64_t)B,
		                   (int64_t)mps_.size(), nbhd_, (double)settings.simulationStart, timeStep_, (double)settings.simulationLength, mode_,
		                   targets, (int64_t)nTarget, targetOffsets, comparison, criteria.data(), nullptr, nullptr));
		lastRunSeconds_ = ekg::wall_seconds() - t0;
	}

	/// "string model": endo delayed minus epi, layer APs only (Simulation::runApproximation)
	void runApproximation(double delay, std::vector<double>& result) {
		result.assign((size_t)(settings.simulationLength / timeStep_), 0.0);
		for (size_t i = 0; i < result.size(); ++i) {
			const double simTime = settings.simulationStart + i * timeStep_;
			result[i] = aps_[0](simTime + delay) - aps_.back()(simTime);
		}
	}

	const std::vector<double>& getMeasurement(size_t n) const { return measurement_[n]; }
	size_t numMeasurements() const { return mps_.size(); }
	const std::vector<PositionVec>& measuringPoints() const { return mps_; }

	void saveMeasurement(const std::string& comment = "") {
		ekg::LogTimer tm(std::cerr, "saving measurement                          ");
		saveMeasurements(settings.outputFilename, comment);
	}
	void saveMeasurementAs(const char* fname, const std::string& comment = "") {
		ekg::LogTimer tm(std::cerr, "saving measurement                          ");
		saveMeasurements(fname, comment);
	}

	/// voxel (layer, delay, class index) by raster index, after the excitation sequence exists
	ShapeElement shapeElementAtIndex(size_t i) {
		ensureModel();
		ShapeElement e{(size_t)(layers_[i] & 0x0fff), 0.0, (size_t)(layers_[i] & 0x0fff)};
		if (haveActivation_) {
			if (delayCache_.empty()) { delayCache_.resize(layers_.size()); check(ekg_model_get_activation(model_, delayCache_.data())); }
			e.excitationDelay = delayCache_[i];
		}
		return e;
	}
	void modelSize(int64_t& Z, int64_t& Y, int64_t& X) const { Z = Z_; Y = Y_; X = X_; }

private:
	std::vector<double> delayCache_;
	std::vector<int64_t> classFirstVoxel_;
	bool cellApsPending_ = false;

	void saveMeasurements(const std::string& filename, const std::string& comment) {
		std::vector<saveVecElement> v(mps_.size());
		for (size_t i = 0; i < mps_.size(); ++i) {
			std::ostringstream name;
			name << mps_[i][2] << "," << mps_[i][1] << "," << mps_[i][0];
			v[i].name = name.str();
			v[i].data = &measurement_[i];
		}
		if (!mps_.empty()) exportVectors(v, filename, startTime_, timeStep_, comment);
	}

	/// after run() the reference's aps hold one AP per (layer, delay) class in first-seen raster
	/// order (Simulation::setApIndices, simulator.cpp:561-621); `-out cell_aps n` indexes that table
	void buildCellAps() {
		if (classFirstVoxel_.empty()) {   // once per excitation sequence: first voxel of every class, raster order
			std::vector<int64_t> idx(layers_.size());
			int64_t K = 0;
			check(ekg_model_ap_classes(model_, idx.data(), &K));
			if (delayCache_.empty()) { delayCache_.resize(layers_.size()); check(ekg_model_get_activation(model_, delayCache_.data())); }
			classFirstVoxel_.assign((size_t)K, -1);
			for (size_t i = 0; i < layers_.size(); ++i)
				if (idx[i] >= 0 && classFirstVoxel_[(size_t)idx[i]] < 0) classFirstVoxel_[(size_t)idx[i]] = (int64_t)i;
		}
		std::vector<ActionPotential> cells(classFirstVoxel_.size());
		for (size_t c = 0; c < cells.size(); ++c) {
			const size_t i = (size_t)classFirstVoxel_[c];
			cells[c].init(aps_[(layers_[i] & 0x0fff) - 1], delayCache_[i]);
		}
		aps_.swap(cells);
		cellApsPending_ = false;
	}

	EkgSim(const EkgSim&);
	void operator=(const EkgSim&);
};

}  // namespace SimLib

typedef SimLib::EkgSim EkgSim;
