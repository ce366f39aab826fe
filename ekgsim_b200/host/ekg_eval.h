// ekgsim_b200/host/ekg_eval.h -- evaluation glue of the B200 build: parameter vector -> 24 layer
// APs -> simulation (GPU) -> criteria.  Newly written host C++; the arithmetic follows the
// reference step by step so that criteria and violations agree with the reference's own
// (paths relative to synergy-twinning/ekgsim):
//
//   settings                       SimSettings.h:158-356
//   eval                           sim.cpp:443-491      (SimImplementation::eval)
//   border / mid layer APs         sim.cpp:750-916      (simUsingBorderAps, simUsingBorderAndMidAps)
//   connector points of the fit    sim.cpp:91-313       (WohlfartInterpolationEvaluator)
//   steepest descent               nonlinearFit.h:92-168
//   criteria                       sim.cpp:600-702      (calculateFitness), vectorMath.h:209-404
//   targets                        sim.cpp:999-1042     (loadTargets)
//   outputs                        sim.cpp:918-991      (OutputAps, writeOutputs)
//
// New relative to the reference: evalBatch() evaluates many parameter vectors at once -- the 21
// per-layer fits of every individual run on host threads, all simulations go to the GPU in one
// batched launch (ekg_simulate with B > 1), criteria are computed per individual afterwards.
#pragma once

#include <set>
#include <thread>

#include "sim_lib.h"

namespace ekg {

// ---- vector statistics used by the criteria (vectorMath.h) ---------------------------------------------
inline double sqr(double v) { return v * v; }

inline void mean_and_var(double& mean, double& var, const std::vector<double>& src, size_t start, size_t end) {
	const double inverseSize = 1.0 / double(end - start);
	double s = 0;
	for (size_t i = start; i < end; ++i) s += src[i];
	mean = s * inverseSize;
	var = 0;
	for (size_t i = start; i < end; ++i) var += sqr(src[i] - mean);
	var *= inverseSize;
}

inline void min_and_max(double& mn, double& mx, const std::vector<double>& src) {
	if (src.empty()) return;
	mn = mx = src[0];
	for (size_t i = 1; i < src.size(); ++i) {
		if (src[i] < mn) mn = src[i];
		else if (src[i] > mx) mx = src[i];
	}
}

/// linear resampling by `factor`; a factor of exactly 1 copies and ignores the offset (vectorMath.h:211-212)
inline void resample(const std::vector<double>& src, std::vector<double>& dest, double factor, int startOffset = 0) {
	if (factor == 1) { dest = src; return; }
	dest.resize(1u + (size_t)std::ceil((src.size() - 1 - startOffset) * factor));
	for (size_t i = 0; i < dest.size(); ++i) {
		const double oldPos = i / factor + startOffset;
		const size_t fPos = (size_t)std::floor(oldPos);
		const size_t cPos = (size_t)(1.0 + fPos);
		if (cPos < src.size()) dest[i] = src[cPos] * (oldPos - fPos) + src[fPos] * (cPos - oldPos);
		else dest[i] = src.back();
	}
}

struct Match { double value = 0; int offset = 0; };

inline size_t overlap_len(const std::vector<double>& a, const std::vector<double>& b, int offs) {
	return std::min(a.size(), b.size() + offs) - std::max(0, offs);
}

/// Pearson correlation of the overlapping parts, b shifted by offs >= 0 (vectorMath.h:287-317)
inline Match pearson(const std::vector<double>& a, const std::vector<double>& b, int offs) {
	Match r; r.offset = offs;
	double meanA, varA, meanB, varB;
	mean_and_var(meanA, varA, a, std::max(0, offs), std::min(a.size(), b.size() + offs));
	const double sdA = std::sqrt(varA);
	mean_and_var(meanB, varB, b, std::max(0, -offs), std::min(b.size(), a.size() - offs));
	const double sdB = std::sqrt(varB);
	const size_t len = overlap_len(a, b, offs);
	double cov = 0;
	for (size_t i = std::max(0, offs), j = std::max(0, -offs), n = 0; n < len; ++i, ++j, ++n) cov += (a[i] - meanA) * (b[j] - meanB);
	cov /= double(len);
	r.value = cov / (sdA * sdB);
	return r;
}

inline Match rms_match(const std::vector<double>& a, const std::vector<double>& b, int offs) {
	Match r; r.offset = offs;
	const size_t len = overlap_len(a, b, offs);
	std::vector<double> d(len);
	for (size_t i = 0; i < len; ++i) { d[i] = a[std::max(0, offs) + i] - b[std::max(0, -offs) + i]; d[i] *= d[i]; }
	double mean, var;
	mean_and_var(mean, var, d, 0, len);
	r.value = std::sqrt(mean);
	return r;
}

inline Match vector_correlation(const std::vector<double>& a, const std::vector<double>& b, int offs) {
	Match r; r.offset = offs;
	const size_t len = overlap_len(a, b, offs);
	const size_t ia = std::max(0, offs), ib = std::max(0, -offs);
	double ab = 0, aa = 0, bb = 0;
	for (size_t n = 0; n < len; ++n) ab += a[ia + n] * b[ib + n];
	for (size_t n = 0; n < len; ++n) aa += a[ia + n] * a[ia + n];
	for (size_t n = 0; n < len; ++n) bb += b[ib + n] * b[ib + n];
	r.value = ab / (std::sqrt(aa) * std::sqrt(bb));
	return r;
}

inline Match dev_from_linear(const std::vector<double>& a, const std::vector<double>& b, double bOfs, int offs) {
	Match r; r.offset = offs;
	const size_t len = overlap_len(a, b, offs);
	std::vector<double> d(len);
	double aMin = 0, aMax = 0;
	min_and_max(aMin, aMax, a);
	const double aMult = 1.0 / (aMax - aMin);
	for (size_t i = std::max(0, offs), j = std::max(0, -offs), k = 0; k < len; ++k, ++i, ++j) d[k] = (a[i] * aMult + bOfs) / (b[j] + bOfs);
	double mean, var;
	mean_and_var(mean, var, d, 0, len);
	r.value = var == 0 ? 1 / 1e-30 : std::sqrt(var);
	return r;
}

// ---- optimisation / evaluation settings (SimSettings.h) ---------------------------------------------------
struct EvalSettings {
	enum CriteriaMode { every_lead = 1, leads_sum = 9 };
	enum ComparisonMode { cmp_rms = 1, cmp_correlation = 2, cmp_norm_offset_div_var = 3, cmp_vector_correlation = 4 };

	std::vector<SimLib::WohlfartPlus> baseAps;
	std::vector<double> kMin, kMax;
	std::string interpolationTypeString;
	std::vector<int> freeKs;
	bool measuringPointsDisplacementIsInput = false;
	std::vector<int> displacementMin, displacementMax;
	std::string optimizationTargetsFname;
	ComparisonMode comparisonMode = cmp_correlation;
	CriteriaMode criteriaMode = every_lead;
	double endoEpiMinCriterionDelay = -1;
	size_t numGenerations = 100, populationSize = 0;
	int queueSize = 1;
	double midPosition = 0.5;
	bool peakPositionIsCriterion = false, fastApproxIsCriterion = false;
	double fastApproxEpiDelay = 0, fastApproxLimit = 2;

	template <class T>
	static void bracket_list(std::ostream& o, const std::vector<T>& v) {
		o << "[";
		for (size_t i = 0; i < v.size(); ++i) o << (i ? "," : "") << v[i];
		o << "]";
	}

	void load(const char* fname) {
		std::cerr << "***** loading additional simulation parameters *************\n";
		IniFile ini(fname);
		if (ini.found()) std::cerr << " from file: " << fname << "\n";
		else std::cerr << " file " << fname << "  not found, \n";
		const int ap = ini.section("wohlfart ap");
		for (size_t i = 0;; ++i) {
			std::ostringstream name;
			name << "base ap " << (i + 1);
			std::vector<double> k;
			if (ini.load_array(k, name.str(), ap, 9) && k.size() == 9) {
				baseAps.push_back(SimLib::WohlfartPlus());
				baseAps.back().setK(k);
				for (size_t j = 0; j < 9; ++j) std::cerr << " k" << j << "=" << k[j];
				std::cerr << "\n";
			} else break;
		}
		if (baseAps.size() < 2) throw std::runtime_error("need at least 2 base APs");
		if (!ini.load_array(kMin, "k min", ap, 9) || kMin.size() != 9) throw std::runtime_error("could not read kMin");
		std::cerr << " k min = "; bracket_list(std::cerr, kMin); std::cerr << "\n";
		if (!ini.load_array(kMax, "k max", ap, 9) || kMax.size() != 9) throw std::runtime_error("could not read kMax");
		std::cerr << " k max = "; bracket_list(std::cerr, kMax); std::cerr << "\n";
		ini.load(interpolationTypeString, "interpolation", ap);
		std::cerr << " AP interpolation set to " << interpolationTypeString << " (string validity not checked yet)\n";
		midPosition = 0.5;
		ini.load(midPosition, "mid AP position", ap);
		if (midPosition > 1.0 || midPosition < 0.0) {
			std::cerr << " warning, mid AP position must be in range [0..1] but is set to " << midPosition
			          << " in settings; using default value of 0.5 instead\n";
			midPosition = 0.5;
		}
		ini.load_array(freeKs, "free k", ap);
		std::cerr << " free Wohlfart koefficients "; bracket_list(std::cerr, freeKs); std::cerr << "\n\n";

		const int mp = ini.section("measuring points");
		if (!ini.load_array(displacementMin, "displacement min", mp)) throw std::runtime_error("could not read displacement min");
		std::cerr << " displacement min = "; bracket_list(std::cerr, displacementMin); std::cerr << "\n";
		if (!ini.load_array(displacementMax, "displacement max", mp)) throw std::runtime_error("could not read displacement max");
		std::cerr << " displacement max = "; bracket_list(std::cerr, displacementMax); std::cerr << "\n";

		std::cerr << "***** loading optimization parameters **********************\n";
		const int opt = ini.section("optimization");
		ini.load(optimizationTargetsFname, "targets filename", opt);
		std::cerr << " targets filename = " << optimizationTargetsFname << "\n";
		int crit = 1;
		ini.load(crit, "mode", opt);
		if (crit != every_lead && crit != leads_sum) throw std::runtime_error("invalid optimization mode");
		criteriaMode = (CriteriaMode)crit;
		std::cerr << " criterization mode = " << criteriaMode << "\n";
		endoEpiMinCriterionDelay = -1;
		ini.load(endoEpiMinCriterionDelay, "endo-epi minimization criterion epi delay", opt);
		if (endoEpiMinCriterionDelay > 0)
			std::cerr << "enabling additional criterion - endo-epi minimization with epi delay of " << endoEpiMinCriterionDelay << "\n";
		int tmp = 0;
		ini.load(tmp, "peak position is criterion", opt);
		if (tmp > 0) std::cerr << "enabling additional criteria - peak position for every base\n";
		peakPositionIsCriterion = tmp > 0;
		measuringPointsDisplacementIsInput = false;
		ini.load(measuringPointsDisplacementIsInput, "optimize measuring points", opt);
		std::cerr << " optimize measuring points = " << measuringPointsDisplacementIsInput << "\n";
		tmp = 0;
		ini.load(tmp, "fast approximation is criterion", opt);
		if (tmp > 0) std::cerr << "enabling additional criteron - fast approximation of an ECG\n";
		fastApproxIsCriterion = tmp > 0;
		fastApproxEpiDelay = 0;
		ini.load(fastApproxEpiDelay, "fast approximation epi delay", opt);
		fastApproxLimit = 2;
		ini.load(fastApproxLimit, "fast approximation limit", opt);
		int cmode = 2;
		ini.load(cmode, "comparison mode", opt);
		if (cmode < 1 || cmode > 4) throw std::runtime_error("invalid comparison mode");
		comparisonMode = (ComparisonMode)cmode;
		std::cerr << " ECG comparison mode = " << comparisonMode << "\n";
		numGenerations = 100;
		ini.load(numGenerations, "number of generations", opt);
		std::cerr << " number of generations = " << numGenerations << "\n";
		populationSize = 0;
		ini.load(populationSize, "population size", opt);
		std::cerr << " population size = ";
		if (populationSize == 0) std::cerr << "automatic\n"; else std::cerr << populationSize << "\n";
		queueSize = 1;
		ini.load(queueSize, "queue size", opt);
		std::cerr << " queue size = " << queueSize << "\n\n";
	}
};

struct OutputSettings {
	std::vector<size_t> outputCellAps;
	bool layerAps = false, excitationSequence = false, result = false;
};

// ---- layer-AP construction ----------------------------------------------------------------------------------
typedef SimLib::ActionPotential AP;

inline double ap_slope(double x, const AP& ap) { return 100 * (ap(x + 0.01) - ap(x)); }

/// Target points for the per-layer fit: 15 straight connectors between matching arc-length
/// positions on two border APs; a layer at `ratio` between them should pass through the points at
/// that ratio along the connectors (sim.cpp:91-313).
class LayerFitTarget {
	struct Connector {
		double k, n, x1, x2;
		double at(double x) const { return k * x + n; }
		double x_at(double rel) const { return x1 + (x2 - x1) * rel; }
	};
	std::vector<Connector> connectors_;
	std::vector<double> points_;  // x0,y0,x1,y1,...

	static size_t clamp700(double apd) {
		const double c = std::ceil(apd);
		// an AP that never repolarises reports apd90 = -1; the reference's size_t cast of that wraps to a
		// huge value, i.e. the 700-sample cap applies
		return c < 0 ? (size_t)700 : std::min((size_t)700, (size_t)c);
	}

public:
	void setBorderAps(const AP& ap1, const AP& ap2) {
		connectors_.clear();
		const size_t numPoints = 15, startX = 10;
		const double xScale = 0.1;
		const double apd1 = ap1.wohl.apd90(), apd2 = ap2.wohl.apd90();
		// one spare element: the arc-length walk below reads y[i+1] with i up to 699
		std::vector<double> y1(701), y2(701);
		for (size_t i = 0; i < 701; ++i) { y1[i] = ap1((double)i); y2[i] = ap2((double)i); }
		double len1 = 0, len2 = 0;
		for (size_t i = startX + 1; i < clamp700(apd1); ++i) len1 += std::sqrt(sqr(y1[i] - y1[i - 1]) + xScale);
		for (size_t i = startX; i < clamp700(apd2); ++i) len2 += std::sqrt(sqr(y2[i] - y2[i - 1]) + xScale);
		double cur1 = 0, cur2 = 0;
		size_t i1 = startX, i2 = startX;
		for (size_t i = 0; i < numPoints; ++i) {
			double target = i * len1 / (numPoints - 2);
			for (double l = cur1; l < target && i1 < 700; ++i1) l += std::sqrt(sqr(y1[i1 + 1] - y1[i1]) + xScale);
			cur1 = target;
			target = i * len2 / (numPoints - 2);
			for (double l = cur2; l < target && i2 < 700; ++i2) l += std::sqrt(sqr(y2[i2 + 1] - y2[i2]) + xScale);
			cur2 = target;
			Connector c;
			c.x1 = (double)i1;
			c.x2 = (double)i2;
			if (i1 == i2) c.x2 += 0.001;
			c.k = (y1[i1] - y2[i2]) / (c.x1 - c.x2);
			c.n = y2[i2] - c.k * i2;
			connectors_.push_back(c);
		}
	}

	void setupRatio(double ratio) {
		points_.clear();
		for (const Connector& c : connectors_) {
			points_.push_back(c.x_at(ratio));
			points_.push_back(c.at(points_.back()));
		}
	}

	const std::vector<double>& points() const { return points_; }

	/// sum of squared misses
	double operator()(const AP& ap) const {
		double sum = 0;
		for (size_t i = 0; i < points_.size(); i += 2) sum += sqr(ap(points_[i]) - points_[i + 1]);
		return sum;
	}
};

/// Per-point factors of the AP at the current iterate, kept so that a one-coefficient perturbation
/// only recomputes the factor that coefficient enters.  Every factor is evaluated by exactly the
/// expression WohlfartPlus::operator[] uses and the factors are multiplied in its order
/// ((A*B)*C)+k0, so the values are bit-identical to a full evaluation.
struct FitCache {
	std::vector<double> A, E4, E5, Q;   // 1/(1+e^{-k1 t}),  e^{-k4 t},  e^{-k5 t},  (1+e^{-k7(t-k8)+c})^{-k6/k7}
	void resize(size_t n) { A.resize(n); E4.resize(n); E5.resize(n); Q.resize(n); }
};

inline double fit_tail_c(const double* k) { return std::log(std::pow(2, (k[7] / k[6])) - 1); }
inline double fit_A(const double* k, double t) { return 1.0 / (1.0 + std::exp(-k[1] * t)); }
inline double fit_Q(const double* k, double c, double t) { return std::pow((1 + std::exp(-k[7] * (t - k[8]) + c)), -(k[6] / k[7])); }
inline double fit_value(const double* k, double A, double E4, double E5, double Q) {
	return A * (k[2] * ((1.0 - k[3]) * E4 + k[3])) * (E5 * (1 - Q)) + k[0];
}

/// f(x) with all factors computed (and remembered in `c`)
inline double fit_eval_full(const LayerFitTarget& f, const AP& x, FitCache& c) {
	const std::vector<double>& p = f.points();
	const double* k = x.getK();
	const double tc = fit_tail_c(k);
	c.resize(p.size() / 2);
	double sum = 0;
	for (size_t i = 0, n = 0; i < p.size(); i += 2, ++n) {
		const double t = p[i] - x.at;
		c.A[n] = fit_A(k, t);
		c.E4[n] = std::exp(-k[4] * t);
		c.E5[n] = std::exp(-k[5] * t);
		c.Q[n] = fit_Q(k, tc, t);
		sum += sqr(fit_value(k, c.A[n], c.E4[n], c.E5[n], c.Q[n]) - p[i + 1]);
	}
	return sum;
}

/// f(x1) where x1 differs from the cached iterate only in coefficient `which`
inline double fit_eval_perturbed(const LayerFitTarget& f, const AP& x1, size_t which, const FitCache& c) {
	const std::vector<double>& p = f.points();
	const double* k = x1.getK();
	const double tc = (which == 6 || which == 7) ? fit_tail_c(k) : 0.0;
	double sum = 0;
	for (size_t i = 0, n = 0; i < p.size(); i += 2, ++n) {
		const double t = p[i] - x1.at;
		double A = c.A[n], E4 = c.E4[n], E5 = c.E5[n], Q = c.Q[n];
		switch (which) {
		case 1: A = fit_A(k, t); break;
		case 4: E4 = std::exp(-k[4] * t); break;
		case 5: E5 = std::exp(-k[5] * t); break;
		case 6: case 7: Q = fit_Q(k, tc, t); break;
		case 8: Q = fit_Q(k, std::log(std::pow(2, (k[7] / k[6])) - 1), t); break;
		default: break;  // k0, k2, k3 only enter fit_value
		}
		sum += sqr(fit_value(k, A, E4, E5, Q) - p[i + 1]);
	}
	return sum;
}

/// sign-following steepest descent with per-coefficient step adaptation (nonlinearFit.h:92-168)
inline int steepest_descent(const LayerFitTarget& f, AP& x0, const AP& d, double stepSize, double epsilon, int iterations) {
	AP grad = x0, oldGrad = x0;
	for (size_t i = 0; i < 9; ++i) oldGrad[i] = 0;
	FitCache cache, trial;
	double y0 = fit_eval_full(f, x0, cache);
	AP move = d;
	for (; iterations > 0 && y0 > epsilon; --iterations) {
		bool stepChange = false;
		for (size_t i = 0; i < 9; ++i) {
			if (d[i] != 0) {
				AP x1 = x0;
				x1[i] += d[i] * .001;
				grad[i] = (fit_eval_perturbed(f, x1, i, cache) - y0) / (d[i] * .001);
				if (grad[i] * oldGrad[i] < 0) { move[i] *= 0.5; stepChange = true; }
				else if (std::fabs(grad[i]) > 0.75 * std::fabs(oldGrad[i])) move[i] *= 1.5;
			} else grad[i] = 0;
		}
		oldGrad = grad;
		AP x1 = x0;
		for (size_t i = 0; i < 9; ++i) x1[i] -= stepSize * ((grad[i] > 0) ? move[i] : -move[i]);
		const double y1 = fit_eval_full(f, x1, trial);
		if (y1 < y0) { y0 = y1; x0 = x1; std::swap(cache, trial); }
		else if (!stepChange) stepSize *= 0.5;
	}
	return iterations;
}

/// "all" | "<count>" | "<id>,<id>,..." -> device ids (new; `-devices` of the CLI, EKGSIM_B200_DEVICES).  An id may repeat:
/// two replicas on one GPU (that is how a single-GPU box exercises the multi-device code).
inline std::vector<int> parse_device_list(const std::string& spec) {
	std::vector<int> ids;
	const int have = ekg_device_count();
	if (spec.empty()) return ids;
	if (spec == "all") { for (int i = 0; i < have; ++i) ids.push_back(i); return ids; }
	if (spec.find(',') == std::string::npos) {
		const int n = atoi(spec.c_str());
		if (n <= 0) throw std::runtime_error("bad device list: " + spec);
		for (int i = 0; i < n; ++i) ids.push_back(i);
	} else {
		std::istringstream in(spec);
		for (std::string tok; std::getline(in, tok, ',');) if (!tok.empty()) ids.push_back(atoi(tok.c_str()));
	}
	for (int id : ids) if (id < 0 || id >= have) throw std::runtime_error("no such CUDA device in device list: " + spec);
	return ids;
}

// ---- the evaluator ---------------------------------------------------------------------------------------------
class Evaluator {
public:
	typedef std::vector<double> Input;
	typedef std::vector<double> Value;

	EvalSettings settings;
	OutputSettings outSettings;
	size_t deducedNumOfCriteria = 0;

private:
	std::unique_ptr<EkgSim> sim;
	/// further GPUs (new): evalBatch splits its batch over `sim` and these, one host thread per device -- individuals
	/// over GPUs without a collective, the in-process form of the reference's MPI task farm (ParallelFramework.h:388-429)
	std::vector<std::unique_ptr<EkgSim>> replicas;
	std::vector<std::vector<double>> targets;
	std::vector<double> targetOffsets;
	enum Interp { unknown = 0, endo_epi = 1, endo_mid_epi = 2 } interp = unknown;
	std::set<char> freeK;
	size_t numDisplacementParams = 0, numWohlfartParams = 0;
	size_t evalCounter = 0;
	/// where the inner-layer fits run: on the device (ekg_fit_layers) whenever the evaluator owns one, unless
	/// EKGSIM_B200_FIT=host asks for the host restatement (which is also what the GPU-less tooling uses)
	bool fitOnDevice = false;

	struct Individual {
		std::vector<AP> layerAps;
		std::vector<EkgSim::PositionVec> leads;
		std::vector<double> approxEcg;
		double approxCriteria = 2;
		double violation = 0;
		bool simulate = true;
	};

public:
	/// withDevice = false builds everything that lives on the host (settings, leads, targets, layer-AP
	/// construction) but never touches the GPU: eval()/evalBatch() then throw, layerCoefficients() works.
	/// `devices`: GPUs to evaluate batches on (empty: EKGSIM_B200_DEVICES if set -- "all", a count or a list --, else the one
	/// device of EKGSIM_B200_DEVICE / 0).  eval() always runs on the first.
	explicit Evaluator(const char* ini = "simulator.ini", bool withDevice = true, std::vector<int> devices = std::vector<int>()) {
		std::cerr << "***** setting up Ekg Simulator *****************************\n";
		settings.load(ini);
		sim.reset(new EkgSim);
		if (withDevice && devices.empty()) if (const char* e = std::getenv("EKGSIM_B200_DEVICES")) devices = parse_device_list(e);
		if (!devices.empty() && !sim->sharded()) sim->setDevice(devices[0]);
		sim->loadSettings(ini);
		sim->loadTransferMatrix();
		sim->loadMeasuringPoints();
		numDisplacementParams = settings.measuringPointsDisplacementIsInput ? sim->numMeasurements() * 2 : 0;
		sim->loadShape();
		if (withDevice) sim->simExcitationSequence();
		const char* fitEnv = std::getenv("EKGSIM_B200_FIT");
		fitOnDevice = withDevice && !(fitEnv && std::string(fitEnv) == "host");
		std::cerr << "\n***** applying (and checking) settings *********************\n";
		sim->applySettings();
		if (settings.interpolationTypeString == "endo-epi") interp = endo_epi;
		else if (settings.interpolationTypeString == "endo-mid-epi") interp = endo_mid_epi;
		else throw std::runtime_error("unknown interpolation type [" + settings.interpolationTypeString + "]");
		// a mid AP position that rounds onto the first or the last layer: the later border AP overwrites the earlier one and
		// the fit proceeds (sim.cpp:833-902); the device fit does not implement that corner, the host restatement does
		if (fitOnDevice && interp == endo_mid_epi && sim->requiredAps() > 2 && (midLayer() == 0 || midLayer() == sim->requiredAps() - 1)) fitOnDevice = false;
		for (int k : settings.freeKs) freeK.insert((char)k);
		numWohlfartParams = freeK.size() * (interp == endo_epi ? 2 : 3);
		std::cerr << " ok\n";
		std::cerr << "\n***** printout of the simulator setup **********************\n";
		sim->printSettings();
		std::cerr << "\n***** loading targets and setting up optimization **********\n";
		loadTargets(settings.optimizationTargetsFname.c_str());
		size_t base = 0;
		if (settings.criteriaMode == EvalSettings::every_lead) base = std::min(sim->numMeasurements(), targets.size());
		else if (settings.criteriaMode == EvalSettings::leads_sum) base = 1;
		deducedNumOfCriteria = base;
		std::cerr << " base number of criteria = " << deducedNumOfCriteria << " (every lead is criterion)\n";
		if (settings.peakPositionIsCriterion) { deducedNumOfCriteria += base; std::cerr << "   peak positions (per every lead) are also criteria\n"; }
		if (settings.fastApproxIsCriterion) { ++deducedNumOfCriteria; std::cerr << "   fast approximation is also a criterion\n"; }
		if (settings.endoEpiMinCriterionDelay >= 0) {
			++deducedNumOfCriteria;
			std::cerr << "   endo-epi min delay is also a criterion (" << settings.endoEpiMinCriterionDelay << ")\n";
		}
		std::cerr << " total number of criteria = " << deducedNumOfCriteria << "\n\n";
		if (withDevice && sim->sharded()) {
			std::cerr << " one model on " << sim->numSlabs() << " z-slabs (" << (sim->slabAutomatonLinked() ? "peer-linked" : "replicated") << " excitation sequence)\n\n";
			devices.clear();   // slabs and batch devices are alternatives
		}
		if (withDevice && devices.size() > 1) {
			std::vector<std::string> errors(devices.size());
			std::vector<std::thread> pool;
			replicas.resize(devices.size() - 1);
			for (size_t d = 1; d < devices.size(); ++d)
				pool.emplace_back([&, d]() {
					try { replicas[d - 1] = sim->replicate(devices[d]); } catch (std::exception& e) { errors[d] = e.what(); }
				});
			for (std::thread& t : pool) t.join();
			for (const std::string& e : errors) if (!e.empty()) throw std::runtime_error(e);
			std::cerr << " evaluating batches on " << devices.size() << " GPUs\n\n";
		}
	}

	EkgSim& simulator() { return *sim; }
	size_t numDevices() const { return 1 + replicas.size(); }
	size_t numGenes() const { return numWohlfartParams + numDisplacementParams; }

	/// one individual: returns the violation, fills `result` with the criteria (SimImplementation::eval)
	double eval(const Input& solution, Value& result) {
		std::cout << "\reval " << ++evalCounter << "  ";
		Individual ind;
		prepare(solution, ind);
		sim->moveMeasuringPointsTo(ind.leads);
		std::vector<std::vector<double>> ecg;
		bool done = false;
		if (ind.simulate) {
			std::vector<AP> aps = ind.layerAps;
			sim->setApsDestructive(aps);
			sim->run();
			ecg.resize(sim->numMeasurements());
			for (size_t l = 0; l < ecg.size(); ++l) ecg[l] = sim->getMeasurement(l);
			done = true;
		} else {
			std::vector<AP> aps = ind.layerAps;
			sim->setApsDestructive(aps);
		}
		const double violation = ind.violation + fitness(ind, ecg, done, result);
		writeOutputs(solution, ind);
		return violation;
	}

	/// many individuals.  Default: border APs + leads are assembled on the host (microseconds), then ONE device
	/// pass fits the inner layers and simulates (ekg_evaluate); criteria per individual from the returned ECGs.
	/// With the host fit selected (EKGSIM_B200_FIT=host) the fits run on host threads and one ekg_simulate follows.
	void evalBatch(const std::vector<Input>& solutions, std::vector<Value>& results, std::vector<double>& violations, int threads = 0) {
		const size_t B = solutions.size();
		results.assign(B, Value());
		violations.assign(B, 0.0);
		std::vector<Individual> inds(B);
		if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
		threads = (int)std::min<size_t>(threads, std::max<size_t>(B, 1));
		const bool gateActive = settings.fastApproxIsCriterion || settings.fastApproxLimit < 2;
		if (fitOnDevice && !gateActive) {
			// nothing here is worth a thread: ~1 us per individual
			for (size_t i = 0; i < B; ++i) { prepareBorders(solutions[i], inds[i]); inds[i].approxCriteria = 2; inds[i].simulate = !(2 > settings.fastApproxLimit); }
		} else {
			std::vector<std::string> errors(threads);
			std::vector<std::thread> pool;
			for (int w = 0; w < threads; ++w)
				pool.emplace_back([&, w]() {
					try {
						for (size_t i = w; i < B; i += threads) {
							prepareBorders(solutions[i], inds[i]);
							if (!fitOnDevice) fitLayersOnHost(inds[i]);
							approximationGate(inds[i]);
						}
					} catch (std::exception& e) { errors[w] = e.what(); }
				});
			for (std::thread& t : pool) t.join();
			for (const std::string& e : errors) if (!e.empty()) throw std::runtime_error(e);
		}
		// gather the individuals that pass the fast-approximation gate into one launch
		std::vector<size_t> run;
		for (size_t i = 0; i < B; ++i) if (inds[i].simulate) run.push_back(i);
		const size_t nl = sim->requiredAps(), L = sim->numMeasurements(), nb = numBorders();
		std::vector<double> k(run.size() * (fitOnDevice ? nb : nl) * 9), leads(run.size() * L * 3), ecg;
		for (size_t r = 0; r < run.size(); ++r) {
			const Individual& ind = inds[run[r]];
			if (fitOnDevice) borderCoefficients(ind, &k[r * nb * 9]);
			else for (size_t l = 0; l < nl; ++l) std::copy(ind.layerAps[l].getK(), ind.layerAps[l].getK() + 9, k.begin() + (r * nl + l) * 9);
			for (size_t m = 0; m < L; ++m) for (int c = 0; c < 3; ++c) leads[(r * L + m) * 3 + c] = ind.leads[m][c];
		}
		size_t T = 0;
		if (!run.empty() && fitOnDevice && criteriaOnDevice()) {
			// plain settings (one criterion per lead, no peak-position criterion): the comparison runs on the device too,
			// B x leads doubles come back instead of the ECGs
			std::vector<double> tg(L * targets[0].size()), crit(run.size() * L);
			for (size_t m = 0; m < L; ++m) std::copy(targets[m].begin(), targets[m].end(), tg.begin() + m * targets[0].size());
			onDevices(run.size(), [&](EkgSim& s, size_t b0, size_t b1) {
				s.moveMeasuringPointsTo(inds[run[b0]].leads);
				std::vector<double> part;
				s.evaluateBatchCriteria(&k[b0 * nb * 9], nb, interp == endo_epi ? 0 : midLayer(), fitOffsets(), 0.5, 1e-3, 100, &leads[b0 * L * 3],
				                        b1 - b0, tg.data(), targets[0].size(), targetOffsets.data(), (int)settings.comparisonMode, part);
				std::copy(part.begin(), part.end(), crit.begin() + b0 * L);
			});
			std::vector<size_t> slotOf(B, (size_t)-1);
			for (size_t r = 0; r < run.size(); ++r) slotOf[run[r]] = r;
			for (size_t i = 0; i < B; ++i) {
				if (slotOf[i] == (size_t)-1) { violations[i] = inds[i].violation + fitness(inds[i], {}, false, results[i]); continue; }
				violations[i] = inds[i].violation + fitnessFromCriteria(inds[i], &crit[slotOf[i] * L], results[i]);
			}
			evalCounter += B;
			return;
		}
		if (!run.empty()) {
			const size_t kPer = (fitOnDevice ? nb : nl) * 9;
			const EkgSim& s0 = *sim;
			T = (size_t)std::ceil(s0.settingsView().simulationLength / s0.settingsView().simulationTimeStep);
			ecg.assign(run.size() * L * T, 0.0);
			onDevices(run.size(), [&](EkgSim& s, size_t b0, size_t b1) {
				s.moveMeasuringPointsTo(inds[run[b0]].leads);  // lead count / bookkeeping; positions travel in `leads`
				std::vector<double> part;
				if (fitOnDevice) s.evaluateBatch(&k[b0 * kPer], nb, interp == endo_epi ? 0 : midLayer(), fitOffsets(), 0.5, 1e-3, 100, &leads[b0 * L * 3], b1 - b0, part, nullptr);
				else s.runBatch(&k[b0 * kPer], &leads[b0 * L * 3], b1 - b0, part);
				std::copy(part.begin(), part.end(), ecg.begin() + b0 * L * T);
			});
		}
		std::vector<size_t> slot(B, (size_t)-1);
		for (size_t r = 0; r < run.size(); ++r) slot[run[r]] = r;
		for (size_t i = 0; i < B; ++i) {
			std::vector<std::vector<double>> e;
			if (slot[i] != (size_t)-1) {
				e.resize(L);
				for (size_t m = 0; m < L; ++m) e[m].assign(ecg.begin() + (slot[i] * L + m) * T, ecg.begin() + (slot[i] * L + m + 1) * T);
			}
			violations[i] = inds[i].violation + fitness(inds[i], e, slot[i] != (size_t)-1, results[i]);
		}
		evalCounter += B;
	}

	/// the 24x9 layer coefficients the glue derives from a parameter vector (for tests / tooling)
	void layerCoefficients(const Input& solution, std::vector<double>& k, std::vector<double>& leadsZyx, double& violation) {
		Individual ind;
		prepare(solution, ind);
		k.clear();
		for (const AP& a : ind.layerAps) k.insert(k.end(), a.getK(), a.getK() + 9);
		leadsZyx.clear();
		for (const EkgSim::PositionVec& p : ind.leads) for (int c = 0; c < 3; ++c) leadsZyx.push_back(p[c]);
		violation = ind.violation;
	}

	void getGeneParams(size_t& nGenes, size_t& nCriteria, std::vector<double>& gMin, std::vector<double>& gMax) {
		const size_t reps = interp == endo_epi ? 2 : 3;
		gMin.assign(freeK.size() * reps, 0);
		gMax.assign(freeK.size() * reps, 0);
		size_t g = 0;
		std::cout << "-----------------\n-- gene limits --\n num free params = " << freeK.size() << "\n";
		for (size_t i = 0; i < 9; ++i) {
			if (!freeK.count((char)i)) continue;
			for (size_t r = 0; r < reps; ++r) { gMin[g + r * freeK.size()] = settings.kMin[i]; gMax[g + r * freeK.size()] = settings.kMax[i]; }
			std::cout << " k" << i << "=[" << gMin[g] << "..." << gMax[g] << "]\n";
			++g;
		}
		std::cout << "\n-----------------\n";
		nGenes = freeK.size() * reps;
		nCriteria = deducedNumOfCriteria;
		if (settings.measuringPointsDisplacementIsInput) {
			nGenes += numDisplacementParams;
			if (settings.displacementMin.empty()) settings.displacementMin.push_back(0);
			if (settings.displacementMax.empty()) settings.displacementMax.push_back(0);
			for (int i = 0; i < (int)numDisplacementParams; ++i) {
				gMin.push_back(settings.displacementMin[std::min(i, (int)settings.displacementMin.size() - 1)]);
				gMax.push_back(settings.displacementMax[std::min(i, (int)settings.displacementMax.size() - 1)]);
			}
		}
	}

private:
	/// fn(simulator, begin, end) for a contiguous, balanced share of [0, n) on every device, one host thread per
	/// further device (the handles are independent; the C ABI keeps its error string per thread)
	template <class Fn>
	void onDevices(size_t n, Fn fn) {
		const size_t D = std::min(numDevices(), std::max<size_t>(n, 1));
		if (D <= 1) { fn(*sim, 0, n); return; }
		auto share = [&](size_t d) { return n * d / D; };
		std::vector<std::string> errors(D);
		std::vector<std::thread> pool;
		for (size_t d = 1; d < D; ++d)
			pool.emplace_back([&, d]() {
				try { fn(*replicas[d - 1], share(d), share(d + 1)); } catch (std::exception& e) { errors[d] = e.what(); }
			});
		try { fn(*sim, share(0), share(1)); } catch (std::exception& e) { errors[0] = e.what(); }
		for (std::thread& t : pool) t.join();
		for (const std::string& e : errors) if (!e.empty()) throw std::runtime_error(e);
	}

	double kViolation(const AP& ap, size_t i) const {
		if (ap[i] < settings.kMin[i]) return settings.kMin[i] - ap[i];
		if (ap[i] > settings.kMax[i]) return ap[i] - settings.kMax[i];
		return 0;
	}

	/// fills one border AP from the base coefficients + the next free genes; accumulates violation
	void borderAp(AP& dst, const SimLib::WohlfartPlus& base, const Input& solution, size_t& cursor, double& violation) const {
		AP tmp;
		tmp.init(base.getK(), 0);
		for (size_t i = 0; i < 9; ++i)
			if (freeK.count((char)i)) {
				tmp[i] = solution[cursor++];
				violation += kViolation(tmp, i);
			}
		dst.init(tmp, 0);
	}

	static void blend(AP& dst, const AP& a, const AP& b, double ratio) {
		for (size_t i = 0; i < 9; ++i) dst[i] = a[i] * (1 - ratio) + b[i] * ratio;
	}

	/// genes -> displaced leads, border APs in their layer slots (the other layers untouched), violation
	/// (the head of simUsingBorderAps / simUsingBorderAndMidAps, sim.cpp:758-781, :831-872)
	void prepareBorders(const Input& solution, Individual& ind) const {
		if (solution.size() < numWohlfartParams + numDisplacementParams) throw std::runtime_error("Solution does not contain enough values");
		const size_t L = sim->numMeasurements();
		ind.leads.resize(L);
		for (size_t i = 0; i < L; ++i) {
			if (settings.measuringPointsDisplacementIsInput)
				ind.leads[i] = sim->displaced(i, solution[numWohlfartParams + i * 2], solution[numWohlfartParams + i * 2 + 1]);
			else ind.leads[i] = sim->measuringPoints()[i];
		}
		if (solution.size() != numDisplacementParams + numWohlfartParams) {
			std::ostringstream t;
			t << "chromosome size does not agree with the combination of the number of free Wohlfart parameters and the selected "
			     "interpolation procedure (" << solution.size() << " != " << numDisplacementParams << "+" << numWohlfartParams << ")";
			throw std::runtime_error(t.str());
		}
		const size_t n = sim->requiredAps();
		std::vector<AP>& aps = ind.layerAps;
		aps.resize(n);
		for (AP& a : aps) a.at = 0;
		size_t cursor = 0;
		double violation = 0;
		if (interp == endo_epi) {
			borderAp(aps.front(), settings.baseAps.front(), solution, cursor, violation);
			borderAp(aps.back(), settings.baseAps.back(), solution, cursor, violation);
		} else {
			borderAp(aps.front(), settings.baseAps.front(), solution, cursor, violation);
			borderAp(aps[midLayer()], settings.baseAps.back(), solution, cursor, violation);   // mid and epi both start from the LAST base ap
			borderAp(aps.back(), settings.baseAps.back(), solution, cursor, violation);
		}
		ind.violation = violation;
	}

	size_t midLayer() const { return (size_t)std::floor(settings.midPosition * (sim->requiredAps() - 1) + 0.5); }
	size_t numBorders() const { return interp == endo_epi ? 2 : 3; }
	static const double* fitOffsets() {   // "experimentally set d", sim.cpp:796, :877
		static const double kd[9] = {0, 0, 0, 0.001, 0, 0.00005, 0.0005, 0.01, 0.2};
		return kd;
	}
	/// the 2 or 3 border APs of an individual, [nBorder][9]
	void borderCoefficients(const Individual& ind, double* out) const {
		const std::vector<AP>& aps = ind.layerAps;
		std::copy(aps.front().getK(), aps.front().getK() + 9, out);
		if (interp != endo_epi) std::copy(aps[midLayer()].getK(), aps[midLayer()].getK() + 9, out + 9);
		std::copy(aps.back().getK(), aps.back().getK() + 9, out + 9 * (numBorders() - 1));
	}

	/// inner layers on the host (the reference's procedure operation by operation; bit-identical results)
	void fitLayersOnHost(Individual& ind) const {
		const size_t n = sim->requiredAps();
		std::vector<AP>& aps = ind.layerAps;
		AP d;
		d.init(fitOffsets(), 0);
		LayerFitTarget fit;
		if (interp == endo_epi) {
			fit.setBorderAps(aps.front(), aps.back());
			for (size_t i = 1; i + 1 < n; ++i) {
				const double ratio = i / double(n - 1);
				fit.setupRatio(ratio);
				aps[i] = aps[i - 1];
				blend(aps[i], aps.front(), aps.back(), ratio);
				steepest_descent(fit, aps[i], d, 0.5, 1e-3, 100);
			}
		} else {
			const size_t mid = midLayer();
			// (the reference rebuilds the endo-mid connectors for every layer below mid; they only depend on
			// the two border APs, so once is enough)
			bool lowerReady = false;
			for (size_t i = 1; i + 1 < n; ++i) {
				if (i < mid) {
					const double ratio = i / double(mid);
					if (!lowerReady) { fit.setBorderAps(aps.front(), aps[mid]); lowerReady = true; }
					fit.setupRatio(ratio);
					aps[i] = aps[i - 1];
					blend(aps[i], aps.front(), aps[mid], ratio);
				} else if (i > mid) {
					const double ratio = (i - mid) / double(n - mid - 1);
					if (i - mid == 1) fit.setBorderAps(aps[mid], aps.back());
					fit.setupRatio(ratio);
					aps[i] = aps[i - 1];
					blend(aps[i], aps[mid], aps.back(), ratio);
				}
				if (i != mid) steepest_descent(fit, aps[i], d, 0.5, 1e-3, 100);
			}
		}
	}

	/// inner layers of one individual through ekg_fit_layers
	void fitLayersOnDevice(Individual& ind) const {
		const size_t n = sim->requiredAps();
		std::vector<double> border(numBorders() * 9), k;
		borderCoefficients(ind, border.data());
		sim->fitLayers(border.data(), 1, numBorders(), interp == endo_epi ? 0 : midLayer(), fitOffsets(), 0.5, 1e-3, 100, k);
		for (size_t l = 0; l < n; ++l) ind.layerAps[l].init(&k[9 * l], 0);
	}

	/// string-model approximation and the gate in front of the full simulation (sim.cpp:712-747); only the
	/// first and the last layer AP enter, i.e. it can run before the inner layers exist
	void approximationGate(Individual& ind) const {
		const std::vector<AP>& aps = ind.layerAps;
		const SimLib::Settings& s = sim->getSettings();
		ind.approxEcg.assign((size_t)(s.simulationLength / s.simulationTimeStep), 0.0);
		for (size_t i = 0; i < ind.approxEcg.size(); ++i) {
			const double t = s.simulationStart + i * s.simulationTimeStep;
			ind.approxEcg[i] = aps[0](t + settings.fastApproxEpiDelay) - aps.back()(t);
		}
		ind.approxCriteria = 2;
		if (settings.fastApproxIsCriterion || settings.fastApproxLimit < 2) ind.approxCriteria = compare(ind.approxEcg, 0);
		ind.simulate = !(ind.approxCriteria > settings.fastApproxLimit);
	}

	/// everything of eval() that happens before Simulation::run: displacement, layer APs, approximation gate
	void prepare(const Input& solution, Individual& ind) const {
		prepareBorders(solution, ind);
		if (fitOnDevice) fitLayersOnDevice(ind);
		else fitLayersOnHost(ind);
		approximationGate(ind);
	}

	/// the device-side comparison (ekg_criteria_kernel) covers the settings where every criterion is one lead compared with
	/// its target: no peak-position criterion, as many equally long targets as leads
	bool criteriaOnDevice() const {
		const size_t L = sim->numMeasurements();
		if (getenv("EKGSIM_B200_HOST_CRITERIA")) return false;
		if (sim->sharded()) return false;   // z-slabs: the ECG only exists once the partial sums have been added on the host
		if (settings.criteriaMode != EvalSettings::every_lead || settings.peakPositionIsCriterion) return false;
		if (targets.size() < L || L == 0) return false;
		for (size_t m = 1; m < L; ++m) if (targets[m].size() != targets[0].size()) return false;
		return (int)settings.comparisonMode >= 1 && (int)settings.comparisonMode <= 4 && !targets[0].empty();
	}

	/// calculateFitness for an individual whose per-lead comparison values were computed on the device
	double fitnessFromCriteria(const Individual& ind, const double* crit, Value& result) const {
		if (result.empty()) result.resize(deducedNumOfCriteria);
		if (result.size() != deducedNumOfCriteria) throw std::runtime_error("error in result.size() - doesn't match deducedNumOfCriteria");
		result[0] = 0;
		const size_t nCrit = std::min(sim->numMeasurements(), targets.size());
		if (settings.fastApproxIsCriterion) result[nCrit] = ind.approxCriteria;
		for (size_t i = 0; i < nCrit; ++i) {
			result[i] = crit[i];
			if (result[i] < 0 || !std::isfinite(result[i]) || !std::isnormal(result[i])) {
				std::cerr << " warning, ConvolutionResult returned wierd value: " << crit[i] << ", causing fitness to be " << result[i]
				          << "; making correction - setting fitness to a large number\n";
				result[i] = 2e10;
			}
		}
		if (settings.endoEpiMinCriterionDelay >= 0) result[result.size() - 1] = endoEpiCriterion(ind);
		return 0.0;
	}

	double endoEpiCriterion(const Individual& ind) const {
		double sd = 0.0;
		for (double t = 1.0; t < 700.0; t += 1.0) {
			const double dd = ind.layerAps.front()(t + settings.endoEpiMinCriterionDelay) - ind.layerAps.back()(t);
			sd += dd * dd;
		}
		return sd;
	}

	double compare(const std::vector<double>& simResult, size_t target) const {
		switch (settings.comparisonMode) {
		case EvalSettings::cmp_correlation: return 1.0 - pearson(simResult, targets[target], 0).value;
		case EvalSettings::cmp_vector_correlation: return 1.0 - vector_correlation(simResult, targets[target], 0).value;
		case EvalSettings::cmp_rms: return rms_match(simResult, targets[target], 0).value;
		default: return dev_from_linear(simResult, targets[target], targetOffsets[target], 0).value;
		}
	}

	/// criteria of one individual (calculateFitness); returns the extra violation (always 0)
	double fitness(const Individual& ind, const std::vector<std::vector<double>>& ecg, bool simulationDone, Value& result) const {
		if (settings.criteriaMode != EvalSettings::every_lead) throw std::runtime_error("selected criteria mode is either not implemented yet or invalid");
		if (result.empty()) result.resize(deducedNumOfCriteria);
		if (result.size() != deducedNumOfCriteria) throw std::runtime_error("error in result.size() - doesn't match deducedNumOfCriteria");
		result[0] = 0;
		const size_t nCrit = std::min(sim->numMeasurements(), targets.size());
		if (settings.fastApproxIsCriterion) result[nCrit] = ind.approxCriteria;
		for (size_t i = 0; i < nCrit; ++i) {
			const std::vector<double>& simResult = simulationDone ? ecg[i] : ind.approxEcg;
			const double value = simulationDone ? compare(simResult, i) : ind.approxCriteria;
			size_t at = i;
			if (settings.peakPositionIsCriterion) {
				at = 2 * i;
				const int rp = int(std::max_element(simResult.begin(), simResult.end()) - simResult.begin());
				const int tp = int(std::max_element(targets[i].begin(), targets[i].end()) - targets[i].begin());
				result[at + 1] = std::abs(rp - tp);
			}
			result[at] = value;
			if (result[at] < 0 || !std::isfinite(result[at]) || !std::isnormal(result[at])) {
				std::cerr << " warning, ConvolutionResult returned wierd value: " << value << ", causing fitness to be " << result[at]
				          << "; making correction - setting fitness to a large number\n";
				result[at] = 2e10;
			}
		}
		if (settings.endoEpiMinCriterionDelay >= 0) result[result.size() - 1] = endoEpiCriterion(ind);
		return 0.0;
	}

	void saveAps(const std::string& fname, const std::vector<AP>& aps, const std::vector<size_t>* indices) const {
		const size_t maxTime = 700;
		std::vector<size_t> which;
		if (indices) { for (size_t i : *indices) if (i < aps.size()) which.push_back(i); }
		else for (size_t i = 0; i < aps.size(); ++i) which.push_back(i);
		std::vector<std::vector<double>> data(which.size(), std::vector<double>(maxTime));
		std::vector<NamedColumn> cols(which.size());
		for (size_t n = 0; n < which.size(); ++n) {
			const AP& ap = aps[which[n]];
			std::ostringstream c, nm;
			c << "k = [" << ap[0];
			for (size_t k = 1; k < 9; ++k) c << ", " << ap[k];
			c << "]";
			nm << "ap_" << which[n];
			for (size_t t = 0; t < maxTime; ++t) data[n][t] = ap((double)t);
			cols[n].comment = c.str();
			cols[n].name = nm.str();
			cols[n].data = &data[n];
		}
		if (!cols.empty()) export_columns(cols, fname, 0, 1);
	}

	void writeOutputs(const Input& solution, const Individual& ind) {
		if (!outSettings.outputCellAps.empty()) saveAps("cell_aps.column", sim->getAps(), &outSettings.outputCellAps);
		if (outSettings.layerAps) saveAps("layer_aps.column", ind.layerAps, nullptr);
		if (outSettings.result) sim->saveMeasurement("input=" + angle_list(solution) + "; ");
	}

	void loadTargets(const char* fname) {
		std::vector<std::vector<double>> file = load_column_file(fname);
		targets.assign(std::max<size_t>(1, file.size() - 1), std::vector<double>());
		targetOffsets.assign(targets.size(), 0.0);
		double targetTimeStep = sim->getSettings().inputActionPotentialsTimeStep;
		const size_t ecgColumns = std::min<size_t>(1, file.size() - 1);
		if (ecgColumns > 0) targetTimeStep = file[0][1] - file[0][0];
		const int resampledStart = (int)(sim->getSettings().simulationStart / targetTimeStep);
		for (size_t i = ecgColumns; i < file.size(); ++i) {
			double tMin = 0, tMax = 0;
			min_and_max(tMin, tMax, file[i]);
			const double m = 1.0 / (tMax - tMin);
			for (double& v : file[i]) v *= m;
			resample(file[i], targets[i - ecgColumns], targetTimeStep / sim->getSettings().simulationTimeStep, resampledStart);
			targetOffsets[i - ecgColumns] = 1.0 - (tMin / (tMax - tMin));
		}
		std::vector<NamedColumn> chk(targets.size());
		for (size_t i = 0; i < chk.size(); ++i) {
			std::ostringstream name;
			name << "target " << (i + 1);
			chk[i].name = name.str();
			chk[i].data = &targets[i];
		}
		export_columns(chk, "target_chk.column", sim->getSettings().simulationStart, sim->getSettings().simulationTimeStep);
	}
};

}  // namespace ekg
