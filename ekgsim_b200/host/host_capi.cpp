// ekgsim_b200/host/host_capi.cpp -- C shim over the host-side evaluation glue (ekg_eval.h) so that
// tests and tools can drive it without a process boundary.  Not part of the device ABI.
#include "ekg_eval.h"

extern "C" {

static thread_local std::string g_host_error;
const char* ekg_host_last_error() { return g_host_error.c_str(); }

/// AP formula (host f64), exposed for cross-checks in the tests
double ekg_host_wohlfart_plus(const double* k, double t) {
	SimLib::WohlfartPlus w(k);
	return w[t];
}

double ekg_host_apd90(const double* k) {
	SimLib::WohlfartPlus w(k);
	return w.apd90();
}

/// EkgSim::runApproximation (Simulation::runApproximation, simulator.cpp:552-559) for layer APs layer_k [n_layers][9]:
/// out must hold (size_t)(length / step) values; returns that count.  Host only, no device needed.
int64_t ekg_host_run_approximation(const double* layer_k, int n_layers, int start, int length, double step, double delay, double* out) {
	try {
		SimLib::EkgSim sim;
		sim.getSettings().simulationStart = start;
		sim.getSettings().simulationLength = length;
		sim.getSettings().simulationTimeStep = step;
		sim.getSettings().neighbourhoodType = "3D4";
		sim.applySettings();
		std::vector<SimLib::ActionPotential> aps((size_t)n_layers);
		for (int l = 0; l < n_layers; ++l) aps[(size_t)l].init(layer_k + 9 * l, 0);
		sim.setApsDestructive(aps);
		std::vector<double> r;
		sim.runApproximation(delay, r);
		std::copy(r.begin(), r.end(), out);
		return (int64_t)r.size();
	} catch (std::exception& e) { g_host_error = e.what(); return -1; }
}

/// the built-in test shape (InputLoader::generateTestShape): layers_out[160 * 120] u16, dims_out = {Z, Y, X}; export_as may be ""
int ekg_host_generate_test_shape(uint16_t* layers_out, int64_t* dims_out, const char* export_as) {
	try {
		std::vector<uint16_t> layers;
		int64_t Z, Y, X;
		ekg::generate_test_shape(layers, Z, Y, X, export_as);
		std::copy(layers.begin(), layers.end(), layers_out);
		dims_out[0] = Z; dims_out[1] = Y; dims_out[2] = X;
		return 0;
	} catch (std::exception& e) { g_host_error = e.what(); return -1; }
}

/// .matrix shape file -> layers (u16, start flag 0x1000), the parser the facade's loadShape uses; returns the voxel count or -1
int64_t ekg_host_load_shape(const char* fname, uint16_t* layers_out, int64_t capacity, int64_t* dims_out) {
	try {
		std::vector<uint16_t> layers;
		int64_t Z, Y, X;
		ekg::load_shape_matrix(fname, layers, Z, Y, X);
		if ((int64_t)layers.size() > capacity) throw std::runtime_error("buffer too small");
		std::copy(layers.begin(), layers.end(), layers_out);
		dims_out[0] = Z; dims_out[1] = Y; dims_out[2] = X;
		return (int64_t)layers.size();
	} catch (std::exception& e) { g_host_error = e.what(); return -1; }
}

/// EkgSim::balancedSlabs: the z-slab cuts of the C++ host's `-slabs` mode; out = n x (z_begin, z_end).  Host only.
int ekg_host_balanced_slabs(const uint16_t* layers, int64_t Z, int64_t Y, int64_t X, int n, int64_t* out) {
	try {
		std::vector<uint16_t> l(layers, layers + Z * Y * X);
		const std::vector<std::pair<int64_t, int64_t>> s = SimLib::EkgSim::balancedSlabs(l, Z, Y, X, (size_t)n);
		for (int i = 0; i < n; ++i) { out[2 * i] = s[(size_t)i].first; out[2 * i + 1] = s[(size_t)i].second; }
		return 0;
	} catch (std::exception& e) { g_host_error = e.what(); return -1; }
}

/// Evaluator over the simulator.ini of the CURRENT directory (like the CLI).  Needs a GPU.
void* ekg_host_evaluator_create(const char* ini, int with_device) {
	try { return new ekg::Evaluator(ini, with_device != 0); }
	catch (std::exception& e) { g_host_error = e.what(); return nullptr; }
}
/// the same on several GPUs: devices = "all" | "<count>" | "0,1,..." (ekg::parse_device_list); evalBatch splits its batch
void* ekg_host_evaluator_create_on(const char* ini, const char* devices) {
	try { return new ekg::Evaluator(ini, true, ekg::parse_device_list(devices ? devices : "")); }
	catch (std::exception& e) { g_host_error = e.what(); return nullptr; }
}
int ekg_host_num_devices(void* ev) { return (int)static_cast<ekg::Evaluator*>(ev)->numDevices(); }
void ekg_host_evaluator_destroy(void* ev) { delete static_cast<ekg::Evaluator*>(ev); }
int ekg_host_num_criteria(void* ev) { return (int)static_cast<ekg::Evaluator*>(ev)->deducedNumOfCriteria; }

int ekg_host_eval(void* ev, const double* genes, int n, double* criteria, double* violation) {
	try {
		std::vector<double> sol(genes, genes + n), res;
		*violation = static_cast<ekg::Evaluator*>(ev)->eval(sol, res);
		std::copy(res.begin(), res.end(), criteria);
		return 0;
	} catch (std::exception& e) { g_host_error = e.what(); return -1; }
}

int ekg_host_eval_batch(void* ev, const double* genes, int n_genes, int B, int threads, double* criteria, double* violations) {
	try {
		std::vector<std::vector<double>> sols(B), res;
		for (int b = 0; b < B; ++b) sols[b].assign(genes + (size_t)b * n_genes, genes + (size_t)(b + 1) * n_genes);
		std::vector<double> viol;
		ekg::Evaluator* e = static_cast<ekg::Evaluator*>(ev);
		e->evalBatch(sols, res, viol, threads);
		for (int b = 0; b < B; ++b) {
			std::copy(res[b].begin(), res[b].end(), criteria + (size_t)b * e->deducedNumOfCriteria);
			violations[b] = viol[b];
		}
		return 0;
	} catch (std::exception& e) { g_host_error = e.what(); return -1; }
}

/// glue only (no GPU work after construction): genes -> layer coefficients [layers*9], leads [L*3]
int ekg_host_layer_coefficients(void* ev, const double* genes, int n, double* k_out, double* leads_out, double* violation) {
	try {
		std::vector<double> sol(genes, genes + n), k, leads;
		static_cast<ekg::Evaluator*>(ev)->layerCoefficients(sol, k, leads, *violation);
		std::copy(k.begin(), k.end(), k_out);
		std::copy(leads.begin(), leads.end(), leads_out);
		return 0;
	} catch (std::exception& e) { g_host_error = e.what(); return -1; }
}

}  // extern "C"
