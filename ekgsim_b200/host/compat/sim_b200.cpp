// ekgsim_b200/host/compat/sim_b200.cpp -- the reference's in-process evaluation interface on the B200 evaluator.
//
// The reference's main.cpp drives everything through `struct OptimizationFunction : VirtualOptimizationFunction`
// (sim.h:97-123; VirtualOptimizationFunction: AMS-DEMO/GeneralOptimizationAlgorithm.h:83-93): the embedded AMS-DEMO
// optimizer (`ekgSim` without arguments, main.cpp:283-362, with or without MPI) hands it to
// GeneralOptimizationAlgorithm as `internalFunction`, `ekgSim test -sim ...` calls it once (main.cpp:367-384).
// This translation unit implements that struct -- same members, same semantics as sim.cpp:1083-1130 -- on top of
// ekg::Evaluator (device-side layer fit + SEPARABLE simulation), so that the reference's UNMODIFIED main.cpp
// links against it instead of sim.cpp + simlib:
//
//     g++ ... -I<reference> -I<reference>/copyOfLibs -I<repo> -include ekgsim_b200/host/compat/refglue_shim.h \
//         <reference>/main.cpp ekgsim_b200/host/compat/sim_b200.cpp <reference>/copyOfLibs/{Ini,Random,Arguments}.cpp \
//         -lekgsim_b200                      (the test harness builds it as ekgSim_refmain_b200)
//
// It is compiled against the reference's own headers where they lie; nothing of them is copied here.
#include "sim.h"   // the reference's sim.h (include path), declares OutputSettings / OptimizationFunction

#include "ekgsim_b200/host/ekg_eval.h"

class SimImplementation : public ekg::Evaluator {
public:
	SimImplementation() : ekg::Evaluator("simulator.ini", true) {}
};

void testWohlInterpolation() {}   // a developer's scratch test in the reference (writes to a fixed Windows path)

OptimizationFunction::OptimizationFunction(size_t propertiesL) : impl(new SimImplementation) {
	valueLen = impl->deducedNumOfCriteria;
	propertiesLen = propertiesL;
	geneBounds = false;
}

OptimizationFunction::~OptimizationFunction() { delete impl; }

void OptimizationFunction::setup(const OutputSettings& outSet) {
	impl->outSettings.outputCellAps = outSet.outputCellAps;
	impl->outSettings.layerAps = outSet.layerAps;
	impl->outSettings.result = outSet.result;
}

void OptimizationFunction::operator()(const Input& solution, Value& result, double& violation, Properties& properties) const {
	if (valueLen > 0) TypeWrapper::resize(result, valueLen);
	if (propertiesLen > 0) TypeWrapper::resize(properties, propertiesLen);
	violation = impl->eval(solution, result);
}

void OptimizationFunction::getGeneParams(size_t& numGenes, size_t& numCriteria, std::vector<double>& gMin, std::vector<double>& gMax) {
	impl->getGeneParams(numGenes, numCriteria, gMin, gMax);
	geneMin = gMin;
	geneMax = gMax;
}

void OptimizationFunction::getEvolutionaryParams(size_t& numGenerations, size_t& popSize, int& queueSize) {
	numGenerations = impl->settings.numGenerations;
	popSize = impl->settings.populationSize;
	queueSize = impl->settings.queueSize;
}

void OptimizationFunction::normalize(Input& solution) const {
	// clip every gene into [geneMin, geneMax] (as far as bounds are known)
	for (size_t i = 0; i < solution.size(); ++i) {
		if (i < geneMin.size()) solution[i] = std::max(solution[i], geneMin[i]);
		if (i < geneMax.size()) solution[i] = std::min(solution[i], geneMax[i]);
	}
}
