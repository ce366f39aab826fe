// ekgsim_b200/host/ekgsim_main.cpp -- the `ekgSim` command line of the B200 build.
//
// Same contract as the reference's main.cpp (synergy-twinning/ekgsim main.cpp:114-277, :367-414):
//
//     ekgSim test -sim p1,p2,...,p16 -out result        single simulation in the current directory
//                                                       (reads simulator.ini and the files it names)
//
// with the same console transcript (" parameters for single simulator run: <...>", the u/v lines,
// the model / neighbourhood / simulation banner, "\reval N  ", " simulation done in X seconds",
// " criteria = <...>, violation = V", "All done") and the same `result.column`.  Errors are caught
// and printed to stdout as "runtime error caught: ..."; the process still exits 0 (main.cpp:405-413).
//
// Additions of this build (not in the reference):
//     ekgSim -batch vectors.txt [-batchout criteria.txt]     evaluate many parameter vectors (one per
//         line, comma/space separated) in one GPU batch; prints one " criteria = <..>, violation = V"
//         line per vector.  This is what a population-based optimizer should call per generation.  -evalout <file> also
//         writes the batch in the format of the reference optimizer's evaluations.txt (AMS-DEMO/Individual.h:361-380).
//     ekgSim -extern <homeDir>     AMS-DEMO ExternalEvaluation protocol (ExternalEvaluation.h:95-151):
//         reads <homeDir>/input.txt (one gene per line, '#' comments), writes <homeDir>/output.txt
//         ("# violation v", then the criteria).  With -server <socket> or EKGSIM_B200_SERVER=<socket> the genes go to a
//         resident server instead of a fresh process.
//     ekgSim -serve <socket>       resident evaluation service (ekg_server.h): one Evaluator with the model on the GPU,
//         concurrent requests are evaluated together in one batch.  ekgSim -shutdown <socket> ends it.
//     -devices all|<count>|<id,id,...>   with -batch / -serve: one model replica per GPU, every batch split over them by
//         one host thread per device (individuals over GPUs, no collective; default EKGSIM_B200_DEVICES, else one GPU:
//         EKGSIM_B200_DEVICE or 0).  With MPI workers of the reference's optimizer set EKGSIM_B200_DEVICE=<rank % gpus>.
//     -slabs all|<count>|<id,id,...>     with any mode: ONE model spread over several GPUs as z-slabs (for models too large or
//         too slow for one GPU): every GPU sums the ECG over its slab, the partial ECGs are added; the excitation sequence is
//         computed by all GPUs together (peer-linked automaton, NVLink).  Same as EKGSIM_B200_SLABS.  An alternative to
//         -devices (which replicates the model and splits the batch).
// The optimizer itself (`ekgSim` without arguments, AMS-DEMO over MPI) is outside the hot path and is
// not part of this build; use the reference's optimizer with `-extern` or `-batch` as its evaluator.

#include <cctype>
#include <map>

#include "ekg_eval.h"
#include "ekg_server.h"

namespace {

struct DashArgs {
	std::multimap<std::string, std::string> dash;
	std::vector<std::string> free_args;

	DashArgs(int argc, char** argv) {
		// "-name value" pairs; "-<digit>..." is a value, not a flag (copyOfLibs/Arguments.cpp:122-137)
		std::multimap<std::string, std::string>::iterator last = dash.end();
		for (int i = 1; i < argc; ++i) {
			const std::string a = argv[i];
			if (a.size() > 1 && a[0] == '-' && !isdigit((unsigned char)a[1])) last = dash.insert(std::make_pair(a, std::string()));
			else if (!dash.empty() && last->second.empty()) last->second = a;
			else free_args.push_back(a);
		}
	}
	bool is_set(const std::string& n) const { return dash.find(n) != dash.end(); }
	std::string get(const std::string& n) const {
		std::multimap<std::string, std::string>::const_iterator it = dash.find(n);
		return it == dash.end() ? std::string() : it->second;
	}
	std::vector<std::string> all(const std::string& n) const {
		std::vector<std::string> v;
		auto r = dash.equal_range(n);
		for (auto it = r.first; it != r.second; ++it) v.push_back(it->second);
		return v;
	}
};

std::vector<double> parse_vector(const std::string& s) {
	std::istringstream in(s);
	std::vector<double> v;
	for (double t; in >> t;) {
		v.push_back(t);
		const int c = in.peek();
		if (c == ',' || c == ';') in.ignore(1);
	}
	return v;
}

void parse_outputs(const DashArgs& args, ekg::OutputSettings& out) {
	for (const std::string& val : args.all("-out")) {
		std::istringstream st(val);
		std::string name;
		st >> name;
		if (name == "cell_aps") {
			while (st) {
				st.ignore(1);
				size_t n;
				st >> n;
				if (st) out.outputCellAps.push_back(n);
			}
		} else if (name == "layer_aps") out.layerAps = true;
		else if (name == "result") out.result = true;
		else std::cerr << "skipping an unrecognized output opition: " << name << "\n";
	}
	if (args.is_set("-out")) std::cout << "\n";
}

void run_single(const std::vector<double>& params, const ekg::OutputSettings& out) {
	std::cerr << "##### Running a single simulation experiment ####################\n";
	ekg::Evaluator ev("simulator.ini");
	ev.outSettings = out;
	std::vector<double> result;
	const double t0 = ekg::wall_seconds();
	const double violation = ev.eval(params, result);
	const double secs = ekg::wall_seconds() - t0;
	std::cout << " simulation done in " << secs << " seconds\n";
	std::cout << " criteria = " << ekg::angle_list(result) << ", violation = " << violation << "\n";
}

/// the reference optimizer's evaluation log (AMS-DEMO/Individual.h:361-380, `evaluations.txt`): one row per individual,
/// "evaluation_number \t[input] \t violation \t[properties] \t[output] \t evaluator_rank \t evaluation_time \t life_time",
/// the chromosome with 26 significant digits (the stream keeps that precision for the rest of the row, like the reference's)
void write_evaluations(const std::string& fname, const std::vector<std::vector<double>>& sols, const std::vector<std::vector<double>>& results,
                       const std::vector<double>& violations, double secondsPerEvaluation) {
	std::ofstream file(fname.c_str());
	if (!file.is_open()) throw std::runtime_error("could not open " + fname);
	file << "# format of this file:\n"
	     << "# evaluation_number \t[function_input_vector] \t violation \t[properties] \t[function_output_vector] \t evaluator_rank \t evaluation_time \t life_time \n";
	for (size_t i = 0; i < sols.size(); ++i) {
		file << i << "\t" << std::setprecision(26) << ekg::angle_list(sols[i], 26);
		file << "\t" << violations[i] << "\t" << "<>" << "\t" << ekg::angle_list(results[i], 26) << "\t" << 0 << "\t" << secondsPerEvaluation << "\t"
		     << secondsPerEvaluation << "\n";
	}
}

void run_batch(const std::string& file, const std::string& outfile, int threads, const std::string& devices, const std::string& evalfile) {
	std::cerr << "##### Running a batch of simulations ############################\n";
	std::ifstream in(file.c_str());
	if (!in.is_open()) throw std::runtime_error("could not open " + file);
	std::vector<std::vector<double>> sols;
	for (std::string line; std::getline(in, line);) {
		for (char& c : line) if (c == ',' || c == ';') c = ' ';
		std::vector<double> v = parse_vector(line);
		if (!v.empty()) sols.push_back(v);
	}
	ekg::Evaluator ev("simulator.ini", true, ekg::parse_device_list(devices));
	std::vector<std::vector<double>> results;
	std::vector<double> violations;
	const double t0 = ekg::wall_seconds();
	ev.evalBatch(sols, results, violations, threads);
	const double secs = ekg::wall_seconds() - t0;
	std::ofstream of;
	if (!outfile.empty()) of.open(outfile.c_str());
	for (size_t i = 0; i < sols.size(); ++i) {
		std::cout << " criteria = " << ekg::angle_list(results[i]) << ", violation = " << violations[i] << "\n";
		if (of.is_open()) {
			of.precision(17);
			for (double c : results[i]) of << c << " ";
			of << violations[i] << "\n";
		}
	}
	if (!evalfile.empty()) write_evaluations(evalfile, sols, results, violations, sols.empty() ? 0.0 : secs / (double)sols.size());
	std::cout << " batch of " << sols.size() << " simulations done in " << secs << " seconds on " << ev.numDevices() << " GPU(s) (GPU part "
	          << ev.simulator().lastRunSeconds() << " s)\n";
}

void write_extern_output(const std::string& dir, const std::vector<double>& result, double violation) {
	// ExternalEvaluation::readOut (ExternalEvaluation.h:118-151) only looks for a '#' where it expects the next value
	// and stops after the last value: comment lines -- the violation among them -- must come BEFORE the criteria
	std::ofstream out((dir + "output.txt").c_str());
	out.precision(17);
	out << "# written by ekgSim -extern (B200)\n# violation " << violation << "\n";
	for (double c : result) out << c << "\n";
}

void run_extern(const std::string& home, std::string server) {
	const std::string dir = home.empty() ? std::string() : home + "/";
	std::ifstream in((dir + "input.txt").c_str());
	if (!in.is_open()) throw std::runtime_error("could not open " + dir + "input.txt");
	std::vector<double> genes;
	for (std::string line; std::getline(in, line);) {
		if (line.empty() || line[0] == '#') continue;
		std::istringstream ls(line);
		double d;
		while (ls >> d) genes.push_back(d);
	}
	std::vector<double> result;
	double violation = 0;
	if (server.empty()) if (const char* e = getenv("EKGSIM_B200_SERVER")) server = e;
	if (!server.empty()) {
		// a resident `ekgSim -serve` holds the model on the GPU: milliseconds instead of a process start-up per evaluation
		if (ekg::remote_eval(server, genes, result, violation)) { write_extern_output(dir, result, violation); return; }
		std::cerr << "no evaluation server at " << server << ", evaluating in this process\n";
	}
	ekg::Evaluator ev("simulator.ini");
	violation = ev.eval(genes, result);
	write_extern_output(dir, result, violation);
}

void run_serve(const std::string& path, const std::string& devices) {
	std::cerr << "##### Starting the evaluation server ############################\n";
	ekg::Evaluator ev("simulator.ini", true, ekg::parse_device_list(devices));
	ekg::serve(ev, path);
}

}  // namespace

int main(int argc, char** argv) {
	try {
		DashArgs args(argc, argv);
		std::cerr << "***** parsing program arguments ****************************\n";
		if (args.is_set("-slabs")) setenv("EKGSIM_B200_SLABS", args.get("-slabs").c_str(), 1);   // read by every EkgSim of this process
		ekg::OutputSettings out;
		parse_outputs(args, out);
		if (args.is_set("-?")) {
			std::cout << "Argument list:\n   -? \tshow this help screen\n   -sim \tjust run single a simulation with parameters provided after -sim\n"
			             "   -out \tspecify outputs of the program; possible values include result, layer_aps, cell_aps <num> [<num>]*\n"
			             "   -batch \tevaluate every parameter vector of a text file in one GPU batch [-batchout file] [-evalout evaluations.txt] [-threads n]\n"
			             "   -devices \twith -batch / -serve: GPUs to split the batches over: all, a count, or a list 0,1,2 (default: EKGSIM_B200_DEVICES, else one)\n"
			             "   -slabs \tone model over several GPUs as z-slabs: all, a count, or a list 0,1,2 (default: EKGSIM_B200_SLABS, else off)\n"
			             "   -extern \tAMS-DEMO ExternalEvaluation protocol: <homeDir>/input.txt -> <homeDir>/output.txt [-server socket]\n"
			             "   -serve \tresident evaluation server on a unix socket (clients: -extern with -server or EKGSIM_B200_SERVER)\n"
			             "   -shutdown \task the server on the given socket to exit\n";
		} else if (args.is_set("-sim")) {
			const std::vector<double> params = parse_vector(args.get("-sim"));
			std::cout << " parameters for single simulator run: " << ekg::angle_list(params) << "\n";
			if (params.empty()) throw std::runtime_error(" Error: not enough parameters to run simulation: " + args.get("-sim"));
			std::cerr << "\n";
			run_single(params, out);
		} else if (args.is_set("-batch")) {
			std::cerr << "\n";
			run_batch(args.get("-batch"), args.get("-batchout"), atoi(args.get("-threads").c_str()), args.get("-devices"), args.get("-evalout"));
		} else if (args.is_set("-extern")) {
			std::cerr << "\n";
			run_extern(args.get("-extern"), args.get("-server"));
		} else if (args.is_set("-serve")) {
			std::cerr << "\n";
			if (args.get("-serve").empty()) throw std::runtime_error("-serve needs a socket path");
			run_serve(args.get("-serve"), args.get("-devices"));
		} else if (args.is_set("-shutdown")) {
			ekg::remote_shutdown(args.get("-shutdown"));
		} else {
			std::cout << "the optimizer (AMS-DEMO) is not part of the B200 build; use -sim, -batch or -extern as its evaluator\n";
		}
		std::cout << "All done\n";
	} catch (const char* e) {
		std::cout << "exception caught: " << e << "\n";
	} catch (std::runtime_error& e) {
		std::cout << "runtime error caught: " << e.what() << "\n";
	} catch (std::exception& e) {
		std::cout << "std exception caught: " << e.what() << "\n";
	} catch (...) {
		std::cout << "unknown exception caught, execution halted\n";
	}
	return 0;
}
