// tests/native/libm_port_check.cpp -- host build of ekgsim_b200/csrc/fit_math.cuh against the machine's libm, bit for bit.
// Usage: libm_port_check <n per range> ; prints one line per range "name n mismatches first-mismatch", exit code = #ranges with mismatches.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include "../../ekgsim_b200/csrc/fit_math.cuh"

static std::mt19937_64 rng(12345);
static double uni(double a, double b) { return std::uniform_real_distribution<double>(a, b)(rng); }
static double logu(double a, double b) { return std::exp(uni(std::log(a), std::log(b))); }
static bool same(double a, double b) { return ekg_fm::bits_(a) == ekg_fm::bits_(b) || (a != a && b != b); }

template <class Gen, class F, class G>
static int run(const char* name, long n, Gen gen, F mine, G ref) {
	long bad = 0; double bx = 0, by = 0;
	for (long i = 0; i < n; ++i) {
		double x, y; gen(x, y);
		if (!same(mine(x, y), ref(x, y))) { if (!bad) { bx = x; by = y; } ++bad; }
	}
	printf("%-28s %ld %ld", name, n, bad);
	if (bad) printf("  first: x=%a y=%a mine=%a ref=%a", bx, by, mine(bx, by), ref(bx, by));
	printf("\n");
	return bad != 0;
}

int main(int argc, char** argv) {
	const long n = argc > 1 ? atol(argv[1]) : 1000000;
	int fails = 0;
	auto e1 = [](double x, double) { return ekg_fm::exp_(x); };
	auto e2 = [](double x, double) { return std::exp(x); };
	auto l1 = [](double x, double) { return ekg_fm::log_(x); };
	auto l2 = [](double x, double) { return std::log(x); };
	auto p1 = [](double x, double y) { return ekg_fm::pow_(x, y); };
	auto p2 = [](double x, double y) { return std::pow(x, y); };
	fails += run("exp uniform [-120,10]", n, [](double& x, double& y) { x = uni(-120, 10); y = 0; }, e1, e2);
	fails += run("exp -k*t (fit arguments)", n, [](double& x, double& y) { x = -uni(1e-4, 3.5) * uni(0, 1000); y = 0; }, e1, e2);
	fails += run("exp log-uniform +-", n, [](double& x, double& y) { x = logu(1e-20, 800) * (uni(0, 1) < 0.5 ? -1 : 1); y = 0; }, e1, e2);
	fails += run("exp specials", 8, [](double& x, double& y) { static int c = 0; const double v[8] = {0.0, -0.0, 1e-300, -1e-300, 709.9, -745.2, 1e308, -1e308}; x = v[c++ & 7]; y = 0; }, e1, e2);
	fails += run("log log-uniform", n, [](double& x, double& y) { x = logu(1e-300, 1e300); y = 0; }, l1, l2);
	fails += run("log near 1", n, [](double& x, double& y) { x = uni(0.9, 1.1); y = 0; }, l1, l2);
	fails += run("log 2^q-1 (fit tail)", n, [](double& x, double& y) { x = std::pow(2.0, uni(0.05, 12)) - 1.0; y = 0; }, l1, l2);
	fails += run("log specials", 8, [](double& x, double& y) { static int c = 0; const double v[8] = {1.0, 0.0, -1.0, 4.9e-324, 1e-310, INFINITY, 0.9375, 1.064697265625}; x = v[c++ & 7]; y = 0; }, l1, l2);
	fails += run("pow (1+e)^(-k6/k7)", n, [](double& x, double& y) { x = 1.0 + std::exp(uni(-80, 80)); y = -uni(0.01, 0.1) / uni(0.01, 0.1); }, p1, p2);
	fails += run("pow 2^(k7/k6)", n, [](double& x, double& y) { x = 2.0; y = uni(0.01, 0.1) / uni(0.01, 0.1); }, p1, p2);
	fails += run("pow wide", n, [](double& x, double& y) { x = logu(1e-200, 1e200); y = uni(-3, 3); }, p1, p2);
	fails += run("pow specials", 8, [](double& x, double& y) { static int c = 0; const double vx[8] = {1.0, 0.0, -2.0, 2.0, 1e-310, INFINITY, 3.0, 1.0000000000000002};
	                                                      const double vy[8] = {5.0, 2.0, 3.0, 1e-70, 0.5, -1.0, 700.0, 1e15}; x = vx[c & 7]; y = vy[c & 7]; ++c; }, p1, p2);
	// the N-wide entry points (what the fit kernel calls) against the scalar ones
	{
		long bad = 0;
		for (long i = 0; i < n / 4; ++i) {
			double x[3], y[3], o[3], e[3];
			for (int j = 0; j < 3; ++j) { x[j] = 1.0 + std::exp(uni(-90, 90)); y[j] = -uni(0.01, 0.1) / uni(0.01, 0.1); e[j] = uni(-800, 100); }
			if (i % 97 == 0) { x[1] = -1.0; e[2] = 1e-30; }
			ekg_fm::pow_n<3>(x, y, o);
			for (int j = 0; j < 3; ++j) bad += !same(o[j], std::pow(x[j], y[j]));
			ekg_fm::exp_n<3>(e, o);
			for (int j = 0; j < 3; ++j) bad += !same(o[j], std::exp(e[j]));
		}
		printf("%-28s %ld %ld\n", "pow_n<3> / exp_n<3>", n / 4 * 6, bad);
		fails += bad != 0;
	}
	return fails;
}
