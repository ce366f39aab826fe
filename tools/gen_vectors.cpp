// tools/gen_vectors.cpp -- generates the BASELINE config-3 batch: 256 parameter vectors,
// std::mt19937_64 seed 20240817, individual-major / gene-minor, x = min + (max-min)*U[0,1),
// bounds from the reference README.md:102-103 (k5,k6,k7,k8 for endo/mid/epi + 4 displacements).
// Output: one vector per line, comma separated, %.17g  (tests/golden/vectors256.txt).
#include <cstdio>
#include <cstdlib>
#include <random>
int main(int argc, char** argv) {
	int n = argc > 1 ? atoi(argv[1]) : 256;
	const double lo[16] = {3e-4, 0.01, 0.01, 200, 3e-4, 0.01, 0.01, 200, 3e-4, 0.01, 0.01, 200, -50, -50, -50, -50};
	const double hi[16] = {1e-3, 0.1, 0.1, 400, 1e-3, 0.1, 0.1, 400, 1e-3, 0.1, 0.1, 400, 50, 50, 50, 50};
	std::mt19937_64 rng(20240817);
	std::uniform_real_distribution<double> U(0.0, 1.0);
	for (int i = 0; i < n; ++i) {
		for (int g = 0; g < 16; ++g) printf("%s%.17g", g ? "," : "", lo[g] + (hi[g] - lo[g]) * U(rng));
		printf("\n");
	}
	return 0;
}
