// ekgsim_b200/csrc/fit.cu -- per-layer action-potential construction on the device.
//
// The reference derives the Wohlfart coefficients of every inner layer from two border APs
// (endo/mid or mid/epi): 15 straight connectors join matching arc-length positions on the two
// border curves (WohlfartInterpolationEvaluator, sim.cpp:91-313), the layer at `ratio` should pass
// through the points at that ratio along the connectors, and a sign-following steepest descent
// (nonlinearFit.h:92-168; <= 100 iterations, 5 free coefficients) moves the linearly blended
// coefficients towards them (sim.cpp:751-821 endo-epi, :825-916 endo-mid-epi).  That is 21 fits x
// ~9000 f64 AP evaluations per parameter vector, 39 ms in the reference and the end-to-end
// bottleneck once the simulation itself takes a fraction of a millisecond.
//
// Here: two kernels, all arithmetic in f64 with explicit round-to-nearest intrinsics (no FMA
// contraction), every expression in the reference's association order:
//   fit_setup_kernel   one CTA per (vector, border AP): 1000 AP samples in shared memory, apd90,
//                      arc lengths and the two connector walks (kept serial: their running sums
//                      decide integer sample indices)
//   fit_descent_kernel one half-warp per (vector, inner layer): lane = connector point; the 15
//                      squared misses are summed in the reference's order through shuffles, so all
//                      lanes of a fit carry identical descent state and branch identically
// exp / log / pow are NOT CUDA's: fit_math.cuh restates glibc's own algorithms (the reference's libm) operation
// by operation, so every accept/reject and sign decision of the descent is taken on the bits the reference's
// host glue sees -- and the three functions cost 15-40 f64 operations each instead of CUDA's 40-200
// (tests/test_libm_port.py: bit-identical to the host libm on 10^7 arguments; tests/test_gpu_fit.py: all 256
// seeded vectors bit-identical to the reference glue's 24x9 coefficients).

#include <cmath>

#include "ekg_internal.cuh"
#include "fit_math.cuh"

namespace ekg {

namespace {

constexpr int kFitPoints = 15;      // numPoints, sim.cpp:185
constexpr int kFitSamples = 1000;   // WohlfartPlus::apd90 samples 0..999 (Wohlfart.h:206-223)
constexpr int kFitStartX = 10;      // sim.cpp:192

struct FitConn {
	int32_t idx[2][kFitPoints + 1];  // role 0: AP is the first of a pair, role 1: the second
	double y[2][kFitPoints + 1];
};

struct FitArgs {
	const double* border_k;  // [B][nb][9]
	FitConn* conn;           // [B][nb]
	double* layer_k;         // [B][nl][9]
	int32_t B, nb, nl, mid, iterations;
	double d[9];
	double step, eps;
};

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double sqr(double a) { return __dmul_rn(a, a); }

// The factors of WohlfartPlus::operator[] (Wohlfart.h:195-203), each by the reference's expression.
__device__ __forceinline__ double f_tail_c(double k6, double k7) { return ekg_fm::log_(sub(ekg_fm::pow_(2.0, dvd(k7, k6)), 1.0)); }
__device__ __forceinline__ double f_exp(double k, double t) { return ekg_fm::exp_(mul(-k, t)); }
__device__ __forceinline__ double f_A(double k1, double t) { return dvd(1.0, add(1.0, f_exp(k1, t))); }
__device__ __forceinline__ double f_Q(double k6, double k7, double k8, double c, double t) {
	return ekg_fm::pow_(add(1.0, ekg_fm::exp_(add(mul(-k7, sub(t, k8)), c))), -dvd(k6, k7));
}
__device__ __forceinline__ double f_value(double k0, double k2, double k3, double A, double E4, double E5, double Q) {
	return add(mul(mul(A, mul(k2, add(mul(sub(1.0, k3), E4), k3))), mul(E5, sub(1.0, Q))), k0);
}
__device__ double ap_value(const double* k, double c, double t) {
	return f_value(k[0], k[2], k[3], f_A(k[1], t), f_exp(k[4], t), f_exp(k[5], t), f_Q(k[6], k[7], k[8], c, t));
}

// ---- connectors ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fit_setup_kernel(FitArgs a) {
	__shared__ double y[kFitSamples];
	__shared__ double seg[704];
	__shared__ int cross;
	const int bj = blockIdx.x;  // b * nb + j
	const int j = bj % a.nb;
	const int b = bj / a.nb;
	double k[9];
#pragma unroll
	for (int q = 0; q < 9; ++q) k[q] = a.border_k[(size_t)bj * 9 + q];
	if (threadIdx.x == 0) cross = 0;
	// the border AP itself is the AP of its layer: endo -> 0, mid -> `mid`, epi -> nl-1, later ones win
	// (sim.cpp:843-872 writes front, [absoluteMidPos], back in this order)
	if (threadIdx.x < 9) {
		const int layer = (j == 0) ? 0 : (j == a.nb - 1) ? a.nl - 1 : a.mid;
		bool owner = true;
		for (int j2 = j + 1; j2 < a.nb; ++j2) {
			const int l2 = (j2 == a.nb - 1) ? a.nl - 1 : a.mid;
			if (l2 == layer) owner = false;
		}
		if (owner) a.layer_k[((size_t)b * a.nl + layer) * 9 + threadIdx.x] = a.border_k[(size_t)bj * 9 + threadIdx.x];
	}
	const double c = f_tail_c(k[6], k[7]);
	for (int i = threadIdx.x; i < kFitSamples; i += blockDim.x) y[i] = ap_value(k, c, (double)i);
	__syncthreads();
	// apd90: last downward crossing of k0 + 0.1 k2 on the 1 ms grid (Wohlfart.h:206-223)
	const double target = add(k[0], mul(k[2], 0.1));
	int mine = 0;
	for (int i = threadIdx.x + 1; i < kFitSamples; i += blockDim.x)
		if (y[i - 1] > target && y[i] <= target) mine = i;
	if (mine) atomicMax(&cross, mine);
	for (int i = threadIdx.x + 1; i <= 700; i += blockDim.x) seg[i] = __dsqrt_rn(add(sqr(sub(y[i], y[i - 1])), 0.1));
	__syncthreads();
	const int role = threadIdx.x >> 5;
	if ((threadIdx.x & 31) != 0 || role > 1) return;
	int end = 700;  // apd90 = -1 (never repolarises): the reference's size_t cast wraps, the 700 cap applies
	if (cross > 0) {
		const int i = cross;
		const double wa = sub(y[i - 1], target), wb = sub(target, y[i]);
		const double apd = dvd(add(mul((double)(i - 1), wa), mul((double)i, wb)), add(wa, wb));
		const double ce = ceil(apd);
		end = (ce >= 0.0 && ce <= 700.0) ? (int)ce : 700;
	}
	// arc length from 10 ms to apd90; the first AP of a pair starts one sample later (sim.cpp:206-209)
	double len = 0.0;
	for (int i = kFitStartX + 1 - role; i < end; ++i) len = add(len, seg[i]);
	FitConn& out = a.conn[bj];
	double cur = 0.0;
	int idx = kFitStartX;
	for (int i = 0; i < kFitPoints; ++i) {
		const double tgt = dvd(mul((double)i, len), (double)(kFitPoints - 2));
		for (double l = cur; l < tgt && idx < 700; ++idx) l = add(l, seg[idx + 1]);
		cur = tgt;
		out.idx[role][i] = idx;
		out.y[role][i] = y[idx];
	}
}

// ---- descent --------------------------------------------------------------------------------------
constexpr int kFitsPerCta = 8;    // 128 threads, one half-warp per fit
constexpr int kRowStride = 18;    // doubles per row of squared misses: 16-byte aligned, rows 4 banks apart

// State of one fit that is the same for all its 15 points, in shared memory (every lane of the half-warp reads the same
// word: a broadcast), plus the rows the ordered sums go through.
template <int NROWS>
struct FitShared {
	double x0[9];                       // iterate
	double kt[9];                       // trial point; coefficients that are not fitted stay equal to x0
	double move[9], grad[9];            // per-coefficient step factor and the previous difference quotient
	double rows[NROWS][kRowStride];     // squared misses of one evaluation per row, [row][point]
};

// sum of the 15 squared misses of a row in index order ((e0 + e1) + e2) + ... exactly like the reference's loop
__device__ __forceinline__ double row_sum(const double* row) {
	const double2* r2 = reinterpret_cast<const double2*>(row);
	double2 v = r2[0];
	double s = add(v.x, v.y);   // 0 + e0 == e0
#pragma unroll
	for (int n = 1; n < 7; ++n) { v = r2[n]; s = add(add(s, v.x), v.y); }
	return add(s, row[14]);
}

// DM: compile-time set of fitted coefficients (bit q <=> d[q] != 0), 0 = decide at run time.  The reference's
// set {k3, k5, k6, k7, k8} has its own instantiation.
//
// One half-warp per (vector, inner layer): lane = connector point for everything that is evaluated per point (the AP
// factors), lane = fitted coefficient for everything that is evaluated per coefficient (the ordered sums of the squared
// misses, the difference quotients, the step-factor updates, the trial point) -- the five perturbed evaluations of an
// iteration cost one pass of 14 dependent additions and one division instead of five.  The three perturbations that go
// through the repolarisation term (k6, k7, k8: exp + pow each) are evaluated together by exp_n / pow_n: three
// independent dependency chains in one straight line of code.  Per iteration (nonlinearFit.h:104-164):
//   1. one-sided differences: only the AP factor a coefficient enters is recomputed (like the host glue does),
//   2. sign-following update of the step factors and the trial point,
//   3. ln(2^(k7/k6) - 1) for the trial point and its k6- / k7-perturbed neighbours (needed by the NEXT iteration if the
//      trial point is accepted) side by side in three lanes of ONE call,
//   4. full evaluation of the trial point, accept / reject.
// After a rejected trial point the iterate, and with it every difference quotient, is unchanged: the next iteration
// reuses them (grad[] holds exactly those values).  Every value is produced by the same function from the same inputs
// as in the reference's serial loop -- only WHO computes it changed -- so the bits are the reference's.
template <unsigned DM>
__global__ void __launch_bounds__(16 * kFitsPerCta, 5) fit_descent_kernel(FitArgs a) {
	constexpr int NROWS = DM ? (int)__builtin_popcount(DM) : 9;
	__shared__ __align__(16) FitShared<NROWS> s_fit[kFitsPerCta];
	const int per_b = a.nl - a.nb;  // inner layers per vector
	const int fit = blockIdx.x * kFitsPerCta + (threadIdx.x >> 4);
	if (fit >= a.B * per_b) return;  // whole half-warps leave together
	const unsigned mask = 0xFFFFu << (threadIdx.x & 16);
	const int lane16 = threadIdx.x & 15;
	FitShared<NROWS>& S = s_fit[threadIdx.x >> 4];
	const int b = fit / per_b;
	int layer = fit % per_b + 1;
	int ja = 0, jb = a.nb - 1;
	double ratio;
	if (a.nb == 2) {
		ratio = dvd((double)layer, (double)(a.nl - 1));                     // sim.cpp:802
	} else {
		if (layer >= a.mid) ++layer;                                         // skip the mid layer
		if (layer < a.mid) { jb = (a.mid == a.nl - 1) ? 2 : 1; ratio = dvd((double)layer, (double)a.mid); }   // sim.cpp:886
		else { ja = 1; ratio = dvd((double)(layer - a.mid), (double)(a.nl - a.mid - 1)); }                     // sim.cpp:892-893
	}
	const int pt = min(lane16, kFitPoints - 1);  // lane 15 shadows point 14, its miss is never read
	const FitConn& ca = a.conn[b * a.nb + ja];
	const FitConn& cb = a.conn[b * a.nb + jb];
	// connector through (i1, ap1(i1)) and (i2, ap2(i2)), sim.cpp:231-240
	const int i1 = ca.idx[0][pt], i2 = cb.idx[1][pt];
	const double x1 = (double)i1;
	double x2 = (double)i2;
	if (i1 == i2) x2 = add(x2, 0.001);
	const double ck = dvd(sub(ca.y[0][pt], cb.y[1][pt]), sub(x1, x2));
	const double cn = sub(cb.y[1][pt], mul(ck, (double)i2));
	const double px = add(x1, mul(sub(x2, x1), ratio));   // LineConnector::getX, sim.cpp:107-109
	const double py = add(mul(ck, px), cn);               // sim.cpp:102-104

	auto fitted = [&](int q) { return DM ? ((DM >> q) & 1u) != 0 : a.d[q] != 0; };
	// the coefficient this lane looks after: the lane16-th fitted one; its row; the number of fitted coefficients
	int myq = -1, n_fitted = 0;
#pragma unroll
	for (int q = 0; q < 9; ++q) if (fitted(q)) { if (n_fitted == lane16) myq = q; ++n_fitted; }
	const int my_row = min(lane16, max(n_fitted - 1, 0));
	const double my_h = myq >= 0 ? mul(a.d[myq], .001) : 1.0;
	if (lane16 < 9) {
		const double* ka = a.border_k + (size_t)(b * a.nb + ja) * 9;
		const double* kb = a.border_k + (size_t)(b * a.nb + jb) * 9;
		const double v = add(mul(ka[lane16], sub(1.0, ratio)), mul(kb[lane16], ratio));  // combineAps, sim.cpp:705-709
		S.x0[lane16] = v;
		S.kt[lane16] = v;
		S.move[lane16] = fitted(lane16) ? a.d[lane16] : 0.0;
		S.grad[lane16] = 0.0;
	}
	__syncwarp(mask);
	const double h6 = mul(a.d[6], .001), h7 = mul(a.d[7], .001);
	const bool tails = fitted(6) || fitted(7);
	// lanes 0, 1, 2 of the half-warp: ln(2^(k7/k6) - 1) at (k6, k7), (k6 + h6, k7), (k6, k7 + h7)
	auto tail3 = [&](double k6, double k7, double& c, double& c6, double& c7) {
		const double t = f_tail_c(lane16 == 1 ? add(k6, h6) : k6, lane16 == 2 ? add(k7, h7) : k7);
		c = __shfl_sync(mask, t, 0, 16);
		c6 = tails ? __shfl_sync(mask, t, 1, 16) : c;
		c7 = tails ? __shfl_sync(mask, t, 2, 16) : c;
	};
	// per-point factors of the AP at the iterate (Wohlfart.h:195-203)
	double c0, c6, c7;
	tail3(S.x0[6], S.x0[7], c0, c6, c7);
	double A = f_A(S.x0[1], px);
	double E4 = f_exp(S.x0[4], px);
	double E5 = f_exp(S.x0[5], px);
	double Q = f_Q(S.x0[6], S.x0[7], S.x0[8], c0, px);
	S.rows[0][lane16] = sqr(sub(f_value(S.x0[0], S.x0[2], S.x0[3], A, E4, E5, Q), py));
	__syncwarp(mask);
	double y0 = row_sum(S.rows[0]);
	__syncwarp(mask);
	double step = a.step;
	bool reuse = false;   // the last trial point was rejected: iterate and difference quotients are unchanged
	for (int it = a.iterations; it > 0 && y0 > a.eps; --it) {
		double g = 0.0;   // this lane's coefficient: (f(x0 + h e_q) - f(x0)) / h
		if (!reuse) {
			const double k0 = S.x0[0], k2 = S.x0[2], k3 = S.x0[3], k6 = S.x0[6], k7 = S.x0[7], k8 = S.x0[8];
			double Qp[3] = {Q, Q, Q};
			if (fitted(6) || fitted(7) || fitted(8)) {
				const double k6v[3] = {add(k6, h6), k6, k6}, k7v[3] = {k7, add(k7, h7), k7};
				const double k8v[3] = {k8, k8, add(k8, mul(a.d[8], .001))}, cv[3] = {c6, c7, c0};
				double arg[3], e[3], base[3], ex[3];
#pragma unroll
				for (int i = 0; i < 3; ++i) arg[i] = add(mul(-k7v[i], sub(px, k8v[i])), cv[i]);
				ekg_fm::exp_n<3>(arg, e);
#pragma unroll
				for (int i = 0; i < 3; ++i) { base[i] = add(1.0, e[i]); ex[i] = -dvd(k6v[i], k7v[i]); }
				ekg_fm::pow_n<3>(base, ex, Qp);
			}
			const double Ap = fitted(1) ? f_A(add(S.x0[1], mul(a.d[1], .001)), px) : A;
			const double E4p = fitted(4) ? f_exp(add(S.x0[4], mul(a.d[4], .001)), px) : E4;
			const double E5p = fitted(5) ? f_exp(add(S.x0[5], mul(a.d[5], .001)), px) : E5;
			int row = 0;
#pragma unroll
			for (int q = 0; q < 9; ++q) {
				if (fitted(q)) {
					const double h = mul(a.d[q], .001);
					const double v = f_value(q == 0 ? add(k0, h) : k0, q == 2 ? add(k2, h) : k2, q == 3 ? add(k3, h) : k3, q == 1 ? Ap : A,
					                         q == 4 ? E4p : E4, q == 5 ? E5p : E5, q == 6 ? Qp[0] : q == 7 ? Qp[1] : q == 8 ? Qp[2] : Q);
					S.rows[row][lane16] = sqr(sub(v, py));
					++row;
				}
			}
			__syncwarp(mask);
			g = dvd(sub(row_sum(S.rows[my_row]), y0), my_h);
			__syncwarp(mask);   // the rows are free again
		} else if (myq >= 0) g = S.grad[myq];
		bool sc = false;
		if (myq >= 0) {
			double mv = S.move[myq];
			const double gr = S.grad[myq];
			if (mul(g, gr) < 0) { mv = mul(mv, 0.5); sc = true; }
			else if (fabs(g) > mul(0.75, fabs(gr))) mv = mul(mv, 1.5);
			S.move[myq] = mv;
			S.grad[myq] = g;   // oldGrad = grad, nonlinearFit.h:144 (only its own component is ever read)
			S.kt[myq] = sub(S.x0[myq], mul(step, (g > 0) ? mv : -mv));   // nonlinearFit.h:148-150
		}
		const bool step_change = __any_sync(mask, sc);
		__syncwarp(mask);
		double c1, c61, c71;
		tail3(S.kt[6], S.kt[7], c1, c61, c71);
		const double At = fitted(1) ? f_A(S.kt[1], px) : A;
		const double E4t = fitted(4) ? f_exp(S.kt[4], px) : E4;
		double E5t = E5, Qt = Q;
		if (fitted(5) || fitted(6) || fitted(7) || fitted(8)) {
			// exp(-k5 t) and the exponential inside the repolarisation term together, then the power
			const double arg[2] = {mul(-S.kt[5], px), add(mul(-S.kt[7], sub(px, S.kt[8])), c1)};
			double e[2];
			ekg_fm::exp_n<2>(arg, e);
			E5t = e[0];
			Qt = ekg_fm::pow_(add(1.0, e[1]), -dvd(S.kt[6], S.kt[7]));
		}
		S.rows[0][lane16] = sqr(sub(f_value(S.kt[0], S.kt[2], S.kt[3], At, E4t, E5t, Qt), py));
		__syncwarp(mask);
		const double y1 = row_sum(S.rows[0]);
		if (y1 < y0) {
			y0 = y1;
			c0 = c1; c6 = c61; c7 = c71;
			A = At; E4 = E4t; E5 = E5t; Q = Qt;
			if (myq >= 0) S.x0[myq] = S.kt[myq];
			reuse = false;
		} else {
			if (!step_change) step = mul(step, 0.5);
			reuse = true;
		}
		__syncwarp(mask);
	}
	if (lane16 < 9) a.layer_k[((size_t)b * a.nl + layer) * 9 + lane16] = S.x0[lane16];
}

}  // namespace

int run_fit(ekg_model* m, const double* d_border_k, int64_t B, int64_t n_border, int64_t n_layers, int64_t mid, const double* d9,
            double step, double eps, int64_t iterations, double* d_layer_k, cudaStream_t st) {
	if (B <= 0 || n_layers <= 0) return fail(EKG_E_INVALID, "bad sizes");
	if (n_border != 2 && n_border != 3) return fail(EKG_E_INVALID, "n_border must be 2 (endo-epi) or 3 (endo-mid-epi)");
	if (n_layers < n_border) return fail(EKG_E_INVALID, "fewer layers than border APs");
	if (n_border == 3 && (mid < 0 || mid >= n_layers)) return fail(EKG_E_INVALID, "mid layer out of range");
	if (n_border == 3 && (mid == 0 || mid == n_layers - 1) && n_layers > 2)
		return fail(EKG_E_UNSUPPORTED, "mid layer coincides with a border layer");
	if (B * n_layers > (int64_t)1 << 28) return fail(EKG_E_UNSUPPORTED, "batch too large for one fit launch");
	int rc;
	FitConn* conn = nullptr;
	{
		int64_t need = (B * n_border * (int64_t)sizeof(FitConn) + 7) / 8;
		if ((rc = ensure(&m->d_fit_conn, &m->fit_conn_cap, need))) return rc;
		conn = reinterpret_cast<FitConn*>(m->d_fit_conn);
	}
	FitArgs a;
	a.border_k = d_border_k;
	a.conn = conn;
	a.layer_k = d_layer_k;
	a.B = (int32_t)B; a.nb = (int32_t)n_border; a.nl = (int32_t)n_layers; a.mid = (int32_t)(n_border == 3 ? mid : -1);
	a.iterations = (int32_t)iterations;
	for (int q = 0; q < 9; ++q) a.d[q] = d9[q];
	a.step = step; a.eps = eps;
	fit_setup_kernel<<<(unsigned)(B * n_border), 256, 0, st>>>(a);
	EKG_CUDA(cudaGetLastError());
	m->last_launches = 1;
	const int64_t fits = B * (n_layers - n_border);
	if (fits > 0) {
		unsigned dm = 0;
		for (int q = 0; q < 9; ++q) if (d9[q] != 0) dm |= 1u << q;
		if (dm == 0x1E8u) fit_descent_kernel<0x1E8u><<<(unsigned)((fits + kFitsPerCta - 1) / kFitsPerCta), 16 * kFitsPerCta, 0, st>>>(a);   // k3, k5, k6, k7, k8 (sim.cpp:877)
		else fit_descent_kernel<0u><<<(unsigned)((fits + kFitsPerCta - 1) / kFitsPerCta), 16 * kFitsPerCta, 0, st>>>(a);
		EKG_CUDA(cudaGetLastError());
		++m->last_launches;
	}
	return EKG_OK;
}

}  // namespace ekg
