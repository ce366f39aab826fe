// ekgsim_b200/csrc/fit_math.cuh -- exp / log / pow in double precision with the bits of the reference's libm.
//
// Why: the layer-AP fit (fit.cu; reference sim.cpp:91-313 + nonlinearFit.h:92-168) is a chain of ~100 accept/reject
// and sign decisions per layer, each taken on sums of WohlfartPlus values (Wohlfart.h:195-203: 4 exp, 2 pow, 1 log per
// value).  The reference evaluates them with glibc's libm; CUDA's own exp/log/pow are as accurate (1-2 ulp) but not
// bit-identical, ~2x more instructions, and full of special-case branches that keep independent evaluations from
// overlapping.  glibc (>= 2.28, sysdeps/ieee754/dbl-64/e_exp.c, e_log.c, e_pow.c -- the table-driven algorithms of
// ARM's optimized-routines) needs 15-40 f64 operations per function.  This header restates those algorithms operation
// by operation -- every fused multiply-add where the FMA code path of x86-64 glibc has one (the variant its ifunc
// resolver selects on any CPU with FMA3; transcribed from the disassembly of glibc 2.39's __exp_fma / __log_fma /
// __pow_fma, which is what fixes the association order) -- so that the device takes the SAME decisions as the
// reference's host glue.  tests/test_libm_port.py compares the host build of this header with the machine's libm on
// millions of arguments, bit for bit.
//
// Scope: the arguments the fit produces.  Operands outside the main path of the algorithms (|x| >= 512 or < 2^-54 for exp,
// non-positive / subnormal / non-finite for log, non-finite / non-positive base or tiny / huge exponent for pow)
// fall back to the platform's own function: such results are 0, inf or 1 +- 1 ulp and only ever enter `1 + e`.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define EKG_LIBM_TABLE __device__ const double
#define EKG_LIBM_TABLE_U64 __device__ const uint64_t
#define EKG_FM_INLINE __device__ __forceinline__
#else
#define EKG_FM_INLINE static inline
#endif
#if defined(__CUDACC__) && !defined(__CUDA_ARCH__)
// host pass of nvcc: the tables are also needed as device symbols
#undef EKG_LIBM_TABLE
#undef EKG_LIBM_TABLE_U64
#define EKG_LIBM_TABLE __device__ const double
#define EKG_LIBM_TABLE_U64 __device__ const uint64_t
#undef EKG_FM_INLINE
#define EKG_FM_INLINE __device__ __forceinline__
#endif
#include "libm_tables.h"

namespace ekg_fm {

#if defined(__CUDA_ARCH__) || defined(__CUDACC__)
EKG_FM_INLINE double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }
EKG_FM_INLINE double add_(double a, double b) { return __dadd_rn(a, b); }
EKG_FM_INLINE double sub_(double a, double b) { return __dsub_rn(a, b); }
EKG_FM_INLINE double mul_(double a, double b) { return __dmul_rn(a, b); }
EKG_FM_INLINE uint64_t bits_(double x) { return (uint64_t)__double_as_longlong(x); }
EKG_FM_INLINE double dbl_(uint64_t u) { return __longlong_as_double((long long)u); }
EKG_FM_INLINE double tabd_(const double* t, int i) { return __ldg(t + i); }
EKG_FM_INLINE uint64_t tabu_(const uint64_t* t, int i) { return (uint64_t)__ldg((const unsigned long long*)t + i); }
#else
EKG_FM_INLINE double fma_(double a, double b, double c) { return __builtin_fma(a, b, c); }
EKG_FM_INLINE double add_(double a, double b) { return a + b; }
EKG_FM_INLINE double sub_(double a, double b) { return a - b; }
EKG_FM_INLINE double mul_(double a, double b) { return a * b; }
EKG_FM_INLINE uint64_t bits_(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
EKG_FM_INLINE double dbl_(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
EKG_FM_INLINE double tabd_(const double* t, int i) { return t[i]; }
EKG_FM_INLINE uint64_t tabu_(const uint64_t* t, int i) { return t[i]; }
#endif

// ---- exp (e_exp.c: __exp; with `xtail` = exp_inline of e_pow.c) ---------------------------------------------------------
// x = k ln2/128 + r, exp(x) = 2^(k/128) exp(r); 2^(k/128) = scale (1 + tail) from the table, exp(r) - 1 by a degree-5 polynomial.
// Valid for 2^-54 <= |x| < 512 (the caller checks).
EKG_FM_INLINE double exp_core(double x, double xtail, bool with_tail) {
	const double InvLn2N = ekg_libm_exp_c[0], Shift = ekg_libm_exp_c[1], NegLn2hiN = ekg_libm_exp_c[2], NegLn2loN = ekg_libm_exp_c[3];
	const double C2 = ekg_libm_exp_c[4], C3 = ekg_libm_exp_c[5], C4 = ekg_libm_exp_c[6], C5 = ekg_libm_exp_c[7];
	double kd = fma_(x, InvLn2N, Shift);                 // z + Shift, contracted
	const uint64_t ki = bits_(kd);
	kd = sub_(kd, Shift);
	double r = fma_(kd, NegLn2loN, fma_(kd, NegLn2hiN, x));
	if (with_tail) r = add_(xtail, r);
	const int idx = 2 * (int)(ki & 127u);
	const double tail = dbl_(tabu_(ekg_libm_exp_tab, idx));
	const uint64_t sbits = tabu_(ekg_libm_exp_tab, idx + 1) + (ki << 45);
	const double p23 = fma_(r, C3, C2);
	const double tr = add_(r, tail);
	const double r2 = mul_(r, r);
	const double p45 = fma_(r, C5, C4);
	const double t1 = fma_(p23, r2, tr);
	const double tmp = fma_(mul_(r2, r2), p45, t1);
	const double scale = dbl_(sbits);
	return fma_(scale, tmp, scale);
}

EKG_FM_INLINE bool exp_special(double x) {                // |x| < 2^-54 or |x| >= 512 (or not finite)
	const uint32_t abstop = (uint32_t)(bits_(x) >> 52) & 0x7ffu;
	return abstop - 0x3c9u >= 0x3fu;
}
EKG_FM_INLINE double exp_slow(double x) {
	const uint32_t abstop = (uint32_t)(bits_(x) >> 52) & 0x7ffu;
	if (abstop < 0x3c9u) return add_(1.0, x);
	return exp(x);
}
EKG_FM_INLINE double exp_(double x) {
	if (exp_special(x)) return exp_slow(x);
	return exp_core(x, 0.0, false);
}
// N independent evaluations in one straight line of code: the main paths first (no branch between them, so that the
// instruction scheduler interleaves the N dependency chains), the rare special operands patched afterwards
template <int N>
EKG_FM_INLINE void exp_n(const double (&x)[N], double (&out)[N]) {
#pragma unroll
	for (int i = 0; i < N; ++i) out[i] = exp_core(x[i], 0.0, false);
#pragma unroll
	for (int i = 0; i < N; ++i) if (exp_special(x[i])) out[i] = exp_slow(x[i]);
}

// ---- log (e_log.c: __log) -----------------------------------------------------------------------------------------------
EKG_FM_INLINE double log_(double x) {
	const uint64_t ix = bits_(x);
	const double* A = ekg_libm_log_c + 2;   // A[0..4]
	const double* B = ekg_libm_log_c + 7;   // B[0..10]
	if (ix - 0x3fee000000000000ull < 0x3ff1090000000000ull - 0x3fee000000000000ull) {   // 1 - 2^-4 <= x < 1 + 0x1.09p-4
		if (ix == 0x3ff0000000000000ull) return 0.0;
		const double r = sub_(x, 1.0);
		const double r2 = mul_(r, r);
		const double r3 = mul_(r, r2);
		const double p1 = fma_(r2, B[3], fma_(r, B[2], B[1]));
		const double p2 = fma_(r2, B[6], fma_(r, B[5], B[4]));
		double p3 = fma_(r2, B[9], fma_(r, B[8], B[7]));
		p3 = fma_(r3, B[10], p3);
		double P = fma_(p3, r3, p2);
		P = fma_(P, r3, p1);
		const double two27 = 0x1p27;
		const double rw = fma_(r, two27, r);               // r + w, w = r 2^27
		const double rhi = fma_(-two27, r, rw);            // (r + w) - w
		const double rhi2 = mul_(rhi, rhi);
		const double rlo = sub_(r, rhi);
		const double hi = fma_(rhi2, B[0], r);
		double lo = fma_(rhi2, B[0], sub_(r, hi));
		lo = fma_(mul_(B[0], rlo), add_(r, rhi), lo);
		const double y = fma_(P, r3, lo);
		return add_(hi, y);
	}
	const uint32_t top = (uint32_t)(ix >> 48);
	if (top - 0x0010u >= 0x7ff0u - 0x0010u) return log(x);   // <= 0, subnormal, inf, nan
	const uint64_t tmp = ix - 0x3fe6000000000000ull;
	const int i = (int)((tmp >> 45) & 127u);
	const int k = (int)((int64_t)tmp >> 52);
	const uint64_t iz = ix - (tmp & (0xfffull << 52));
	const double invc = tabd_(ekg_libm_log_tab, 2 * i), logc = tabd_(ekg_libm_log_tab, 2 * i + 1);
	const double z = dbl_(iz);
	const double kd = (double)k;
	const double Ln2hi = ekg_libm_log_c[0], Ln2lo = ekg_libm_log_c[1];
	const double w = fma_(kd, Ln2hi, logc);
	const double r = fma_(z, invc, -1.0);
	const double p12 = fma_(r, A[2], A[1]);
	const double hi = add_(r, w);
	const double r2 = mul_(r, r);
	const double lo = fma_(kd, Ln2lo, add_(sub_(w, hi), r));
	const double r3 = mul_(r, r2);
	const double p34 = fma_(r, A[4], A[3]);
	const double q = fma_(r2, A[0], lo);
	const double pp = fma_(p34, r2, p12);
	return add_(fma_(r3, pp, q), hi);
}

// ---- pow (e_pow.c: __pow = log_inline in double-double, then exp_inline) ---------------------------------------------
// main path; `special` = the operands (or y log x) are outside it and the general function has to be used instead
EKG_FM_INLINE double pow_core(double x, double y, bool& special) {
	const uint64_t ix = bits_(x), iy = bits_(y);
	const uint32_t topx = (uint32_t)(ix >> 52), topy = (uint32_t)(iy >> 52);
	// x subnormal / zero / negative / inf / nan, or |y| outside [2^-65, 2^63)
	special = topx - 0x001u >= 0x7ffu - 0x001u || (topy & 0x7ffu) - 0x3beu >= 0x43eu - 0x3beu;
	// log_inline
	const double* A = ekg_libm_powlog_c + 2;   // A[0..6]
	const uint64_t tmp = ix - 0x3fe6955500000000ull;
	const int i = (int)((tmp >> 45) & 127u);
	const int k = (int)((int64_t)tmp >> 52);
	const uint64_t iz = ix - (tmp & (0xfffull << 52));
	const double z = dbl_(iz);
	const double kd = (double)k;
	const double invc = tabd_(ekg_libm_powlog_tab, 4 * i), logc = tabd_(ekg_libm_powlog_tab, 4 * i + 2), logctail = tabd_(ekg_libm_powlog_tab, 4 * i + 3);
	const double Ln2hi = ekg_libm_powlog_c[0], Ln2lo = ekg_libm_powlog_c[1];
	const double t1 = fma_(kd, Ln2hi, logc);
	const double lo1 = fma_(kd, Ln2lo, logctail);
	const double r = fma_(z, invc, -1.0);
	const double ar = mul_(r, A[0]);
	const double p12 = fma_(r, A[2], A[1]);
	const double p34 = fma_(r, A[4], A[3]);
	const double t2 = add_(r, t1);
	const double lo2 = add_(sub_(t1, t2), r);
	const double ar2 = mul_(r, ar);
	const double ar3 = mul_(r, ar2);
	const double lo3 = fma_(ar, r, -ar2);
	const double hi = add_(t2, ar2);
	const double p56 = fma_(r, A[6], A[5]);
	const double lo4 = add_(sub_(t2, hi), ar2);
	const double inner = fma_(ar2, fma_(p56, ar2, p34), p12);
	double lo = add_(lo1, lo2);
	lo = add_(lo, lo3);
	lo = add_(lo, lo4);
	lo = fma_(ar3, inner, lo);
	const double lhi = add_(hi, lo);
	const double llo = add_(sub_(hi, lhi), lo);
	// ehi + elo = y log(x)
	const double ehi = mul_(y, lhi);
	const double elo = fma_(y, llo, fma_(lhi, y, -ehi));
	// exp_inline (sign_bias = 0: x > 0)
	const uint32_t abstop = (uint32_t)(bits_(ehi) >> 52) & 0x7ffu;
	const bool tiny = abstop < 0x3c9u;                         // |y log x| < 2^-54: 1 + y log x (e_pow.c, exp_inline)
	special = special || (abstop - 0x3c9u >= 0x3fu && !tiny);  // overflow / underflow range: the general function
	const double e = exp_core(ehi, elo, true);
	return tiny ? add_(1.0, ehi) : e;
}
EKG_FM_INLINE double pow_slow(double x, double y) { return pow(x, y); }
EKG_FM_INLINE double pow_(double x, double y) {
	bool special;
	const double r = pow_core(x, y, special);
	if (special) return pow_slow(x, y);
	return r;
}
template <int N>
EKG_FM_INLINE void pow_n(const double (&x)[N], const double (&y)[N], double (&out)[N]) {
	bool sp[N];
#pragma unroll
	for (int i = 0; i < N; ++i) out[i] = pow_core(x[i], y[i], sp[i]);
#pragma unroll
	for (int i = 0; i < N; ++i) if (sp[i]) out[i] = pow_slow(x[i], y[i]);
}

}  // namespace ekg_fm
