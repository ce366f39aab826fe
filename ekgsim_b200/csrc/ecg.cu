// ekgsim_b200/csrc/ecg.cu -- fused AP + stencil + lead-reduction kernels (sm_100a).
//
// Replaces Simulation::run (reference simlib/simulator.cpp:452-550), which for every occupied
// voxel c and time sample t evaluates the voxel's action potential V_c(t) and those of its
// occupied neighbours (WohlfartPlus::operator[], Wohlfart.h:195-203, through
// ActionPotential::operator(), simulator.cpp:154-170), forms the dipole
// D_c(t) = sum_dif dif * (V_{c-dif}(t) - V_c(t)) (:511-525) and projects it on every lead with
// w_{l,c} = (mp_l - p_c)/|mp_l - p_c|^3 (:530-537).
//
// Design:
//  * The double sum is re-associated per voxel:  ECG_l(t) = sum_c G_{l,c} V_c(t) with the
//    time-invariant lead-field/stencil coefficient
//        G_{l,c} = - [ (sum_{k in occ(c)} dif_k) . w_{l,c} + sum_{k in occ(c)} dif_k . w_{l,c-dif_k} ]
//    so each AP is evaluated ONCE per voxel and time sample (the reference evaluates it 8.65x)
//    and the stencil never touches the time loop.  G is computed on the fly in the kernel's
//    phase A from the voxel's packed coordinates and its 26-bit neighbour-occupancy mask; the AP
//    field and G are never materialised in HBM.
//  * One thread owns one time sample; a CTA owns one segment (a run of voxels of one layer) of
//    one parameter vector.  Per-layer AP coefficients live in registers, per-voxel data
//    (activation time, G) is staged through shared memory and read as warp-wide broadcasts, the
//    per-lead sums stay in registers for the whole voxel loop (fp32 for 32 voxels, then folded
//    into fp64), and one f64 partial per (segment, vector, lead, sample) is written at the end.
//    A second tiny kernel adds the partials in a fixed order -> bitwise run-to-run determinism.
//  * DIRECT variant: the full 9-coefficient AP per voxel and sample = 5 ex2 + 1 lg2 + 1 rcp on
//    the MUFU pipe (the binding unit) + ~20 FP32 ops.  HOISTED variant: the factors that do not
//    depend on the voxel (repolarisation sigmoid, exp(-k t)) come from a per-(vector, layer,
//    sample) table computed in f64, the factors that do not depend on time (exp(+k at)) are
//    computed once per voxel in phase A; only the depolarisation sigmoid 1/(1+exp(-k1 (t-at)))
//    remains per voxel and sample (1 ex2 + 1 rcp).
//
// Algorithmic traffic: 16 B per voxel (pos, mask, at) per CTA, re-read by the B CTAs of a
// segment from L2; 0.04 B per voxel-timestep at T = 400 -> MUFU-bound, not HBM-bound.

#include <cstdlib>
#include <type_traits>

#include "ekg_internal.cuh"
#include "fit_math.cuh"   // f64 exp / log / pow: glibc's table-driven algorithms, 2-3x fewer operations than CUDA's (the f64 table kernels below)

namespace ekg {

// ---- MUFU wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rsq(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// 1/|q|^3 with one Newton step on the reciprocal square root (error ~1 ulp)
__device__ __forceinline__ float inv_cube(float sq) {
	float y = mufu_rsq(sq);
	y = y * fmaf(-0.5f * sq, y * y, 1.5f);
	return y * y * y;
}

// 1/x for x in [1, 2^127] on the FMA pipe: integer-subtraction seed (12 % off) + three Newton steps
// (error 0.12 -> 1.4e-2 -> 2e-4 -> 4e-8).  Used for the depolarisation sigmoid so that the XU pipe,
// the binding unit of the DIRECT kernel, carries six instead of seven operations per voxel-timestep.
__device__ __forceinline__ float fma_rcp(float x) {
	float y = __int_as_float(0x7EF311C7 - __float_as_int(x));
	y = fmaf(y, fmaf(-x, y, 1.f), y);
	y = fmaf(y, fmaf(-x, y, 1.f), y);
	y = fmaf(y, fmaf(-x, y, 1.f), y);
	return y;
}

constexpr int MODE_DIRECT = 1;
constexpr int MODE_HOISTED = 2;

template <int MODE, int NL>
struct RowLayout {
	static constexpr int kUsed = (MODE == MODE_HOISTED ? 3 : 1) + NL;
	static constexpr int kFloats = kUsed <= 4 ? 4 : 8;
};

// ---- the fused kernel ---------------------------------------------------------------------------
// Work decomposition: grid.x = segment (a run of voxels of ONE layer), grid.y = pair tile.  The
// (parameter vector b, time sample t) pairs of the whole batch are flattened, p = b*T + t, and cut
// into tiles of kEcgThreads = 256 consecutive pairs (8 warps = 2 per SM sub-partition, every lane
// busy; a 400-sample trace is 12.5 warps, which left one sub-partition with 33 % more MUFU work
// in the first version of this kernel).  A tile therefore spans up to kMaxVecPerTile parameter
// vectors; phase A computes the lead-field coefficients for each of them.
//
// Voxel slices (small batches): when B*T is not a multiple of the tile size the last tile would be
// part empty (B = 1, T = 400: tiles of 256 + 144 pairs, 22 % of the lanes idle).  The pair index is
// therefore generalised to (slice s, vector b, sample t), p = (s*B + b)*T + t with S = 256/gcd(B*T, 256)
// slices, so that S*B*T is a whole number of tiles (T = 400, B = 1: S = 16, 25 full tiles).  A "virtual
// vector" vv = s*B + b evaluates vector b on the s-th of S equal sub-ranges of the CTA's segment; the
// threads of one tile belong to at most kMaxVecPerTile virtual vectors, each with its own sub-range
// staged in shared memory (rows of one chunk position sit side by side, so the two halves of a warp
// that straddles a slice boundary read two adjacent 16-byte rows: still one wavefront).  Every thread
// walks the same number (+-1) of voxels; the partial sums carry the slice in their segment index.
template <int MODE, int NL>
__global__ void __launch_bounds__(kEcgThreads, 4) ecg_kernel(const EcgArgs a) {
	constexpr int ROW = RowLayout<MODE, NL>::kFloats;
	__shared__ __align__(16) float s_vox[kSmemRows * ROW];
	__shared__ float2 s_lead[kMaxVecPerTile * NL * 3];  // lead coordinate as fp32 hi + lo
	__shared__ int s_atmax[2];                          // HOISTED: latest activation time of the chunk (float bits)
	__shared__ int s_sub[kMaxVecPerTile][2];            // voxel sub-range [begin, end) of every virtual vector of the tile

	const Segment sg = a.segs[blockIdx.x];
	const PairTile tile = a.tiles[blockIdx.y];
	const int vv0 = tile.begin / a.T;
	const int nb = (tile.end - 1) / a.T - vv0 + 1;   // virtual vectors touched by this tile (<= kMaxVecPerTile)
	const int p = tile.begin + threadIdx.x;
	const bool live = p < tile.end;
	const int vv = live ? p / a.T : vv0;
	const int t = live ? p - vv * a.T : 0;
	const int bl = vv - vv0;
	const int sl = vv / a.B;                          // voxel slice of this thread
	const int b = vv - sl * a.B;                      // parameter vector of this thread
	const int chunk = min(kChunk, kSmemRows / nb);
	const int seg_n = sg.end - sg.begin;
	// sub-range of slice s: [begin + n s / S, begin + n (s + 1) / S)  (the whole segment for S = 1)
	const int my_sub_b = sg.begin + (int)((int64_t)seg_n * sl / a.S);
	const int my_sub_n = sg.begin + (int)((int64_t)seg_n * (sl + 1) / a.S) - my_sub_b;
	int n_iter = 0;                                   // chunks of the longest sub-range of the tile
	for (int v = 0; v < nb; ++v) {
		const int s = (vv0 + v) / a.B;
		const int sb = sg.begin + (int)((int64_t)seg_n * s / a.S), se = sg.begin + (int)((int64_t)seg_n * (s + 1) / a.S);
		if (threadIdx.x == 0) { s_sub[v][0] = sb; s_sub[v][1] = se; }
		n_iter = max(n_iter, (se - sb + chunk - 1) / chunk);
	}

	if (threadIdx.x < 2) s_atmax[threadIdx.x] = 0;
	if (threadIdx.x < nb * NL * 3) {
		const int v = threadIdx.x / (NL * 3), r = threadIdx.x % (NL * 3);
		const int l = a.lead0 + r / 3;
		const double c = l < a.L ? a.leads[((int64_t)((vv0 + v) % a.B) * a.L + l) * 3 + r % 3] : 0.0;
		const float hi = (float)c;
		s_lead[threadIdx.x] = make_float2(hi, (float)(c - (double)hi));
	}

	// per-thread constants of (vector b, layer, sample t)
	const float* P = a.params + ((int64_t)b * a.n_layers + (sg.layer - 1)) * kParamStride;
	const float a1 = __ldg(P + 0), k0 = __ldg(P + 8);
	float a4 = 0.f, a5 = 0.f, a7 = 0.f, c2 = 0.f, np = 0.f, A = 0.f, Bc = 0.f, k8hi = 0.f, k8lo = 0.f, F1 = 0.f, F2 = 0.f;
	if (MODE == MODE_DIRECT) {
		a4 = __ldg(P + 1); a5 = __ldg(P + 2);
		a7 = __ldg(P + 3); c2 = __ldg(P + 4); np = __ldg(P + 5); A = __ldg(P + 6); Bc = __ldg(P + 7);
		k8hi = __ldg(P + 9); k8lo = __ldg(P + 10);
	} else if (live) {
		const float* F = a.ftab + (((int64_t)b * a.n_layers + (sg.layer - 1)) * 2) * a.T + t;
		F1 = __ldg(F);
		F2 = __ldg(F + a.T);
	}
	const float thi = live ? __ldg(a.t_hi + t) : 0.f;
	const float tlo = live ? __ldg(a.t_lo + t) : 0.f;

	// per-lead running sums: fp32 for 32 voxels at a time, folded into an fp32 (sum, compensation)
	// pair with an error-free TwoSum -- f64-grade accumulation without conversions on the XU pipe
	float sum[NL], comp[NL];
#pragma unroll
	for (int l = 0; l < NL; ++l) { sum[l] = 0.f; comp[l] = 0.f; }

	int parity = 0;
	for (int it = 0; it < n_iter; ++it, parity ^= 1) {
		const int n = max(0, min(chunk, my_sub_n - it * chunk));   // voxels of this thread's sub-range in this chunk
		__syncthreads();  // phase B of the previous chunk is done with s_vox; s_lead and s_sub are visible

		// ---- phase A: per-(voxel, virtual vector) time-invariant data -> shared memory ----
		for (int idx = threadIdx.x; idx < chunk * nb; idx += kEcgThreads) {
			const int j = idx / nb, v = idx - j * nb;
			const int base = s_sub[v][0] + it * chunk;
			if (base + j >= s_sub[v][1]) continue;
			const uint32_t pos = __ldg(a.pos + base + j);
			const uint32_t mask = __ldg(a.mask + base + j);
			const float at = __ldg(a.at32 + base + j);
			// voxel position = its index in the zero-bordered matrix (simulator.cpp:487, :531);
			// integer -> float through the 2^23 mantissa trick (keeps I2F off the MUFU/XU pipe)
			const float pz = __uint_as_float(0x4B000000u | ((pos >> 22) + 1u)) - 8388608.f;
			const float py = __uint_as_float(0x4B000000u | (((pos >> 11) & 0x7ffu) + 1u)) - 8388608.f;
			const float px = __uint_as_float(0x4B000000u | ((pos & 0x7ffu) + 1u)) - 8388608.f;
			const float2* lead = s_lead + v * NL * 3;
			float G[NL];
#pragma unroll
			for (int l = 0; l < NL; ++l) {
				const float rz = (lead[3 * l].x - pz) + lead[3 * l].y;      // mp - p_c, exact difference + low part
				const float ry = (lead[3 * l + 1].x - py) + lead[3 * l + 1].y;
				const float rx = (lead[3 * l + 2].x - px) + lead[3 * l + 2].y;
				float g = 0.f;
				float sz = 0.f, sy = 0.f, sx = 0.f;
				for (int k = 0; k < a.nbr.n; ++k) {
					if ((mask >> a.nbr.bit[k]) & 1u) {
						const float dz = a.nbr.fz[k], dy = a.nbr.fy[k], dx = a.nbr.fx[k];
						const float qz = rz + dz, qy = ry + dy, qx = rx + dx;  // mp - p_{c-dif}
						const float sq = fmaf(qx, qx, fmaf(qy, qy, qz * qz));
						const float dot = fmaf(dz, qz, fmaf(dy, qy, dx * qx));
						g = fmaf(dot, inv_cube(sq), g);
						sz += dz; sy += dy; sx += dx;
					}
				}
				const float sqc = fmaf(rx, rx, fmaf(ry, ry, rz * rz));
				g = fmaf(fmaf(sz, rz, fmaf(sy, ry, sx * rx)), inv_cube(sqc), g);
				G[l] = (a.lead0 + l < a.L) ? -g : 0.f;
			}
			float* row = s_vox + idx * ROW;
			if (MODE == MODE_DIRECT) {
				row[0] = at;
#pragma unroll
				for (int l = 0; l < NL; ++l) row[1 + l] = G[l];
			} else {
				// time-invariant AP factors exp((k4+k5)(at-t0)), exp(k5 (at-t0)); clamped so that
				// the products with the (also clamped) table entries can never be inf*0
				const float* Pv = a.params + ((int64_t)((vv0 + v) % a.B) * a.n_layers + (sg.layer - 1)) * kParamStride;
				const float v4 = __ldg(Pv + 1), v5 = __ldg(Pv + 2), t0 = __ldg(Pv + 11);
				const float da = at - t0;
				// row = (h1, h2, G_0..G_{NL-1}, at): the saturated loop only needs the leading part
				row[0] = mufu_ex2(fminf(-(v4 + v5) * da, 60.f));
				row[1] = mufu_ex2(fminf(-v5 * da, 60.f));
#pragma unroll
				for (int l = 0; l < NL; ++l) row[2 + l] = G[l];
				row[2 + NL] = at;
				// activation times are >= 0, so their float bit patterns order like ints
				const int amax = __reduce_max_sync(__activemask(), __float_as_int(at));
				if ((threadIdx.x & 31) == (__ffs(__activemask()) - 1)) atomicMax(&s_atmax[parity], amax);
			}
		}
		__syncthreads();
		// HOISTED: once every sample of this warp is later than the chunk's last activation by enough
		// that exp(-k1 (t-at)) < 2^-25, the depolarisation sigmoid is exactly 1 in fp32 and the two
		// MUFU ops per voxel can be skipped (always the case for T/U-wave runs that start at 100 ms)
		bool saturated = false;
		if (MODE == MODE_HOISTED) {
			const float tau_min = (thi - __int_as_float(s_atmax[parity])) + tlo;
			saturated = __all_sync(0xffffffffu, !live || a1 * tau_min < -25.f);
			if (threadIdx.x == 0) s_atmax[parity ^ 1] = 0;  // for the next chunk (written after the barrier above)
		}

		// ---- phase B: time loop, one (vector, sample) pair per thread, voxel rows from shared memory ----
		if (live) {
			const float* my_rows = s_vox + bl * ROW;
			const int stride = nb * ROW;
			// SAT = true is the HOISTED loop without the depolarisation sigmoid (see above)
			auto time_loop = [&](auto sat_tag) {
				constexpr bool SAT = decltype(sat_tag)::value;
				for (int j0 = 0; j0 < n; j0 += 32) {
					const int j1 = min(j0 + 32, n);
					float acc[NL];
#pragma unroll
					for (int l = 0; l < NL; ++l) acc[l] = 0.f;
					const float4* rp = reinterpret_cast<const float4*>(my_rows + j0 * stride);
#pragma unroll 4
					for (int j = j0; j < j1; ++j, rp += stride / 4) {
						const float4 r0 = rp[0];
						float V;
						float G[NL];
						if (MODE == MODE_DIRECT) {
							const float tau = (thi - r0.x) + tlo;               // t - at   (simulator.cpp:169)
							const float S = fma_rcp(1.f + mufu_ex2(fminf(a1 * tau, 126.f)));  // 1/(1+exp(-k1 t'))
							const float k8s = k8hi - r0.x;                   // k8 - at  (simulator.cpp:156)
							const float u = (tau - k8s) - k8lo;              // t' - k8'
							// (1+e)^(-k6/k7) with e = exp(-k7 (t'-k8') + ln(2^(k7/k6)-1)) = 2^z, evaluated as
							// 2^(-(k6/k7) * log2(1+2^z)),  log2(1+2^z) = max(z,0) + log2(1 + 2^-|z|):
							// the same two MUFU ops, but 2^z may exceed the fp32 range (small k6/k7 make
							// (huge)^(-small) a perfectly ordinary number)
							const float z = fmaf(a7, u, c2);
							const float Q = mufu_ex2(np * (fmaxf(z, 0.f) + mufu_lg2(1.f + mufu_ex2(-fabsf(z)))));
							// k2((1-k3) exp(-k4 t') + k3); the clamp only matters far before activation, where S is 0 and
							// an overflowing exponential would turn 0 * inf into NaN (the f64 reference stays finite there)
							const float Pp = fmaf(A, mufu_ex2(fminf(a4 * tau, 100.f)), Bc);
							const float E5 = mufu_ex2(a5 * tau);
							V = fmaf((S * Pp) * E5, 1.f - Q, k0);
							G[0] = r0.y;
							if (NL >= 2) G[1] = r0.z;
							if (NL >= 3) G[2] = r0.w;
							if (NL >= 4) G[3] = rp[1].x;
						} else {
							G[0] = r0.z;
							if (NL >= 2) G[1] = r0.w;
							if (NL >= 3) G[2] = rp[1].x;
							if (NL >= 4) G[3] = rp[1].y;
							if (SAT) {
								V = fmaf(F1, r0.x, fmaf(F2, r0.y, k0));
							} else {
								const float at = NL <= 2 ? rp[1].x : rp[1].z;
								const float tau = (thi - at) + tlo;
								const float S = mufu_rcp(1.f + mufu_ex2(a1 * tau));
								V = fmaf(S, fmaf(F1, r0.x, F2 * r0.y), k0);
							}
						}
#pragma unroll
						for (int l = 0; l < NL; ++l) acc[l] = fmaf(G[l], V, acc[l]);
					}
#pragma unroll
					for (int l = 0; l < NL; ++l) {
						const float s1 = sum[l] + acc[l];
						const float bp = s1 - sum[l];
						comp[l] += (sum[l] - (s1 - bp)) + (acc[l] - bp);
						sum[l] = s1;
					}
				}
			};
			if (MODE == MODE_HOISTED && saturated) time_loop(std::true_type{});
			else time_loop(std::false_type{});
		}
	}

	if (live) {
#pragma unroll
		for (int l = 0; l < NL; ++l) {
			const int lead = a.lead0 + l;
			if (lead < a.L) a.partial[((((int64_t)blockIdx.x * a.S + sl) * a.B + b) * a.L + lead) * a.T + t] = (double)sum[l] + (double)comp[l];
		}
	}
}

// ---- SEPARABLE path -------------------------------------------------------------------------------
// Once the depolarisation sigmoid has saturated for every voxel (all of a run that starts after the QRS
// complex, like the reference's testRun: activation ends at 39 ms, the simulation starts at 100 ms), the AP
// is  V_c(t) = k0 + F1(t) h1_c + F2(t) h2_c  with F1, F2 functions of (vector, layer, t) only and
// h1_c = exp((k4+k5)(at_c - t0)), h2_c = exp(k5 (at_c - t0)) functions of (vector, layer, voxel) only --
// the HOISTED kernel's saturated loop evaluates exactly this per voxel and sample.  The sum over voxels
// then leaves the time loop as well:
//     ECG_l(t) = sum_layers [ k0 M0 + F1(t) M1 + F2(t) M2 ],   (M0, M1, M2) = sum_{c in layer} G_{l,c} (1, h1_c, h2_c)
// i.e. O(voxels) work per simulation instead of O(voxels x samples).  ecg_moment_kernel computes the
// moments for any stencil (the phase A arithmetic of ecg_kernel + two ex2 per voxel and vector) -- the reference's
// default stencil has the two specialised kernels further down --, ecg_combine_kernel evaluates F1, F2 in f64 and
// sums the 3 x layers terms per sample.
//
// Threads of a CTA = VB parameter vectors x 256/VB voxel lanes (VB = 32 for batches: a warp works on ONE
// voxel for 32 vectors, voxel loads are broadcasts and the occupancy-mask branches are warp-uniform; VB = 1
// for a single simulation: all 256 threads stride over voxels).  Lead positions and layer constants live in
// registers; fp32 sums over 16 voxels are folded into f64; the voxel lanes are reduced in a fixed order.
template <int NL>
__global__ void __launch_bounds__(256) ecg_moment_kernel(const MomentArgs a) {
	__shared__ double s_red[256 * NL * 3];
	const Segment sg = a.segs[blockIdx.x];
	const int vb = 1 << a.vb_shift;
	const int vs = threadIdx.x & (vb - 1), vl = threadIdx.x >> a.vb_shift, lanes = 256 >> a.vb_shift;
	const int b = blockIdx.y * vb + vs;
	const int bb = min(b, a.B - 1);   // surplus vector slots shadow the last vector, their sums are dropped

	float lh[NL * 3], ll[NL * 3];
#pragma unroll
	for (int r = 0; r < NL * 3; ++r) {
		const int l = a.lead0 + r / 3;
		const double c = l < a.L ? a.leads[((int64_t)bb * a.L + l) * 3 + r % 3] : 0.0;
		lh[r] = (float)c;
		ll[r] = (float)(c - (double)lh[r]);
	}
	const float* Pv = a.params + ((int64_t)bb * a.n_layers + (sg.layer - 1)) * kParamStride;
	const float v4 = __ldg(Pv + 1), v5 = __ldg(Pv + 2), t0 = __ldg(Pv + 11);
	const float v45 = -(v4 + v5), nv5 = -v5;

	float f0[NL], f1[NL], f2[NL];
	double d0[NL], d1[NL], d2[NL];
#pragma unroll
	for (int l = 0; l < NL; ++l) { f0[l] = f1[l] = f2[l] = 0.f; d0[l] = d1[l] = d2[l] = 0.0; }
	int pending = 0;
	for (int j = sg.begin + vl; j < sg.end; j += lanes) {
		const uint32_t pos = __ldg(a.pos + j);
		const uint32_t mask = __ldg(a.mask + j);
		const float da = __ldg(a.at32 + j) - t0;
		const float pz = __uint_as_float(0x4B000000u | ((pos >> 22) + 1u)) - 8388608.f;
		const float py = __uint_as_float(0x4B000000u | (((pos >> 11) & 0x7ffu) + 1u)) - 8388608.f;
		const float px = __uint_as_float(0x4B000000u | ((pos & 0x7ffu) + 1u)) - 8388608.f;
		const float h1 = mufu_ex2(fminf(v45 * da, 60.f));   // same clamp as the HOISTED kernel and its table
		const float h2 = mufu_ex2(fminf(nv5 * da, 60.f));
#pragma unroll
		for (int l = 0; l < NL; ++l) {
			const float rz = (lh[3 * l] - pz) + ll[3 * l];
			const float ry = (lh[3 * l + 1] - py) + ll[3 * l + 1];
			const float rx = (lh[3 * l + 2] - px) + ll[3 * l + 2];
			float g = 0.f, sz = 0.f, sy = 0.f, sx = 0.f;
			for (int k = 0; k < a.nbr.n; ++k) {
				if ((mask >> a.nbr.bit[k]) & 1u) {
					const float dz = a.nbr.fz[k], dy = a.nbr.fy[k], dx = a.nbr.fx[k];
					const float qz = rz + dz, qy = ry + dy, qx = rx + dx;
					const float sq = fmaf(qx, qx, fmaf(qy, qy, qz * qz));
					const float dot = fmaf(dz, qz, fmaf(dy, qy, dx * qx));
					g = fmaf(dot, inv_cube(sq), g);
					sz += dz; sy += dy; sx += dx;
				}
			}
			const float sqc = fmaf(rx, rx, fmaf(ry, ry, rz * rz));
			g = fmaf(fmaf(sz, rz, fmaf(sy, ry, sx * rx)), inv_cube(sqc), g);
			const float G = -g;
			f0[l] += G;
			f1[l] = fmaf(G, h1, f1[l]);
			f2[l] = fmaf(G, h2, f2[l]);
		}
		if (++pending == 16) {
			pending = 0;
#pragma unroll
			for (int l = 0; l < NL; ++l) {
				d0[l] += (double)f0[l]; d1[l] += (double)f1[l]; d2[l] += (double)f2[l];
				f0[l] = f1[l] = f2[l] = 0.f;
			}
		}
	}
#pragma unroll
	for (int l = 0; l < NL; ++l) {
		s_red[(threadIdx.x * NL + l) * 3 + 0] = d0[l] + (double)f0[l];
		s_red[(threadIdx.x * NL + l) * 3 + 1] = d1[l] + (double)f1[l];
		s_red[(threadIdx.x * NL + l) * 3 + 2] = d2[l] + (double)f2[l];
	}
	__syncthreads();
	// thread (vector slot vs, lead l, moment q) adds the voxel lanes in lane order
	for (int o = threadIdx.x; o < vb * NL * 3; o += 256) {
		const int ovs = o / (NL * 3), r = o - ovs * (NL * 3);
		const int ob = blockIdx.y * vb + ovs, lead = a.lead0 + r / 3;
		if (ob >= a.B || lead >= a.L) continue;
		double s = 0.0;
		for (int lane = 0; lane < lanes; ++lane) s += s_red[((lane << a.vb_shift) + ovs) * (NL * 3) + r];
		a.mom[(((int64_t)blockIdx.x * a.B + ob) * a.L + lead) * 3 + r % 3] = s;
	}
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2, two lanes per issued instruction) --------
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 mk2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo2(f2 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); return x; }
__device__ __forceinline__ float hi2(f2 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); return y; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
// 1/|q|^3 for two squared distances at once (two MUFU.RSQ + one packed Newton step + packed cube)
__device__ __forceinline__ f2 inv_cube2(f2 sq) {
	f2 y = mk2(mufu_rsq(lo2(sq)), mufu_rsq(hi2(sq)));
	y = mul2(y, fma2(mul2(sq, mk2(-0.5f, -0.5f)), mul2(y, y), mk2(1.5f, 1.5f)));
	return mul2(mul2(y, y), y);
}

// The stencil sum of an INTERIOR voxel (all 8 cube corners occupied) as a series.  With q = lead - voxel and
// f = 1/|q|:   g = sum_{d in {+-1}^3} d.(q+d)/|q+d|^3 = -d/ds [ sum_d f(q + s d) ]_{s=1},
//              sum_d f(q + s d) = 8 cosh(s d/dx) cosh(s d/dy) cosh(s d/dz) f.
// f is harmonic, so the s^2 term vanishes and the rest are cubic harmonics:  with u = 1/|q|^2, xi_i = q_i^2 u,
// e2 = xi_z xi_y + xi_y xi_x + xi_x xi_z, e3 = xi_z xi_y xi_x
//     g = |q|^-5 [ (112 - 560 e2) + u (3024 e2 - 33264 e3 - 288) + u^2 (-51480 e2^2 + 14256 e2 + 41184 e3 - 792) + O(u^3) ]
// The 8 corner terms are O(|q|^-2) each and cancel down to this O(|q|^-5) remainder -- summed directly in fp32 they
// leave 1e-4..1e-2 relative noise on g (it averages out over the model, which is why the direct sum passes its
// tolerance); the series has no cancellation (3e-6 relative in fp32), a truncation error below 7e-8 relative for
// |q| >= 32 voxels (3e-8 measured at 35, falling like |q|^-6) and costs 24 packed operations + 2 MUFU per lead
// pair instead of ~110 + 16.  Segments with a lead closer than sqrt(kSeriesMinR2) take the direct sum.
constexpr float kSeriesMinR2 = 1024.f;
__device__ __forceinline__ f2 c2(float c) { return mk2(c, c); }
__device__ __forceinline__ f2 corner_series2(f2 qz, f2 qy, f2 qx, f2 r2) {
	// MUFU.RSQ as it comes (2 ulp): with nothing to cancel, 1e-6 relative on g is far inside the budget (the direct sum
	// needs its Newton step because its terms cancel by 4..6 orders of magnitude)
	const f2 y = mk2(mufu_rsq(lo2(r2)), mufu_rsq(hi2(r2)));
	const f2 u = mul2(y, y);
	const f2 xz = mul2(qz, u), xy = mul2(qy, u), xx = mul2(qx, u);
	const f2 zy = mul2(xz, xy);
	const f2 e2 = fma2(add2(xz, xy), xx, zy);
	const f2 e3 = mul2(zy, xx);
	const f2 a4 = fma2(c2(-560.f), e2, c2(112.f));
	const f2 a6 = fma2(c2(3024.f), e2, fma2(c2(-33264.f), e3, c2(-288.f)));
	const f2 a8 = fma2(fma2(c2(-51480.f), e2, c2(14256.f)), e2, fma2(c2(41184.f), e3, c2(-792.f)));
	return mul2(mul2(mul2(u, u), y), fma2(fma2(a8, u, a6), u, a4));
}

// ---- the moment kernels of the reference's default stencil "3D4" (= the 8 cube corners, sim_lib.h:133-141) ----------
// Two leads ride in the two lanes of packed fp32x2 instructions (NP = lead pairs per pass).  Every layer's part of the
// voxel list starts with its interior voxels (capi.cu, build_ecg_list) and the moment segments do not mix the two kinds:
//   ecg_moment_interior_kernel  interior segments by the series above: one short dependent chain per (voxel, vector),
//                               48 registers.  A CTA that meets a lead closer than sqrt(kSeriesMinR2) raises its
//                               near_flag and stores nothing;
//   ecg_moment_corners_kernel   boundary segments -- and the flagged (or, with EKG_FLAG_CORNER_SUM, all) interior segments --
//                               by the direct sum over the occupied corners + the centre term.
// Both: threads of a CTA = VB parameter vectors x 256/VB voxel lanes; one 16-byte record per voxel (capi.cu,
// gather_at_kernel), the record of the next iteration requested before this one is worked on (an iteration is one
// dependent chain; the load latency would otherwise be the larger part of it); fp32 sums over 16 voxels in registers,
// their f64 totals in the thread's own column of s_red ([row][thread], row = lead * 3 + moment), voxel lanes reduced in
// a fixed order -> deterministic.
template <int NP>
struct MomentAcc {
	f2 f0[NP], f1[NP], f2s[NP];
	__device__ __forceinline__ void init(double* s_red) {
#pragma unroll
		for (int p = 0; p < NP; ++p) f0[p] = f1[p] = f2s[p] = mk2(0.f, 0.f);
#pragma unroll
		for (int r = 0; r < NP * 6; ++r) s_red[r * 256 + threadIdx.x] = 0.0;
	}
	// G = -g
	__device__ __forceinline__ void add(int p, f2 g, float h1, float h2) {
		f0[p] = sub2(f0[p], g);
		f1[p] = fma2(g, mk2(-h1, -h1), f1[p]);
		f2s[p] = fma2(g, mk2(-h2, -h2), f2s[p]);
	}
	__device__ __forceinline__ void fold(double* s_red) {
#pragma unroll
		for (int p = 0; p < NP; ++p) {
			double* o = s_red + (p * 6) * 256 + threadIdx.x;
			o[0 * 256] += (double)lo2(f0[p]); o[1 * 256] += (double)lo2(f1[p]); o[2 * 256] += (double)lo2(f2s[p]);
			o[3 * 256] += (double)hi2(f0[p]); o[4 * 256] += (double)hi2(f1[p]); o[5 * 256] += (double)hi2(f2s[p]);
			f0[p] = f1[p] = f2s[p] = mk2(0.f, 0.f);
		}
	}
};

// lead coordinates (z, y, x) of the two leads of a pair, rounded to fp32: a lead displaced by < 2^-17 voxel for ALL
// voxels alike (the time-loop kernels carry a lo part as well; it changes the ECG by ~1e-7 of its peak)
template <int NP>
__device__ __forceinline__ void load_lead_pairs(const MomentArgs& a, int bb, f2 (&lh)[NP][3]) {
#pragma unroll
	for (int p = 0; p < NP; ++p)
#pragma unroll
		for (int c = 0; c < 3; ++c) {
			float h[2];
#pragma unroll
			for (int e = 0; e < 2; ++e) {
				const int lead = a.lead0 + 2 * p + e;
				h[e] = (float)(lead < a.L ? a.leads[((int64_t)bb * a.L + lead) * 3 + c] : 1e6);   // a far-away dummy lead
			}
			lh[p][c] = mk2(h[0], h[1]);
		}
}

// after the loop: thread (vector slot, lead, moment) adds the voxel lanes in lane order
template <int NP>
__device__ __forceinline__ void store_moments(const MomentArgs& a, const double* s_red) {
	constexpr int NL = NP * 2;
	const int vb = 1 << a.vb_shift, lanes = 256 >> a.vb_shift;
	for (int o = threadIdx.x; o < vb * NL * 3; o += 256) {
		const int ovs = o / (NL * 3), r = o - ovs * (NL * 3);
		const int ob = blockIdx.y * vb + ovs, lead = a.lead0 + r / 3;
		if (ob >= a.B || lead >= a.L) continue;
		double s = 0.0;
		for (int lane = 0; lane < lanes; ++lane) s += s_red[r * 256 + (lane << a.vb_shift) + ovs];
		a.mom[(((int64_t)blockIdx.x * a.B + ob) * a.L + lead) * 3 + r % 3] = s;
	}
}

// The voxels of a segment, thread by thread: lane vl of `lanes` takes the voxels begin + vl, begin + vl + lanes, ...  in
// blocks of 16 with a fixed trip count (no per-voxel bookkeeping: the loads of an unrolled group are issued together, the
// fp32 sums are folded into their f64 totals after every block), then the remainder.
template <int UNROLL, class Body, class Fold>
__device__ __forceinline__ void walk_voxels(const float4* __restrict__ vox, int begin, int end, int vl, int lanes, Body&& body, Fold&& fold) {
	const int n = end - begin;
	int n_mine = vl < n ? (n - vl + lanes - 1) / lanes : 0;
	const float4* pv = vox + begin + vl;
	for (; n_mine >= 16; n_mine -= 16) {
#pragma unroll UNROLL
		for (int u = 0; u < 16; ++u, pv += lanes) body(__ldg(pv));
		fold();
	}
	for (; n_mine > 0; --n_mine, pv += lanes) body(__ldg(pv));
	fold();
}

// series walk over an interior segment; returns the smallest squared lead distance seen by this thread
template <int NP>
__device__ __forceinline__ float moment_series_walk(const MomentArgs& a, const Segment& sg, const f2 (&lh)[NP][3], float v45, float nv5, float t0,
                                                    int vl, int lanes, MomentAcc<NP>& acc, double* s_red) {
	float r2_min = 3.0e38f;
	walk_voxels<4>(a.vox, sg.begin, sg.end, vl, lanes,
		[&](const float4 vx) {
			const float da = vx.w - t0;
			const float h1 = mufu_ex2(fminf(v45 * da, 60.f));   // same clamp as the HOISTED kernel and its table
			const float h2 = mufu_ex2(fminf(nv5 * da, 60.f));
			const f2 PZ = mk2(vx.x, vx.x), PY = mk2(vx.y, vx.y), PX = mk2(vx.z, vx.z);
#pragma unroll
			for (int p = 0; p < NP; ++p) {
				const f2 rz = sub2(lh[p][0], PZ), ry = sub2(lh[p][1], PY), rx = sub2(lh[p][2], PX);
				const f2 qz = mul2(rz, rz), qy = mul2(ry, ry), qx = mul2(rx, rx);
				const f2 r2 = add2(add2(qz, qy), qx);
				r2_min = fminf(r2_min, fminf(lo2(r2), hi2(r2)));
				acc.add(p, corner_series2(qz, qy, qx, r2), h1, h2);
			}
		},
		[&]() { acc.fold(s_red); });
	return r2_min;
}

template <int NP>
__global__ void __launch_bounds__(256, NP == 1 ? 4 : 3) ecg_moment_interior_kernel(const MomentArgs a) {
	__shared__ double s_red[256 * NP * 6];
	const Segment sg = a.segs[blockIdx.x];
	if (sg.kind != kSegInterior) return;
	const int vb = 1 << a.vb_shift;
	const int vs = threadIdx.x & (vb - 1), vl = threadIdx.x >> a.vb_shift, lanes = 256 >> a.vb_shift;
	const int bb = min(blockIdx.y * vb + vs, a.B - 1);   // surplus vector slots shadow the last vector, their sums are dropped
	f2 lh[NP][3];
	load_lead_pairs<NP>(a, bb, lh);
	const float* Pv = a.params + ((int64_t)bb * a.n_layers + (sg.layer - 1)) * kParamStride;
	const float v4 = __ldg(Pv + 1), v5 = __ldg(Pv + 2), t0 = __ldg(Pv + 11);
	MomentAcc<NP> acc;
	acc.init(s_red);
	const float r2_min = moment_series_walk<NP>(a, sg, lh, -(v4 + v5), -v5, t0, vl, lanes, acc, s_red);
	// a lead inside the series' validity radius (NaN counts as near): leave this (segment, vector group) to the direct sum
	const int near = __syncthreads_or(!(r2_min >= kSeriesMinR2));
	if (threadIdx.x == 0) a.near_flag[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = near;
	if (near) return;
	store_moments<NP>(a, s_red);
}

// The direct sum.  With d = (+-1, +-1, +-1) the neighbour offsets fold into six per-axis terms p = r + 1, m = 1 - r:
// |r + d|^2 = (p|m)_z^2 + (p|m)_y^2 + (p|m)_x^2 and d . (r + d) = (p|m)_z + (p|m)_y + (p|m)_x, 4 packed instructions + 2 MUFU +
// 7 packed for the inverse cube per corner and lead pair (the generic loop issues ~35 scalar instructions per corner and lead).
template <int NP>
__device__ __forceinline__ void moment_corner_walk(const MomentArgs& a, const Segment& sg, const bool interior, const f2 (&lh)[NP][3], float v45, float nv5,
                                                   float t0, int vl, int lanes, MomentAcc<NP>& acc, double* s_red) {
	const f2 one = mk2(1.f, 1.f);
	walk_voxels<2>(a.vox, sg.begin, sg.end, vl, lanes,
		[&](const float4 vx) {
			const float da = vx.w - t0;
			// boundary voxels carry the occupancy of their corners, order (dz, dy, dx) = (-,-,-), (-,-,+), ... (+,+,+), in the low
			// 8 mantissa bits of y (zero for an integer <= 2048)
			const uint32_t yb = __float_as_uint(vx.y);
			const uint32_t mask = interior ? 0xffu : (yb & 0xffu);
			const float pz = vx.x, py = __uint_as_float(yb & ~0xffu), px = vx.z;
			const float h1 = mufu_ex2(fminf(v45 * da, 60.f));
			const float h2 = mufu_ex2(fminf(nv5 * da, 60.f));
			constexpr uint32_t kZp = 0xf0u, kYp = 0xccu, kXp = 0xaau;   // corners with dz (dy, dx) = +1
			const f2 PZ = mk2(pz, pz), PY = mk2(py, py), PX = mk2(px, px);
#pragma unroll
			for (int p = 0; p < NP; ++p) {
				const f2 rz = sub2(lh[p][0], PZ);
				const f2 ry = sub2(lh[p][1], PY);
				const f2 rx = sub2(lh[p][2], PX);
				const f2 zt[2] = {sub2(one, rz), add2(rz, one)};   // [0]: d = -1 -> -(r - 1),  [1]: d = +1 -> r + 1
				const f2 yt[2] = {sub2(one, ry), add2(ry, one)};
				const f2 xt[2] = {sub2(one, rx), add2(rx, one)};
				const f2 zq[2] = {mul2(zt[0], zt[0]), mul2(zt[1], zt[1])};
				const f2 yq[2] = {mul2(yt[0], yt[0]), mul2(yt[1], yt[1])};
				const f2 xq[2] = {mul2(xt[0], xt[0]), mul2(xt[1], xt[1])};
				f2 g = mk2(0.f, 0.f);
				auto corners = [&](auto all_tag) {
					constexpr bool ALL = decltype(all_tag)::value;
#pragma unroll
					for (int zy = 0; zy < 4; ++zy) {
						const f2 sq_zy = add2(zq[zy >> 1], yq[zy & 1]);
						const f2 dt_zy = add2(zt[zy >> 1], yt[zy & 1]);
#pragma unroll
						for (int x = 0; x < 2; ++x) {
							if (ALL || (mask & (1u << (zy * 2 + x))))
								g = fma2(add2(dt_zy, xt[x]), inv_cube2(add2(sq_zy, xq[x])), g);
						}
					}
				};
				if (interior) {
					// the offsets of all 8 corners add up to zero: no centre term
					corners(std::true_type{});
				} else {
					corners(std::false_type{});
					// S = sum of the offsets of the occupied corners, per axis: (# with +1) - (# with -1), as floats via the
					// 2^23 trick (values -8..8, no I2F)
					const int n_occ = __popc(mask);
					const float szf = __uint_as_float(0x4B000000u + (uint32_t)(8 + 2 * __popc(mask & kZp) - n_occ)) - 8388616.f;
					const float syf = __uint_as_float(0x4B000000u + (uint32_t)(8 + 2 * __popc(mask & kYp) - n_occ)) - 8388616.f;
					const float sxf = __uint_as_float(0x4B000000u + (uint32_t)(8 + 2 * __popc(mask & kXp) - n_occ)) - 8388616.f;
					const f2 sqc = fma2(rx, rx, fma2(ry, ry, mul2(rz, rz)));
					const f2 sdot = fma2(mk2(szf, szf), rz, fma2(mk2(syf, syf), ry, mul2(mk2(sxf, sxf), rx)));
					g = fma2(sdot, inv_cube2(sqc), g);
				}
				acc.add(p, g, h1, h2);
			}
		},
		[&]() { acc.fold(s_red); });
}

template <int NP>
__global__ void __launch_bounds__(256, NP == 1 ? 3 : 2) ecg_moment_corners_kernel(const MomentArgs a) {
	__shared__ double s_red[256 * NP * 6];
	const Segment sg = a.segs[blockIdx.x];
	const bool interior = sg.kind == kSegInterior;   // all 8 corners of every voxel occupied: no centre term, no per-corner tests
	if (interior && !a.force_sum && !a.near_flag[(int64_t)blockIdx.y * gridDim.x + blockIdx.x]) return;
	const int vb = 1 << a.vb_shift;
	const int vs = threadIdx.x & (vb - 1), vl = threadIdx.x >> a.vb_shift, lanes = 256 >> a.vb_shift;
	const int bb = min(blockIdx.y * vb + vs, a.B - 1);
	f2 lh[NP][3];
	load_lead_pairs<NP>(a, bb, lh);
	const float* Pv = a.params + ((int64_t)bb * a.n_layers + (sg.layer - 1)) * kParamStride;
	const float v4 = __ldg(Pv + 1), v5 = __ldg(Pv + 2), t0 = __ldg(Pv + 11);
	MomentAcc<NP> acc;
	acc.init(s_red);
	moment_corner_walk<NP>(a, sg, interior, lh, -(v4 + v5), -v5, t0, vl, lanes, acc, s_red);
	__syncthreads();
	store_moments<NP>(a, s_red);
}

// Both kinds of segment in ONE launch, for small batches (a single simulation is bound by launch latencies, not by
// arithmetic): interior segments by the series, falling back to the direct sum inside the same CTA when a lead is near.
template <int NP>
__global__ void __launch_bounds__(256, NP == 1 ? 3 : 2) ecg_moment_fused_kernel(const MomentArgs a) {
	__shared__ double s_red[256 * NP * 6];
	const Segment sg = a.segs[blockIdx.x];
	const bool interior = sg.kind == kSegInterior;
	const int vb = 1 << a.vb_shift;
	const int vs = threadIdx.x & (vb - 1), vl = threadIdx.x >> a.vb_shift, lanes = 256 >> a.vb_shift;
	const int bb = min(blockIdx.y * vb + vs, a.B - 1);
	f2 lh[NP][3];
	load_lead_pairs<NP>(a, bb, lh);
	const float* Pv = a.params + ((int64_t)bb * a.n_layers + (sg.layer - 1)) * kParamStride;
	const float v4 = __ldg(Pv + 1), v5 = __ldg(Pv + 2), t0 = __ldg(Pv + 11);
	MomentAcc<NP> acc;
	acc.init(s_red);
	if (interior && !a.force_sum) {
		const float r2_min = moment_series_walk<NP>(a, sg, lh, -(v4 + v5), -v5, t0, vl, lanes, acc, s_red);
		if (!__syncthreads_or(!(r2_min >= kSeriesMinR2))) { store_moments<NP>(a, s_red); return; }
		acc.init(s_red);   // every thread resets its own column only
	}
	moment_corner_walk<NP>(a, sg, interior, lh, -(v4 + v5), -v5, t0, vl, lanes, acc, s_red);
	__syncthreads();
	store_moments<NP>(a, s_red);
}

// ECG[b][l][t] for the samples t >= t_off from the moments: one CTA per (vector, block of 32 samples), threads =
// 32 samples x 8 layer lanes (lane g takes the layers g, g + 8, ...; a single simulation still yields 13 CTAs x 256
// threads of f64 pow/exp work instead of 4 x 128).  Shared memory: the per-layer moments of this vector
// (ecg_layer_moments_kernel: segments added in order) and the lanes' partial sums, added in lane order -> deterministic.  F1, F2 by the expressions of
// ecg_ftab_kernel, kept in f64; ln(2^(k7/k6)-1) per (vector, layer) comes from ecg_params_kernel.
constexpr int kCombSamples = 32, kCombLanes = 8;

// per-(vector, layer) moments, once (not once per combine CTA): a warp per (layer, lead, moment); lane i adds the segments
// i, i + 32, ... of the layer in that order, then the 32 lane sums are combined by a fixed butterfly -- the loads are
// independent (a serial sum over ~30 segments is ~30 dependent L2 round trips) and the order is fixed -> deterministic
__global__ void __launch_bounds__(256) ecg_layer_moments_kernel(const double* __restrict__ mom, const int32_t* __restrict__ seg_first,
                                                                double* __restrict__ lmom, int B, int L, int n_layers) {
	const int b = blockIdx.x, lane = threadIdx.x & 31;
	for (int o = threadIdx.x >> 5; o < n_layers * L * 3; o += blockDim.x >> 5) {
		const int layer = o / (L * 3), r = o - layer * (L * 3);
		double s = 0.0;
		for (int sgi = seg_first[layer] + lane; sgi < seg_first[layer + 1]; sgi += 32) s += mom[((int64_t)sgi * B + b) * L * 3 + r];
#pragma unroll
		for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
		if (lane == 0) lmom[(int64_t)b * n_layers * L * 3 + o] = s;
	}
}

__global__ void __launch_bounds__(kCombSamples * kCombLanes) ecg_combine_kernel(const double* __restrict__ layer_k, const double* __restrict__ tail,
                                                                               const double* __restrict__ times, const double* __restrict__ lmom,
                                                                               const int32_t* __restrict__ seg_first, double* __restrict__ ecg,
                                                                               int B, int L, int n_layers, int T, int t_off, double t0) {
	extern __shared__ double s_dyn[];
	double* s_mom = s_dyn;                       // [n_layers][L][3]
	double* s_acc = s_dyn + n_layers * L * 3;    // [kCombLanes][4][kCombSamples + 1]
	const int b = blockIdx.y;
	for (int o = threadIdx.x; o < n_layers * L * 3; o += blockDim.x) s_mom[o] = lmom[(int64_t)b * n_layers * L * 3 + o];
	__syncthreads();
	const int ts = threadIdx.x & (kCombSamples - 1), g = threadIdx.x / kCombSamples;
	const int t = t_off + blockIdx.x * kCombSamples + ts;
	const bool valid = t < T;
	const double tt = valid ? times[t] : 0.0;
	const double lim = 60.0 * 0.69314718055994530942;
	for (int l0 = 0; l0 < L; l0 += 4) {
		double acc[4] = {0.0, 0.0, 0.0, 0.0};
		if (valid) {
			for (int layer = g; layer < n_layers; layer += kCombLanes) {
				if (seg_first[layer] == seg_first[layer + 1]) continue;   // no voxels in this layer (inside the slab)
				const double* k = layer_k + ((int64_t)b * n_layers + layer) * 9;
				const double R = 1.0 - ekg_fm::pow_(1.0 + ekg_fm::exp_(-k[7] * (tt - k[8]) + tail[(int64_t)b * n_layers + layer]), -(k[6] / k[7]));
				const double F1 = k[2] * (1.0 - k[3]) * R * ekg_fm::exp_(fmin(-(k[4] + k[5]) * (tt - t0), lim));
				const double F2 = k[2] * k[3] * R * ekg_fm::exp_(fmin(-k[5] * (tt - t0), lim));
				const double* M = s_mom + layer * L * 3;
#pragma unroll
				for (int l = 0; l < 4; ++l)
					if (l0 + l < L) acc[l] += k[0] * M[(l0 + l) * 3] + F1 * M[(l0 + l) * 3 + 1] + F2 * M[(l0 + l) * 3 + 2];
			}
		}
		if (l0) __syncthreads();   // the previous pass is done with s_acc
#pragma unroll
		for (int l = 0; l < 4; ++l) s_acc[(g * 4 + l) * (kCombSamples + 1) + ts] = acc[l];
		__syncthreads();
		if (threadIdx.x < 4 * kCombSamples) {
			const int l = threadIdx.x / kCombSamples;   // one (lead, sample) per thread of the first four warps
			const int tw = t_off + blockIdx.x * kCombSamples + ts;
			if (l0 + l < L && tw < T) {
				double r = s_acc[l * (kCombSamples + 1) + ts];
#pragma unroll
				for (int q = 1; q < kCombLanes; ++q) r += s_acc[(q * 4 + l) * (kCombSamples + 1) + ts];
				ecg[((int64_t)b * L + l0 + l) * T + tw] = r;
			}
		}
	}
}

// ---- partial sums -> ECG, fixed order ------------------------------------------------------------
// partial is [n_rows][n_out] (n_rows = segments x slices).  A CTA owns 32 outputs and one block of kRedRows rows
// (grid.y); its 8 warps each add every 8th row of the block (coalesced 256-byte reads, 8 independent chains instead of
// one chain of n_rows dependent loads -- a single simulation has only 800 outputs but ~10^4 rows), the 8 sub-sums are
// added in warp order.  More than one row block: the block sums go to a scratch array and further launches add those.
// Every order is fixed -> bitwise run-to-run determinism.
// (the time loop may cover only the first T_loop of T samples: row stride T on the output side of the final pass)
constexpr int kRedGroups = 8;
constexpr int kRedRows = 128;
__global__ void __launch_bounds__(32 * kRedGroups) ecg_reduce_kernel(const double* __restrict__ partial, double* __restrict__ out, int n_rows,
                                                                   int64_t n_out, int T_loop, int T, int final_pass, int rows_per_block) {
	__shared__ double s_part[kRedGroups][33];
	const int o = threadIdx.x & 31, g = threadIdx.x >> 5;
	const int64_t i = (int64_t)blockIdx.x * 32 + o;
	const int r0 = blockIdx.y * rows_per_block, r1 = min(n_rows, r0 + rows_per_block);
	double s = 0.0;
	if (i < n_out) {
#pragma unroll 4
		for (int k = r0 + g; k < r1; k += kRedGroups) s += __ldg(partial + (int64_t)k * n_out + i);
	}
	s_part[g][o] = s;
	__syncthreads();
	if (g == 0 && i < n_out) {
		double r = s_part[0][o];
#pragma unroll
		for (int q = 1; q < kRedGroups; ++q) r += s_part[q][o];
		if (final_pass) out[(i / T_loop) * T + i % T_loop] = r;
		else out[(int64_t)blockIdx.y * n_out + i] = r;
	}
}

// ---- curve comparison on the device (calculateFitness, sim.cpp:600-702; vectorMath.h) -----------------
// One CTA per (vector, lead); f64 block reductions over the n overlapping samples, offset 0.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	__syncthreads();
	if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
	__syncthreads();
	double s = 0.0;
	for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += scratch[w];
	return s;
}

__device__ __forceinline__ double block_min(double v, double* scratch) {
	for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_down_sync(0xffffffffu, v, o));
	__syncthreads();
	if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
	__syncthreads();
	double s = scratch[0];
	for (int w = 1; w < (int)(blockDim.x >> 5); ++w) s = fmin(s, scratch[w]);
	return s;
}

__global__ void __launch_bounds__(256) ecg_criteria_kernel(const double* __restrict__ ecg, const double* __restrict__ targets,
                                                           const double* __restrict__ offsets, double* __restrict__ crit,
                                                           int L, int T, int n_target, int comparison) {
	__shared__ double scratch[8];
	const int l = blockIdx.x % L;
	const double* a = ecg + (int64_t)blockIdx.x * T;
	const double* b = targets + (int64_t)l * n_target;
	const int n = min(T, n_target);
	const double inv = 1.0 / (double)n;
	double out;
	if (comparison == 2) {   // 1 - Pearson (statisticalCorrelationCoeff with offset 0)
		double sa = 0, sb = 0;
		for (int i = threadIdx.x; i < n; i += blockDim.x) { sa += a[i]; sb += b[i]; }
		const double ma = block_sum(sa, scratch) * inv, mb = block_sum(sb, scratch) * inv;
		double va = 0, vb = 0, cv = 0;
		for (int i = threadIdx.x; i < n; i += blockDim.x) { const double da = a[i] - ma, db = b[i] - mb; va += da * da; vb += db * db; cv += da * db; }
		va = block_sum(va, scratch) * inv; vb = block_sum(vb, scratch) * inv; cv = block_sum(cv, scratch) * inv;
		out = 1.0 - cv / (sqrt(va) * sqrt(vb));
	} else if (comparison == 1) {   // RMS
		double s = 0;
		for (int i = threadIdx.x; i < n; i += blockDim.x) { const double d = a[i] - b[i]; s += d * d; }
		out = sqrt(block_sum(s, scratch) * inv);
	} else if (comparison == 4) {   // 1 - vector correlation
		double ab = 0, aa = 0, bb = 0;
		for (int i = threadIdx.x; i < n; i += blockDim.x) { ab += a[i] * b[i]; aa += a[i] * a[i]; bb += b[i] * b[i]; }
		ab = block_sum(ab, scratch); aa = block_sum(aa, scratch); bb = block_sum(bb, scratch);
		out = 1.0 - ab / (sqrt(aa) * sqrt(bb));
	} else {   // deviation from linear: sqrt(var((a * aMult + ofs) / (b + ofs))), min/max of a over ALL T samples
		double mn = 1e300, mx = -1e300;
		for (int i = threadIdx.x; i < T; i += blockDim.x) { mn = fmin(mn, a[i]); mx = fmax(mx, a[i]); }
		mn = block_min(mn, scratch); mx = -block_min(-mx, scratch);
		const double mult = 1.0 / (mx - mn), ofs = offsets[l];
		double s = 0;
		for (int i = threadIdx.x; i < n; i += blockDim.x) s += (a[i] * mult + ofs) / (b[i] + ofs);
		const double mean = block_sum(s, scratch) * inv;
		double v = 0;
		for (int i = threadIdx.x; i < n; i += blockDim.x) { const double d = (a[i] * mult + ofs) / (b[i] + ofs) - mean; v += d * d; }
		v = block_sum(v, scratch) * inv;
		out = v == 0 ? 1.0 / 1e-30 : sqrt(v);
	}
	if (threadIdx.x == 0) crit[blockIdx.x] = out;
}

int run_criteria(ekg_model* m, const double* d_ecg, const double* d_targets, const double* d_offsets, double* d_crit,
                 int64_t B, int64_t L, int64_t T, int64_t n_target, int comparison, cudaStream_t st) {
	if (comparison < 1 || comparison > 4) return fail(EKG_E_INVALID, "invalid comparison mode");
	if (comparison == 3 && !d_offsets) return fail(EKG_E_INVALID, "comparison mode 3 needs target offsets");
	ecg_criteria_kernel<<<(unsigned)(B * L), 256, 0, st>>>(d_ecg, d_targets, d_offsets, d_crit, (int)L, (int)T, (int)n_target, comparison);
	EKG_CUDA(cudaGetLastError());
	++m->last_launches;
	return EKG_OK;
}

// ---- per-(vector, layer) coefficient tables, f64 -> f32 -------------------------------------------
// P[0..11] = -k1 log2e, -k4 log2e, -k5 log2e, -k7 log2e, log2(2^(k7/k6)-1), -k6/k7, k2(1-k3), k2 k3,
//            k0, hi(k8), lo(k8), t0
__global__ void ecg_params_kernel(const double* __restrict__ layer_k, float* __restrict__ P, double* __restrict__ tail, int64_t n, float t0,
                                  int* __restrict__ k1min) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double* k = layer_k + i * 9;
	// smallest depolarisation rate of the batch (float bits order like ints for positive values; anything
	// else -- k1 <= 0, NaN -- becomes 0 = "never saturates")
	if (k1min) {
		atomicMin(k1min, k[1] > 0 ? __float_as_int(__double2float_rd(k[1])) : 0);
		const double decay = fmax(fabs(k[4] + k[5]), fabs(k[5]));
		atomicMax(k1min + 1, decay == decay ? __float_as_int(__double2float_ru(fmin(decay, 3e38))) : 0x7f7fffff);   // NaN -> "too large"
	}
	const double log2e = 1.4426950408889634074;
	float* p = P + i * kParamStride;
	p[0] = (float)(-k[1] * log2e);
	p[1] = (float)(-k[4] * log2e);
	p[2] = (float)(-k[5] * log2e);
	p[3] = (float)(-k[7] * log2e);
	const double tc = ekg_fm::log_(ekg_fm::pow_(2.0, k[7] / k[6]) - 1.0);   // ln(2^(k7/k6) - 1), Wohlfart.h:200
	tail[i] = tc;
	p[4] = (float)(tc * log2e);
	p[5] = (float)(-(k[6] / k[7]));
	p[6] = (float)(k[2] * (1.0 - k[3]));
	p[7] = (float)(k[2] * k[3]);
	p[8] = (float)k[0];
	const float k8hi = (float)k[8];
	p[9] = k8hi;
	p[10] = (float)(k[8] - (double)k8hi);
	p[11] = t0;
}

// HOISTED table: F1 = k2(1-k3) R(t) exp(-(k4+k5)(t-t0)),  F2 = k2 k3 R(t) exp(-k5 (t-t0)),
// R(t) = 1 - (1 + exp(-k7 (t-k8) + ln(2^(k7/k6)-1)))^(-k6/k7)        (Wohlfart.h:199-201; the
// activation time cancels in t' - k8' = (t-at) - (k8-at), simulator.cpp:156,:169)
__global__ void ecg_ftab_kernel(const double* __restrict__ layer_k, const double* __restrict__ times, float* __restrict__ F,
                                int64_t n_bl, int T, double t0) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_bl * T) return;
	const int64_t bl = i / T;
	const int t = (int)(i % T);
	const double* k = layer_k + bl * 9;
	const double tt = times[t];
	const double R = 1.0 - ekg_fm::pow_(1.0 + ekg_fm::exp_(-k[7] * (tt - k[8]) + ekg_fm::log_(ekg_fm::pow_(2.0, k[7] / k[6]) - 1.0)), -(k[6] / k[7]));
	const double lim = 60.0 * 0.69314718055994530942;  // same clamp as phase A (2^60)
	const double f1 = k[2] * (1.0 - k[3]) * R * ekg_fm::exp_(fmin(-(k[4] + k[5]) * (tt - t0), lim));
	const double f2 = k[2] * k[3] * R * ekg_fm::exp_(fmin(-k[5] * (tt - t0), lim));
	F[(bl * 2) * T + t] = (float)f1;
	F[(bl * 2 + 1) * T + t] = (float)f2;
}

// ---- host side ------------------------------------------------------------------------------------
int make_nbr_table(int nbhd, NbrTable* out) {
	// same enumeration order as Neighbourhood::create (simulator.h:376-384); bit = index in the cube list
	int n = 0, cube = 0;
	for (int a0 = 0; a0 < 3; ++a0) for (int a1 = 0; a1 < 3; ++a1) for (int a2 = 0; a2 < 3; ++a2) {
		const int d0 = abs(a0 - 1), d1 = abs(a1 - 1), d2 = abs(a2 - 1);
		const int mx3 = std::max(std::max(d0, d1), d2), mn3 = std::min(std::min(d0, d1), d2);
		const bool in_cube = mx3 == 1;
		bool keep = false;
		switch (nbhd) {
		case EKG_NBHD_2D4: keep = std::min(d1, d2) == 1 && std::max(d1, d2) == 1 && d0 == 0; break;  // callback4N :338
		case EKG_NBHD_2D8: keep = std::max(d1, d2) == 1 && d0 == 0; break;                            // callback8N :348
		case EKG_NBHD_3D4: keep = mn3 == 1 && mx3 == 1; break;                                          // callback3x2N :357
		case EKG_NBHD_3D8: keep = mx3 == 1; break;                                                      // callbackCube :367
		default: return fail(EKG_E_INVALID, "unknown neighbourhood (only know of these: 2D4, 3D4, 2D8, 3D8 (cube))");
		}
		if (keep) {
			out->dz[n] = (int8_t)(a0 - 1); out->dy[n] = (int8_t)(a1 - 1); out->dx[n] = (int8_t)(a2 - 1);
			out->fz[n] = (float)(a0 - 1); out->fy[n] = (float)(a1 - 1); out->fx[n] = (float)(a2 - 1);
			out->bit[n] = (int8_t)cube;
			++n;
		}
		if (in_cube) ++cube;
	}
	out->n = n;
	return EKG_OK;
}

template <class T>
int ensure(T** p, int64_t* cap, int64_t need) {
	if (need <= *cap) return EKG_OK;
	if (*p) cudaFree(*p);
	*p = nullptr; *cap = 0;
	cudaError_t e = cudaMalloc((void**)p, (size_t)need * sizeof(T));
	if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(scratch)", __FILE__, __LINE__);
	*cap = need;
	return EKG_OK;
}
template int ensure<float>(float**, int64_t*, int64_t);
template int ensure<double>(double**, int64_t*, int64_t);
template int ensure<Segment>(Segment**, int64_t*, int64_t);
template int ensure<PairTile>(PairTile**, int64_t*, int64_t);

static int build_segments(ekg_model* m, int64_t seg_len, cudaStream_t st) {
	if (m->seg_len == seg_len && m->n_segs > 0) return EKG_OK;
	std::vector<Segment> segs;
	for (int l = 1; l <= m->n_layers; ++l) {
		const int64_t b0 = m->layer_off[l - 1], b1 = m->layer_off[l];
		const int64_t cnt = b1 - b0;
		if (cnt <= 0) continue;
		const int64_t pieces = (cnt + seg_len - 1) / seg_len;
		for (int64_t p = 0; p < pieces; ++p) {
			Segment s;
			s.begin = (int32_t)(b0 + cnt * p / pieces);
			s.end = (int32_t)(b0 + cnt * (p + 1) / pieces);
			s.layer = l;
			s.kind = 0;
			if (s.end > s.begin) segs.push_back(s);
		}
	}
	int rc = ensure(&m->d_segs, &m->segs_cap, (int64_t)segs.size());
	if (rc) return rc;
	// the pageable source is consumed before cudaMemcpyAsync returns
	EKG_CUDA(cudaMemcpyAsync(m->d_segs, segs.data(), segs.size() * sizeof(Segment), cudaMemcpyHostToDevice, st));
	EKG_CUDA(cudaStreamSynchronize(st));
	m->n_segs = (int64_t)segs.size();
	m->seg_len = seg_len;
	return EKG_OK;
}

template <int MODE, int NL>
static int launch_ecg(const EcgArgs& a, dim3 grid, int threads, cudaStream_t st) {
	ecg_kernel<MODE, NL><<<grid, threads, 0, st>>>(a);
	EKG_CUDA(cudaGetLastError());
	return EKG_OK;
}

template int ensure<int32_t>(int32_t**, int64_t*, int64_t);

// segment table of the SEPARABLE path + the first segment of every layer
static int build_moment_segments(ekg_model* m, int64_t seg_len, cudaStream_t st) {
	if (m->mseg_len == seg_len && m->n_msegs > 0) return EKG_OK;
	std::vector<Segment> segs;
	std::vector<int32_t> first(m->n_layers + 1, 0);
	// pieces of one kind: interior voxels (all 8 cube corners occupied; they lead every layer's range of the list) and
	// boundary voxels never share a segment.  A boundary voxel costs ~3x an interior one: half-length pieces.
	auto cut = [&](int64_t b0, int64_t cnt, int64_t len, int l, int32_t kind) {
		if (cnt <= 0) return;
		const int64_t pieces = (cnt + len - 1) / len;
		for (int64_t p = 0; p < pieces; ++p) {
			Segment sg;
			sg.begin = (int32_t)(b0 + cnt * p / pieces);
			sg.end = (int32_t)(b0 + cnt * (p + 1) / pieces);
			sg.layer = l;
			sg.kind = kind;
			if (sg.end > sg.begin) segs.push_back(sg);
		}
	};
	for (int l = 1; l <= m->n_layers; ++l) {
		first[l - 1] = (int32_t)segs.size();
		const int64_t b0 = m->layer_off[l - 1], cnt = m->layer_off[l] - b0, n_in = m->interior_cnt[l - 1];
		cut(b0, n_in, seg_len, l, kSegInterior);
		cut(b0 + n_in, cnt - n_in, std::max<int64_t>(seg_len / 2, 1), l, 0);
	}
	first[m->n_layers] = (int32_t)segs.size();
	int rc;
	if ((rc = ensure(&m->d_msegs, &m->msegs_cap, (int64_t)std::max<size_t>(segs.size(), 1)))) return rc;
	if ((rc = ensure(&m->d_mseg_first, &m->mseg_first_cap, (int64_t)first.size()))) return rc;
	EKG_CUDA(cudaMemcpyAsync(m->d_msegs, segs.data(), segs.size() * sizeof(Segment), cudaMemcpyHostToDevice, st));
	EKG_CUDA(cudaMemcpyAsync(m->d_mseg_first, first.data(), first.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
	EKG_CUDA(cudaStreamSynchronize(st));
	m->n_msegs = (int64_t)segs.size();
	m->mseg_len = seg_len;
	return EKG_OK;
}

static int run_ecg_one(ekg_model* m, const double* d_layer_k, const double* d_leads, int64_t B, int64_t L, int nbhd,
                       double t_start, double t_step, double total_time, int flags, double* d_ecg, cudaStream_t st, KHints hints);

// Large batches are cut into sub-batches so that the f64 partial-sum scratch stays below ~1 GiB.
int run_ecg(ekg_model* m, const double* d_layer_k, const double* d_leads, int64_t B, int64_t L, int nbhd,
            double t_start, double t_step, double total_time, int flags, double* d_ecg, cudaStream_t st, KHints hints) {
	if (B <= 0 || L <= 0) return fail(EKG_E_INVALID, "B and n_leads must be positive");
	if (!(t_step > 0) || !(total_time > 0)) return fail(EKG_E_INVALID, "t_step and total_time must be positive");
	const int64_t T = (int64_t)ceil(total_time / t_step);
	if (T <= 0 || T > (1 << 24)) return fail(EKG_E_INVALID, "bad number of time steps");
	// at ~100 waves the segment count is about (592 * 96) / (B * T / 256); partial bytes = segs * B * L * T * 8
	const int64_t per_vector = std::max<int64_t>(L * T * 8, 1);
	int64_t sub = std::max<int64_t>(1, std::min<int64_t>(B, ((int64_t)1 << 30) / (per_vector * 160)));
	sub = std::min<int64_t>(sub, 16384);
	int64_t launches = 0;
	const bool timed = (flags & EKG_FLAG_TIME_KERNEL) != 0;
	m->ev_recorded = false;
	if (timed) {
		if (!m->ev_k0) { EKG_CUDA(cudaEventCreate(&m->ev_k0)); EKG_CUDA(cudaEventCreate(&m->ev_k1)); }
	}
	for (int64_t b0 = 0; b0 < B; b0 += sub) {
		const int64_t nb = std::min(sub, B - b0);
		int rc = run_ecg_one(m, d_layer_k + b0 * m->n_layers * 9, d_leads + b0 * L * 3, nb, L, nbhd, t_start, t_step, total_time, flags,
		                     d_ecg + b0 * L * T, st, hints);
		if (rc) return rc;
		launches += m->last_launches;
	}
	m->last_launches = launches;
	return EKG_OK;
}

// exp(-k (t - at)) = exp(-k (t - t0)) exp(k (at - t0)), both factors clamped at 2^60: fine as long as the largest decay rate
// times the largest distance from t0 (of an activation time, or of the first sample) stays below 55 octaves
bool decay_within_clamp(const ekg_model* m, double decay_max, double t_start) {
	const double span = std::max(std::max(m->at_max - m->t0, m->t0 - m->at_min), std::max(m->t0 - t_start, 0.0));
	return decay_max * span * 1.4426950408889634074 <= 55.0;
}

static int run_ecg_one(ekg_model* m, const double* d_layer_k, const double* d_leads, int64_t B, int64_t L, int nbhd,
                       double t_start, double t_step, double total_time, int flags, double* d_ecg, cudaStream_t st, KHints hints) {
	if (!m->have_activation) return fail(EKG_E_STATE, "no excitation sequence: call ekg_model_activation or ekg_model_set_activation first");
	if (B <= 0 || L <= 0) return fail(EKG_E_INVALID, "B and n_leads must be positive");
	if (B > 65535) return fail(EKG_E_UNSUPPORTED, "at most 65535 parameter vectors per call");
	if (!(t_step > 0) || !(total_time > 0)) return fail(EKG_E_INVALID, "t_step and total_time must be positive");
	const int64_t T = (int64_t)ceil(total_time / t_step);  // simulator.cpp:471
	if (T <= 0 || T > (1 << 24)) return fail(EKG_E_INVALID, "bad number of time steps");
	int mode = flags & 0xff;
	if (mode == EKG_MODE_DEFAULT) mode = EKG_MODE_SEPARABLE;
	if (mode != EKG_MODE_DIRECT && mode != EKG_MODE_HOISTED && mode != EKG_MODE_SEPARABLE) return fail(EKG_E_INVALID, "unknown ECG mode");

	EcgArgs a{};
	int rc = make_nbr_table(nbhd, &a.nbr);
	if (rc) return rc;
	m->last_launches = 0;

	// time samples by repeated addition (simulator.cpp:496-500), split hi/lo for the fp32 kernel
	if ((rc = ensure(&m->d_times, &m->times_cap, 2 * T + 2 * T))) return rc;  // [hi | lo | f64 times]
	double* d_t64 = reinterpret_cast<double*>(m->d_times + 2 * T);
	if (m->times_T != T || m->times_t_start != t_start || m->times_t_step != t_step) {
		std::vector<float> h_t(2 * T);
		m->h_times.resize(T);
		double st_ = t_start;
		for (int64_t i = 0; i < T; ++i) {
			m->h_times[i] = st_;
			h_t[i] = (float)st_;
			h_t[T + i] = (float)(st_ - (double)h_t[i]);
			st_ += t_step;
		}
		EKG_CUDA(cudaMemcpyAsync(m->d_times, h_t.data(), 2 * T * sizeof(float), cudaMemcpyHostToDevice, st));
		EKG_CUDA(cudaMemcpyAsync(d_t64, m->h_times.data(), T * sizeof(double), cudaMemcpyHostToDevice, st));
		EKG_CUDA(cudaStreamSynchronize(st));  // pageable sources go out of scope here
		m->times_T = T; m->times_t_start = t_start; m->times_t_step = t_step;
	}

	const int64_t n_bl = B * m->n_layers;
	if ((rc = ensure(&m->d_params, &m->params_cap, n_bl * kParamStride))) return rc;
	if ((rc = ensure(&m->d_tail, &m->tail_cap, n_bl))) return rc;
	const bool read_k1 = mode == EKG_MODE_SEPARABLE && !(hints.k1_min > 0);
	const bool want_k1 = read_k1 || hints.verify;
	if (want_k1) {
		if (!m->d_k1min) EKG_CUDA(cudaMalloc(&m->d_k1min, 2 * sizeof(int)));
		EKG_CUDA(cudaMemsetAsync(m->d_k1min, 0x7f, sizeof(int), st));  // 0x7f7f7f7f = 3.4e38f
		EKG_CUDA(cudaMemsetAsync(m->d_k1min + 1, 0, sizeof(int), st));
	}
	ecg_params_kernel<<<(int)((n_bl + 127) / 128), 128, 0, st>>>(d_layer_k, m->d_params, m->d_tail, n_bl, (float)m->t0, want_k1 ? m->d_k1min : nullptr);
	EKG_CUDA(cudaGetLastError());
	++m->last_launches;
	if (read_k1) {
		int bits[2] = {0, 0};
		EKG_CUDA(cudaMemcpyAsync(bits, m->d_k1min, sizeof bits, cudaMemcpyDeviceToHost, st));
		EKG_CUDA(cudaStreamSynchronize(st));
		float f[2]; memcpy(f, bits, 8);
		hints.k1_min = (double)f[0];
		hints.decay_max = (double)f[1];
	}
	// The HOISTED and SEPARABLE kernels factor exp(-k (t - at)) = exp(-k (t - t0)) exp(k (at - t0)) and clamp both exponents
	// at 2^60; a batch whose decay rates or whose distance from t0 would reach the clamp runs through DIRECT instead (known
	// rates only: an explicit HOISTED request with raw device pointers stays asynchronous and unchecked)
	if (mode != EKG_MODE_DIRECT && hints.k1_min > 0 && !decay_within_clamp(m, hints.decay_max, t_start)) mode = EKG_MODE_DIRECT;

	// SEPARABLE: the time-loop kernel handles the first T_loop samples (those before every voxel's
	// depolarisation sigmoid is exactly 1 in fp32: exp(-k1 (t - at)) < 2^-25, the HOISTED kernel's own
	// saturation test), the moment path the rest.  T_loop = 0 for runs that start after the QRS complex.
	int64_t T_loop = T;
	int loop_mode = mode;
	if (mode == EKG_MODE_SEPARABLE) {
		loop_mode = EKG_MODE_HOISTED;
		const double k1_min = hints.k1_min;
		const size_t smem_need = (size_t)(m->n_layers * L * 3 + kCombLanes * 4 * (kCombSamples + 1)) * sizeof(double);
		if (k1_min > 0 && smem_need <= 48 * 1024) {
			// k1 log2e (t - at) > 25 (+ a margin of 1e-3 ms for the fp32 evaluation of the same test)
			const double t_sat = m->at_max + 25.0 / (k1_min * 1.4426950408889634074) + 1e-3;
			T_loop = 0;
			while (T_loop < T && !(m->h_times[T_loop] > t_sat)) ++T_loop;
		}
	}

	const bool timed = (flags & EKG_FLAG_TIME_KERNEL) != 0;
	bool need_k0 = timed && !m->ev_recorded;   // first sub-batch of a timed call: event before the first main kernel

	if (T_loop > 0) {
		// voxel slices: S * B * T_loop a whole number of tiles when the batch alone would leave a noticeable part of its
		// last tile empty (see ecg_kernel); S = 1 for anything with more than a few dozen tiles
		int64_t S = 1;
		{
			const int64_t P1 = B * T_loop, tiles1 = (P1 + kEcgThreads - 1) / kEcgThreads;
			static const int64_t max_s = getenv("EKGSIM_B200_ECG_SLICES") ? std::max(1, atoi(getenv("EKGSIM_B200_ECG_SLICES"))) : 16;
			if (P1 % kEcgThreads != 0 && tiles1 <= 32) {
				int64_t g = P1, h = kEcgThreads;
				while (h) { const int64_t r = g % h; g = h; h = r; }
				S = std::min<int64_t>(kEcgThreads / g, max_s);
			}
		}
		const int64_t VB = S * B;   // virtual vectors
		// work decomposition: pair tiles (<= kEcgThreads pairs, <= kMaxVecPerTile virtual vectors each) x segments
		if (m->tiles_B != VB || m->tiles_T != T_loop) {
			std::vector<PairTile> tiles;
			const int64_t P = VB * T_loop;
			if (P >= ((int64_t)1 << 31)) return fail(EKG_E_UNSUPPORTED, "B * n_steps must be below 2^31");
			for (int64_t p = 0; p < P;) {
				int64_t e = std::min<int64_t>(p + kEcgThreads, P);
				const int64_t b0 = p / T_loop;
				if ((e - 1) / T_loop - b0 + 1 > kMaxVecPerTile) e = (b0 + kMaxVecPerTile) * T_loop;
				tiles.push_back(PairTile{(int32_t)p, (int32_t)e});
				p = e;
			}
			if ((rc = ensure(&m->d_tiles, &m->tiles_cap, (int64_t)tiles.size()))) return rc;
			EKG_CUDA(cudaMemcpyAsync(m->d_tiles, tiles.data(), tiles.size() * sizeof(PairTile), cudaMemcpyHostToDevice, st));
			EKG_CUDA(cudaStreamSynchronize(st));
			m->n_tiles = (int64_t)tiles.size(); m->tiles_B = VB; m->tiles_T = T_loop;
		}
		// ~100 waves of CTAs: the tail of the last wave costs about 1/waves of the launch.  A thread walks seg_len / S
		// voxels: at least kMinSub of them (a CTA pays ~1 us of fixed cost: parameter loads, phase A, two barriers)
		static const int64_t kMinSub = getenv("EKGSIM_B200_ECG_SUB") ? std::max(32, atoi(getenv("EKGSIM_B200_ECG_SUB"))) : 128;
		const int64_t target_ctas = (int64_t)m->sm_count * 4 * 96;
		int64_t want_segs = (target_ctas + m->n_tiles - 1) / m->n_tiles;
		int64_t seg_len = (m->n_ecg + want_segs - 1) / std::max<int64_t>(want_segs, 1);
		if (S == 1) {
			seg_len = std::max<int64_t>(kChunk, std::min<int64_t>(seg_len, 16384));
			seg_len = (seg_len + kChunk - 1) / kChunk * kChunk;
		} else {
			seg_len = std::max<int64_t>(kMinSub * S, std::min<int64_t>(seg_len, 16384));
			seg_len = (seg_len + S - 1) / S * S;
		}
		if ((rc = build_segments(m, seg_len, st))) return rc;
		if (m->n_tiles > 65535) return fail(EKG_E_UNSUPPORTED, "too many (vector, sample) pairs for one launch (B * n_steps <= 16.7 M)");

		const int64_t n_out = B * L * T_loop;
		if ((rc = ensure(&m->d_partial, &m->partial_cap, m->n_segs * S * n_out))) return rc;
		if (loop_mode == EKG_MODE_HOISTED) {
			if ((rc = ensure(&m->d_ftab, &m->ftab_cap, n_bl * 2 * T_loop))) return rc;
			ecg_ftab_kernel<<<(int)((n_bl * T_loop + 255) / 256), 256, 0, st>>>(d_layer_k, d_t64, m->d_ftab, n_bl, (int)T_loop, (double)(float)m->t0);
			EKG_CUDA(cudaGetLastError());
			++m->last_launches;
		}

		a.pos = m->d_pos; a.mask = m->d_mask; a.at32 = m->d_at32; a.segs = m->d_segs; a.tiles = m->d_tiles;
		a.params = m->d_params; a.ftab = m->d_ftab; a.leads = d_leads;
		a.t_hi = m->d_times; a.t_lo = m->d_times + T;
		a.partial = m->d_partial;
		a.n_segs = (int32_t)m->n_segs; a.B = (int32_t)B; a.L = (int32_t)L; a.T = (int32_t)T_loop; a.n_layers = m->n_layers;
		a.S = (int32_t)S;

		const dim3 grid((unsigned)m->n_segs, (unsigned)m->n_tiles, 1);
		const int threads = kEcgThreads;
		if (need_k0) { EKG_CUDA(cudaEventRecord(m->ev_k0, st)); need_k0 = false; }
		for (int lead0 = 0; lead0 < L; lead0 += kMaxLeadsPerPass) {
			a.lead0 = lead0;
			const int nl = (int)std::min<int64_t>(kMaxLeadsPerPass, L - lead0);
			if (loop_mode == EKG_MODE_DIRECT) {
				if (nl <= 2) rc = launch_ecg<MODE_DIRECT, 2>(a, grid, threads, st);
				else rc = launch_ecg<MODE_DIRECT, 4>(a, grid, threads, st);
				m->last_kernel = "ecg_kernel<DIRECT>";
			} else {
				if (nl <= 2) rc = launch_ecg<MODE_HOISTED, 2>(a, grid, threads, st);
				else rc = launch_ecg<MODE_HOISTED, 4>(a, grid, threads, st);
				m->last_kernel = "ecg_kernel<HOISTED>";
			}
			if (rc) return rc;
			++m->last_launches;
		}
		if (T_loop == T && timed) { EKG_CUDA(cudaEventRecord(m->ev_k1, st)); m->ev_recorded = true; }
		{
			// row blocks of kRedRows rows are added in parallel, then the block sums, ... until one block is left (ping-pong
			// between the scratch array and the partial array, whose contents have been consumed by then)
			int64_t n_rows = m->n_segs * S;
			const unsigned gx = (unsigned)((n_out + 31) / 32);
			const double* src = m->d_partial;
			if (n_rows > 4 * kRedRows) {   // a few hundred rows are one pass
				if ((rc = ensure(&m->d_partial2, &m->partial2_cap, ((n_rows + kRedRows - 1) / kRedRows) * n_out))) return rc;
				double* dst = m->d_partial2;
				while (n_rows > 4 * kRedRows) {
					const int64_t n_blocks = (n_rows + kRedRows - 1) / kRedRows;
					if (n_blocks > 65535) return fail(EKG_E_UNSUPPORTED, "too many partial sums for the reduction");
					ecg_reduce_kernel<<<dim3(gx, (unsigned)n_blocks), 32 * kRedGroups, 0, st>>>(src, dst, (int)n_rows, n_out, (int)T_loop, (int)T, 0, kRedRows);
					EKG_CUDA(cudaGetLastError());
					++m->last_launches;
					src = dst;
					dst = dst == m->d_partial2 ? m->d_partial : m->d_partial2;
					n_rows = n_blocks;
				}
			}
			ecg_reduce_kernel<<<dim3(gx, 1), 32 * kRedGroups, 0, st>>>(src, d_ecg, (int)n_rows, n_out, (int)T_loop, (int)T, 1, (int)n_rows);
			EKG_CUDA(cudaGetLastError());
			++m->last_launches;
		}
	}

	if (T_loop < T) {
		// threads of a CTA = vb vectors x 256/vb voxel lanes
		int vb_shift = 0;
		while (vb_shift < 5 && (1 << vb_shift) < B) ++vb_shift;
		const int64_t groups = (B + (1 << vb_shift) - 1) >> vb_shift;
		const int64_t lanes = 256 >> vb_shift;
		// segments differ in cost by up to 3x (boundary voxels take the 9-term sum, the first and last layer are all
		// boundary): ~48 CTAs per SM keep the tail of the launch short; at least 4 voxels per thread
		static const int64_t ctas_per_sm = getenv("EKGSIM_B200_MOMENT_CTAS") ? std::max(1, atoi(getenv("EKGSIM_B200_MOMENT_CTAS"))) : 48;
		const int64_t want_segs = std::max<int64_t>(1, ((int64_t)m->sm_count * ctas_per_sm + groups - 1) / groups);
		int64_t seg_len = (m->n_ecg + want_segs - 1) / want_segs;
		seg_len = std::max<int64_t>(lanes * 4, std::min<int64_t>(seg_len, (int64_t)1 << 20));
		seg_len = (seg_len + 255) / 256 * 256;
		if ((rc = build_moment_segments(m, seg_len, st))) return rc;
		if (groups > 65535) return fail(EKG_E_UNSUPPORTED, "too many parameter vectors for one launch");
		if ((rc = ensure(&m->d_mom, &m->mom_cap, std::max<int64_t>(m->n_msegs, 1) * B * L * 3))) return rc;
		if ((rc = ensure(&m->d_near, &m->near_cap, std::max<int64_t>(m->n_msegs, 1) * groups))) return rc;
		MomentArgs ma{};
		ma.pos = m->d_pos; ma.mask = m->d_mask; ma.at32 = m->d_at32; ma.segs = m->d_msegs; ma.params = m->d_params; ma.leads = d_leads;
		ma.vox = m->d_vox; ma.near_flag = m->d_near;
		// EKG_FLAG_CORNER_SUM: interior voxels through the direct corner sum as well (the cross-check of the series)
		ma.force_sum = (flags & EKG_FLAG_CORNER_SUM) ? 1 : 0;
		ma.mom = m->d_mom; ma.B = (int32_t)B; ma.L = (int32_t)L; ma.n_layers = m->n_layers; ma.vb_shift = vb_shift; ma.nbr = a.nbr;
		if (need_k0) { EKG_CUDA(cudaEventRecord(m->ev_k0, st)); need_k0 = false; }
		// the 8-corner stencil ("3D4") has its own packed-fp32x2 kernels; make sure the table is what they hard-code
		static const int corner_bit[8] = {0, 2, 6, 8, 17, 19, 23, 25};
		bool corners = a.nbr.n == 8 && getenv("EKGSIM_B200_GENERIC_MOMENTS") == nullptr;
		for (int k = 0; corners && k < 8; ++k)
			corners = a.nbr.bit[k] == corner_bit[k] && a.nbr.dz[k] == ((k & 4) ? 1 : -1) && a.nbr.dy[k] == ((k & 2) ? 1 : -1) && a.nbr.dx[k] == ((k & 1) ? 1 : -1);
		// small batches: one launch for both kinds of segment (EKGSIM_B200_MOMENT_FUSED=0/1 overrides)
		static const int fused_env = getenv("EKGSIM_B200_MOMENT_FUSED") ? atoi(getenv("EKGSIM_B200_MOMENT_FUSED")) : -1;
		const bool fused = fused_env >= 0 ? fused_env != 0 : groups <= 2;
		if (m->n_msegs > 0) {
			const dim3 grid((unsigned)m->n_msegs, (unsigned)groups, 1);
			for (int lead0 = 0; lead0 < L; lead0 += kMaxLeadsPerPass) {
				ma.lead0 = lead0;
				const int nl = (int)std::min<int64_t>(kMaxLeadsPerPass, L - lead0);
				if (corners && fused) {
					if (nl <= 2) ecg_moment_fused_kernel<1><<<grid, 256, 0, st>>>(ma);
					else ecg_moment_fused_kernel<2><<<grid, 256, 0, st>>>(ma);
				} else if (corners) {
					// interior segments by the series (raising near_flag where a lead is too close), then the boundary
					// segments and whatever was flagged by the direct sum
					if (!ma.force_sum) {
						if (nl <= 2) ecg_moment_interior_kernel<1><<<grid, 256, 0, st>>>(ma);
						else ecg_moment_interior_kernel<2><<<grid, 256, 0, st>>>(ma);
						EKG_CUDA(cudaGetLastError());
						++m->last_launches;
					}
					if (nl <= 2) ecg_moment_corners_kernel<1><<<grid, 256, 0, st>>>(ma);
					else ecg_moment_corners_kernel<2><<<grid, 256, 0, st>>>(ma);
				}
				else if (nl <= 2) ecg_moment_kernel<2><<<grid, 256, 0, st>>>(ma);
				else ecg_moment_kernel<4><<<grid, 256, 0, st>>>(ma);
				EKG_CUDA(cudaGetLastError());
				++m->last_launches;
			}
		}
		if (timed) { EKG_CUDA(cudaEventRecord(m->ev_k1, st)); m->ev_recorded = true; }
		const int64_t n_late = T - T_loop;
		const size_t smem = (size_t)(m->n_layers * L * 3 + kCombLanes * 4 * (kCombSamples + 1)) * sizeof(double);
		const dim3 cgrid((unsigned)((n_late + kCombSamples - 1) / kCombSamples), (unsigned)B, 1);
		if ((rc = ensure(&m->d_lmom, &m->lmom_cap, B * m->n_layers * L * 3))) return rc;
		ecg_layer_moments_kernel<<<(unsigned)B, 256, 0, st>>>(m->d_mom, m->d_mseg_first, m->d_lmom, (int)B, (int)L, m->n_layers);
		EKG_CUDA(cudaGetLastError());
		++m->last_launches;
		ecg_combine_kernel<<<cgrid, kCombSamples * kCombLanes, smem, st>>>(d_layer_k, m->d_tail, d_t64, m->d_lmom, m->d_mseg_first, d_ecg, (int)B, (int)L,
		                                                                  m->n_layers, (int)T, (int)T_loop, (double)(float)m->t0);
		EKG_CUDA(cudaGetLastError());
		++m->last_launches;
		m->last_kernel = T_loop > 0 ? "ecg_kernel<HOISTED> + ecg_moment_kernel" : "ecg_moment_kernel";
		(void)corners;
	}
	return EKG_OK;
}

}  // namespace ekg
