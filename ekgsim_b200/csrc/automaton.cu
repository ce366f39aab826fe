// ekgsim_b200/csrc/automaton.cu -- activation-time automaton on the GPU.
//
// Replaces Simulation::calculateExcitationSequence + exciteElement (reference
// simlib/simulator.cpp:212-286), a priority-queue Dijkstra over the occupied voxels with the
// 26-cube (3-D) / 8 (2-D) neighbourhood.  The reference's result is the least fixed point of
//
//     d(start) = 1,   d(v) = min_u  fl( d(u) + fl( T[layer_u][layer_v] * sqrt(|dif|^2) ) )
//
// (fl = IEEE-754 double rounding).  Floating-point addition is monotone, so any label-correcting
// schedule converges to exactly the same bits as Dijkstra as long as the device performs the same
// two roundings: the edge weight table is computed on the host (IEEE sqrt and multiply, same as
// simulator.cpp:239-240) and the device adds with __dadd_rn (no FMA contraction possible).
//
// Two kernels, both persistent cooperative grids:
//
//  * automaton_brick_kernel (default) -- frontier based.  The grid is tiled by 8x8x8 bricks; a work
//    queue holds the bricks whose surroundings changed in the previous round (initially the bricks
//    of the start voxels).  A CTA takes a brick, stages it with a one-voxel halo in shared memory
//    (1000 f64 times + 1000 u8 layers), relaxes it to a LOCAL fixed point with in-place sweeps in
//    shared memory, writes the improved times back and queues the neighbouring bricks that touch a
//    changed face/edge/corner voxel for the next round.  One grid.sync per round; the number of
//    rounds is the number of brick hops of the slowest wavefront path, and only the bricks on the
//    wavefront are touched in a round.
//  * automaton_kernel (EKGSIM_B200_AUTOMATON=sweep) -- the plain label-correcting sweep over all
//    occupied voxels, kept as the simple cross-check.
//
// Reads of the time field bypass L1 (ld.global.cg) since other SMs update it; 8-byte accesses do not
// tear; values only ever decrease, so stale halo reads are harmless (the writer re-queues us).
// Working set on model_24: 1.6 MB padded u8 layers + 13 MB padded f64 times -> L2 resident.

#include <cooperative_groups.h>
#include <cstdlib>
#include <cstring>

#include "ekg_internal.cuh"

namespace cg = cooperative_groups;

namespace ekg {

__global__ void __launch_bounds__(256) automaton_kernel(AutoArgs a) {
	cg::grid_group grid = cg::this_grid();
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const double inf = __longlong_as_double(0x7ff0000000000000LL);
	int sweep = 0;
	for (; sweep < a.max_sweeps; ++sweep) {
		bool changed = false;
		for (int64_t i = gtid; i < a.n; i += stride) {
			const uint32_t p = __ldg(a.pidx + i);
			const int lv = __ldg(a.layer + p);
			const double tv = __ldcg(a.time + p);
			double best = tv;
#pragma unroll 1
			for (int k = 0; k < a.n_nbr; ++k) {
				const uint32_t q = p - a.off[k];
				const int lu = __ldg(a.layer + q);
				if (lu == 0) continue;
				const double tu = __ldcg(a.time + q);
				if (tu >= best) continue;  // cannot improve (weights are >= 0), also skips +inf
				const double w = __ldg(a.wtab + ((int64_t)lu * a.nl1 + lv) * 3 + a.sq[k]);
				const double cand = __dadd_rn(tu, w);
				if (cand < best) best = cand;
			}
			if (best < tv) {
				__stcg(a.time + p, best);
				changed = true;
			}
		}
		if (__syncthreads_or(changed) && threadIdx.x == 0) a.flags[sweep] = 1;
		__threadfence();
		grid.sync();
		if (__ldcg(a.flags + sweep) == 0) break;
	}
	if (gtid == 0) *a.sweeps_out = sweep + 1;
	(void)inf;
}

__global__ void __launch_bounds__(256) automaton_brick_kernel(BrickArgs a) {
	cg::grid_group grid = cg::this_grid();
	__shared__ double s_t[kBrickCells];
	__shared__ uint8_t s_l[kBrickCells];
	__shared__ uint8_t s_todo[2][kBrickCells];  // cell must be re-evaluated in the sweep of that parity
	__shared__ unsigned s_mask;
	extern __shared__ double s_w[];  // edge-weight table [nl1][nl1][3] when it fits (a.w_in_smem)
	const int tid = threadIdx.x;
	if (a.w_in_smem) {
		for (int i = tid; i < a.nl1 * a.nl1 * 3; i += 256) s_w[i] = __ldg(a.wtab + i);
		__syncthreads();
	}
	const double* __restrict__ wt = a.w_in_smem ? s_w : a.wtab;
	constexpr int kOwn = kBrick * kBrick * kBrick / 256;  // interior voxels per thread (2)
	constexpr int kInnerCap = 96;
	int loc[kOwn], cz[kOwn], cy[kOwn], cx[kOwn];
#pragma unroll
	for (int o = 0; o < kOwn; ++o) {
		const int v = tid + o * 256;
		cz[o] = v >> 6; cy[o] = (v >> 3) & 7; cx[o] = v & 7;
		loc[o] = ((cz[o] + 1) * kBrickHalo + (cy[o] + 1)) * kBrickHalo + (cx[o] + 1);
	}
	const double inf = __longlong_as_double(0x7ff0000000000000LL);
	int round = 0;
	int visits = 0;
	for (; round < a.max_rounds; ++round) {
		const int cur = round % 3, nxt = (round + 1) % 3, old = (round + 2) % 3;
		const int par = round & 1;
		const int n_active = __ldcg(a.counters + cur);
		if (n_active == 0) break;
		if (blockIdx.x == 0 && tid == 0) a.counters[old] = 0;  // consumed last round, appended to again next round
		for (int q = blockIdx.x; q < n_active; q += gridDim.x) {
			const int b = __ldcg(a.queue + (size_t)cur * a.n_live + q);
			const uint32_t origin = __ldg(a.origin + b);
			if (tid == 0) { a.flag[(size_t)par * a.n_live + b] = 0; s_mask = 0; ++visits; }
			for (int i = tid; i < kBrickCells; i += 256) {
				const int lz = i / (kBrickHalo * kBrickHalo), ly = (i / kBrickHalo) % kBrickHalo, lx = i % kBrickHalo;
				const uint32_t p = origin + (uint32_t)(((lz - 1) * a.pY + (ly - 1)) * a.pX + (lx - 1));
				s_l[i] = __ldg(a.layer + p);
				s_t[i] = __ldcg(a.time + p);
				s_todo[0][i] = 1;  // first sweep looks at every voxel, later ones only next to what changed
				s_todo[1][i] = 0;
			}
			__syncthreads();
			double t_init[kOwn];
			int lv[kOwn];
#pragma unroll
			for (int o = 0; o < kOwn; ++o) { t_init[o] = s_t[loc[o]]; lv[o] = s_l[loc[o]]; }
			int it = 0;
			for (; it < kInnerCap; ++it) {
				bool ch = false;
				uint8_t* todo = s_todo[it & 1];
				uint8_t* todo_next = s_todo[(it & 1) ^ 1];
#pragma unroll
				for (int o = 0; o < kOwn; ++o) {
					if (lv[o] == 0 || !todo[loc[o]]) continue;
					todo[loc[o]] = 0;
					const double tv = s_t[loc[o]];
					double best = tv;
#pragma unroll 2
					for (int k = 0; k < a.n_nbr; ++k) {
						const int qq = loc[o] - a.loff[k];
						const int lu = s_l[qq];
						if (lu == 0) continue;
						const double tu = s_t[qq];
						if (tu >= best) continue;
						const double cand = __dadd_rn(tu, wt[(lu * a.nl1 + lv[o]) * 3 + a.sq[k]]);
						if (cand < best) best = cand;
					}
					if (best < tv) {
						s_t[loc[o]] = best;
						ch = true;
						// a voxel's minimum can only move when a neighbour's time moved: wake the neighbours
						for (int k = 0; k < a.n_nbr; ++k) todo_next[loc[o] - a.loff[k]] = 1;
					}
				}
				if (!__syncthreads_or(ch)) break;
			}
			// write back what improved; queue a neighbouring brick only if one of ITS cells (our halo copy of
			// it, never smaller than its current value) would actually improve through one of our changed
			// voxels -- otherwise bricks that already hold better times would be revisited for nothing.
			// In round 0 every reached voxel counts as changed (the start voxel itself never "improves").
			unsigned bits = 0;
#pragma unroll
			for (int o = 0; o < kOwn; ++o) {
				const double tf = s_t[loc[o]];
				if (lv[o] == 0) continue;
				const bool improved = tf < t_init[o];
				if (improved) __stcg(a.time + origin + (uint32_t)((cz[o] * a.pY + cy[o]) * a.pX + cx[o]), tf);
				const bool on_face = cz[o] == 0 || cz[o] == kBrick - 1 || cy[o] == 0 || cy[o] == kBrick - 1 || cx[o] == 0 || cx[o] == kBrick - 1;
				if (!on_face || !(improved || (round == 0 && tf < inf))) continue;
				for (int k = 0; k < a.n_nbr; ++k) {
					const int qq = loc[o] - a.loff[k];
					const int nz = cz[o] - a.dz[k], ny = cy[o] - a.dy[k], nx = cx[o] - a.dx[k];  // neighbour = index - dif
					const int dz = nz < 0 ? -1 : nz >= kBrick ? 1 : 0;
					const int dy = ny < 0 ? -1 : ny >= kBrick ? 1 : 0;
					const int dx = nx < 0 ? -1 : nx >= kBrick ? 1 : 0;
					if (!(dz | dy | dx)) continue;  // interior cell
					const int lu = s_l[qq];
					if (lu == 0) continue;
					const double cand = __dadd_rn(tf, wt[(lv[o] * a.nl1 + lu) * 3 + a.sq[k]]);  // we excite it: T[ours][its]
					if (cand < s_t[qq]) {
						const int id = (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1);
						bits |= 1u << (id > 13 ? id - 1 : id);
					}
				}
			}
			if (bits) atomicOr(&s_mask, bits);
			__syncthreads();
			const int np = par ^ 1;
			if (tid < 26 && ((s_mask >> tid) & 1u)) {
				const int nb = __ldg(a.nbr + (size_t)b * 26 + tid);
				if (nb >= 0 && atomicExch(a.flag + (size_t)np * a.n_live + nb, 1) == 0)
					a.queue[(size_t)nxt * a.n_live + atomicAdd(a.counters + nxt, 1)] = nb;
			}
			if (tid == 26 && it == kInnerCap) {  // not yet locally converged: come back next round
				if (atomicExch(a.flag + (size_t)np * a.n_live + b, 1) == 0)
					a.queue[(size_t)nxt * a.n_live + atomicAdd(a.counters + nxt, 1)] = b;
			}
			__syncthreads();  // s_mask / s_t are reused by the next brick
		}
		__threadfence();
		grid.sync();
	}
	if (tid == 0 && visits) atomicAdd(a.counters + 4, visits);
	if (blockIdx.x == 0 && tid == 0) a.counters[3] = round;
}

// Negative weights would make "tu >= best -> skip" wrong and Dijkstra itself ill-defined; the
// reference's conduction matrices hold delays (>= 0) for every layer pair that can touch.

__global__ void init_time_kernel(double* time, int64_t n) {
	const double inf = __longlong_as_double(0x7ff0000000000000LL);
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) time[i] = inf;
}

__global__ void set_start_kernel(double* time, const uint32_t* starts, int n) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) time[starts[i]] = 1.0;  // simulator.cpp:263: starts.push(PriorityQueueEl(1, index))
}

static int run_automaton_bricks(ekg_model* m, int64_t* rounds_out);

int run_automaton(ekg_model* m, int64_t* sweeps_out) {
	cudaStream_t st = m->stream;
	const int64_t npad = m->pZ * m->pY * m->pX;
	// Small models (time field resident in L2) relax fastest with plain sweeps; once the dense field
	// outgrows L2 every sweep streams it from HBM and the brick frontier wins.  EKGSIM_B200_AUTOMATON
	// = sweep | bricks overrides the choice.
	const char* sel = getenv("EKGSIM_B200_AUTOMATON");
	bool sweep = (size_t)npad * 9 <= (size_t)96 << 20;
	if (sel && std::string(sel) == "sweep") sweep = true;
	if (sel && std::string(sel) == "bricks") sweep = false;
	init_time_kernel<<<m->sm_count * 4, 256, 0, st>>>(m->d_time_pad, npad);
	EKG_CUDA(cudaGetLastError());

	std::vector<uint32_t> h_starts(m->h_starts.size());
	for (size_t i = 0; i < h_starts.size(); ++i) {
		int64_t r = m->h_starts[i];
		int64_t z = r / (m->Y * m->X), y = (r / m->X) % m->Y, x = r % m->X;
		h_starts[i] = (uint32_t)(((z + 1) * m->pY + (y + 1)) * m->pX + (x + 1));
	}
	uint32_t* d_starts = nullptr;
	EKG_CUDA(cudaMalloc(&d_starts, h_starts.size() * sizeof(uint32_t)));
	EKG_CUDA(cudaMemcpyAsync(d_starts, h_starts.data(), h_starts.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
	set_start_kernel<<<(int)((h_starts.size() + 127) / 128), 128, 0, st>>>(m->d_time_pad, d_starts, (int)h_starts.size());
	EKG_CUDA(cudaGetLastError());
	if (!sweep) {
		int rc = run_automaton_bricks(m, sweeps_out);
		cudaFree(d_starts);
		return rc;
	}
	EKG_CUDA(cudaMemsetAsync(m->d_flags, 0, (size_t)(m->max_sweeps + 1) * sizeof(int), st));

	AutoArgs a{};
	a.layer = m->d_layer_pad;
	a.time = m->d_time_pad;
	a.pidx = m->d_auto_pidx;
	a.wtab = m->d_wtab;
	a.flags = m->d_flags;
	a.sweeps_out = m->d_flags + m->max_sweeps;
	a.n = m->n_occ;
	a.nl1 = m->n_layers + 1;
	a.max_sweeps = m->max_sweeps;
	NbrTable nb;
	// simulator.cpp:251-254: cube for 3-D shapes, 8-neighbourhood for 2-D ones
	make_nbr_table(m->Z > 1 ? EKG_NBHD_3D8 : EKG_NBHD_2D8, &nb);
	a.n_nbr = nb.n;
	for (int k = 0; k < nb.n; ++k) {
		a.off[k] = (int32_t)((nb.dz[k] * m->pY + nb.dy[k]) * m->pX + nb.dx[k]);
		a.sq[k] = nb.dz[k] * nb.dz[k] + nb.dy[k] * nb.dy[k] + nb.dx[k] * nb.dx[k] - 1;
	}

	int per_sm = 0;
	EKG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, automaton_kernel, 256, 0));
	if (per_sm < 1) return fail(EKG_E_CUDA, "automaton kernel does not fit on the device");
	int64_t want = (m->n_occ + 255) / 256;
	int grid = (int)std::min<int64_t>((int64_t)per_sm * m->sm_count, std::max<int64_t>(want, 1));
	void* kargs[] = {&a};
	EKG_CUDA(cudaLaunchCooperativeKernel((void*)automaton_kernel, dim3(grid), dim3(256), kargs, 0, st));
	int sweeps = 0;
	EKG_CUDA(cudaMemcpyAsync(&sweeps, a.sweeps_out, sizeof(int), cudaMemcpyDeviceToHost, st));
	EKG_CUDA(cudaStreamSynchronize(st));
	EKG_CUDA(cudaFree(d_starts));
	if (sweeps_out) *sweeps_out = sweeps;
	if (sweeps > m->max_sweeps) return fail(EKG_E_STATE, "activation automaton did not converge");
	return EKG_OK;
}

static int run_automaton_bricks(ekg_model* m, int64_t* rounds_out) {
	cudaStream_t st = m->stream;
	const int64_t n = m->n_bricks;
	int* flag = m->d_brick_state;
	int* queue = flag + 2 * n;
	int* counters = queue + 3 * n;
	EKG_CUDA(cudaMemsetAsync(m->d_brick_state, 0, ((size_t)n * 5 + 8) * sizeof(int), st));
	// round 0 processes the bricks of the start voxels
	std::vector<int> q0(m->h_start_bricks.begin(), m->h_start_bricks.end());
	const int n0 = (int)q0.size();
	EKG_CUDA(cudaMemcpyAsync(queue, q0.data(), q0.size() * sizeof(int), cudaMemcpyHostToDevice, st));
	EKG_CUDA(cudaMemcpyAsync(counters, &n0, sizeof(int), cudaMemcpyHostToDevice, st));
	EKG_CUDA(cudaStreamSynchronize(st));  // pageable sources

	BrickArgs a{};
	a.layer = m->d_layer_pad; a.time = m->d_time_pad; a.wtab = m->d_wtab;
	a.origin = m->d_brick_origin; a.nbr = m->d_brick_nbr;
	a.flag = flag; a.queue = queue; a.counters = counters;
	a.n_live = (int32_t)n; a.nl1 = m->n_layers + 1; a.pY = (int32_t)m->pY; a.pX = (int32_t)m->pX;
	a.max_rounds = 1 << 20;
	NbrTable nb;
	make_nbr_table(m->Z > 1 ? EKG_NBHD_3D8 : EKG_NBHD_2D8, &nb);  // simulator.cpp:251-254
	a.n_nbr = nb.n;
	for (int k = 0; k < nb.n; ++k) {
		a.loff[k] = (nb.dz[k] * kBrickHalo + nb.dy[k]) * kBrickHalo + nb.dx[k];
		a.dz[k] = nb.dz[k]; a.dy[k] = nb.dy[k]; a.dx[k] = nb.dx[k];
		a.sq[k] = nb.dz[k] * nb.dz[k] + nb.dy[k] * nb.dy[k] + nb.dx[k] * nb.dx[k] - 1;
	}
	const size_t w_bytes = (size_t)a.nl1 * a.nl1 * 3 * sizeof(double);
	a.w_in_smem = w_bytes <= 32 * 1024;
	const size_t dyn = a.w_in_smem ? w_bytes : 0;
	int per_sm = 0;
	EKG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, automaton_brick_kernel, 256, dyn));
	if (per_sm < 1) return fail(EKG_E_CUDA, "automaton kernel does not fit on the device");
	const int grid = (int)std::min<int64_t>((int64_t)per_sm * m->sm_count, std::max<int64_t>(n, 1));
	void* kargs[] = {&a};
	EKG_CUDA(cudaLaunchCooperativeKernel((void*)automaton_brick_kernel, dim3(grid), dim3(256), kargs, dyn, st));
	int h[8] = {0};
	EKG_CUDA(cudaMemcpyAsync(h, counters, sizeof h, cudaMemcpyDeviceToHost, st));
	EKG_CUDA(cudaStreamSynchronize(st));
	if (rounds_out) *rounds_out = h[3];
	m->last_brick_visits = h[4];
	if (h[3] >= a.max_rounds) return fail(EKG_E_STATE, "activation automaton did not converge");
	return EKG_OK;
}

}  // namespace ekg
