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
// Kernel: persistent cooperative grid; every sweep each thread pulls over its voxels
// (min over up to 26 neighbour candidates, in place), then the grid synchronises and stops at
// the first sweep that changed nothing.  Reads of the time field bypass L1 (ld.global.cg) since
// other SMs update it; the layer map is read-only (ld.global.nc).
//
// Working set on model_24: 1.5 MB padded u8 layers + 12 MB padded f64 times -> L2 resident.

#include <cooperative_groups.h>

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

int run_automaton(ekg_model* m, int64_t* sweeps_out) {
	cudaStream_t st = m->stream;
	const int64_t npad = m->pZ * m->pY * m->pX;
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

}  // namespace ekg
