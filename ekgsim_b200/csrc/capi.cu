// ekgsim_b200/csrc/capi.cu -- the extern "C" entry points of libekgsim_b200.so and the host-side
// preparation of the device-resident model (compaction, neighbour masks, edge-weight table).
// Interface and the reference symbols each entry point replaces: include/ekgsim_b200.h.

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <unordered_map>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include "ekg_internal.cuh"

namespace ekg {

static thread_local std::string g_error;

void set_error(const std::string& msg) { g_error = msg; }
int fail(int code, const std::string& msg) { g_error = msg; return code; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
	char buf[512];
	snprintf(buf, sizeof buf, "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
	g_error = buf;
	cudaGetLastError();  // clear the sticky-less error state
	return EKG_E_CUDA;
}

constexpr uint16_t kStartFlag = 0x1000;  // ShapeElement::layerStartingPoint (matrix.h:90)

__global__ void gather_at_kernel(const double* __restrict__ time_pad, const uint32_t* __restrict__ pidx, const uint32_t* __restrict__ pos,
                                 const uint32_t* __restrict__ mask, double* __restrict__ at, float* __restrict__ at32, float4* __restrict__ vox,
                                 int64_t n) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double v = time_pad[pidx[i]];
	if (isinf(v)) v = 0.0;  // never reached -> excitationDelay stays 0 (simulator.cpp:219)
	at[i] = v;
	at32[i] = (float)v;
	// the moment kernels' record: bordered coordinates (simulator.cpp:458-463 adds a border of one voxel); for a boundary
	// voxel the occupancy of the corners (dz, dy, dx) = (-,-,-), (-,-,+), ... (+,+,+) in the low 8 mantissa bits of y, which
	// are zero for an integer <= 2048 (interior voxels sit in segments of their own and keep a clean y)
	const uint32_t p = pos[i], mk = mask[i];
	const int corner_bit[8] = {0, 2, 6, 8, 17, 19, 23, 25};   // their positions in the 26-neighbour cube list (make_nbr_table)
	uint32_t c8 = 0;
#pragma unroll
	for (int k = 0; k < 8; ++k) c8 |= ((mk >> corner_bit[k]) & 1u) << k;
	const float x = (float)((p & 0x7ffu) + 1u);
	const float y = __uint_as_float(__float_as_uint((float)(((p >> 11) & 0x7ffu) + 1u)) | (c8 == 0xffu ? 0u : c8));
	vox[i] = make_float4((float)((p >> 22) + 1u), y, x, (float)v);
}

// Host -> device upload that is COMPLETE on return.  A plain cudaMemcpy from pageable memory may return
// while the DMA is still in flight, and kernels on our non-blocking stream do not wait for the legacy
// default stream -- so uploads go through the model's stream and are synchronised there.
static int upload(ekg_model* m, void* dst, const void* src, size_t bytes) {
	if (!bytes) return EKG_OK;
	EKG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, m->stream));
	EKG_CUDA(cudaStreamSynchronize(m->stream));
	return EKG_OK;
}

static int64_t pad_index(const ekg_model* m, int64_t z, int64_t y, int64_t x) { return ((z + 1) * m->pY + (y + 1)) * m->pX + (x + 1); }

static void free_model(ekg_model* m) {
	if (!m) return;
	cudaSetDevice(m->device);
	shard_unlink(m);
	void* ptrs[] = {m->d_layer_pad, m->d_time_pad, m->d_auto_pidx, m->d_wtab, m->d_flags, m->d_brick_origin, m->d_brick_nbr, m->d_brick_state, m->d_pos, m->d_mask, m->d_ecg_pidx, m->d_at, m->d_at32,
	                m->d_vox, m->d_segs, m->d_tiles, m->d_params, m->d_tail, m->d_ftab, m->d_times, m->d_partial, m->d_partial2, m->d_io_k, m->d_io_leads, m->d_io_ecg, m->d_io_tgt, m->d_io_border, m->d_fit_conn, m->d_msegs, m->d_mseg_first, m->d_mom, m->d_lmom, m->d_near, m->d_k1min, m->d_brick_index, m->d_brick_own, m->d_brick_mark, m->d_improved, m->d_range, m->d_start_pidx, m->d_start_bricks};
	for (void* p : ptrs) if (p) cudaFree(p);
	if (m->h_pin_in) cudaFreeHost(m->h_pin_in);
	if (m->h_pin_out) cudaFreeHost(m->h_pin_out);
	if (m->ev_k0) cudaEventDestroy(m->ev_k0);
	if (m->ev_k1) cudaEventDestroy(m->ev_k1);
	if (m->stream) cudaStreamDestroy(m->stream);
	delete m;
}

// ---- the ECG voxel list, built on the device ------------------------------------------------------------------------
// The list holds the occupied voxels with z in [z0, z1), sorted by layer; inside a layer first the interior voxels (all 8
// cube corners occupied -- the moment kernels treat them by a series, ecg.cu), then the boundary voxels, raster order in
// both parts.  Everything it needs is resident already: the padded layer map and the raster list of occupied voxels
// (d_auto_pidx; padded indices grow with the raster index, so a z-slab is a contiguous run of it).
struct CubeOffsets { int32_t off[kMaxNbr]; int32_t n; };

// key = 2 (layer - 1) + (boundary ? 1 : 0); mask bit k: voxel `index - dif_k` occupied (simulator.cpp:514)
__global__ void ecg_list_keys_kernel(const uint8_t* __restrict__ layer_pad, const uint32_t* __restrict__ pidx, int64_t n, const CubeOffsets co,
                                     uint16_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ masks) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int64_t p = pidx[i];
	uint32_t mk = 0;
#pragma unroll
	for (int k = 0; k < 26; ++k) if (layer_pad[p - co.off[k]]) mk |= 1u << k;
	keys[i] = (uint16_t)(2u * (layer_pad[p] - 1u) + ((mk & kCornerMask) == kCornerMask ? 0u : 1u));
	vals[i] = (uint32_t)i;
	masks[i] = mk;
}

__global__ void ecg_list_gather_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ pidx, const uint32_t* __restrict__ masks,
                                       int64_t n, int64_t pY, int64_t pX, uint32_t* __restrict__ pos, uint32_t* __restrict__ mask,
                                       uint32_t* __restrict__ ecg_pidx) {
	const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	const uint32_t src = order[j];
	const uint32_t p = pidx[src];
	const uint32_t x = p % (uint32_t)pX - 1u, zy = p / (uint32_t)pX, y = zy % (uint32_t)pY - 1u, z = zy / (uint32_t)pY - 1u;
	pos[j] = x | (y << 11) | (z << 22);
	mask[j] = masks[src];
	ecg_pidx[j] = p;
}

// bounds[k] = first position of the sorted key array whose key is >= k, k = 0 .. n_keys
__global__ void ecg_list_bounds_kernel(const uint16_t* __restrict__ keys, int64_t n, int32_t n_keys, int64_t* __restrict__ bounds) {
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k > n_keys) return;
	int64_t lo = 0, hi = n;
	while (lo < hi) {
		const int64_t mid = (lo + hi) >> 1;
		if ((int)keys[mid] < k) lo = mid + 1; else hi = mid;
	}
	bounds[k] = lo;
}

static int build_ecg_list(ekg_model* m, int64_t z0, int64_t z1) {
	const int nl = m->n_layers;
	const int64_t i0 = m->h_occ_before_z[(size_t)z0], n = m->h_occ_before_z[(size_t)z1] - i0;
	if (n >= ((int64_t)1 << 31)) return fail(EKG_E_UNSUPPORTED, "more than 2^31 occupied voxels in one slab");

	for (void* p : {(void*)m->d_pos, (void*)m->d_mask, (void*)m->d_ecg_pidx, (void*)m->d_at, (void*)m->d_at32, (void*)m->d_vox}) if (p) cudaFree(p);
	m->d_pos = m->d_mask = m->d_ecg_pidx = nullptr; m->d_at = nullptr; m->d_at32 = nullptr; m->d_vox = nullptr;
	const size_t nn = (size_t)std::max<int64_t>(n, 1);
	EKG_CUDA(cudaMalloc(&m->d_pos, nn * 4));
	EKG_CUDA(cudaMalloc(&m->d_mask, nn * 4));
	EKG_CUDA(cudaMalloc(&m->d_ecg_pidx, nn * 4));
	EKG_CUDA(cudaMalloc(&m->d_at, nn * 8));
	EKG_CUDA(cudaMalloc(&m->d_at32, nn * 4));
	EKG_CUDA(cudaMalloc(&m->d_vox, nn * sizeof(float4)));

	m->layer_off.assign(nl + 1, 0);
	m->interior_cnt.assign(nl + 1, 0);
	if (n > 0) {
		NbrTable cube;
		make_nbr_table(EKG_NBHD_3D8, &cube);
		CubeOffsets co{};
		co.n = cube.n;
		for (int k = 0; k < cube.n; ++k) co.off[k] = (int32_t)((cube.dz[k] * m->pY + cube.dy[k]) * m->pX + cube.dx[k]);

		// scratch: keys in / out, order in / out, masks in raster order, bounds, radix-sort workspace
		uint16_t *d_keys = nullptr, *d_keys_sorted = nullptr;
		uint32_t *d_vals = nullptr, *d_order = nullptr, *d_rmask = nullptr;
		int64_t* d_bounds = nullptr;
		void* d_tmp = nullptr;
		auto cleanup = [&]() { for (void* p : {(void*)d_keys, (void*)d_keys_sorted, (void*)d_vals, (void*)d_order, (void*)d_rmask, (void*)d_bounds, d_tmp}) if (p) cudaFree(p); };
#define EKG_LIST_CUDA(call)                                                                        \
		do {                                                                                       \
			cudaError_t e__ = (call);                                                              \
			if (e__ != cudaSuccess) { cleanup(); return cuda_fail(e__, #call, __FILE__, __LINE__); } \
		} while (0)
		const int n_keys = 2 * nl;
		EKG_LIST_CUDA(cudaMalloc(&d_keys, nn * 2));
		EKG_LIST_CUDA(cudaMalloc(&d_keys_sorted, nn * 2));
		EKG_LIST_CUDA(cudaMalloc(&d_vals, nn * 4));
		EKG_LIST_CUDA(cudaMalloc(&d_order, nn * 4));
		EKG_LIST_CUDA(cudaMalloc(&d_rmask, nn * 4));
		EKG_LIST_CUDA(cudaMalloc(&d_bounds, (size_t)(n_keys + 1) * 8));
		int end_bit = 1;
		while ((1 << end_bit) < n_keys) ++end_bit;
		size_t tmp_bytes = 0;
		EKG_LIST_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys_sorted, d_vals, d_order, (int)n, 0, end_bit, m->stream));
		EKG_LIST_CUDA(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 1)));
		const unsigned blocks = (unsigned)((n + 255) / 256);
		ecg_list_keys_kernel<<<blocks, 256, 0, m->stream>>>(m->d_layer_pad, m->d_auto_pidx + i0, n, co, d_keys, d_vals, d_rmask);
		EKG_LIST_CUDA(cudaGetLastError());
		// LSD radix sort: stable, so the raster order survives inside every (layer, kind) range
		EKG_LIST_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys_sorted, d_vals, d_order, (int)n, 0, end_bit, m->stream));
		ecg_list_gather_kernel<<<blocks, 256, 0, m->stream>>>(d_order, m->d_auto_pidx + i0, d_rmask, n, m->pY, m->pX, m->d_pos, m->d_mask, m->d_ecg_pidx);
		EKG_LIST_CUDA(cudaGetLastError());
		ecg_list_bounds_kernel<<<(n_keys + 1 + 127) / 128, 128, 0, m->stream>>>(d_keys_sorted, n, n_keys, d_bounds);
		EKG_LIST_CUDA(cudaGetLastError());
		std::vector<int64_t> bounds((size_t)n_keys + 1);
		EKG_LIST_CUDA(cudaMemcpyAsync(bounds.data(), d_bounds, bounds.size() * 8, cudaMemcpyDeviceToHost, m->stream));
		EKG_LIST_CUDA(cudaStreamSynchronize(m->stream));
#undef EKG_LIST_CUDA
		cleanup();
		for (int l = 0; l < nl; ++l) {
			m->layer_off[l] = bounds[2 * l];
			m->interior_cnt[l] = bounds[2 * l + 1] - bounds[2 * l];
		}
		m->layer_off[nl] = bounds[n_keys];
	}
	m->n_ecg = n;
	m->slab_z0 = z0; m->slab_z1 = z1;
	m->n_segs = 0; m->seg_len = 0;  // segment tables depend on the list
	m->n_msegs = 0; m->mseg_len = 0;
	return EKG_OK;
}

static int gather_at(ekg_model* m) {
	if (m->n_ecg == 0) return EKG_OK;
	gather_at_kernel<<<(int)((m->n_ecg + 255) / 256), 256, 0, m->stream>>>(m->d_time_pad, m->d_ecg_pidx, m->d_pos, m->d_mask, m->d_at, m->d_at32, m->d_vox, m->n_ecg);
	EKG_CUDA(cudaGetLastError());
	EKG_CUDA(cudaStreamSynchronize(m->stream));
	return EKG_OK;
}

// order-preserving map of a double onto an unsigned 64-bit key (and back), for atomicMin / atomicMax
__host__ __device__ inline unsigned long long range_key(double v) {
	long long b;
	memcpy(&b, &v, 8);
	return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ULL));
}
static double range_unkey(unsigned long long k) {
	long long b = (long long)k;
	b = (b < 0) ? (b ^ (long long)0x8000000000000000ULL) : ~b;
	double v;
	memcpy(&v, &b, 8);
	return v;
}

// smallest / largest activation time over the occupied voxels of the WHOLE model (never reached counts as 0, like the
// excitationDelay the reference leaves untouched, simulator.cpp:219): range[0] = min key, range[1] = max key
__global__ void activation_range_kernel(const double* __restrict__ time_pad, const uint32_t* __restrict__ pidx, int64_t n,
                                        unsigned long long* __restrict__ range) {
	unsigned long long lo = ~0ULL, hi = 0ULL;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		double v = time_pad[pidx[i]];
		if (isinf(v)) v = 0.0;
		const unsigned long long k = range_key(v);
		lo = min(lo, k);
		hi = max(hi, k);
	}
	for (int o = 16; o; o >>= 1) {
		lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
		hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
	}
	if ((threadIdx.x & 31) == 0 && lo <= hi) {
		atomicMin(range, lo);
		atomicMax(range + 1, hi);
	}
}

// raster elements [i0, i0 + cnt) of the activation map out of the padded grid: 0 for empty and never-reached voxels
__global__ void unpad_activation_kernel(const double* __restrict__ time_pad, const uint8_t* __restrict__ layer_pad,
                                        double* __restrict__ out, int64_t i0, int64_t cnt, int64_t Y, int64_t X, int64_t pY, int64_t pX) {
	const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= cnt) return;
	const int64_t i = i0 + j;
	const int64_t x = i % X, zy = i / X, y = zy % Y, z = zy / Y;
	const int64_t p = ((z + 1) * pY + (y + 1)) * pX + (x + 1);
	double v = 0.0;
	if (layer_pad[p]) {
		v = time_pad[p];
		if (isinf(v)) v = 0.0;
	}
	out[j] = v;
}

// The raster host copy of the device-resident map (Z*Y*X doubles), made only when a caller asks for it: unpadded on the
// device, brought over in chunks through two pinned buffers so that the copy of one chunk overlaps the memcpy of the
// previous one into the caller's (pageable) array.
static int download_activation(ekg_model* m, double* dst) {
	const int64_t n = m->Z * m->Y * m->X;
	const int64_t chunk = std::min<int64_t>(n, (int64_t)4 << 20);  // 32 MB of doubles
	double* d_buf[2] = {nullptr, nullptr};
	double* h_buf[2] = {nullptr, nullptr};
	cudaEvent_t done[2] = {nullptr, nullptr};
	auto cleanup = [&]() {
		for (int i = 0; i < 2; ++i) {
			if (d_buf[i]) cudaFree(d_buf[i]);
			if (h_buf[i]) cudaFreeHost(h_buf[i]);
			if (done[i]) cudaEventDestroy(done[i]);
		}
	};
#define EKG_DL_CUDA(call)                                                                          \
	do {                                                                                           \
		cudaError_t e__ = (call);                                                                  \
		if (e__ != cudaSuccess) { cleanup(); return cuda_fail(e__, #call, __FILE__, __LINE__); }   \
	} while (0)
	const int n_buf = n > chunk ? 2 : 1;
	for (int i = 0; i < n_buf; ++i) {
		EKG_DL_CUDA(cudaMalloc(&d_buf[i], (size_t)chunk * 8));
		EKG_DL_CUDA(cudaMallocHost(&h_buf[i], (size_t)chunk * 8));
		EKG_DL_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
	}
	const int64_t n_chunks = (n + chunk - 1) / chunk;
	for (int64_t c = 0; c <= n_chunks; ++c) {
		if (c < n_chunks) {
			const int b = (int)(c % n_buf);
			const int64_t i0 = c * chunk, cnt = std::min(chunk, n - i0);
			unpad_activation_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, m->stream>>>(m->d_time_pad, m->d_layer_pad, d_buf[b], i0, cnt, m->Y, m->X, m->pY, m->pX);
			EKG_DL_CUDA(cudaGetLastError());
			EKG_DL_CUDA(cudaMemcpyAsync(h_buf[b], d_buf[b], (size_t)cnt * 8, cudaMemcpyDeviceToHost, m->stream));
			EKG_DL_CUDA(cudaEventRecord(done[b], m->stream));
		}
		if (c > 0) {  // chunk c - 1 has arrived (its buffers are reused by chunk c + 1, enqueued after this memcpy)
			const int b = (int)((c - 1) % n_buf);
			const int64_t i0 = (c - 1) * chunk, cnt = std::min(chunk, n - i0);
			EKG_DL_CUDA(cudaEventSynchronize(done[b]));
			memcpy(dst + i0, h_buf[b], (size_t)cnt * 8);
		}
	}
#undef EKG_DL_CUDA
	cleanup();
	return EKG_OK;
}

// h_delay on demand (ekg_model_get_activation, ekg_model_ap_classes)
static int ensure_host_delay(ekg_model* m) {
	if (m->h_delay_valid) return EKG_OK;
	EKG_CUDA(cudaSetDevice(m->device));
	m->h_delay.resize((size_t)(m->Z * m->Y * m->X));
	int rc = download_activation(m, m->h_delay.data());
	if (rc) return rc;
	m->h_delay_valid = true;
	return EKG_OK;
}

// after d_time_pad holds a map: the range of activation times (t0 of the hoisted forms, the saturation time of the
// SEPARABLE path) by a device reduction, then the ECG-list gather.  The map itself stays on the device.
static int publish_activation(ekg_model* m) {
	double lo = 0.0, hi = 0.0;
	if (m->n_occ > 0) {
		unsigned long long h_range[2] = {~0ULL, 0ULL};
		if (!m->d_range) EKG_CUDA(cudaMalloc(&m->d_range, 2 * sizeof(unsigned long long)));
		EKG_CUDA(cudaMemcpyAsync(m->d_range, h_range, sizeof h_range, cudaMemcpyHostToDevice, m->stream));
		const int blocks = (int)std::min<int64_t>((m->n_occ + 255) / 256, (int64_t)m->sm_count * 8);
		activation_range_kernel<<<blocks, 256, 0, m->stream>>>(m->d_time_pad, m->d_auto_pidx, m->n_occ, m->d_range);
		EKG_CUDA(cudaGetLastError());
		EKG_CUDA(cudaMemcpyAsync(h_range, m->d_range, sizeof h_range, cudaMemcpyDeviceToHost, m->stream));
		EKG_CUDA(cudaStreamSynchronize(m->stream));
		lo = range_unkey(h_range[0]);
		hi = range_unkey(h_range[1]);
	}
	m->t0 = 0.5 * (lo + hi);
	m->at_max = hi;
	m->at_min = lo;
	m->have_activation = true;
	return gather_at(m);
}


// ---- model creation on the device ---------------------------------------------------------------------------------------
// ekg_model_create receives the layer map as the .matrix reader left it (u16 raster, 0x1000 = start flag).  One upload,
// then everything the automaton and the ECG list need is derived by kernels: the zero-bordered u8 layer map, the
// occupied voxels per z-plane, the start voxels, the raster list of occupied voxels (stream compaction), the live 4^3
// bricks with their origins and 26 neighbours.  (Round 1 made three host passes over the dense grid: 0.46 s at 4x.)
struct CreateStats {
	int32_t max_layer;      // highest layer number
	int32_t bad_layer;      // a layer number > 255
	int32_t bad_start;      // a start voxel without a layer
	int32_t n_starts;       // start voxels seen (the first kMaxStarts are recorded)
};
constexpr int kMaxStarts = 4096;

__global__ void __launch_bounds__(256) create_pad_kernel(const uint16_t* __restrict__ raw, int64_t n, int64_t Y, int64_t X, int64_t pY, int64_t pX,
                                                         uint8_t* __restrict__ layer_pad, unsigned long long* __restrict__ occ_z,
                                                         CreateStats* __restrict__ stats, long long* __restrict__ starts) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	int l = 0;
	int64_t z = -1;
	if (i < n) {
		l = raw[i];
		const int64_t x = i % X, zy = i / X, y = zy % Y;
		z = zy / Y;
		if (l & kStartFlag) {   // simulator.cpp:261-264
			l -= kStartFlag;
			if (l == 0) atomicExch(&stats->bad_start, 1);
			const int slot = atomicAdd(&stats->n_starts, 1);
			if (slot < kMaxStarts) starts[slot] = i;
		}
		if (l > 255) { atomicExch(&stats->bad_layer, 1); l = 0; }
		if (l) layer_pad[((z + 1) * pY + (y + 1)) * pX + (x + 1)] = (uint8_t)l;
	}
	// occupied voxels per plane: a warp usually sits inside one plane
	const unsigned occ = __ballot_sync(0xffffffffu, l != 0);
	const int64_t z0 = __shfl_sync(0xffffffffu, z, 0);
	if (__all_sync(0xffffffffu, z == z0)) {
		if ((threadIdx.x & 31) == 0 && occ && z0 >= 0) atomicAdd(occ_z + z0, (unsigned long long)__popc(occ));
	} else if (l) atomicAdd(occ_z + z, 1ull);
	const int mx = __reduce_max_sync(0xffffffffu, l);
	if ((threadIdx.x & 31) == 0 && mx) atomicMax(&stats->max_layer, mx);
}

// padded indices [p0, p0 + cnt) of occupied voxels, ascending (= raster order), appended at out + *n_out
struct OccupiedAt {
	const uint8_t* layer_pad;
	__device__ bool operator()(uint32_t p) const { return layer_pad[p] != 0; }
};

// live[c] = 1 if brick cell c (dense brick grid [bZ][bY][bX]) holds an occupied voxel
__global__ void brick_live_kernel(const uint8_t* __restrict__ layer_pad, int64_t n_cells, int bY, int bX, int64_t pY, int64_t pX, int32_t* __restrict__ live) {
	const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n_cells) return;
	const int64_t bx = c % bX, t = c / bX, by = t % bY, bz = t / bY;
	const uint8_t* base = layer_pad + ((bz * kBrick + 1) * pY + (by * kBrick + 1)) * pX + (bx * kBrick + 1);   // the padded extents cover whole bricks
	int any = 0;
	for (int z = 0; z < kBrick; ++z) for (int y = 0; y < kBrick; ++y) {
		const uint8_t* r = base + (z * pY + y) * pX;
		any |= r[0] | r[1] | r[2] | r[3];
	}
	live[c] = any ? 1 : 0;
}

// index[c] = live brick id (exclusive scan of live) or -1; origin of every live brick
__global__ void brick_index_kernel(const int32_t* __restrict__ live, const int32_t* __restrict__ scan, int64_t n_cells, int bY, int bX, int64_t pY, int64_t pX,
                                   int32_t* __restrict__ index, uint32_t* __restrict__ origin) {
	const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n_cells) return;
	if (!live[c]) { index[c] = -1; return; }
	const int64_t bx = c % bX, t = c / bX, by = t % bY, bz = t / bY;
	index[c] = scan[c];
	origin[scan[c]] = (uint32_t)(((bz * kBrick + 1) * pY + (by * kBrick + 1)) * pX + (bx * kBrick + 1));
}

__global__ void brick_nbr_kernel(const int32_t* __restrict__ index, int64_t n_cells, int bZ, int bY, int bX, int32_t* __restrict__ nbr) {
	const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n_cells) return;
	const int32_t bi = index[c];
	if (bi < 0) return;
	const int bx = (int)(c % bX), by = (int)((c / bX) % bY), bz = (int)(c / ((int64_t)bX * bY));
	int k = 0;
	for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
		if (!dz && !dy && !dx) continue;
		const int cz = bz + dz, cy = by + dy, cx = bx + dx;
		int32_t v = -1;
		if (cz >= 0 && cz < bZ && cy >= 0 && cy < bY && cx >= 0 && cx < bX) v = index[((int64_t)cz * bY + cy) * bX + cx];
		nbr[(size_t)bi * 26 + k++] = v;
	}
}

// raster activation map -> padded grid (ekg_model_set_activation)
__global__ void pad_activation_kernel(const double* __restrict__ raster, int64_t n, int64_t Y, int64_t X, int64_t pY, int64_t pX, double* __restrict__ time_pad) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int64_t x = i % X, zy = i / X, y = zy % Y, z = zy / Y;
	time_pad[((z + 1) * pY + (y + 1)) * pX + (x + 1)] = raster[i];
}

// raster u8 layer map (start flags stripped) out of the padded one: the lazily made host copy behind ekg_model_ap_classes
__global__ void unpad_layer_kernel(const uint8_t* __restrict__ layer_pad, int64_t n, int64_t Y, int64_t X, int64_t pY, int64_t pX, uint8_t* __restrict__ out) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int64_t x = i % X, zy = i / X, y = zy % Y, z = zy / Y;
	out[i] = layer_pad[((z + 1) * pY + (y + 1)) * pX + (x + 1)];
}

static int ensure_host_layer(ekg_model* m) {
	const int64_t n = m->Z * m->Y * m->X;
	if ((int64_t)m->h_layer.size() == n) return EKG_OK;
	EKG_CUDA(cudaSetDevice(m->device));
	uint8_t* d = nullptr;
	EKG_CUDA(cudaMalloc(&d, (size_t)n));
	unpad_layer_kernel<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(m->d_layer_pad, n, m->Y, m->X, m->pY, m->pX, d);
	m->h_layer.resize((size_t)n);
	cudaError_t e = cudaMemcpyAsync(m->h_layer.data(), d, (size_t)n, cudaMemcpyDeviceToHost, m->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
	cudaFree(d);
	if (e != cudaSuccess) { m->h_layer.clear(); return cuda_fail(e, "download of the layer map", __FILE__, __LINE__); }
	return EKG_OK;
}

}  // namespace ekg

using namespace ekg;

extern "C" {

int ekg_abi_version(void) { return EKG_ABI_VERSION; }
const char* ekg_last_error(void) { return g_error.c_str(); }

int ekg_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int ekg_model_create(const uint16_t* layers, int64_t Z, int64_t Y, int64_t X,
                     const double* transfer, int64_t t_rows, int64_t t_cols, int device, ekg_model** out) {
	if (!out) return fail(EKG_E_INVALID, "out is NULL");
	*out = nullptr;
	if (!layers || Z <= 0 || Y <= 0 || X <= 0) return fail(EKG_E_INVALID, "bad shape");
	if (!transfer || t_rows <= 0 || t_cols <= 0) return fail(EKG_E_INVALID, "bad transfer matrix");
	if (X > 2047 || Y > 2047 || Z > 1023) return fail(EKG_E_UNSUPPORTED, "grid exceeds the packed coordinate layout (X,Y <= 2047, Z <= 1023)");
	if ((Z + 10) * (Y + 10) * (X + 10) >= (int64_t)1 << 32) return fail(EKG_E_UNSUPPORTED, "grid exceeds 2^32 padded voxels");
	if (ekg_device_count() <= device || device < 0) return fail(EKG_E_CUDA, "no such CUDA device (libekgsim_b200 has no CPU fallback)");

	ekg_model* m = new ekg_model();
	m->device = device;
	m->Z = Z; m->Y = Y; m->X = X;
	// zero border of one voxel, extents rounded up to a multiple of 8 (whole bricks for the frontier automaton)
	m->pZ = (Z + 7) / 8 * 8 + 2; m->pY = (Y + 7) / 8 * 8 + 2; m->pX = (X + 7) / 8 * 8 + 2;
	const int64_t n = Z * Y * X;
	m->h_transfer.assign(transfer, transfer + t_rows * t_cols);
	m->t_rows = t_rows; m->t_cols = t_cols;

#define EKG_CREATE_CUDA(call)                                                                   \
	do {                                                                                        \
		cudaError_t e__ = (call);                                                               \
		if (e__ != cudaSuccess) { int rc__ = cuda_fail(e__, #call, __FILE__, __LINE__); cleanup(); free_model(m); return rc__; } \
	} while (0)

	// scratch of this function
	uint16_t* d_raw = nullptr;
	unsigned long long* d_occ_z = nullptr;
	CreateStats* d_stats = nullptr;
	long long* d_starts = nullptr;
	int32_t *d_live = nullptr, *d_scan = nullptr;
	uint32_t* d_nsel = nullptr;
	void* d_tmp = nullptr;
	auto cleanup = [&]() { for (void* p : {(void*)d_raw, (void*)d_occ_z, (void*)d_stats, (void*)d_starts, (void*)d_live, (void*)d_scan, (void*)d_nsel, d_tmp}) if (p) cudaFree(p); };

	EKG_CREATE_CUDA(cudaSetDevice(device));
	EKG_CREATE_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
	EKG_CREATE_CUDA(cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device));
	cudaStream_t st = m->stream;

	// (1) upload, padded u8 layer map, per-plane counts, start voxels
	const int64_t npad = m->pZ * m->pY * m->pX;
	EKG_CREATE_CUDA(cudaMalloc(&d_raw, (size_t)n * 2));
	EKG_CREATE_CUDA(cudaMalloc(&d_occ_z, (size_t)Z * 8));
	EKG_CREATE_CUDA(cudaMalloc(&d_stats, sizeof(CreateStats)));
	EKG_CREATE_CUDA(cudaMalloc(&d_starts, kMaxStarts * 8));
	EKG_CREATE_CUDA(cudaMalloc(&m->d_layer_pad, (size_t)npad));
	EKG_CREATE_CUDA(cudaMalloc(&m->d_time_pad, (size_t)npad * 8));
	EKG_CREATE_CUDA(cudaMalloc(&m->d_flags, (size_t)(m->max_sweeps + 1) * sizeof(int)));
	EKG_CREATE_CUDA(cudaMemcpyAsync(d_raw, layers, (size_t)n * 2, cudaMemcpyHostToDevice, st));
	EKG_CREATE_CUDA(cudaMemsetAsync(m->d_layer_pad, 0, (size_t)npad, st));
	EKG_CREATE_CUDA(cudaMemsetAsync(m->d_time_pad, 0, (size_t)npad * 8, st));
	EKG_CREATE_CUDA(cudaMemsetAsync(d_occ_z, 0, (size_t)Z * 8, st));
	EKG_CREATE_CUDA(cudaMemsetAsync(d_stats, 0, sizeof(CreateStats), st));
	create_pad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_raw, n, Y, X, m->pY, m->pX, m->d_layer_pad, d_occ_z, d_stats, d_starts);
	EKG_CREATE_CUDA(cudaGetLastError());
	CreateStats stats;
	std::vector<unsigned long long> occ_z((size_t)Z);
	EKG_CREATE_CUDA(cudaMemcpyAsync(&stats, d_stats, sizeof stats, cudaMemcpyDeviceToHost, st));
	EKG_CREATE_CUDA(cudaMemcpyAsync(occ_z.data(), d_occ_z, (size_t)Z * 8, cudaMemcpyDeviceToHost, st));
	EKG_CREATE_CUDA(cudaStreamSynchronize(st));   // also: the caller's (pageable) layer array has been consumed
	if (stats.bad_layer) { cleanup(); free_model(m); return fail(EKG_E_UNSUPPORTED, "more than 255 layers"); }
	// the text format cannot express "start voxel of layer 0" (matrix.h:124-129 maps -v to 0x1000 + v, v >= 1); through the
	// raw ABI it would be a start voxel outside the model (no brick, no layer to conduct from)
	if (stats.bad_start) { cleanup(); free_model(m); return fail(EKG_E_INVALID, "start voxel without a layer (value 0x1000)"); }
	const int max_layer = stats.max_layer;
	m->n_layers = max_layer;  // targetNumOfAps = highest layer number (simulator.cpp:186-198)
	if (t_rows < max_layer || t_cols < max_layer) { cleanup(); free_model(m); return fail(EKG_E_TRANSFER, "loaded transfer matrix too small"); }  // simulator.cpp:203-205
	m->h_occ_before_z.assign((size_t)Z + 1, 0);
	for (int64_t z = 0; z < Z; ++z) m->h_occ_before_z[(size_t)z + 1] = m->h_occ_before_z[(size_t)z] + (int64_t)occ_z[(size_t)z];
	m->n_occ = m->h_occ_before_z[(size_t)Z];
	if (stats.n_starts <= kMaxStarts) {
		std::vector<long long> hs((size_t)stats.n_starts);
		if (!hs.empty()) EKG_CREATE_CUDA(cudaMemcpy(hs.data(), d_starts, hs.size() * 8, cudaMemcpyDeviceToHost));
		std::sort(hs.begin(), hs.end());
		m->h_starts.assign(hs.begin(), hs.end());
	} else {   // a model made of start voxels: read them off the caller's array
		for (int64_t i = 0; i < n; ++i) if (layers[i] & kStartFlag) m->h_starts.push_back(i);
	}

	// (2) raster list of the occupied voxels (padded indices): stream compaction, in chunks CUB's 32-bit item counts can take
	EKG_CREATE_CUDA(cudaMalloc(&m->d_auto_pidx, (size_t)std::max<int64_t>(m->n_occ, 1) * 4));
	EKG_CREATE_CUDA(cudaMalloc(&d_nsel, 4));
	{
		const int64_t chunk = (int64_t)1 << 30;
		size_t tmp_bytes = 0;
		OccupiedAt pred{m->d_layer_pad};
		EKG_CREATE_CUDA(cub::DeviceSelect::If(nullptr, tmp_bytes, cub::CountingInputIterator<uint32_t>(0u), m->d_auto_pidx, d_nsel,
		                                      (int)std::min<int64_t>(npad, chunk), pred, st));
		EKG_CREATE_CUDA(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 1)));
		int64_t done = 0;
		for (int64_t p0 = 0; p0 < npad; p0 += chunk) {
			const int cnt = (int)std::min<int64_t>(chunk, npad - p0);
			EKG_CREATE_CUDA(cub::DeviceSelect::If(d_tmp, tmp_bytes, cub::CountingInputIterator<uint32_t>((uint32_t)p0), m->d_auto_pidx + done, d_nsel, cnt, pred, st));
			if (p0 + chunk < npad) {
				uint32_t got = 0;
				EKG_CREATE_CUDA(cudaMemcpyAsync(&got, d_nsel, 4, cudaMemcpyDeviceToHost, st));
				EKG_CREATE_CUDA(cudaStreamSynchronize(st));
				done += got;
			}
		}
		cudaFree(d_tmp); d_tmp = nullptr;
	}
	// edge weights: lag = T[layer][neighbour layer] * sqrt(sqrLength(dif)) (simulator.cpp:239-240), host IEEE arithmetic
	{
		const int nl1 = max_layer + 1;
		std::vector<double> w((size_t)nl1 * nl1 * 3, INFINITY);
		for (int lu = 1; lu < nl1; ++lu) for (int lv = 1; lv < nl1; ++lv) {
			if (lu >= t_rows || lv >= t_cols) continue;
			for (int sq = 1; sq <= 3; ++sq) {
				double lag = transfer[(int64_t)lu * t_cols + lv];
				lag *= std::sqrt((double)sq);
				w[((size_t)lu * nl1 + lv) * 3 + (sq - 1)] = lag;
			}
		}
		EKG_CREATE_CUDA(cudaMalloc(&m->d_wtab, w.size() * 8));
		if (upload(m, m->d_wtab, w.data(), w.size() * 8)) { cleanup(); free_model(m); return EKG_E_CUDA; }
		// time-bucket width of the automaton's work queue: what the wave needs to cross one brick inside the fastest layer
		// (any positive value gives the same activation times, this one the fewest brick visits; EKGSIM_B200_AUTOMATON_DELTA
		// overrides, 0 = plain FIFO)
		double fastest = INFINITY;
		for (int l = 1; l < nl1; ++l) { const double v = w[((size_t)l * nl1 + l) * 3]; if (v > 0 && v < fastest) fastest = v; }
		if (!std::isfinite(fastest)) for (double v : w) if (v > 0 && v < fastest) fastest = v;
		m->brick_delta = std::isfinite(fastest) ? (float)(fastest * kBrick) : 0.f;
		if (const char* e = getenv("EKGSIM_B200_AUTOMATON_DELTA")) m->brick_delta = (float)std::max(0.0, atof(e));
	}
	// (3) live bricks (kBrick^3 tiles holding at least one occupied voxel) in brick-raster order, their origins and 26 neighbours
	{
		const int64_t kb = kBrick;
		const int64_t bZ = (Z + kb - 1) / kb, bY = (Y + kb - 1) / kb, bX = (X + kb - 1) / kb, n_cells = bZ * bY * bX;
		m->bZ = bZ; m->bY = bY; m->bX = bX;
		EKG_CREATE_CUDA(cudaMalloc(&d_live, (size_t)n_cells * 4));
		EKG_CREATE_CUDA(cudaMalloc(&d_scan, (size_t)n_cells * 4));
		EKG_CREATE_CUDA(cudaMalloc(&m->d_brick_index, (size_t)n_cells * 4));
		const unsigned cb = (unsigned)((n_cells + 255) / 256);
		brick_live_kernel<<<cb, 256, 0, st>>>(m->d_layer_pad, n_cells, (int)bY, (int)bX, m->pY, m->pX, d_live);
		EKG_CREATE_CUDA(cudaGetLastError());
		size_t tmp_bytes = 0;
		EKG_CREATE_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_live, d_scan, (int)n_cells, st));
		EKG_CREATE_CUDA(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 1)));
		EKG_CREATE_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_live, d_scan, (int)n_cells, st));
		int32_t last[2] = {0, 0};
		EKG_CREATE_CUDA(cudaMemcpyAsync(&last[0], d_live + (n_cells - 1), 4, cudaMemcpyDeviceToHost, st));
		EKG_CREATE_CUDA(cudaMemcpyAsync(&last[1], d_scan + (n_cells - 1), 4, cudaMemcpyDeviceToHost, st));
		EKG_CREATE_CUDA(cudaStreamSynchronize(st));
		const int64_t nb = (int64_t)last[0] + last[1];
		m->n_bricks = nb;
		EKG_CREATE_CUDA(cudaMalloc(&m->d_brick_origin, (size_t)std::max<int64_t>(nb, 1) * 4));
		EKG_CREATE_CUDA(cudaMalloc(&m->d_brick_nbr, (size_t)std::max<int64_t>(nb, 1) * 26 * 4));
		// flag[n] | first_visit[n] | rings | counters   (automaton.cu, run_automaton_bricks)
		EKG_CREATE_CUDA(cudaMalloc(&m->d_brick_state, (size_t)brick_state_ints(nb) * sizeof(int)));
		brick_index_kernel<<<cb, 256, 0, st>>>(d_live, d_scan, n_cells, (int)bY, (int)bX, m->pY, m->pX, m->d_brick_index, m->d_brick_origin);
		EKG_CREATE_CUDA(cudaGetLastError());
		brick_nbr_kernel<<<cb, 256, 0, st>>>(m->d_brick_index, n_cells, (int)bZ, (int)bY, (int)bX, m->d_brick_nbr);
		EKG_CREATE_CUDA(cudaGetLastError());
		for (int64_t r : m->h_starts) {
			const int64_t z = r / (Y * X), y = (r / X) % Y, x = r % X;
			int32_t bi = -1;
			EKG_CREATE_CUDA(cudaMemcpyAsync(&bi, m->d_brick_index + (((z / kBrick) * bY + y / kBrick) * bX + x / kBrick), 4, cudaMemcpyDeviceToHost, st));
			EKG_CREATE_CUDA(cudaStreamSynchronize(st));
			if (bi < 0) { cleanup(); free_model(m); return fail(EKG_E_INVALID, "start voxel outside every occupied brick"); }   // cannot happen: a start voxel is occupied
			if (std::find(m->h_start_bricks.begin(), m->h_start_bricks.end(), bi) == m->h_start_bricks.end()) {
				m->h_start_bricks.push_back(bi);
				m->h_start_brick_bz.push_back(z / kBrick);
			}
		}
		// the start voxels (padded indices) and their bricks stay on the device: every automaton run begins with them
		{
			std::vector<uint32_t> sp(m->h_starts.size());
			for (size_t i = 0; i < sp.size(); ++i) {
				const int64_t r = m->h_starts[i];
				sp[i] = (uint32_t)pad_index(m, r / (Y * X), (r / X) % Y, r % X);
			}
			EKG_CREATE_CUDA(cudaMalloc(&m->d_start_pidx, std::max<size_t>(sp.size(), 1) * sizeof(uint32_t)));
			EKG_CREATE_CUDA(cudaMalloc(&m->d_start_bricks, std::max<size_t>(m->h_start_bricks.size(), 1) * sizeof(int32_t)));
			EKG_CREATE_CUDA(cudaMemcpyAsync(m->d_start_pidx, sp.data(), sp.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
			EKG_CREATE_CUDA(cudaMemcpyAsync(m->d_start_bricks, m->h_start_bricks.data(), m->h_start_bricks.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
			EKG_CREATE_CUDA(cudaStreamSynchronize(st));
		}
	}
	cleanup();
#undef EKG_CREATE_CUDA
	int rc = build_ecg_list(m, 0, Z);
	if (rc) { free_model(m); return rc; }
	*out = m;
	return EKG_OK;
}

void ekg_model_destroy(ekg_model* m) { free_model(m); }

int ekg_model_set_slab(ekg_model* m, int64_t z_begin, int64_t z_end) {
	if (!m) return fail(EKG_E_INVALID, "model is NULL");
	if (z_begin < 0 || z_end > m->Z || z_begin > z_end) return fail(EKG_E_INVALID, "bad z range");
	EKG_CUDA(cudaSetDevice(m->device));
	int rc = build_ecg_list(m, z_begin, z_end);
	if (rc) return rc;
	if (m->have_activation) return gather_at(m);
	return EKG_OK;
}

int64_t ekg_model_num_voxels(const ekg_model* m) { return m ? m->n_ecg : 0; }
int64_t ekg_model_num_layers(const ekg_model* m) { return m ? m->n_layers : 0; }

int ekg_model_activation(ekg_model* m, double* delay_out, int64_t* sweeps_out) {
	if (!m) return fail(EKG_E_INVALID, "model is NULL");
	if (m->h_starts.empty()) return fail(EKG_E_NO_START, "Could not find starting point for excitation sequence");  // simulator.cpp:274-277
	if (m->n_layers >= m->t_cols || m->n_layers >= m->t_rows) {
		char buf[128];
		snprintf(buf, sizeof buf, "transfer (conduction) matrix does not define layer %d", m->n_layers);  // simulator.cpp:234-238
		return fail(EKG_E_TRANSFER, buf);
	}
	EKG_CUDA(cudaSetDevice(m->device));
	cudaEvent_t e0, e1;
	EKG_CUDA(cudaEventCreate(&e0));
	EKG_CUDA(cudaEventCreate(&e1));
	EKG_CUDA(cudaEventRecord(e0, m->stream));
	int rc = run_automaton(m, sweeps_out);
	if (rc == EKG_OK) {
		cudaEventRecord(e1, m->stream);
		cudaEventSynchronize(e1);
		cudaEventElapsedTime(&m->activation_ms, e0, e1);
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	if (rc) return rc;
	m->h_delay_valid = false;
	std::vector<double>().swap(m->h_delay);
	rc = publish_activation(m);
	if (rc) return rc;
	if (delay_out) return download_activation(m, delay_out);
	return EKG_OK;
}

// ---- z-slab sharded automaton ---------------------------------------------------------------------------------
int ekg_model_activation_begin(ekg_model* m) {
	if (!m) return fail(EKG_E_INVALID, "model is NULL");
	EKG_CUDA(cudaSetDevice(m->device));
	return shard_begin(m);
}

int ekg_model_activation_relax(ekg_model* m, int64_t* brick_visits_out) {
	if (!m) return fail(EKG_E_INVALID, "model is NULL");
	EKG_CUDA(cudaSetDevice(m->device));
	return shard_relax(m, 0, brick_visits_out, nullptr);
}

int ekg_model_activation_relax_bounded(ekg_model* m, int64_t max_brick_visits, int64_t* brick_visits_out, int64_t* bricks_left_out) {
	if (!m) return fail(EKG_E_INVALID, "model is NULL");
	if (max_brick_visits < 0) return fail(EKG_E_INVALID, "negative visit bound");
	EKG_CUDA(cudaSetDevice(m->device));
	return shard_relax(m, max_brick_visits, brick_visits_out, bricks_left_out);
}

int64_t ekg_model_plane_elems(const ekg_model* m) { return m ? m->pY * m->pX : 0; }

int ekg_model_activation_export(ekg_model* m, int64_t z_begin, int64_t z_end, double* d_planes, void* stream) {
	if (!m || !d_planes) return fail(EKG_E_INVALID, "NULL argument");
	EKG_CUDA(cudaSetDevice(m->device));
	return shard_export(m, z_begin, z_end, d_planes, (cudaStream_t)stream);
}

int ekg_model_activation_merge(ekg_model* m, int64_t z_begin, int64_t z_end, const double* d_planes, int64_t* improved_out, void* stream) {
	if (!m || !d_planes) return fail(EKG_E_INVALID, "NULL argument");
	EKG_CUDA(cudaSetDevice(m->device));
	return shard_merge(m, z_begin, z_end, d_planes, improved_out, nullptr, (cudaStream_t)stream);
}

int ekg_model_activation_merge_async(ekg_model* m, int64_t z_begin, int64_t z_end, const double* d_planes, uint64_t* d_improved_accum, void* stream) {
	if (!m || !d_planes || !d_improved_accum) return fail(EKG_E_INVALID, "NULL argument");
	EKG_CUDA(cudaSetDevice(m->device));
	m->merge_stream = (cudaStream_t)stream;
	m->merge_pending = true;
	return shard_merge(m, z_begin, z_end, d_planes, nullptr, reinterpret_cast<unsigned long long*>(d_improved_accum), (cudaStream_t)stream);
}

int ekg_model_activation_end(ekg_model* m, double* delay_out) {
	if (!m) return fail(EKG_E_INVALID, "model is NULL");
	if (!m->shard_active) return fail(EKG_E_STATE, "ekg_model_activation_begin has not been called");
	EKG_CUDA(cudaSetDevice(m->device));
	if (m->merge_pending) { EKG_CUDA(cudaStreamSynchronize(m->merge_stream)); m->merge_pending = false; }
	m->shard_active = false;
	m->h_delay_valid = false;
	std::vector<double>().swap(m->h_delay);
	int rc = publish_activation(m);
	if (rc) return rc;
	if (delay_out) return download_activation(m, delay_out);
	return EKG_OK;
}

// ---- peer-linked sharded automaton (automaton.cu) --------------------------------------------------------------
int ekg_model_activation_link_info(ekg_model* m, void* info_out) {
	if (!m || !info_out) return fail(EKG_E_INVALID, "NULL argument");
	EKG_CUDA(cudaSetDevice(m->device));
	return shard_link_info(m, info_out);
}

int ekg_model_activation_link(ekg_model* m, int rank, int n_ranks, const void* infos, const int64_t* slabs) {
	if (!m || !infos || !slabs) return fail(EKG_E_INVALID, "NULL argument");
	EKG_CUDA(cudaSetDevice(m->device));
	return shard_link(m, rank, n_ranks, infos, slabs);
}

int ekg_model_activation_linked_launch(ekg_model* m, int max_ctas) {
	if (!m) return fail(EKG_E_INVALID, "model is NULL");
	EKG_CUDA(cudaSetDevice(m->device));
	return shard_linked_launch(m, max_ctas);
}

int ekg_model_activation_linked_wait(ekg_model* m, int64_t* brick_visits_out, int64_t* remote_out) {
	if (!m) return fail(EKG_E_INVALID, "model is NULL");
	EKG_CUDA(cudaSetDevice(m->device));
	return shard_linked_wait(m, brick_visits_out, remote_out);
}

int ekg_model_activation_linked_gather(ekg_model* m) {
	if (!m) return fail(EKG_E_INVALID, "model is NULL");
	EKG_CUDA(cudaSetDevice(m->device));
	return shard_linked_gather(m);
}

int ekg_model_activation_unlink(ekg_model* m) {
	if (!m) return fail(EKG_E_INVALID, "model is NULL");
	EKG_CUDA(cudaSetDevice(m->device));
	return shard_unlink(m);
}

int ekg_model_set_activation(ekg_model* m, const double* delay) {
	if (!m || !delay) return fail(EKG_E_INVALID, "NULL argument");
	EKG_CUDA(cudaSetDevice(m->device));
	const int64_t n = m->Z * m->Y * m->X;
	m->h_delay.assign(delay, delay + n);  // the caller's values as given (also those of empty voxels)
	m->h_delay_valid = true;
	double* d_raster = nullptr;
	EKG_CUDA(cudaMalloc(&d_raster, (size_t)n * 8));
	cudaError_t e = cudaMemcpyAsync(d_raster, delay, (size_t)n * 8, cudaMemcpyHostToDevice, m->stream);
	if (e == cudaSuccess) e = cudaMemsetAsync(m->d_time_pad, 0, (size_t)(m->pZ * m->pY * m->pX) * 8, m->stream);
	if (e == cudaSuccess) {
		pad_activation_kernel<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(d_raster, n, m->Y, m->X, m->pY, m->pX, m->d_time_pad);
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
	cudaFree(d_raster);
	if (e != cudaSuccess) return cuda_fail(e, "upload of the activation map", __FILE__, __LINE__);
	return publish_activation(m);
}

int ekg_model_get_activation(const ekg_model* m, double* delay_out) {
	if (!m || !delay_out) return fail(EKG_E_INVALID, "NULL argument");
	if (!m->have_activation) return fail(EKG_E_STATE, "no excitation sequence yet");
	if (m->h_delay_valid) { memcpy(delay_out, m->h_delay.data(), m->h_delay.size() * 8); return EKG_OK; }
	ekg_model* mm = const_cast<ekg_model*>(m);  // the handle's scratch, not its state
	EKG_CUDA(cudaSetDevice(mm->device));
	return download_activation(mm, delay_out);
}

int64_t ekg_model_activation_brick_visits(const ekg_model* m) { return m ? m->last_brick_visits : 0; }
double ekg_model_activation_ms(const ekg_model* m) { return m ? (double)m->activation_ms : 0.0; }

int ekg_model_ap_classes(const ekg_model* m, int64_t* ap_index_out, int64_t* n_classes_out) {
	if (!m || !ap_index_out || !n_classes_out) return fail(EKG_E_INVALID, "NULL argument");
	if (!m->have_activation) return fail(EKG_E_STATE, "no excitation sequence yet");
	if (int rc = ensure_host_delay(const_cast<ekg_model*>(m))) return rc;  // lazily cached host copies
	if (int rc = ensure_host_layer(const_cast<ekg_model*>(m))) return rc;
	// one (layer, exact delay) -> index map, indices handed out in first-seen raster order
	// (simulator.cpp:566-590 keeps one std::map<double,size_t> per layer with a shared counter)
	struct Key { uint64_t bits; uint32_t layer; bool operator==(const Key& o) const { return bits == o.bits && layer == o.layer; } };
	struct Hash { size_t operator()(const Key& k) const { uint64_t x = k.bits ^ ((uint64_t)k.layer << 52); x ^= x >> 31; x *= 0x9e3779b97f4a7c15ULL; x ^= x >> 29; return (size_t)x; } };
	std::unordered_map<Key, int64_t, Hash> map;
	map.reserve((size_t)m->n_occ / 4 + 16);
	const int64_t n = m->Z * m->Y * m->X;
	int64_t next = 0;
	for (int64_t i = 0; i < n; ++i) {
		if (!m->h_layer[(size_t)i]) { ap_index_out[i] = -1; continue; }
		Key k; memcpy(&k.bits, &m->h_delay[(size_t)i], 8); k.layer = m->h_layer[(size_t)i];
		auto it = map.find(k);
		if (it == map.end()) { map.emplace(k, next); ap_index_out[i] = next++; }
		else ap_index_out[i] = it->second;
	}
	*n_classes_out = next;
	return EKG_OK;
}

int ekg_simulate_device(ekg_model* m, const double* d_layer_k, const double* d_leads_zyx, int64_t B, int64_t n_leads, int nbhd,
                        double t_start, double t_step, double total_time, int flags, double* d_ecg_out, void* stream) {
	if (!m || !d_layer_k || !d_leads_zyx || !d_ecg_out) return fail(EKG_E_INVALID, "NULL argument");
	EKG_CUDA(cudaSetDevice(m->device));
	// the caller's stream as given: NULL is CUDA's default stream (what torch uses unless told otherwise)
	return run_ecg(m, d_layer_k, d_leads_zyx, B, n_leads, nbhd, t_start, t_step, total_time, flags, d_ecg_out, (cudaStream_t)stream);
}

int ekg_simulate_device_hinted(ekg_model* m, const double* d_layer_k, const double* d_leads_zyx, int64_t B, int64_t n_leads, int nbhd,
                               double t_start, double t_step, double total_time, int flags, double k1_min, double decay_max, double* d_ecg_out,
                               void* stream) {
	if (!m || !d_layer_k || !d_leads_zyx || !d_ecg_out) return fail(EKG_E_INVALID, "NULL argument");
	if (!(k1_min > 0) || !std::isfinite(k1_min) || !(decay_max >= 0) || !std::isfinite(decay_max)) return fail(EKG_E_INVALID, "bad coefficient hints");
	EKG_CUDA(cudaSetDevice(m->device));
	KHints hints;
	hints.k1_min = k1_min;
	hints.decay_max = decay_max;
	return run_ecg(m, d_layer_k, d_leads_zyx, B, n_leads, nbhd, t_start, t_step, total_time, flags, d_ecg_out, (cudaStream_t)stream, hints);
}

// border APs + descent settings of the device-side layer fit (ekg_evaluate); border_k == NULL: layer_k is given
struct FitRequest {
	const double* border_k = nullptr;
	int64_t n_border = 0, mid = 0, iterations = 0;
	const double* d9 = nullptr;
	double step = 0, eps = 0;
	double* layer_k_out = nullptr;
};

static int simulate_host(ekg_model* m, const double* layer_k, const double* leads_zyx, int64_t B, int64_t n_leads, int nbhd,
                         double t_start, double t_step, double total_time, int flags, double* ecg_out,
                         const double* targets, int64_t n_target, const double* target_offsets, int comparison, double* criteria_out,
                         const FitRequest* fit = nullptr);

int ekg_simulate(ekg_model* m, const double* layer_k, const double* leads_zyx, int64_t B, int64_t n_leads, int nbhd,
                 double t_start, double t_step, double total_time, int flags, double* ecg_out) {
	if (!ecg_out) return fail(EKG_E_INVALID, "NULL argument");
	return simulate_host(m, layer_k, leads_zyx, B, n_leads, nbhd, t_start, t_step, total_time, flags, ecg_out, nullptr, 0, nullptr, 0, nullptr);
}

int ekg_simulate_criteria(ekg_model* m, const double* layer_k, const double* leads_zyx, int64_t B, int64_t n_leads, int nbhd,
                          double t_start, double t_step, double total_time, int flags,
                          const double* targets, int64_t n_target, const double* target_offsets, int comparison,
                          double* criteria_out, double* ecg_out) {
	if (!targets || n_target <= 0 || !criteria_out) return fail(EKG_E_INVALID, "NULL argument");
	return simulate_host(m, layer_k, leads_zyx, B, n_leads, nbhd, t_start, t_step, total_time, flags, ecg_out, targets, n_target,
	                     target_offsets, comparison, criteria_out);
}

int ekg_fit_layers_device(ekg_model* m, const double* d_border_k, int64_t B, int64_t n_border, int64_t mid, const double* d9,
                          double step_size, double epsilon, int64_t iterations, double* d_layer_k_out, void* stream) {
	if (!m || !d_border_k || !d9 || !d_layer_k_out) return fail(EKG_E_INVALID, "NULL argument");
	EKG_CUDA(cudaSetDevice(m->device));
	return run_fit(m, d_border_k, B, n_border, m->n_layers, mid, d9, step_size, epsilon, iterations, d_layer_k_out, (cudaStream_t)stream);
}

int ekg_fit_layers(ekg_model* m, const double* border_k, int64_t B, int64_t n_border, int64_t mid, const double* d9,
                   double step_size, double epsilon, int64_t iterations, double* layer_k_out) {
	if (!m || !border_k || !d9 || !layer_k_out) return fail(EKG_E_INVALID, "NULL argument");
	if (B <= 0 || (n_border != 2 && n_border != 3)) return fail(EKG_E_INVALID, "bad sizes");
	EKG_CUDA(cudaSetDevice(m->device));
	const int64_t nb = B * n_border * 9, nk = B * m->n_layers * 9;
	int rc;
	if ((rc = ensure(&m->d_io_border, &m->io_border_cap, nb))) return rc;
	if ((rc = ensure(&m->d_io_k, &m->io_k_cap, nk))) return rc;
	if ((rc = upload(m, m->d_io_border, border_k, (size_t)nb * 8))) return rc;
	if ((rc = run_fit(m, m->d_io_border, B, n_border, m->n_layers, mid, d9, step_size, epsilon, iterations, m->d_io_k, m->stream))) return rc;
	EKG_CUDA(cudaMemcpyAsync(layer_k_out, m->d_io_k, (size_t)nk * 8, cudaMemcpyDeviceToHost, m->stream));
	EKG_CUDA(cudaStreamSynchronize(m->stream));
	return EKG_OK;
}

int ekg_evaluate(ekg_model* m, const double* border_k, int64_t n_border, int64_t mid, const double* d9, double step_size, double epsilon,
                 int64_t iterations, const double* leads_zyx, int64_t B, int64_t n_leads, int nbhd, double t_start, double t_step,
                 double total_time, int flags, const double* targets, int64_t n_target, const double* target_offsets, int comparison,
                 double* criteria_out, double* layer_k_out, double* ecg_out) {
	if (!border_k || !d9) return fail(EKG_E_INVALID, "NULL argument");
	if (criteria_out && (!targets || n_target <= 0)) return fail(EKG_E_INVALID, "criteria requested without targets");
	if (!criteria_out && !ecg_out && !layer_k_out) return fail(EKG_E_INVALID, "no output requested");
	FitRequest f;
	f.border_k = border_k; f.n_border = n_border; f.mid = mid; f.iterations = iterations;
	f.d9 = d9; f.step = step_size; f.eps = epsilon; f.layer_k_out = layer_k_out;
	return simulate_host(m, nullptr, leads_zyx, B, n_leads, nbhd, t_start, t_step, total_time, flags, ecg_out, targets, n_target,
	                     target_offsets, comparison, criteria_out, &f);
}

static int simulate_host(ekg_model* m, const double* layer_k, const double* leads_zyx, int64_t B, int64_t n_leads, int nbhd,
                         double t_start, double t_step, double total_time, int flags, double* ecg_out,
                         const double* targets, int64_t n_target, const double* target_offsets, int comparison, double* criteria_out,
                         const FitRequest* fit) {
	if (!m || (!layer_k && !fit) || !leads_zyx) return fail(EKG_E_INVALID, "NULL argument");
	if (fit && fit->n_border != 2 && fit->n_border != 3) return fail(EKG_E_INVALID, "n_border must be 2 (endo-epi) or 3 (endo-mid-epi)");
	if (B <= 0 || n_leads <= 0 || !(t_step > 0) || !(total_time > 0)) return fail(EKG_E_INVALID, "bad sizes");
	EKG_CUDA(cudaSetDevice(m->device));
	const int64_t T = (int64_t)ceil(total_time / t_step);
	const int64_t nk_out = B * m->n_layers * 9;                        // layer coefficients on the device
	const int64_t nk = fit ? B * fit->n_border * 9 : nk_out;           // coefficients that travel host -> device
	const int64_t nlead = B * n_leads * 3, necg = B * n_leads * T;
	int rc;
	if ((rc = ensure(&m->d_io_k, &m->io_k_cap, nk_out))) return rc;
	if (fit && (rc = ensure(&m->d_io_border, &m->io_border_cap, nk))) return rc;
	if ((rc = ensure(&m->d_io_leads, &m->io_leads_cap, nlead))) return rc;
	if ((rc = ensure(&m->d_io_ecg, &m->io_ecg_cap, necg))) return rc;
	if (m->pin_in_cap < nk + nlead) {
		if (m->h_pin_in) cudaFreeHost(m->h_pin_in);
		m->h_pin_in = nullptr; m->pin_in_cap = 0;
		EKG_CUDA(cudaMallocHost(&m->h_pin_in, (size_t)(nk + nlead) * 8));
		m->pin_in_cap = nk + nlead;
	}
	if (m->pin_out_cap < necg) {
		if (m->h_pin_out) cudaFreeHost(m->h_pin_out);
		m->h_pin_out = nullptr; m->pin_out_cap = 0;
		EKG_CUDA(cudaMallocHost(&m->h_pin_out, (size_t)necg * 8));
		m->pin_out_cap = necg;
	}
	memcpy(m->h_pin_in, fit ? fit->border_k : layer_k, (size_t)nk * 8);
	memcpy(m->h_pin_in + nk, leads_zyx, (size_t)nlead * 8);
	EKG_CUDA(cudaMemcpyAsync(fit ? m->d_io_border : m->d_io_k, m->h_pin_in, (size_t)nk * 8, cudaMemcpyHostToDevice, m->stream));
	EKG_CUDA(cudaMemcpyAsync(m->d_io_leads, m->h_pin_in + nk, (size_t)nlead * 8, cudaMemcpyHostToDevice, m->stream));
	int64_t fit_launches = 0;
	if (fit) {
		if ((rc = run_fit(m, m->d_io_border, B, fit->n_border, m->n_layers, fit->mid, fit->d9, fit->step, fit->eps, fit->iterations, m->d_io_k, m->stream))) return rc;
		fit_launches = m->last_launches;
	}
	// smallest depolarisation rate of the batch, known here without asking the device: the caller's layer
	// coefficients, or -- with the device fit -- the border APs (k1 of an inner layer is the blend of its two
	// border values unless d9[1] asks the descent to move it)
	KHints hints;
	{
		const double* src = fit ? fit->border_k : layer_k;
		const int64_t n = B * (fit ? fit->n_border : m->n_layers);
		double k1_min = INFINITY, decay = 0.0;
		for (int64_t i = 0; i < n; ++i) {
			k1_min = std::min(k1_min, src[i * 9 + 1]);
			decay = std::max(decay, std::max(std::fabs(src[i * 9 + 4] + src[i * 9 + 5]), std::fabs(src[i * 9 + 5])));
		}
		// with the device fit: k1 and k4 of an inner layer are blends of the border values (unless d9 asks the descent to
		// move them), k5 is fitted and moves by a few percent -- a factor 2 on the decay rate is the working assumption, and
		// the true rates of the fitted layers come back with the results and are checked below (hints.verify)
		const bool known = !fit || (fit->d9[1] == 0 && fit->d9[4] == 0);
		if (known && k1_min > 0 && std::isfinite(k1_min) && std::isfinite(decay)) { hints.k1_min = k1_min; hints.decay_max = fit ? 2.0 * decay : decay; }
		hints.verify = fit != nullptr && hints.k1_min > 0;
	}
	int mode_flags = flags;
	std::vector<double> crit_host;
	for (int pass = 0; pass < 2; ++pass) {
		rc = run_ecg(m, m->d_io_k, m->d_io_leads, B, n_leads, nbhd, t_start, t_step, total_time, mode_flags, m->d_io_ecg, m->stream, hints);
		if (rc) return rc;
		int rate_bits[2] = {0, 0};
		if (hints.verify) EKG_CUDA(cudaMemcpyAsync(rate_bits, m->d_k1min, sizeof rate_bits, cudaMemcpyDeviceToHost, m->stream));
		if (criteria_out) {
			const int64_t ntg = n_leads * n_target;
			if ((rc = ensure(&m->d_io_tgt, &m->io_tgt_cap, ntg + n_leads + B * n_leads))) return rc;
			double* d_off = m->d_io_tgt + ntg;
			double* d_crit = d_off + n_leads;
			EKG_CUDA(cudaMemcpyAsync(m->d_io_tgt, targets, (size_t)ntg * 8, cudaMemcpyHostToDevice, m->stream));
			if (target_offsets) EKG_CUDA(cudaMemcpyAsync(d_off, target_offsets, (size_t)n_leads * 8, cudaMemcpyHostToDevice, m->stream));
			// (pageable sources: cudaMemcpyAsync returns once they have been staged)
			if ((rc = run_criteria(m, m->d_io_ecg, m->d_io_tgt, target_offsets ? d_off : nullptr, d_crit, B, n_leads, T, n_target, comparison, m->stream))) return rc;
			crit_host.resize((size_t)(B * n_leads));
			EKG_CUDA(cudaMemcpyAsync(crit_host.data(), d_crit, crit_host.size() * 8, cudaMemcpyDeviceToHost, m->stream));
		}
		if (ecg_out) EKG_CUDA(cudaMemcpyAsync(m->h_pin_out, m->d_io_ecg, (size_t)necg * 8, cudaMemcpyDeviceToHost, m->stream));
		if (fit && fit->layer_k_out) EKG_CUDA(cudaMemcpyAsync(fit->layer_k_out, m->d_io_k, (size_t)nk_out * 8, cudaMemcpyDeviceToHost, m->stream));
		EKG_CUDA(cudaStreamSynchronize(m->stream));
		if (!hints.verify) break;
		// the fitted layers' true rates: were the estimates the mode decision was taken on good enough?
		float rate[2];
		memcpy(rate, rate_bits, 8);
		const bool k1_ok = (double)rate[0] >= hints.k1_min * (1.0 - 1e-6);
		const bool decay_ok = (flags & 0xff) == EKG_MODE_DIRECT || decay_within_clamp(m, (double)rate[1], t_start) ||
		                      !decay_within_clamp(m, hints.decay_max, t_start);   // (already sent through DIRECT)
		if (k1_ok && decay_ok) break;
		hints.k1_min = (double)rate[0];
		hints.decay_max = (double)rate[1];
		hints.verify = false;   // second pass with the measured rates
	}
	m->last_launches += fit_launches;
	if (ecg_out) memcpy(ecg_out, m->h_pin_out, (size_t)necg * 8);
	if (criteria_out) memcpy(criteria_out, crit_host.data(), crit_host.size() * 8);
	return EKG_OK;
}

double ekg_last_kernel_ms(ekg_model* m) {
	if (!m || !m->ev_recorded) return -1.0;
	float ms = -1.f;
	if (cudaEventSynchronize(m->ev_k1) != cudaSuccess || cudaEventElapsedTime(&ms, m->ev_k0, m->ev_k1) != cudaSuccess) { cudaGetLastError(); return -1.0; }
	return (double)ms;
}

int64_t ekg_last_launch_count(const ekg_model* m) { return m ? m->last_launches : 0; }
const char* ekg_last_kernel_name(const ekg_model* m) { return m ? m->last_kernel : "none"; }

}  // extern "C"
