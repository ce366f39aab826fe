// ekgsim_b200/csrc/ekg_internal.cuh -- internal declarations shared by the CUDA translation
// units of libekgsim_b200.so.  Nothing here is part of the ABI (see include/ekgsim_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/ekgsim_b200.h"

namespace ekg {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define EKG_CUDA(call)                                                             \
	do {                                                                           \
		cudaError_t e__ = (call);                                                  \
		if (e__ != cudaSuccess) return ekg::cuda_fail(e__, #call, __FILE__, __LINE__); \
	} while (0)

// ---- static layout constants --------------------------------------------------------------------
constexpr int kMaxNbr = 26;          // full cube (simulator.h:367-373)
constexpr int kParamStride = 12;     // floats per (individual, layer) in the DIRECT parameter table
constexpr int kEcgThreads = 256;     // (vector, sample) pairs per CTA: 8 warps, 2 per SM sub-partition
constexpr int kMaxVecPerTile = 4;    // parameter vectors one pair tile may span
constexpr int kChunk = 256;          // voxels staged in shared memory per phase (fewer when a tile spans > 2 vectors)
constexpr int kSmemRows = 512;       // (voxel, vector) rows staged per phase
constexpr int kMaxLeadsPerPass = 4;

// Neighbour table handed to kernels by value.
struct NbrTable {
	int n;
	int8_t dz[kMaxNbr], dy[kMaxNbr], dx[kMaxNbr];
	int8_t bit[kMaxNbr];  // position of this neighbour in the 26-bit cube occupancy mask
	float fz[kMaxNbr], fy[kMaxNbr], fx[kMaxNbr];  // the same offsets as floats (no I2F in the kernel)
};

// One unit of ECG work: a run of voxels of ONE layer in the layer-sorted voxel list.
struct Segment {
	int32_t begin, end;  // [begin, end) into the ECG voxel list
	int32_t layer;       // 1-based layer number
	int32_t kind;        // moment segments: kSegInterior = every voxel has all 8 cube corners occupied; else 0
};
constexpr int32_t kSegInterior = 1;

// A run of consecutive (slice s, vector b, sample t) pairs, p = (s*B + b)*T + t.
struct PairTile {
	int32_t begin, end;
};

struct EcgArgs {
	const uint32_t* pos;    // packed x | y<<11 | z<<22 (unpadded voxel coordinates)
	const uint32_t* mask;   // 26-bit occupancy of the cube neighbourhood (bit k: voxel c - dif_k occupied)
	const float* at32;      // activation time per ECG-list voxel (fp32 copy of the f64 map)
	const Segment* segs;
	const PairTile* tiles;
	const float* params;    // [B][n_layers][kParamStride]
	const float* ftab;      // HOISTED: [B][n_layers][2][T]  (F1, F2)
	const double* leads;    // [B][L][3] (z,y,x)
	const float* t_hi;      // [T] time samples, hi/lo split of the f64 value
	const float* t_lo;
	double* partial;        // [n_segs][S][B][L][T]
	int32_t n_segs, B, L, T, n_layers;
	int32_t S;              // voxel slices per segment (ecg.cu: pair index = (slice, vector, sample)); 1 = none
	int32_t lead0;          // first lead handled by this launch (L > kMaxLeadsPerPass -> several passes)
	NbrTable nbr;
};

// SEPARABLE path, moments of the lead field over the voxels of a segment (ecg_moment_kernel)
struct MomentArgs {
	const uint32_t* pos;
	const uint32_t* mask;
	const float* at32;
	const float4* vox;      // corners kernels: {z + 1, y + 1 | corner bits, x + 1, activation time} (capi.cu, gather_at_kernel)
	int32_t* near_flag;     // [groups][n_segs] set by the interior kernel: a lead is too close for the series, redo by direct sum
	const Segment* segs;
	const float* params;    // [B][n_layers][kParamStride]
	const double* leads;    // [B][L][3]
	double* mom;            // [n_segs][B][L][3]: sum G, sum G h1, sum G h2
	int32_t B, L, n_layers, lead0;
	int32_t vb_shift;       // log2 of the parameter vectors a CTA serves (threads = vectors x voxel lanes)
	int32_t force_sum;      // corners kernel: take every interior segment too (EKG_FLAG_CORNER_SUM), not only the flagged ones
	NbrTable nbr;
};

// occupancy-mask bits of the 8 cube corners (the reference's "3D4" stencil) in the 26-neighbour cube list
constexpr uint32_t kCornerMask = (1u << 0) | (1u << 2) | (1u << 6) | (1u << 8) | (1u << 17) | (1u << 19) | (1u << 23) | (1u << 25);

struct AutoArgs {
	const uint8_t* layer;   // padded dense grid, 0 = empty
	double* time;           // padded dense grid, +inf = not reached
	const uint32_t* pidx;   // padded linear index of every occupied voxel, raster order
	const double* wtab;     // [nl1][nl1][3] edge weight: fl(T[lu][lv] * sqrt(|dif|^2)), |dif|^2 in 1..3
	int* flags;             // flags[i] != 0 <=> sweep i changed something
	int* sweeps_out;
	int64_t n;
	int32_t nl1;            // n_layers + 1
	int32_t n_nbr;
	int32_t off[kMaxNbr];   // padded linear offset of neighbour k (p - off[k])
	int32_t sq[kMaxNbr];    // |dif|^2 - 1
	int32_t max_sweeps;
};

// Brick-frontier automaton (automaton.cu): the padded grid is tiled by 4x4x4 bricks, one warp per brick.
constexpr int kBrick = 4;
constexpr int kBrickHalo = kBrick + 2;                                // 6
constexpr int kBrickCells = kBrickHalo * kBrickHalo * kBrickHalo;    // 216 (brick + one-voxel halo)
constexpr int kBrickWarps = 8;                                        // bricks in flight per CTA

// Linked run (automaton.cu, "peer-linked sharded automaton"): the ranks of a z-slab sharded model write the planes at their
// slab faces straight into the neighbouring ranks' grids and queue the neighbours' bricks in the neighbours' rings, from
// inside the frontier kernel, through peer-mapped memory (NVLink); rank 0 detects global termination.
constexpr int kMaxLinkRanks = 16;
// The work queue is kBrickBuckets rings of brick ids, ordered by activation time: a brick is queued in the ring of the
// time bucket (width BrickArgs::delta) of the earliest voxel that changed next to it, warps take from the earliest
// non-empty bucket.  A plain FIFO visits bricks in hop order and has to redo everything downstream whenever a faster
// path arrives later (4x heart: 4.3 visits per brick); in time order most of that never happens (2.9 per brick in a
// brick-level emulation of the schedule).  The buckets are a cyclic window starting at `cur`; later times wait in the last one.
constexpr int kBrickBuckets = 16;
constexpr int kBucketsBehind = 4, kBucketsAhead = 10;   // the window of buckets around `cur`: [cur - 4, cur + 10), two rings spare
constexpr int kCntHead = 16, kCntTail = kCntHead + kBrickBuckets, kCntCount = kCntTail + kBrickBuckets;
constexpr int kBrickCounters = kCntCount + kBrickBuckets;   // ints behind the rings, see BrickArgs::counters
inline int64_t brick_ring_capacity(int64_t n) {   // slots per ring: a brick sits in at most one ring (flag[]), so n can never overflow
	int64_t cap = 1;
	while (cap < (n > 2 ? n : 2)) cap <<= 1;
	return cap;
}
inline int64_t brick_state_ints(int64_t n) { return 2 * n + kBrickBuckets * brick_ring_capacity(n) + kBrickCounters; }
constexpr uint8_t kOwnMe = 1, kOwnBelow = 2, kOwnAbove = 4;
struct BrickLink {
	double* time_dn;         // padded time grid of the rank below (same layout as ours), NULL = none
	double* time_up;
	int* state_dn;           // its brick state: flag[n] | first_visit[n] | rings[kBrickBuckets][qmask + 1] | counters[kBrickCounters]
	int* state_up;
	int* counters_of[kMaxLinkRanks];   // every rank's counters (only rank 0, the termination detector, reads them)
	int32_t rank, n_ranks;
	int32_t z_first, z_last; // our first and last own voxel plane (unpadded z)
	uint32_t plane;          // pY * pX
};

struct BrickArgs {
	const uint8_t* layer;    // padded dense grid
	double* time;            // padded dense grid
	const double* wtab;      // [nl1][nl1][3]
	const uint32_t* origin;  // [n_live] padded linear index of the brick's first voxel
	const int32_t* nbr;      // [n_live][26] live index of the neighbouring brick in cube direction k, -1 = none
	int* flag;               // [n_live] 1 while the brick sits in the ring
	int* first_visit;        // [n_live] 1 for the bricks of the start voxels until their first visit
	const uint8_t* own;      // [n_live] sharded run: kOwnMe for the bricks this rank relaxes (NULL = all), kOwnBelow / kOwnAbove for
	                         // those the neighbouring ranks relax (a brick that straddles a slab face has several owners)
	int* queue;              // kBrickBuckets rings of brick ids, qmask + 1 slots each, -1 = empty slot
	int* counters;           // [0] ring positions claimed (bounded run), [1] cur = earliest bucket that may hold work, [2] pending
	                         // (queued or in work), [3] warps that gave up waiting, [4] visits, [5] inner sweeps, [6] stop, linked
	                         // run: [7] verdict (1 = terminated everywhere, 2 = aborted), [8] bricks queued at other ranks,
	                         // [9] bricks other ranks queued here, [10] cells written to other ranks; [11] lost races for the last
	                         // brick of a bucket, [kCntHead + r] / [kCntTail + r] / [kCntCount + r] head / tail / published - taken of ring r
	float inv_delta;         // 1 / bucket width (ms); 0 = one bucket (plain FIFO)
	uint32_t budget;         // bounded relaxation (sharded run): no further ring position is claimed once `budget` have been; 0 = no bound
	uint32_t qmask;
	int32_t n_live, nl1, n_nbr, pY, pX, w_in_smem;
	int32_t loff[kMaxNbr];   // neighbour offset inside the 10^3 shared-memory cell array
	int32_t sq[kMaxNbr];     // |dif|^2 - 1
	int8_t dz[kMaxNbr], dy[kMaxNbr], dx[kMaxNbr];
	BrickLink link;
};

}  // namespace ekg

// The opaque handle of the C ABI.
struct ekg_model {
	int device = 0;
	cudaStream_t stream = nullptr;
	int64_t Z = 0, Y = 0, X = 0;
	int64_t pZ = 0, pY = 0, pX = 0;  // padded (zero border) dims
	int n_layers = 0;
	int64_t n_occ = 0;               // occupied voxels in the whole model
	int sm_count = 148;

	// host copies
	std::vector<uint8_t> h_layer;    // raster, start flags stripped: a lazily made host copy (ekg_model_ap_classes)
	std::vector<int64_t> h_starts;   // raster indices of start voxels
	std::vector<int64_t> h_occ_before_z;  // [Z + 1] occupied voxels in the planes below z (a z-slab is a run of d_auto_pidx)
	std::vector<double> h_transfer;
	int64_t t_rows = 0, t_cols = 0;
	std::vector<double> h_delay;     // raster activation map (0 = empty / never reached): a lazily made host copy,
	bool h_delay_valid = false;      // the map itself lives in d_time_pad
	unsigned long long* d_range = nullptr;  // [2] min / max key of the activation times (publish_activation)
	bool have_activation = false;
	double t0 = 0.0;                 // centre of the activation-time range (HOISTED kernel)
	double at_max = 0.0;             // latest activation time of the model (SEPARABLE: first saturated sample)
	double at_min = 0.0;
	float activation_ms = 0.f;       // device time of the last automaton run

	// automaton state (device)
	uint8_t* d_layer_pad = nullptr;
	double* d_time_pad = nullptr;
	uint32_t* d_auto_pidx = nullptr;
	double* d_wtab = nullptr;
	int* d_flags = nullptr;
	int max_sweeps = 1 << 16;
	// brick-frontier automaton
	int64_t n_bricks = 0;
	uint32_t* d_brick_origin = nullptr;
	int32_t* d_brick_nbr = nullptr;
	int* d_brick_state = nullptr;        // flag[n] | first_visit[n] | rings | counters (brick_state_ints)
	float brick_delta = 0.f;             // time-bucket width of the work queue (ms), 0 = FIFO
	std::vector<int32_t> h_start_bricks;
	uint32_t* d_start_pidx = nullptr;    // padded indices of the start voxels (h_starts)
	int32_t* d_start_bricks = nullptr;   // h_start_bricks
	std::vector<int64_t> h_start_brick_bz;   // brick z index of every start brick
	// z-slab sharded automaton (ekg_model_activation_begin / _relax / _export / _merge / _end)
	int64_t bZ = 0, bY = 0, bX = 0;
	int32_t* d_brick_index = nullptr;    // dense brick grid [bZ][bY][bX] -> live brick id, -1 = no occupied voxel
	uint8_t* d_brick_own = nullptr;
	int* d_brick_mark = nullptr;         // bricks to queue at the next relax
	unsigned long long* d_improved = nullptr;
	bool shard_active = false;
	cudaStream_t merge_stream = nullptr;  // stream of the last ekg_model_activation_merge_async (the next relax / end waits for it)
	bool merge_pending = false;
	int64_t last_brick_visits = 0;
	// peer-linked sharded automaton (ekg_model_activation_link / _linked_launch / _linked_wait / _linked_gather)
	struct PeerLink {
		bool active = false;
		int rank = 0, n_ranks = 0, below = -1, above = -1;     // below / above: the nearest ranks with a non-empty slab
		int colocated = 1;                                     // ranks of this process on this device (they share its SMs)
		std::vector<int64_t> slabs;                            // [n_ranks][2]
		std::vector<double*> time;                             // every rank's d_time_pad as seen from this device (own entry = ours)
		std::vector<int*> state;                               // every rank's d_brick_state
		std::vector<void*> ipc_opened;                         // what cudaIpcCloseMemHandle has to release
		std::vector<int> peers_acquired;                       // devices whose peer access this link holds a reference on
		bool launched = false;
		bool timed = false;                                    // work-queue mode the link was made for
		cudaEvent_t ev0 = nullptr, ev1 = nullptr;
		float kernel_ms = 0.f;
	} link;

	// ECG voxel list (layer-sorted, restricted to the slab)
	int64_t slab_z0 = 0, slab_z1 = 0;
	int64_t n_ecg = 0;
	uint32_t* d_pos = nullptr;
	uint32_t* d_mask = nullptr;
	uint32_t* d_ecg_pidx = nullptr;  // padded index, for gathering activation times
	double* d_at = nullptr;
	float* d_at32 = nullptr;
	float4* d_vox = nullptr;         // per-voxel record of the moment kernel (MomentArgs::vox), refreshed with d_at
	std::vector<int64_t> layer_off;  // n_layers + 1 offsets into the ECG list
	std::vector<int64_t> interior_cnt;  // per layer: its range of the list starts with this many interior voxels (all 8 cube
	                                    // corners occupied), the boundary voxels follow; raster order inside both parts

	// per-call scratch (grown on demand)
	ekg::Segment* d_segs = nullptr;  int64_t segs_cap = 0;  int64_t n_segs = 0;  int64_t seg_len = 0;
	ekg::PairTile* d_tiles = nullptr; int64_t tiles_cap = 0; int64_t n_tiles = 0; int64_t tiles_B = 0, tiles_T = 0;  // tiles_B = S * B
	// SEPARABLE path: its own segment table (sized for voxel x vector work, not for the time loop)
	ekg::Segment* d_msegs = nullptr; int64_t msegs_cap = 0;  int64_t n_msegs = 0;  int64_t mseg_len = 0;
	int32_t* d_mseg_first = nullptr; int64_t mseg_first_cap = 0;   // first segment of every layer, n_layers + 1 entries
	double* d_mom = nullptr;         int64_t mom_cap = 0;          // [n_msegs][B][L][3] moments
	double* d_lmom = nullptr;        int64_t lmom_cap = 0;         // [B][n_layers][L][3] the same per layer
	int32_t* d_near = nullptr;       int64_t near_cap = 0;         // MomentArgs::near_flag
	int* d_k1min = nullptr;                                        // [2] float bits of min k1 / max decay rate over (vector, layer), by ecg_params_kernel
	float* d_params = nullptr;       int64_t params_cap = 0;
	double* d_tail = nullptr;        int64_t tail_cap = 0;         // ln(2^(k7/k6) - 1) per (vector, layer), f64 (ecg_params_kernel)
	float* d_ftab = nullptr;         int64_t ftab_cap = 0;
	float* d_times = nullptr;        int64_t times_cap = 0;
	double times_t_start = 0, times_t_step = 0;  int64_t times_T = 0;  // what d_times currently holds
	std::vector<double> h_times;     // host copy of the f64 sample times
	double* d_partial = nullptr;     int64_t partial_cap = 0;
	double* d_partial2 = nullptr;    int64_t partial2_cap = 0;   // row-block sums of the two-pass reduction
	double* d_io_k = nullptr;        int64_t io_k_cap = 0;     // staging for the host-buffer entry point
	double* d_io_leads = nullptr;    int64_t io_leads_cap = 0;
	double* d_io_ecg = nullptr;      int64_t io_ecg_cap = 0;
	double* d_io_tgt = nullptr;      int64_t io_tgt_cap = 0;   // targets | offsets | criteria
	double* d_io_border = nullptr;   int64_t io_border_cap = 0; // border APs of the device-side layer fit
	double* d_fit_conn = nullptr;    int64_t fit_conn_cap = 0;  // connector tables of the layer fit (fit.cu)
	double* h_pin_in = nullptr;      int64_t pin_in_cap = 0;   // pinned host staging
	double* h_pin_out = nullptr;     int64_t pin_out_cap = 0;

	cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;  bool ev_recorded = false;
	int64_t last_launches = 0;
	const char* last_kernel = "none";
};

namespace ekg {
// automaton.cu
int run_automaton(ekg_model* m, int64_t* sweeps_out);
int shard_begin(ekg_model* m);
int shard_relax(ekg_model* m, int64_t max_visits, int64_t* visits_out, int64_t* leftover_out);
int shard_export(ekg_model* m, int64_t z_begin, int64_t z_end, double* d_planes, cudaStream_t st);
int shard_merge(ekg_model* m, int64_t z_begin, int64_t z_end, const double* d_planes, int64_t* improved_out, unsigned long long* d_count,
                cudaStream_t st);
int shard_link_info(ekg_model* m, void* info_out);
int shard_link(ekg_model* m, int rank, int n_ranks, const void* infos, const int64_t* slabs);
int shard_unlink(ekg_model* m);
int shard_linked_launch(ekg_model* m, int max_ctas);
int shard_linked_wait(ekg_model* m, int64_t* visits_out, int64_t* remote_out);
int shard_linked_gather(ekg_model* m);
// ecg.cu
// What the caller knows on the host about the coefficients of the batch (all zero: nothing -- the SEPARABLE path
// then reads both back from the device, one stream synchronisation):
//   k1_min     smallest depolarisation rate k1 over all (vector, layer): where the sigmoid saturates
//   decay_max  largest of |k4 + k5|, |k5|: whether the hoisted exponentials exp(+-k (at - t0)), exp(-k (t - t0)) stay
//              inside the 2^60 clamp of the HOISTED / SEPARABLE kernels (otherwise the run goes through DIRECT)
//   verify     the hints are an estimate (device-side layer fit: k5 is a fitted coefficient): ecg_params_kernel also reduces the
//              true values into ekg_model::d_k1min, which the caller reads back with its results and checks (decay_within_clamp)
struct KHints { double k1_min = 0.0, decay_max = 0.0; bool verify = false; };
bool decay_within_clamp(const ekg_model* m, double decay_max, double t_start);
int run_ecg(ekg_model* m, const double* d_layer_k, const double* d_leads, int64_t B, int64_t L, int nbhd,
            double t_start, double t_step, double total_time, int flags, double* d_ecg, cudaStream_t st, KHints hints = KHints());
int run_criteria(ekg_model* m, const double* d_ecg, const double* d_targets, const double* d_offsets, double* d_crit,
                 int64_t B, int64_t L, int64_t T, int64_t n_target, int comparison, cudaStream_t st);
int make_nbr_table(int nbhd, NbrTable* out);
// fit.cu
int run_fit(ekg_model* m, const double* d_border_k, int64_t B, int64_t n_border, int64_t n_layers, int64_t mid, const double* d9,
            double step, double eps, int64_t iterations, double* d_layer_k, cudaStream_t st);
template <class T>
int ensure(T** p, int64_t* cap, int64_t need);
}  // namespace ekg
