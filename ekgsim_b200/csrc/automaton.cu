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
//  * automaton_brick_kernel -- frontier based.  The grid is tiled by 4x4x4 bricks; a work queue holds
//    the bricks whose surroundings changed in the previous round (initially the bricks of the start
//    voxels).  ONE WARP takes a brick, stages it with a one-voxel halo in shared memory (216 f64
//    times + 216 u8 layers, two interior voxels per lane), relaxes it to a LOCAL fixed point with
//    in-place sweeps synchronised by __syncwarp only, writes the improved times back and queues a
//    neighbouring brick if one of its cells would improve through a changed voxel.  There are no
//    grid-wide rounds: warps pull brick ids from a lock-free queue (a `pending` count of queued +
//    in-work bricks for termination), so the wavefront advances at the pace of its own
//    dependencies; write-back uses a 64-bit atomicMin because two warps may hold the same brick at
//    once.  The queue is one FIFO ring for models whose frontier fits the machine and 16 rings of
//    time buckets for larger ones (template parameter TIMED, see "Work queue" below): visiting
//    bricks in the order of their activation times halves the number of visits on a 4x heart.
//  * automaton_kernel (EKGSIM_B200_AUTOMATON=sweep) -- the plain label-correcting sweep over all
//    occupied voxels, kept as the simple cross-check.
//
//  * automaton_brick_kernel<.., LINKED> -- the same kernel on a model sharded into z-slabs over several GPUs: improved
//    voxels on the slab's first / last plane are also written (64-bit atomicMin, system scope) into the neighbouring
//    rank's grid through peer-mapped memory, and the neighbour's bricks that can see them are pushed into the
//    neighbour's ring, all from inside the kernel (NVLink atomics, no host round, no collective).  Rank 0's first warp
//    detects global termination from every rank's (pending, sent, received) counters, see link_detector.
//
// Reads of the time field bypass L1 (ld.global.cg) since other SMs update it; 8-byte accesses do not
// tear; values only ever decrease, so stale halo reads are harmless (the writer re-queues us).
// Working set on model_24: 1.6 MB padded u8 layers + 13 MB padded f64 times -> L2 resident.

#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <unistd.h>

#include <algorithm>
#include <map>
#include <mutex>

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

// Work queue of the frontier automaton: kBrickBuckets rings of brick ids, one per time bucket (ekg_internal.cuh).  A brick
// is in the rings at most once (flag[b] = 1 from push to pop), so the qmask + 1 slots of a ring can never overflow.
// `pending` counts bricks that are queued or being processed; it can only reach 0 when the whole field is at its fixed
// point.  FIFO mode uses ring 0 alone: consumers claim the next position with one atomicAdd on the head and wait there,
// producers hand their bricks to the waiting positions.  Time-bucket mode must not bind a warp to a ring that may stay
// empty for long, so every ring also counts `published - taken`: a consumer first takes one off that count and only
// with a positive result claims a position (whose slot a producer has filled or is about to); losing the race for the
// last brick of a bucket costs two atomics on the count and no ring position.

// n bricks (ids in lanes `pushers`) into the ring of time bucket `bucket` of the queue at (ring, cnt): one tail
// reservation for all of them; SYS = the queue lives in another GPU's memory
template <bool SYS>
__device__ __forceinline__ void ring_publish(int* ring, int* cnt, uint32_t qmask, int bucket, int nb, int lane) {
	constexpr unsigned kFull = 0xffffffffu;
	const int r = bucket & (kBrickBuckets - 1);
	int* slots = ring + (size_t)r * (qmask + 1u);
	const unsigned pushers = __ballot_sync(kFull, nb >= 0);
	const int n = __popc(pushers);
	unsigned base = 0;
	if (lane == 0) base = SYS ? atomicAdd_system((unsigned*)cnt + kCntTail + r, (unsigned)n) : atomicAdd((unsigned*)cnt + kCntTail + r, (unsigned)n);
	base = __shfl_sync(kFull, base, 0);
	if (nb >= 0) {
		int* slot = slots + ((base + __popc(pushers & ((1u << lane) - 1u))) & qmask);
		if (SYS) atomicExch_system(slot, nb); else atomicExch(slot, nb);
	}
	if (SYS) __threadfence_system(); else __threadfence();
	__syncwarp();
	if (lane == 0) { if (SYS) atomicAdd_system(cnt + kCntCount + r, n); else atomicAdd(cnt + kCntCount + r, n); }   // takers may come now
}

__device__ __forceinline__ unsigned long long global_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}

constexpr unsigned long long kLinkTimeoutNs = 12ull * 1000 * 1000 * 1000;   // a linked run that has not ended by then is aborted

// Global termination of a linked run (Mattern's four-counter scheme plus an idle test).  A "message" is a brick queued in
// another rank's ring.  The sender counts it in its own `sent` BEFORE the peer's `pending` goes up, and in the peer's
// `received` AFTER that (link_push).  Rank 0's first warp reads (pending, sent, received) of every rank, lane r = rank r,
// over and over; the run is over when a whole pass saw every rank idle (pending == 0, read before its `sent`) and the sum
// of `sent` of this pass equals the sum of `received` of the PREVIOUS pass: then nothing was in flight when the previous
// pass ended, nobody has sent since, and an idle rank only wakes up through a message.  The verdict is written into every
// rank's counters[7]; waiting warps everywhere watch their own copy.
__device__ void link_detector(const BrickArgs& a, int lane) {
	constexpr unsigned kFull = 0xffffffffu;
	const bool mine = lane < a.link.n_ranks;
	volatile int* c = mine ? a.link.counters_of[lane] : nullptr;
	unsigned r_prev = 0;
	bool have_prev = false;
	const unsigned long long t0 = global_ns();
	for (;;) {
		int pend = 0, gave_up = 0;
		unsigned sent = 0, recv = 0;
		if (mine) { pend = c[2]; gave_up = c[3]; }
		__threadfence_system();
		if (mine) sent = (unsigned)c[8];
		__threadfence_system();
		if (mine) recv = (unsigned)c[9];
		const bool idle = __all_sync(kFull, pend == 0);
		const bool broken = __any_sync(kFull, gave_up != 0);
		const unsigned s_sum = __reduce_add_sync(kFull, sent), r_sum = __reduce_add_sync(kFull, recv);
		int verdict = 0;
		if (broken || global_ns() - t0 > kLinkTimeoutNs) verdict = 2;
		else if (idle && have_prev && s_sum == r_prev) verdict = 1;
		if (verdict) {
			if (mine) c[7] = verdict;
			__threadfence_system();
			return;
		}
		r_prev = r_sum;
		have_prev = true;
		__nanosleep(400);
	}
}

// the 27 bricks (3x3x3 around ours, bit (dz+1)*9 + (dy+1)*3 + (dx+1), 13 = ours) that hold a cell within one voxel of
// the cell (cz, cy, cx) of our brick
__device__ __forceinline__ unsigned reach27(int cz, int cy, int cx) {
	const unsigned mz = (cz == 0 ? 1u : 0u) | 2u | (cz == kBrick - 1 ? 4u : 0u);
	const unsigned my = (cy == 0 ? 1u : 0u) | 2u | (cy == kBrick - 1 ? 4u : 0u);
	const unsigned mx = (cx == 0 ? 1u : 0u) | 2u | (cx == kBrick - 1 ? 4u : 0u);
	const unsigned zs = ((mz & 1u) ? 0x1ffu : 0u) | ((mz & 2u) ? 0x1ffu << 9 : 0u) | ((mz & 4u) ? 0x1ffu << 18 : 0u);
	const unsigned ys = (((my & 1u) ? 7u : 0u) | ((my & 2u) ? 7u << 3 : 0u) | ((my & 4u) ? 7u << 6 : 0u)) * ((1u << 18) | (1u << 9) | 1u);
	const unsigned xs = mx * (((1u << 18) | (1u << 9) | 1u) * ((1u << 6) | (1u << 3) | 1u));
	return zs & ys & xs;
}

// Queue, in the ring of the neighbouring rank whose brick state is `pstate`, those of its bricks among `reach` (reach27
// bits relative to our brick b) that it relaxes.  Our improved times have been written to its grid and fenced before.
__device__ __forceinline__ void link_push(const BrickArgs& a, int* pstate, unsigned reach, uint8_t owner, int b, int lane, float key) {
	constexpr unsigned kFull = 0xffffffffu;
	int* pflag = pstate;
	int* pring = pstate + 2 * (size_t)a.n_live;
	int* pcnt = pring + (size_t)kBrickBuckets * (a.qmask + 1u);
	int nb = -1;
	if (lane < 27 && ((reach >> lane) & 1u)) {
		nb = lane == 13 ? b : __ldg(a.nbr + (size_t)b * 26 + (lane > 13 ? lane - 1 : lane));
		if (nb >= 0 && !(a.own[nb] & owner)) nb = -1;
		if (nb >= 0 && atomicExch_system(pflag + nb, 1) != 0) nb = -1;   // already in its queue: whoever pops it reads our times
	}
	const unsigned pushers = __ballot_sync(kFull, nb >= 0);
	const int n_push = __popc(pushers);
	if (!n_push) return;
	int bucket = 0;
	if (lane == 0) {
		atomicAdd(a.counters + 8, n_push);                                      // sent -- before the work shows up over there
		__threadfence_system();
		atomicAdd_system(pcnt + 2, n_push);                                     // its pending: the rank is busy from here on
		const int cur = *(volatile int*)(pcnt + 1);                             // its window of time buckets
		bucket = min(max((int)(key * a.inv_delta), cur), cur + kBucketsAhead - 1);
	}
	bucket = __shfl_sync(kFull, bucket, 0);
	ring_publish<true>(pring, pcnt, a.qmask, bucket, nb, lane);
	__threadfence_system();
	__syncwarp();
	if (lane == 0) atomicAdd_system(pcnt + 9, n_push);                          // received -- after its pending went up
}

// The automaton's neighbourhoods in the order make_nbr_table (ecg.cu) produces them -- (dz, dy, dx) lexicographic over
// {-1, 0, 1}^3 (NBR = 26) or {0} x {-1, 0, 1}^2 (NBR = 8) without the centre -- as compile-time functions of k, so that the
// unrolled sweep addresses shared memory with immediate offsets (fill_brick_args checks the table against them).
template <int NBR> __host__ __device__ constexpr int nb_idx(int k) { return NBR == 26 ? k + (k >= 13 ? 1 : 0) : k + (k >= 4 ? 1 : 0); }
template <int NBR> __host__ __device__ constexpr int nb_dz(int k) { return NBR == 26 ? nb_idx<NBR>(k) / 9 - 1 : 0; }
template <int NBR> __host__ __device__ constexpr int nb_dy(int k) { return NBR == 26 ? (nb_idx<NBR>(k) / 3) % 3 - 1 : nb_idx<NBR>(k) / 3 - 1; }
template <int NBR> __host__ __device__ constexpr int nb_dx(int k) { return nb_idx<NBR>(k) % 3 - 1; }
template <int NBR> __host__ __device__ constexpr int nb_loff(int k) { return (nb_dz<NBR>(k) * kBrickHalo + nb_dy<NBR>(k)) * kBrickHalo + nb_dx<NBR>(k); }
template <int NBR> __host__ __device__ constexpr int nb_sq(int k) {
	return nb_dz<NBR>(k) * nb_dz<NBR>(k) + nb_dy<NBR>(k) * nb_dy<NBR>(k) + nb_dx<NBR>(k) * nb_dx<NBR>(k) - 1;
}

// WSMEM: the edge-weight table sits in shared memory (it does for up to 36 layers).  A compile-time fact, so that the
// table is read with LDS and 32-bit address arithmetic -- a pointer that may be shared or global made every lookup a
// generic 64-bit load with four instructions of address arithmetic, in the innermost loop.
template <int NBR, bool LINKED, bool TIMED, bool WSMEM>
__global__ void __launch_bounds__(32 * kBrickWarps, TIMED ? 3 : 2) automaton_brick_kernel(BrickArgs a) {
	__shared__ double s_t_all[kBrickWarps][kBrickCells];
	__shared__ uint8_t s_l_all[kBrickWarps][kBrickCells];
	__shared__ uint16_t s_o_all[kBrickWarps][kBrickCells];   // byte offset of the cell's weight row: layer * nl1 * 3 * 8
	extern __shared__ double s_w[];  // edge-weight table [nl1][nl1][3] when it fits (a.w_in_smem)
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (WSMEM) {
		for (int i = tid; i < a.nl1 * a.nl1 * 3; i += blockDim.x) s_w[i] = __ldg(a.wtab + i);
		__syncthreads();
	}
	if (LINKED && a.link.rank == 0 && blockIdx.x == 0 && warp == 0) {   // no block-wide barrier below this line
		link_detector(a, lane);
		return;
	}
	const double* __restrict__ wt = WSMEM ? (const double*)s_w : a.wtab;
	double* s_t = s_t_all[warp];
	uint8_t* s_l = s_l_all[warp];
	uint16_t* s_o = s_o_all[warp];
	constexpr int kOwn = kBrick * kBrick * kBrick / 32;  // interior voxels per lane (2)
	constexpr int kInnerCap = 256;
	constexpr unsigned kFull = 0xffffffffu;
	int loc[kOwn], cz[kOwn], cy[kOwn], cx[kOwn];
#pragma unroll
	for (int o = 0; o < kOwn; ++o) {
		const int v = lane + o * 32;
		cz[o] = v / (kBrick * kBrick); cy[o] = (v / kBrick) % kBrick; cx[o] = v % kBrick;
		loc[o] = ((cz[o] + 1) * kBrickHalo + (cy[o] + 1)) * kBrickHalo + (cx[o] + 1);
	}
	// where this lane's cells of the 6^3 staging cube lie relative to the brick's first voxel (the same for every brick)
	constexpr int kStage = (kBrickCells + 31) / 32;
	int cell_off[kStage];
#pragma unroll
	for (int j = 0; j < kStage; ++j) {
		const int i = lane + 32 * j;
		const int lz = i / (kBrickHalo * kBrickHalo), ly = (i / kBrickHalo) % kBrickHalo, lx = i % kBrickHalo;
		cell_off[j] = ((lz - 1) * a.pY + (ly - 1)) * a.pX + (lx - 1);
	}
	// which of a voxel's neighbours lie outside the brick (bit k: neighbour k), the same for every brick
	unsigned outward[kOwn];
#pragma unroll
	for (int o = 0; o < kOwn; ++o) {
		outward[o] = 0;
#pragma unroll 1
		for (int k = 0; k < NBR; ++k) {
			const int nz = cz[o] - a.dz[k], ny = cy[o] - a.dy[k], nx = cx[o] - a.dx[k];
			if (nz < 0 || nz >= kBrick || ny < 0 || ny >= kBrick || nx < 0 || nx >= kBrick) outward[o] |= 1u << k;
		}
	}
	const double inf = __longlong_as_double(0x7ff0000000000000LL);
	const int nl3 = a.nl1 * 3;
	int visits = 0, sweeps = 0, remote_cells = 0;
	volatile int* vq = a.queue;
	volatile int* vpending = a.counters + 2;
	volatile int* vstop = a.counters + 6;
	volatile int* vverdict = a.counters + 7;
	unsigned long long t_begin = 0;
	if (LINKED) t_begin = global_ns();

	volatile int* vcur = a.counters + 1;
	volatile unsigned* vhead = (volatile unsigned*)a.counters + kCntHead;
	volatile int* vcount = a.counters + kCntCount;
	int given_up = 0;

	for (;;) {   // one warp per brick, no barrier wider than the warp anywhere in here
		int b = -1;
		if (!TIMED) {
			// one FIFO ring: a warp claims the next position and waits there; waiting warps line up behind the tail, whoever
			// queues a brick hands it to the first of them
			if (lane == 0) {
				// bounded run: once `budget` ring positions have been claimed nobody claims another one; the bricks still
				// flagged as queued are carried into the next relaxation by the host (flag[] is authoritative, the ring is rebuilt)
				if (!LINKED && a.budget && (vhead[0] >= a.budget || vstop[0])) { vstop[0] = 1; b = -2; }
				if (LINKED && vverdict[0]) b = -2;
				const unsigned pos = b == -2 ? 0u : atomicAdd((unsigned*)a.counters + kCntHead, 1u);   // head: claim a ring position
				unsigned nap = 40;
				for (unsigned spins = 0; b != -2; ++spins) {
					b = vq[pos & a.qmask];
					if (b >= 0) { vq[pos & a.qmask] = -1; break; }
					if (LINKED) {
						// an idle rank waits for work from its neighbours until rank 0 has seen every rank idle with nothing in flight
						if (vverdict[0]) { b = -2; break; }
						if ((spins & 255u) == 255u && global_ns() - t_begin > kLinkTimeoutNs) { atomicAdd(a.counters + 3, 1); b = -2; break; }
						if (spins > 32u && nap < 1000u) nap += nap >> 2;          // long waits poll less often (<= 1 us)
					} else {
						if (*vpending == 0) { b = -2; break; }                       // nothing queued, nobody working: done
						if (a.budget && vstop[0]) { b = -2; break; }                 // bounded run over: the producers have left
						// a warp that found nothing for seconds retires (never spins forever, whatever happens);
						// the host reports non-convergence if work was still pending when the last warp left
						if (spins > (1u << 25)) { b = -2; break; }
					}
					__nanosleep(nap);
				}
			}
			b = __shfl_sync(kFull, b, 0);
		} else {
			// time buckets: take a brick from the earliest bucket that holds one.  Lane i looks at bucket cur - kBucketsBehind + i:
			// the buckets behind cur catch bricks that were queued with a `cur` read just before the window moved on, new
			// bricks go into [cur, cur + kBucketsAhead), the rings in between are only ever reached by bricks queued with a
			// very stale `cur` (taken like any other, but they do not move the window).
			unsigned nap = 40;
			for (unsigned spins = 0; b == -1; ++spins) {
				int quit = 0;
				if (lane == 0) {
					if (!LINKED && a.budget && (*(volatile unsigned*)a.counters >= a.budget || vstop[0])) { vstop[0] = 1; quit = 1; }   // bounded run
					if (LINKED && vverdict[0]) quit = 1;
				}
				const int cur = vcur[0];
				const int bk = cur - kBucketsBehind + lane;
				bool avail = false;
				if (lane < kBrickBuckets && bk >= 0) avail = vcount[bk & (kBrickBuckets - 1)] > 0;
				const unsigned have = __ballot_sync(kFull, avail);
				quit = __shfl_sync(kFull, quit, 0);
				const int cur0 = __shfl_sync(kFull, cur, 0);
				if (quit) { b = -2; break; }
				if (have) {
					const int sel = __ffs(have) - 1;
					const int sbk = __shfl_sync(kFull, bk, sel), r = sbk & (kBrickBuckets - 1);
					int v = -1;
					if (lane == 0) {
						if (atomicSub(a.counters + kCntCount + r, 1) <= 0) {   // others took what there was
							atomicAdd(a.counters + kCntCount + r, 1);
							++given_up;
						} else {
							const unsigned pos = atomicAdd((unsigned*)a.counters + kCntHead + r, 1u);   // ours: a position of that ring
							volatile int* slot = vq + (size_t)r * (a.qmask + 1u) + (pos & a.qmask);
							// as many bricks have been published as positions handed out, ours is there or about to be (a producer
							// with an earlier reservation may publish after a later one); never wait forever all the same
							for (unsigned w = 0; (v = *slot) < 0 && w < (1u << 22); ++w) __nanosleep(20);
							if (v >= 0) {
								*slot = -1;
								// nothing earlier is queued: the window moves on, one bucket per brick taken (a brick queued with a very
								// stale `cur` shows up in a ring of the future; it must not drag the window along)
								if (sbk > cur0 && sbk < cur0 + kBucketsAhead) atomicMax(a.counters + 1, cur0 + 1);
								if (!LINKED && a.budget) atomicAdd((unsigned*)a.counters, 1u);
							} else {
								atomicAdd(a.counters + 3, 1);   // the queue is broken: give up, the host reports it
								v = -2;
							}
						}
					}
					b = __shfl_sync(kFull, v, 0);
					continue;   // b == -1: look again
				}
				int idle_quit = 0;
				if (lane == 0) {
					if (LINKED) {
						if ((spins & 255u) == 255u && global_ns() - t_begin > kLinkTimeoutNs) { atomicAdd(a.counters + 3, 1); idle_quit = 1; }
					} else {
						if (*vpending == 0) idle_quit = 1;                  // nothing queued, nobody working: done
						if (a.budget && vstop[0]) idle_quit = 1;            // bounded run over: the producers have left
						if (spins > (1u << 22)) idle_quit = 1;              // never spin forever (the host reports non-convergence)
					}
				}
				if (__shfl_sync(kFull, idle_quit, 0)) { b = -2; break; }
				if (spins > 32u && nap < 2000u) nap += nap >> 2;            // long waits poll less often (<= 2 us)
				__nanosleep(nap);
			}
		}
		if (b < 0) break;
		int first_i = 0;   // start bricks: reached voxels count as changed on the first visit
		if (lane == 0) {
			first_i = a.first_visit[b];
			if (first_i) a.first_visit[b] = 0;
			atomicExch(a.flag + b, 0);   // from here on, a neighbour that improves our halo queues us again
			__threadfence();
			++visits;
		}
		const bool first = __shfl_sync(kFull, first_i, 0) != 0;
		const uint32_t origin = __ldg(a.origin + b);
#pragma unroll
		for (int j = 0; j < kStage; ++j) {
			const int i = lane + 32 * j;
			if (i < kBrickCells) {
				const uint32_t p = origin + (uint32_t)cell_off[j];
				const int lc = __ldg(a.layer + p);
				s_l[i] = (uint8_t)lc;
				s_o[i] = (uint16_t)(lc * nl3 * 8);
				s_t[i] = __ldcg(a.time + p);
			}
		}
		__syncwarp();
		double t_init[kOwn];
		int lv[kOwn];
#pragma unroll
		for (int o = 0; o < kOwn; ++o) { t_init[o] = s_t[loc[o]]; lv[o] = s_l[loc[o]]; }
		// relax the brick to its local fixed point: in-place sweeps, values only decrease
		for (int it = 0; it < kInnerCap; ++it) {
			bool ch = false;
#pragma unroll
			for (int o = 0; o < kOwn; ++o) {
				if (lv[o] == 0) continue;
				// branch-free: empty cells have layer 0 whose weight row is +inf, unreached cells hold +inf;
				// all NBR loads are independent so their latencies overlap
				const char* __restrict__ wv = (const char*)(wt + lv[o] * 3);
				const double tv = s_t[loc[o]];
				double best = tv;
#pragma unroll
				for (int k = 0; k < NBR; ++k) {
					// neighbour cell = ours - dif_k; its time, the byte offset of its weight row, the weight w[its layer][ours][|dif|^2]
					const double tu = (s_t + loc[o])[-nb_loff<NBR>(k)];
					const int wo = (s_o + loc[o])[-nb_loff<NBR>(k)];
					const double cand = __dadd_rn(tu, *(const double*)(wv + wo + nb_sq<NBR>(k) * 8));
					best = cand < best ? cand : best;   // (fmin's NaN rules cost four more instructions; there are no NaNs here)
				}
				if (best < tv) { s_t[loc[o]] = best; ch = true; }
			}
			__syncwarp();
			++sweeps;
			if (!__any_sync(kFull, ch)) break;
		}
		// Write back what improved (atomicMin on the bit pattern: positive doubles order like integers, and
		// two warps may hold the same brick at once).  Queue a neighbouring brick only if one of ITS cells
		// (our halo copy of it, never smaller than its current value) would improve through a changed voxel.
		unsigned bits = 0, reach_dn = 0, reach_up = 0;
		float kmin = 3.0e38f;
		int vz0 = 0;
		if (LINKED) vz0 = (int)(origin / a.link.plane) - 1;   // voxel plane of the brick's first cell
#pragma unroll
		for (int o = 0; o < kOwn; ++o) {
			const double tf = s_t[loc[o]];
			if (lv[o] == 0) continue;
			const bool improved = tf < t_init[o];
			if (improved) {
				const uint32_t p = origin + (uint32_t)((cz[o] * a.pY + cy[o]) * a.pX + cx[o]);
				atomicMin((unsigned long long*)(a.time + p), (unsigned long long)__double_as_longlong(tf));
				if (LINKED) {
					// our first / last own plane is the halo of the rank below / above: the value goes straight into its grid, and
					// ITS bricks around the cell get a visit (we cannot test its cells' values from here, so all that can see it)
					const int z = vz0 + cz[o];
					const bool dn = z == a.link.z_first && a.link.time_dn != nullptr, up = z == a.link.z_last && a.link.time_up != nullptr;
					if (dn | up) {
						const unsigned r27 = reach27(cz[o], cy[o], cx[o]);
						if (dn) { atomicMin_system((unsigned long long*)(a.link.time_dn + p), (unsigned long long)__double_as_longlong(tf)); reach_dn |= r27; ++remote_cells; }
						if (up) { atomicMin_system((unsigned long long*)(a.link.time_up + p), (unsigned long long)__double_as_longlong(tf)); reach_up |= r27; ++remote_cells; }
					}
				}
			}
			const bool on_face = cz[o] == 0 || cz[o] == kBrick - 1 || cy[o] == 0 || cy[o] == kBrick - 1 || cx[o] == 0 || cx[o] == kBrick - 1;
			if (!on_face || !(improved || (first && tf < inf))) continue;
			if (TIMED) kmin = fminf(kmin, (float)tf);
			// (rolled: this runs once per visit, and unrolled 2 x 26 times it was 45 of the kernel's 64 KB of code -- the top
			// stall of the kernel was no_instruction, warps waiting for the instruction cache)
#pragma unroll 1
			for (unsigned todo = outward[o]; todo; todo &= todo - 1u) {   // the neighbours of this voxel that lie in another brick
				const int k = __ffs(todo) - 1;
				const int qq = loc[o] - a.loff[k];
				const int lu = s_l[qq];
				if (lu == 0) continue;
				const double cand = __dadd_rn(tf, wt[(lv[o] * a.nl1 + lu) * 3 + a.sq[k]]);  // we excite it: T[ours][its]
				if (cand < s_t[qq]) {
					const int nz = cz[o] - a.dz[k], ny = cy[o] - a.dy[k], nx = cx[o] - a.dx[k];  // neighbour = index - dif
					const int dz = nz < 0 ? -1 : nz >= kBrick ? 1 : 0;
					const int dy = ny < 0 ? -1 : ny >= kBrick ? 1 : 0;
					const int dx = nx < 0 ? -1 : nx >= kBrick ? 1 : 0;
					const int id = (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1);
					bits |= 1u << (id > 13 ? id - 1 : id);
				}
			}
		}
		bits = __reduce_or_sync(kFull, bits);
		// the time bucket everything this visit queues goes into: that of the earliest voxel it changed
		float key = 0.f;
		if (TIMED) key = __uint_as_float(__reduce_min_sync(kFull, __float_as_uint(kmin)));   // positive floats order like their bits
		if (LINKED) {
			reach_dn = __reduce_or_sync(kFull, reach_dn);
			reach_up = __reduce_or_sync(kFull, reach_up);
			if (reach_dn | reach_up) {
				__threadfence_system();   // our times are in the neighbours' grids before their bricks are told to look
				__syncwarp();
				if (reach_dn) link_push(a, a.link.state_dn, reach_dn, kOwnBelow, b, lane, key);
				if (reach_up) link_push(a, a.link.state_up, reach_up, kOwnAbove, b, lane, key);
			}
		}
		// our improved times are visible before anyone is told to look at them -- also before the flag exchange below: a
		// neighbour that is queued already reads our times after ITS flag exchange.  (Nobody to tell: no fence.)
		if (bits) __threadfence();
		// push the neighbours that need a visit: one tail reservation and one pending update per warp
		// (head, tail and pending are single hot addresses; per-brick atomics on them would serialise)
		int nb = -1;
		if (lane < 26 && ((bits >> lane) & 1u)) {
			nb = __ldg(a.nbr + (size_t)b * 26 + lane);
			if (nb >= 0 && a.own && !(a.own[nb] & kOwnMe)) nb = -1;    // sharded run: another rank relaxes that brick
			if (nb >= 0 && atomicExch(a.flag + nb, 1) != 0) nb = -1;   // already queued
		}
		const unsigned pushers = __ballot_sync(kFull, nb >= 0);
		const int n_push = __popc(pushers);
		int bucket = 0;
		if (lane == 0) {
			// (linked run: this is where our brick stops counting as work -- after everything it caused elsewhere is accounted for)
			if (n_push != 1) atomicAdd(a.counters + 2, n_push - 1);        // pending += pushes - (this brick done)
			if (TIMED && n_push) {
				const int cur = vcur[0];
				bucket = min(max((int)(key * a.inv_delta), cur), cur + kBucketsAhead - 1);
			}
		}
		if (n_push) {
			if (TIMED) bucket = __shfl_sync(kFull, bucket, 0);
			ring_publish<false>(a.queue, a.counters, a.qmask, bucket, nb, lane);
		}
	}
	if (lane == 0 && visits) { atomicAdd(a.counters + 4, visits); atomicAdd(a.counters + 5, sweeps); if (given_up) atomicAdd(a.counters + 11, given_up); }
	if (LINKED) {
		remote_cells = __reduce_add_sync(kFull, remote_cells);
		if (lane == 0 && remote_cells) atomicAdd(a.counters + 10, remote_cells);
	}
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
	// The brick frontier is the default (model_24: 1.2 ms vs 3.4 ms for sweeps; 4x heart: 40 ms vs 417 ms);
	// EKGSIM_B200_AUTOMATON=sweep selects the plain sweeps as a cross-check.
	const char* sel = getenv("EKGSIM_B200_AUTOMATON");
	bool sweep = false;
	if (sel && std::string(sel) == "sweep") sweep = true;
	// the frontier kernel stages 16-bit offsets into the (layers + 1)^2 x 3 weight table: up to 51 layers (the reference's
	// models have 24); anything larger takes the plain sweeps
	if ((int64_t)(m->n_layers + 1) * (m->n_layers + 1) * 3 * 8 > 65535) sweep = true;
	init_time_kernel<<<m->sm_count * 4, 256, 0, st>>>(m->d_time_pad, npad);
	EKG_CUDA(cudaGetLastError());

	const int n_starts = (int)m->h_starts.size();
	set_start_kernel<<<(n_starts + 127) / 128, 128, 0, st>>>(m->d_time_pad, m->d_start_pidx, n_starts);
	EKG_CUDA(cudaGetLastError());
	if (!sweep) return run_automaton_bricks(m, sweeps_out);
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
	if (sweeps_out) *sweeps_out = sweeps;
	if (sweeps > m->max_sweeps) return fail(EKG_E_STATE, "activation automaton did not converge");
	return EKG_OK;
}

static int launch_bricks(ekg_model* m, const uint8_t* d_own, int64_t* visits_out, int64_t budget = 0, int64_t* leftover_out = nullptr);

static int64_t ring_capacity(int64_t n) { return brick_ring_capacity(n); }

static bool use_time_buckets(const ekg_model* m);
static int brick_ctas_per_sm(const ekg_model* m) {
	if (const char* e = getenv("EKGSIM_B200_AUTOMATON_CTAS_PER_SM")) return std::max(1, atoi(e));
	return use_time_buckets(m) ? 3 : 2;
}

// Time buckets pay off when the frontier is wider than the machine (then the order of the visits is the queue's choice);
// a small model is bound by the wave's critical path, every queued brick is taken at once, and the FIFO ring's hand-off
// (waiting warps line up behind the tail) has the shorter latency.  EKGSIM_B200_AUTOMATON_QUEUE=timed|fifo overrides.
static bool use_time_buckets(const ekg_model* m) {
	if (!(m->brick_delta > 0.f)) return false;
	if (const char* e = getenv("EKGSIM_B200_AUTOMATON_QUEUE")) {
		if (std::string(e) == "timed") return true;
		if (std::string(e) == "fifo") return false;
	}
	return m->n_bricks >= (int64_t)m->sm_count * 2 * kBrickWarps * 96;   // >= 96 bricks per resident warp (2x heart: 35, FIFO 1.79 ms / buckets 1.98; 4x: 280, 12.3 / 8.75)
}

// start bricks: flagged, first visit pending, queued in ring 0
__global__ void brick_seed_kernel(const int32_t* __restrict__ starts, int n0, int n_live, int* __restrict__ flag, int* __restrict__ first,
                                  int* __restrict__ ring0, int* __restrict__ counters) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n0) {
		const int b = starts[i];
		flag[b] = 1;
		first[b] = 1;
		ring0[i] = b;
	}
	if (i == 0) { counters[kCntTail] = n0; counters[kCntCount] = n0; counters[2] = n0; }
}

static int run_automaton_bricks(ekg_model* m, int64_t* rounds_out) {
	cudaStream_t st = m->stream;
	const int64_t n = m->n_bricks;
	const int64_t cap = ring_capacity(n);
	// state = flag[n] | first_visit[n] | rings[kBrickBuckets][cap] | counters[kBrickCounters], initialised on the device (the
	// rings of a 4x heart are 64 MB); the start bricks go into ring 0
	int* flag = m->d_brick_state;
	int* ring = flag + 2 * n;
	int* counters = ring + kBrickBuckets * cap;
	const int n0 = (int)m->h_start_bricks.size();
	if (n > 0) EKG_CUDA(cudaMemsetAsync(flag, 0, (size_t)(2 * n) * sizeof(int), st));
	EKG_CUDA(cudaMemsetAsync(ring, 0xff, (size_t)((use_time_buckets(m) ? kBrickBuckets : 1) * cap) * sizeof(int), st));   // -1 = empty slot
	EKG_CUDA(cudaMemsetAsync(counters, 0, kBrickCounters * sizeof(int), st));
	brick_seed_kernel<<<(n0 + 127) / 128 + 1, 128, 0, st>>>(m->d_start_bricks, n0, (int)n, flag, flag + n, ring, counters);
	EKG_CUDA(cudaGetLastError());
	return launch_bricks(m, nullptr, rounds_out);
}

template <int NBR, bool LINKED>
static void* pick_brick_kernel2(bool timed, bool wsmem) {
	if (timed) return wsmem ? (void*)automaton_brick_kernel<NBR, LINKED, true, true> : (void*)automaton_brick_kernel<NBR, LINKED, true, false>;
	return wsmem ? (void*)automaton_brick_kernel<NBR, LINKED, false, true> : (void*)automaton_brick_kernel<NBR, LINKED, false, false>;
}
static void* pick_brick_kernel(bool cube, bool linked, bool timed, bool wsmem) {
	if (cube) return linked ? pick_brick_kernel2<26, true>(timed, wsmem) : pick_brick_kernel2<26, false>(timed, wsmem);
	return linked ? pick_brick_kernel2<8, true>(timed, wsmem) : pick_brick_kernel2<8, false>(timed, wsmem);
}

// kernel arguments of the frontier kernel for this model (the link part stays zeroed)
static void fill_brick_args(ekg_model* m, const uint8_t* d_own, int64_t budget, BrickArgs& a) {
	const int64_t n = m->n_bricks;
	const int64_t cap = ring_capacity(n);
	int* flag = m->d_brick_state;
	int* first = flag + n;
	int* ring = first + n;
	int* counters = ring + kBrickBuckets * cap;
	a = BrickArgs{};
	a.layer = m->d_layer_pad; a.time = m->d_time_pad; a.wtab = m->d_wtab;
	a.origin = m->d_brick_origin; a.nbr = m->d_brick_nbr; a.own = d_own;
	a.flag = flag; a.first_visit = first; a.queue = ring; a.counters = counters; a.qmask = (uint32_t)(cap - 1);
	a.n_live = (int32_t)n; a.nl1 = m->n_layers + 1; a.pY = (int32_t)m->pY; a.pX = (int32_t)m->pX;
	a.budget = (uint32_t)std::min<int64_t>(std::max<int64_t>(budget, 0), 0x7fffffff);
	a.inv_delta = use_time_buckets(m) ? 1.0f / m->brick_delta : 0.f;
	NbrTable nb;
	make_nbr_table(m->Z > 1 ? EKG_NBHD_3D8 : EKG_NBHD_2D8, &nb);  // simulator.cpp:251-254
	a.n_nbr = nb.n;
	for (int k = 0; k < nb.n; ++k) {
		a.loff[k] = (nb.dz[k] * kBrickHalo + nb.dy[k]) * kBrickHalo + nb.dx[k];
		a.dz[k] = nb.dz[k]; a.dy[k] = nb.dy[k]; a.dx[k] = nb.dx[k];
		a.sq[k] = nb.dz[k] * nb.dz[k] + nb.dy[k] * nb.dy[k] + nb.dx[k] * nb.dx[k] - 1;
		// the kernel's unrolled sweep hard-codes this enumeration
		const bool same = nb.n == 26 ? (a.loff[k] == nb_loff<26>(k) && a.sq[k] == nb_sq<26>(k)) : (a.loff[k] == nb_loff<8>(k) && a.sq[k] == nb_sq<8>(k));
		if (!same) a.n_nbr = -1;
	}
	const size_t w_bytes = (size_t)a.nl1 * a.nl1 * 3 * sizeof(double);
	a.w_in_smem = w_bytes <= 32 * 1024;
	if ((int64_t)a.nl1 * a.nl1 * 3 * 8 > 65535) a.n_nbr = -2;   // the staged row offsets are 16 bits (up to 52 layers)
}

// the frontier kernel over whatever the ring holds; d_own restricts the pushes (sharded run)
static int launch_bricks(ekg_model* m, const uint8_t* d_own, int64_t* rounds_out, int64_t budget, int64_t* leftover_out) {
	cudaStream_t st = m->stream;
	const int64_t n = m->n_bricks;
	BrickArgs a;
	fill_brick_args(m, d_own, budget, a);
	if (a.n_nbr < 0) return fail(EKG_E_UNSUPPORTED, a.n_nbr == -2 ? "more than 52 layers: use EKGSIM_B200_AUTOMATON=sweep" : "unexpected neighbour enumeration");
	const size_t dyn = a.w_in_smem ? (size_t)a.nl1 * a.nl1 * 3 * sizeof(double) : 0;
	int per_sm = 0;
	const int threads = 32 * kBrickWarps;
	const bool timed = a.inv_delta > 0.f;
	void* kfun = pick_brick_kernel(a.n_nbr == 26, false, timed, a.w_in_smem != 0);
	EKG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)kfun, threads, dyn));
	if (per_sm < 1) return fail(EKG_E_CUDA, "automaton kernel does not fit on the device");
	// every CTA must be resident (waiting warps spin on the ring): never launch more than fit.  The time-bucket kernel (large
	// models, throughput bound) is built for 3 CTAs per SM at 80 registers, the FIFO kernel (small models, bound by the wave's
	// critical path: fewer warps polling the queue, fewer instructions per visit) for 2 at 128
	per_sm = std::min(per_sm, brick_ctas_per_sm(m));
	const int grid = (int)std::min<int64_t>((int64_t)per_sm * m->sm_count, std::max<int64_t>((n + kBrickWarps - 1) / kBrickWarps, 1));
	void* kargs[] = {&a};
	EKG_CUDA(cudaLaunchCooperativeKernel(kfun, dim3(grid), dim3(threads), kargs, dyn, st));
	int hc[kBrickCounters] = {0};
	EKG_CUDA(cudaMemcpyAsync(hc, a.counters, sizeof hc, cudaMemcpyDeviceToHost, st));
	EKG_CUDA(cudaStreamSynchronize(st));
	if (rounds_out) *rounds_out = hc[4];   // brick visits (there are no global rounds in the work-queue scheme)
	m->last_brick_visits = hc[4];
	if (getenv("EKGSIM_B200_DEBUG")) {
		unsigned pushes = 0;
		for (int r = 0; r < kBrickBuckets; ++r) pushes += (unsigned)hc[kCntTail + r];
		fprintf(stderr, "automaton bricks (%s queue, bucket %.3g ms): visits %d inner sweeps %d ring positions %u lost races %d last bucket %d\n",
		        timed ? "time-bucket" : "fifo", timed ? 1.0 / a.inv_delta : 0.0, hc[4], hc[5], pushes, hc[11], hc[1]);
	}
	if (leftover_out) *leftover_out = hc[2];   // bricks still flagged as queued (bounded run)
	else if (hc[2] != 0) return fail(EKG_E_STATE, "activation automaton did not converge");
	return EKG_OK;
}

// ---- z-slab sharded automaton ("automaton on a sharded model") -----------------------------------
// Every rank keeps the whole padded grid but relaxes only the bricks that intersect its z-slab.  Between
// relaxations the ranks exchange the planes next to their slab faces (the 26-neighbourhood reaches one plane
// across); merging is an elementwise minimum, which can only move values towards the least fixed point of
// d(v) = min_u fl(d(u) + w(u,v)) -- the same bits the single-GPU run and the reference's Dijkstra produce --
// and the loop ends when a whole exchange improved nothing anywhere.  Values a rank holds outside its slab are
// upper bounds only (its bricks may straddle the slab face); they are overwritten by the final gather.

// bricks of the slab that can see cell (z, y, x) (voxel coordinates) get marked for the next relaxation
__global__ void shard_merge_kernel(double* __restrict__ time, const double* __restrict__ src, int64_t n_cells, int64_t first_cell,
                                   int pY, int pX, int Z, int Y, int X, const int32_t* __restrict__ bindex, int bZ, int bY, int bX,
                                   const uint8_t* __restrict__ own, int* __restrict__ mark, unsigned long long* __restrict__ improved) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_cells) return;
	const int64_t p = first_cell + i;
	const double v = src[i];
	if (!(v < time[p])) return;
	time[p] = v;
	atomicAdd(improved, 1ull);
	const int z = (int)(p / ((int64_t)pY * pX)) - 1, y = (int)((p / pX) % pY) - 1, x = (int)(p % pX) - 1;
	for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
		const int cz = z + dz, cy = y + dy, cx = x + dx;
		if (cz < 0 || cz >= Z || cy < 0 || cy >= Y || cx < 0 || cx >= X) continue;
		const int32_t b = bindex[((int64_t)(cz / kBrick) * bY + cy / kBrick) * bX + cx / kBrick];
		if (b >= 0 && (own[b] & kOwnMe)) mark[b] = 1;
	}
}

// marked bricks -> ring (as "first visits": their reached voxels count as changed, whoever changed them)
__global__ void shard_enqueue_kernel(int* __restrict__ mark, int n_live, int* __restrict__ flag, int* __restrict__ first,
                                     int* __restrict__ ring, int* __restrict__ counters, uint32_t qmask) {
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n_live || !mark[b]) return;
	mark[b] = 0;
	flag[b] = 1;
	first[b] = 1;
	const unsigned pos = atomicAdd((unsigned*)counters + kCntTail, 1u);   // ring 0: the earliest bucket
	ring[pos & qmask] = b;
	atomicAdd(counters + kCntCount, 1);
	atomicAdd(counters + 2, 1);
}

// which ranks relax brick b: ours if it intersects our slab [z0, z1), the rank below / above if it intersects theirs
// ([b0, b1) / [a0, a1), empty ranges without a linked neighbour)
__global__ void shard_own_kernel(const uint32_t* __restrict__ origin, int n_live, int64_t plane, int64_t z0, int64_t z1, int64_t b0, int64_t b1,
                                 int64_t a0, int64_t a1, uint8_t* __restrict__ own, int* __restrict__ mark) {
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n_live) return;
	const int64_t z = (int64_t)origin[b] / plane - 1;   // first voxel plane of the brick
	own[b] = (uint8_t)(((z < z1 && z + kBrick > z0) ? kOwnMe : 0) | ((z < b1 && z + kBrick > b0) ? kOwnBelow : 0) | ((z < a1 && z + kBrick > a0) ? kOwnAbove : 0));
	mark[b] = 0;
}

// our bricks that can see a start voxel (its own brick and those that have it in their halo) are queued at the start
__global__ void shard_mark_starts_kernel(const uint32_t* __restrict__ starts, int n, int pY, int pX, int Z, int Y, int X,
                                         const int32_t* __restrict__ bindex, int bY, int bX, const uint8_t* __restrict__ own, int* __restrict__ mark) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t p = starts[i];
	const int z = (int)(p / ((uint32_t)pY * pX)) - 1, y = (int)((p / pX) % pY) - 1, x = (int)(p % pX) - 1;
	for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
		const int cz = z + dz, cy = y + dy, cx = x + dx;
		if (cz < 0 || cz >= Z || cy < 0 || cy >= Y || cx < 0 || cx >= X) continue;
		const int32_t b = bindex[((int64_t)(cz / kBrick) * bY + cy / kBrick) * bX + cx / kBrick];
		if (b >= 0 && (own[b] & kOwnMe)) mark[b] = 1;
	}
}

static int shard_check(ekg_model* m) {
	if (m->h_starts.empty()) return fail(EKG_E_NO_START, "Could not find starting point for excitation sequence");
	if (m->n_layers >= m->t_cols || m->n_layers >= m->t_rows) return fail(EKG_E_TRANSFER, "transfer (conduction) matrix does not define every layer");
	return EKG_OK;
}

// bricks a bounded relaxation left flagged as queued -> marked for the next one
__global__ void shard_carry_kernel(const int* __restrict__ flag, int* __restrict__ mark, int n_live) {
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b < n_live && flag[b]) mark[b] = 1;
}

// ring, flags and counters rebuilt from the marked bricks (asynchronous on the model's stream)
static int shard_fill_ring(ekg_model* m) {
	cudaStream_t st = m->stream;
	const int64_t n = m->n_bricks;
	const int64_t cap = ring_capacity(n);
	int* flag = m->d_brick_state;
	int* ring = flag + 2 * n;
	int* counters = ring + kBrickBuckets * cap;
	if (n > 0) {
		shard_carry_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(flag, m->d_brick_mark, (int)n);
		EKG_CUDA(cudaGetLastError());
		EKG_CUDA(cudaMemsetAsync(flag, 0, (size_t)(2 * n) * sizeof(int), st));
	}
	EKG_CUDA(cudaMemsetAsync(ring, 0xff, (size_t)(kBrickBuckets * cap) * sizeof(int), st));   // -1 = empty slot
	EKG_CUDA(cudaMemsetAsync(counters, 0, kBrickCounters * sizeof(int), st));
	if (n > 0) {
		shard_enqueue_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(m->d_brick_mark, (int)n, flag, flag + n, ring, counters, (uint32_t)(cap - 1));
		EKG_CUDA(cudaGetLastError());
	}
	return EKG_OK;
}

int shard_begin(ekg_model* m) {
	int rc = shard_check(m);
	if (rc) return rc;
	cudaStream_t st = m->stream;
	const int64_t n = m->n_bricks;
	const size_t nn = (size_t)std::max<int64_t>(n, 1);
	if (!m->d_brick_own) {
		EKG_CUDA(cudaMalloc(&m->d_brick_own, nn));
		EKG_CUDA(cudaMalloc(&m->d_brick_mark, nn * sizeof(int)));
		EKG_CUDA(cudaMalloc(&m->d_improved, sizeof(unsigned long long)));
	}
	// bricks that intersect the slab [slab_z0, slab_z1) (brick bz covers the voxel planes [4 bz, 4 bz + 4)), and -- linked run --
	// the slabs of the neighbouring ranks
	int64_t nb0 = 0, nb1 = 0, na0 = 0, na1 = 0;
	if (m->link.active) {
		if (m->link.slabs[2 * (size_t)m->link.rank] != m->slab_z0 || m->link.slabs[2 * (size_t)m->link.rank + 1] != m->slab_z1)
			return fail(EKG_E_STATE, "the slab has changed since ekg_model_activation_link");
		if (m->link.below >= 0) { nb0 = m->link.slabs[2 * (size_t)m->link.below]; nb1 = m->link.slabs[2 * (size_t)m->link.below + 1]; }
		if (m->link.above >= 0) { na0 = m->link.slabs[2 * (size_t)m->link.above]; na1 = m->link.slabs[2 * (size_t)m->link.above + 1]; }
	}
	if (n > 0) {
		shard_own_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(m->d_brick_origin, (int)n, m->pY * m->pX, m->slab_z0, m->slab_z1, nb0, nb1, na0, na1,
		                                                        m->d_brick_own, m->d_brick_mark);
		EKG_CUDA(cudaGetLastError());
	}
	const int64_t npad = m->pZ * m->pY * m->pX;
	init_time_kernel<<<m->sm_count * 4, 256, 0, st>>>(m->d_time_pad, npad);
	EKG_CUDA(cudaGetLastError());
	// every rank sets every start voxel (simulator.cpp:263); only the owner's bricks relax from it
	const int n_starts = (int)m->h_starts.size();
	set_start_kernel<<<(n_starts + 127) / 128, 128, 0, st>>>(m->d_time_pad, m->d_start_pidx, n_starts);
	EKG_CUDA(cudaGetLastError());
	if (n > 0) {
		shard_mark_starts_kernel<<<(n_starts + 127) / 128, 128, 0, st>>>(m->d_start_pidx, n_starts, (int)m->pY, (int)m->pX, (int)m->Z, (int)m->Y, (int)m->X,
		                                                                 m->d_brick_index, (int)m->bY, (int)m->bX, m->d_brick_own, m->d_brick_mark);
		EKG_CUDA(cudaGetLastError());
	}
	if (n > 0) EKG_CUDA(cudaMemsetAsync(m->d_brick_state, 0, (size_t)(2 * n) * sizeof(int), st));   // flags: nothing carried over
	m->have_activation = false;
	m->shard_active = true;
	m->link.launched = false;
	// linked run: the neighbours write into our ring as soon as their kernels run, so it is made ready here -- the caller
	// puts a barrier over all ranks between this call and ekg_model_activation_linked_launch
	if (m->link.active) { int rc2 = shard_fill_ring(m); if (rc2) return rc2; }
	EKG_CUDA(cudaStreamSynchronize(st));
	return EKG_OK;
}

// max_visits > 0 bounds the relaxation (the wave is handed to the neighbouring slabs before this slab has reached its
// fixed point, so that the ranks work side by side instead of one after the other); *leftover_out = bricks still queued
int shard_relax(ekg_model* m, int64_t max_visits, int64_t* visits_out, int64_t* leftover_out) {
	if (!m->shard_active) return fail(EKG_E_STATE, "ekg_model_activation_begin has not been called");
	const int64_t n = m->n_bricks;
	if (leftover_out) *leftover_out = 0;
	if (m->merge_pending) { EKG_CUDA(cudaStreamSynchronize(m->merge_stream)); m->merge_pending = false; }   // asynchronous merges have landed
	if (n == 0) { if (visits_out) *visits_out = 0; return EKG_OK; }
	int rc0 = shard_fill_ring(m);
	if (rc0) return rc0;
	int64_t left = 0;
	int rc = launch_bricks(m, m->d_brick_own, visits_out, max_visits, &left);
	if (rc) return rc;
	if (max_visits <= 0 && left != 0) return fail(EKG_E_STATE, "activation automaton did not converge");
	if (leftover_out) *leftover_out = left;
	return EKG_OK;
}

static int plane_range(ekg_model* m, int64_t z_begin, int64_t z_end, int64_t* first_cell, int64_t* n_cells) {
	if (z_begin < -1 || z_end > m->pZ - 1 || z_begin > z_end) return fail(EKG_E_INVALID, "bad plane range");
	*first_cell = (z_begin + 1) * m->pY * m->pX;
	*n_cells = (z_end - z_begin) * m->pY * m->pX;
	return EKG_OK;
}

int shard_export(ekg_model* m, int64_t z_begin, int64_t z_end, double* d_planes, cudaStream_t st) {
	int64_t first = 0, n = 0;
	int rc = plane_range(m, z_begin, z_end, &first, &n);
	if (rc) return rc;
	// asynchronous on the caller's stream (the relaxation that produced the values has completed: shard_relax synchronises
	// the model's stream); a collective the caller enqueues after this on the same stream is ordered behind the copy
	if (n) EKG_CUDA(cudaMemcpyAsync(d_planes, m->d_time_pad + first, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
	return EKG_OK;
}

// d_count != NULL: the number of improved cells is ADDED to that device counter and nothing is read back (asynchronous: the
// driver all-reduces the counter anyway); else it is returned through improved_out (one stream synchronisation)
int shard_merge(ekg_model* m, int64_t z_begin, int64_t z_end, const double* d_planes, int64_t* improved_out, unsigned long long* d_count,
                cudaStream_t st) {
	if (!m->shard_active) return fail(EKG_E_STATE, "ekg_model_activation_begin has not been called");
	int64_t first = 0, n = 0;
	int rc = plane_range(m, z_begin, z_end, &first, &n);
	if (rc) return rc;
	unsigned long long h = 0;
	if (n) {
		unsigned long long* counter = d_count ? d_count : m->d_improved;
		if (!d_count) EKG_CUDA(cudaMemsetAsync(m->d_improved, 0, sizeof(unsigned long long), st));
		shard_merge_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(m->d_time_pad, d_planes, n, first, (int)m->pY, (int)m->pX, (int)m->Z, (int)m->Y,
		                                                         (int)m->X, m->d_brick_index, (int)m->bZ, (int)m->bY, (int)m->bX, m->d_brick_own,
		                                                         m->d_brick_mark, counter);
		EKG_CUDA(cudaGetLastError());
		if (!d_count) {
			EKG_CUDA(cudaMemcpyAsync(&h, m->d_improved, sizeof h, cudaMemcpyDeviceToHost, st));
			EKG_CUDA(cudaStreamSynchronize(st));
		}
	}
	if (improved_out) *improved_out = (int64_t)h;
	return EKG_OK;
}

// ---- peer-linked sharded automaton ------------------------------------------------------------------------------
// The ranks of a z-slab sharded model exchange their slab-face planes from inside the frontier kernel (see
// automaton_brick_kernel<.., LINKED>).  For that every rank maps the other ranks' time grids and brick states: raw
// pointers + cudaDeviceEnablePeerAccess when the handles live in one process (one host thread per device), CUDA IPC
// handles between processes (one process per GPU, the torch.distributed layout).  The exchange needs native 64-bit
// atomics between the devices (NVLink); ekg_model_activation_link refuses anything else and the caller falls back to the
// host-driven rounds (ekg_model_activation_relax / _export / _merge).
struct LinkInfo {
	uint64_t magic;
	int64_t pid;
	int32_t device, ipc_ok;
	int32_t timed, pad;       // work-queue mode (all ranks must agree: they push into each other's queues)
	uint64_t time_ptr, state_ptr;
	int64_t n_bricks, npad;
	cudaIpcMemHandle_t time_h, state_h;
};
static_assert(sizeof(LinkInfo) <= EKG_LINK_INFO_BYTES, "EKG_LINK_INFO_BYTES too small");
constexpr uint64_t kLinkMagic = 0x454b474c494e4b31ull;   // "EKGLINK1"

static int64_t this_pid() { return (int64_t)getpid(); }

// Peer access this library switched on, per (device, peer device), counted over the linked handles of the process: with
// peer access enabled every later cudaMalloc on the device also maps the allocation for the peer (milliseconds each), so
// the last unlink switches it off again.  Access that was already on (NCCL, the application) is left alone.
static std::mutex g_peer_mutex;
static std::map<std::pair<int, int>, int> g_peer_refs;

static cudaError_t peer_acquire(int dev, int peer) {
	std::lock_guard<std::mutex> lock(g_peer_mutex);
	int& refs = g_peer_refs[std::make_pair(dev, peer)];
	if (refs > 0) { ++refs; return cudaSuccess; }
	const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
	if (e == cudaSuccess) { refs = 1; return e; }
	cudaGetLastError();
	g_peer_refs.erase(std::make_pair(dev, peer));
	return e == cudaErrorPeerAccessAlreadyEnabled ? cudaSuccess : e;   // somebody else's: not ours to switch off
}

static void peer_release(int dev, int peer) {
	std::lock_guard<std::mutex> lock(g_peer_mutex);
	auto it = g_peer_refs.find(std::make_pair(dev, peer));
	if (it == g_peer_refs.end()) return;
	if (--it->second == 0) {
		g_peer_refs.erase(it);
		cudaDeviceDisablePeerAccess(peer);
		cudaGetLastError();
	}
}

int shard_link_info(ekg_model* m, void* info_out) {
	LinkInfo li;
	memset(&li, 0, sizeof li);
	li.magic = kLinkMagic;
	li.pid = this_pid();
	li.device = m->device;
	li.time_ptr = (uint64_t)(uintptr_t)m->d_time_pad;
	li.state_ptr = (uint64_t)(uintptr_t)m->d_brick_state;
	li.n_bricks = m->n_bricks;
	li.npad = m->pZ * m->pY * m->pX;
	li.timed = use_time_buckets(m) ? 1 : 0;
	li.ipc_ok = cudaIpcGetMemHandle(&li.time_h, m->d_time_pad) == cudaSuccess && cudaIpcGetMemHandle(&li.state_h, m->d_brick_state) == cudaSuccess;
	if (!li.ipc_ok) cudaGetLastError();   // same-process links do not need the handles
	memset(info_out, 0, EKG_LINK_INFO_BYTES);
	memcpy(info_out, &li, sizeof li);
	return EKG_OK;
}

int shard_unlink(ekg_model* m) {
	if (m->link.launched) { cudaStreamSynchronize(m->stream); m->link.launched = false; }
	for (void* p : m->link.ipc_opened) cudaIpcCloseMemHandle(p);
	for (int peer : m->link.peers_acquired) peer_release(m->device, peer);
	if (m->link.ev0) cudaEventDestroy(m->link.ev0);
	if (m->link.ev1) cudaEventDestroy(m->link.ev1);
	m->link = ekg_model::PeerLink();
	return EKG_OK;
}

int shard_link(ekg_model* m, int rank, int n_ranks, const void* infos, const int64_t* slabs) {
	if (n_ranks < 1 || n_ranks > kMaxLinkRanks || rank < 0 || rank >= n_ranks) return fail(EKG_E_INVALID, "bad rank / number of ranks (at most 16)");
	shard_unlink(m);
	ekg_model::PeerLink L;
	L.rank = rank; L.n_ranks = n_ranks;
	L.slabs.assign(slabs, slabs + 2 * n_ranks);
	for (int r = 0; r < n_ranks; ++r) {
		if (slabs[2 * r] < 0 || slabs[2 * r + 1] > m->Z || slabs[2 * r] > slabs[2 * r + 1]) return fail(EKG_E_INVALID, "bad slab");
		if (r > 0 && slabs[2 * r] != slabs[2 * r - 1]) return fail(EKG_E_INVALID, "slabs must partition [0, Z) in rank order");
	}
	if (slabs[0] != 0 || slabs[2 * n_ranks - 1] != m->Z) return fail(EKG_E_INVALID, "slabs must partition [0, Z) in rank order");
	if (slabs[2 * rank] != m->slab_z0 || slabs[2 * rank + 1] != m->slab_z1) return fail(EKG_E_STATE, "ekg_model_set_slab has not been called with this rank's slab");
	const bool live = slabs[2 * rank + 1] > slabs[2 * rank];
	for (int r = rank - 1; r >= 0 && live; --r) if (slabs[2 * r + 1] > slabs[2 * r]) { L.below = r; break; }
	for (int r = rank + 1; r < n_ranks && live; ++r) if (slabs[2 * r + 1] > slabs[2 * r]) { L.above = r; break; }
	L.time.assign((size_t)n_ranks, nullptr);
	L.state.assign((size_t)n_ranks, nullptr);
	auto bail = [&](int code, const std::string& msg) {
		for (void* p : L.ipc_opened) cudaIpcCloseMemHandle(p);
		for (int peer : L.peers_acquired) peer_release(m->device, peer);
		return fail(code, msg);
	};
	for (int r = 0; r < n_ranks; ++r) {
		LinkInfo p;
		memcpy(&p, (const char*)infos + (size_t)r * EKG_LINK_INFO_BYTES, sizeof p);
		if (p.magic != kLinkMagic) return bail(EKG_E_INVALID, "not a link info record");
		if (p.n_bricks != m->n_bricks || p.npad != m->pZ * m->pY * m->pX) return bail(EKG_E_INVALID, "the ranks hold different models");
		if ((p.timed != 0) != use_time_buckets(m)) return bail(EKG_E_INVALID, "the ranks disagree on the work-queue mode (EKGSIM_B200_AUTOMATON_QUEUE / _DELTA)");
		if (r == rank) { L.time[(size_t)r] = m->d_time_pad; L.state[(size_t)r] = m->d_brick_state; continue; }
		if (p.device == m->device && p.pid == this_pid()) ++L.colocated;
		if (p.device != m->device || p.pid != this_pid()) {
			int can = 0, native = 0;
			int peer_dev = p.device;
			if (p.device != m->device) {
				cudaDeviceCanAccessPeer(&can, m->device, peer_dev);
				if (can) cudaDeviceGetP2PAttribute(&native, cudaDevP2PAttrNativeAtomicSupported, m->device, peer_dev);
				if (!can || !native) return bail(EKG_E_UNSUPPORTED, "no peer access with native atomics between the devices (the linked automaton needs NVLink)");
			}
		}
		if (p.pid == this_pid()) {
			if (p.device != m->device && std::find(L.peers_acquired.begin(), L.peers_acquired.end(), p.device) == L.peers_acquired.end()) {
				const cudaError_t e = peer_acquire(m->device, p.device);
				if (e != cudaSuccess) return bail(EKG_E_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
				L.peers_acquired.push_back(p.device);
			}
			L.time[(size_t)r] = (double*)(uintptr_t)p.time_ptr;
			L.state[(size_t)r] = (int*)(uintptr_t)p.state_ptr;
		} else {
			if (!p.ipc_ok) return bail(EKG_E_UNSUPPORTED, "the peer could not export CUDA IPC handles");
			void *tp = nullptr, *sp = nullptr;
			cudaError_t e = cudaIpcOpenMemHandle(&tp, p.time_h, cudaIpcMemLazyEnablePeerAccess);
			if (e != cudaSuccess) { cudaGetLastError(); return bail(EKG_E_UNSUPPORTED, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); }
			L.ipc_opened.push_back(tp);
			e = cudaIpcOpenMemHandle(&sp, p.state_h, cudaIpcMemLazyEnablePeerAccess);
			if (e != cudaSuccess) { cudaGetLastError(); return bail(EKG_E_UNSUPPORTED, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); }
			L.ipc_opened.push_back(sp);
			L.time[(size_t)r] = (double*)tp;
			L.state[(size_t)r] = (int*)sp;
		}
	}
	L.active = true;
	L.timed = use_time_buckets(m);
	if (cudaEventCreate(&L.ev0) != cudaSuccess || cudaEventCreate(&L.ev1) != cudaSuccess) return bail(EKG_E_CUDA, "cudaEventCreate failed");
	m->link = L;
	return EKG_OK;
}

// Launches the linked frontier kernel and returns at once; all ranks must have passed ekg_model_activation_begin (a
// barrier is the caller's job).  max_ctas > 0 caps the grid: ranks that share ONE device (tests) must all be resident
// at the same time, idle ranks wait for their neighbours inside the kernel.
int shard_linked_launch(ekg_model* m, int max_ctas) {
	if (!m->link.active) return fail(EKG_E_STATE, "ekg_model_activation_link has not been called");
	if (!m->shard_active) return fail(EKG_E_STATE, "ekg_model_activation_begin has not been called");
	if (m->link.launched) return fail(EKG_E_STATE, "the linked automaton is already running");
	cudaStream_t st = m->stream;
	const int64_t n = m->n_bricks;
	const int64_t cap = ring_capacity(n);
	BrickArgs a;
	fill_brick_args(m, m->d_brick_own, 0, a);
	const ekg_model::PeerLink& L = m->link;
	a.link.rank = L.rank; a.link.n_ranks = L.n_ranks;
	a.link.z_first = (int32_t)m->slab_z0; a.link.z_last = (int32_t)m->slab_z1 - 1;
	a.link.plane = (uint32_t)(m->pY * m->pX);
	if (L.below >= 0) { a.link.time_dn = L.time[(size_t)L.below]; a.link.state_dn = L.state[(size_t)L.below]; }
	if (L.above >= 0) { a.link.time_up = L.time[(size_t)L.above]; a.link.state_up = L.state[(size_t)L.above]; }
	if (a.n_nbr < 0) return fail(EKG_E_UNSUPPORTED, a.n_nbr == -2 ? "more than 52 layers: use EKGSIM_B200_AUTOMATON=sweep" : "unexpected neighbour enumeration");
	for (int r = 0; r < L.n_ranks; ++r) a.link.counters_of[r] = L.state[(size_t)r] + 2 * n + kBrickBuckets * cap;
	const size_t dyn = a.w_in_smem ? (size_t)a.nl1 * a.nl1 * 3 * sizeof(double) : 0;
	int per_sm = 0;
	const int threads = 32 * kBrickWarps;
	const bool timed = a.inv_delta > 0.f;
	if (timed != L.timed) return fail(EKG_E_STATE, "the work-queue mode has changed since ekg_model_activation_link");
	void* kfun = pick_brick_kernel(a.n_nbr == 26, true, timed, a.w_in_smem != 0);
	EKG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)kfun, threads, dyn));
	if (per_sm < 1) return fail(EKG_E_CUDA, "automaton kernel does not fit on the device");
	// our own bricks bound the useful grid (a slab is ~1/N of the model); rank 0 gives one warp to the detector
	per_sm = std::min(per_sm, brick_ctas_per_sm(m));
	int64_t grid = std::min<int64_t>((int64_t)per_sm * m->sm_count, std::max<int64_t>((n + kBrickWarps - 1) / kBrickWarps, 1) + 1);
	// ranks of one process on one device: an equal share of the SMs each, so that every rank's kernel is resident
	if (max_ctas <= 0 && L.colocated > 1) max_ctas = std::max(1, per_sm * m->sm_count / L.colocated);
	if (max_ctas > 0) grid = std::min<int64_t>(grid, max_ctas);
	void* kargs[] = {&a};
	EKG_CUDA(cudaEventRecord(L.ev0, st));
	EKG_CUDA(cudaLaunchKernel(kfun, dim3((unsigned)grid), dim3(threads), kargs, dyn, st));
	EKG_CUDA(cudaEventRecord(L.ev1, st));
	m->link.launched = true;
	return EKG_OK;
}

int shard_linked_wait(ekg_model* m, int64_t* visits_out, int64_t* remote_out) {
	if (!m->link.launched) return fail(EKG_E_STATE, "ekg_model_activation_linked_launch has not been called");
	cudaStream_t st = m->stream;
	const int64_t n = m->n_bricks;
	int* counters = m->d_brick_state + 2 * n + kBrickBuckets * ring_capacity(n);
	int hc[kBrickCounters] = {0};
	m->link.launched = false;
	EKG_CUDA(cudaMemcpyAsync(hc, counters, sizeof hc, cudaMemcpyDeviceToHost, st));
	EKG_CUDA(cudaStreamSynchronize(st));
	cudaEventElapsedTime(&m->link.kernel_ms, m->link.ev0, m->link.ev1);
	m->activation_ms = m->link.kernel_ms;
	m->last_brick_visits = hc[4];
	if (visits_out) *visits_out = hc[4];
	if (remote_out) { remote_out[0] = (unsigned)hc[8]; remote_out[1] = (unsigned)hc[9]; remote_out[2] = (unsigned)hc[10]; }
	if (getenv("EKGSIM_B200_DEBUG"))
		fprintf(stderr, "linked automaton rank %d: visits %d inner sweeps %d, bricks queued elsewhere %u / here by others %u, cells written to neighbours %d, last bucket %d, verdict %d, %.3f ms\n",
		        m->link.rank, hc[4], hc[5], (unsigned)hc[8], (unsigned)hc[9], hc[10], hc[1], hc[7], m->link.kernel_ms);
	if (hc[7] != 1 || hc[2] != 0) {
		char buf[160];
		snprintf(buf, sizeof buf, "linked activation automaton did not terminate cleanly (verdict %d, %d bricks pending, %d warps gave up)", hc[7], hc[2], hc[3]);
		return fail(EKG_E_STATE, buf);
	}
	return EKG_OK;
}

// After every rank's kernel has ended (barrier: the caller's job) the slabs of the other ranks are pulled over the links,
// so that every rank holds the whole map like after ekg_model_activation.
int shard_linked_gather(ekg_model* m) {
	if (!m->link.active) return fail(EKG_E_STATE, "ekg_model_activation_link has not been called");
	if (m->link.launched) return fail(EKG_E_STATE, "the linked automaton is still running (ekg_model_activation_linked_wait)");
	const int64_t plane = m->pY * m->pX;
	// rank r starts with rank r + 1: at any moment every source is read by one rank only (its NVLink egress is the limit)
	for (int i = 1; i < m->link.n_ranks; ++i) {
		const int r = (m->link.rank + i) % m->link.n_ranks;
		const int64_t z0 = m->link.slabs[2 * (size_t)r], z1 = m->link.slabs[2 * (size_t)r + 1];
		if (z1 <= z0) continue;
		const size_t off = (size_t)((z0 + 1) * plane), cnt = (size_t)((z1 - z0) * plane);
		EKG_CUDA(cudaMemcpyAsync(m->d_time_pad + off, m->link.time[(size_t)r] + off, cnt * sizeof(double), cudaMemcpyDefault, m->stream));
	}
	EKG_CUDA(cudaStreamSynchronize(m->stream));
	return EKG_OK;
}

}  // namespace ekg
