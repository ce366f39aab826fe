/* include/ekgsim_b200.h -- C ABI of the B200 EkgSim hot path (libekgsim_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  The host-side
 * C++ facade ekgsim_b200/host/sim_lib.h (class SimLib::EkgSim, same surface as the reference's
 * simlib/sim_lib.h:93-287) is a thin wrapper over these entry points; INTEGRATION.md shows the
 * binding a maintainer of the reference would add.
 *
 * What each entry point replaces in the reference (paths relative to synergy-twinning/ekgsim):
 *
 *   ekg_model_create          Simulation::loadShape + loadTransferMatrix   simulator.cpp:181-206
 *                             (the .matrix text parsing stays on the host: matrix.h:133-248)
 *   ekg_model_activation      Simulation::calculateExcitationSequence      simulator.cpp:248-286
 *                             + exciteElement                              simulator.cpp:212-246
 *   ekg_model_set_activation  Simulation::loadExcitationSequence           simulator.cpp:288-367
 *   ekg_model_get_activation  EkgSim::getModelShape()[..].excitationDelay  sim_lib.h:276-282
 *   ekg_model_ap_classes      Simulation::setApIndices (class table)       simulator.cpp:561-621
 *   ekg_simulate              Simulation::run for B parameter sets at once simulator.cpp:452-550
 *                             (AP evaluation Wohlfart.h:195-203 via simulator.cpp:154-170)
 *   ekg_simulate_device       same, device-resident inputs/outputs on a caller stream
 *   ekg_fit_layers            the layer-AP construction of SimImplementation::simUsingBorderAps /
 *                             simUsingBorderAndMidAps (sim.cpp:751-916): connectors of
 *                             WohlfartInterpolationEvaluator (sim.cpp:91-313) + steepestDescend
 *                             (nonlinearFit.h:92-168) for every inner layer, B vectors at once
 *   ekg_evaluate              fit + ekg_simulate + curve comparison without leaving the device:
 *                             border APs and lead positions in, criteria out (SimImplementation::eval
 *                             sim.cpp:443-491 minus the gene unpacking, which stays on the host)
 *   ekg_model_set_slab,       one model over several GPUs (BASELINE config 4; no counterpart in the reference, whose
 *   ekg_model_activation_*    parallelism is one full model per MPI rank, main.cpp:301): the ECG sum restricted to a z-slab,
 *                             and calculateExcitationSequence (simulator.cpp:248-286) computed by all GPUs together --
 *                             peer-linked over NVLink (_link_info / _link / _linked_launch / _linked_wait / _linked_gather)
 *                             or in host-driven rounds (_begin / _relax_bounded / _export / _merge / _end)
 *
 * Conventions
 *   - voxel arrays are raster z,y,x (x fastest): index (z*Y+y)*X+x          (matrix.h:166-173)
 *   - layers[]: uint16, 0 = empty, 1..n = layer, bit 0x1000 marks an excitation start voxel
 *     (ShapeElement::layerStartingPoint, matrix.h:90,128)
 *   - transfer[t_rows][t_cols] row-major, [exciting layer][excited layer], ms per voxel edge
 *   - lead positions are (z,y,x) triples, voxel units                       (simulator.cpp:376-381)
 *   - layer_k[b][layer][9]: WohlfartPlus coefficients per layer, activation time 0
 *   - ecg[b][lead][step], step i is at time t_start + i*t_step (accumulated), n_steps =
 *     ceil(total_time / t_step)                                             (simulator.cpp:471)
 *   - every function returns EKG_OK (0) or a negative EKG_E_* code; ekg_last_error() returns
 *     the message for the calling thread.  The C++ facade turns codes into std::runtime_error,
 *     the reference's own error convention (main.cpp:405-413).
 *   - a handle is bound to one CUDA device; calls on one handle must not overlap (the reference
 *     is single-threaded per process, sim.cpp:444).  Different handles are independent.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     EKG_E_CUDA.
 */
#ifndef EKGSIM_B200_H
#define EKGSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EKG_ABI_VERSION 3

enum {
	EKG_OK = 0,
	EKG_E_INVALID = -1,    /* bad argument */
	EKG_E_NO_START = -2,   /* "Could not find starting point for excitation sequence" (simulator.cpp:276) */
	EKG_E_TRANSFER = -3,   /* transfer matrix too small / does not define a layer (simulator.cpp:204, :236) */
	EKG_E_STATE = -4,      /* call order: e.g. simulate before an activation map exists */
	EKG_E_CUDA = -5,       /* CUDA runtime error or no device */
	EKG_E_NOMEM = -6,
	EKG_E_UNSUPPORTED = -7 /* grid too large for the packed layout, too many leads, ... */
};

/* neighbourhood ids for the ECG stencil (sim_lib.h:133-141; note 3D4 = the 8 cube corners) */
enum { EKG_NBHD_2D4 = 0, EKG_NBHD_2D8 = 1, EKG_NBHD_3D4 = 2, EKG_NBHD_3D8 = 3 };

/* ECG kernel selection (flags argument of ekg_simulate*) */
enum {
	EKG_MODE_DEFAULT = 0,  /* library picks (currently EKG_MODE_SEPARABLE) */
	EKG_MODE_DIRECT = 1,   /* full 9-coefficient AP evaluated per voxel and time sample */
	EKG_MODE_HOISTED = 2,  /* voxel-invariant and time-invariant factors of the AP hoisted */
	EKG_MODE_SEPARABLE = 3 /* samples later than the last depolarisation: per-layer moments of the lead
	                        * field replace the voxel x sample loop; earlier samples run as HOISTED */
};

/* OR-ed into flags: record CUDA events around the ECG kernel launch(es) on the launching stream
 * so that ekg_last_kernel_ms() can report the dominant kernel's device time (bench.py roofline). */
#define EKG_FLAG_TIME_KERNEL 0x100
/* OR-ed into flags: the SEPARABLE path evaluates the stencil sum of interior voxels (all 8 corners of the "3D4" stencil
 * occupied) by its harmonic series (csrc/ecg.cu, corner_series2); this flag makes it add the 8 corner terms like it does
 * for boundary voxels -- slower and, in fp32, noisier; kept as the independent cross-check of the series. */
#define EKG_FLAG_CORNER_SUM 0x200

typedef struct ekg_model ekg_model;

int         ekg_abi_version(void);
const char* ekg_last_error(void);
int         ekg_device_count(void);               /* 0 when no usable CUDA device */

int  ekg_model_create(const uint16_t* layers, int64_t Z, int64_t Y, int64_t X,
                      const double* transfer, int64_t t_rows, int64_t t_cols,
                      int device, ekg_model** out);
void ekg_model_destroy(ekg_model* m);

/* Restrict the ECG sum to voxels with z in [z_begin, z_end): the z-slab shard of one large
 * model (BASELINE config 4).  Partial ECGs of all slabs add up to the full ECG.  The automaton
 * always runs on the whole model.  Default range is [0, Z). */
int  ekg_model_set_slab(ekg_model* m, int64_t z_begin, int64_t z_end);

int64_t ekg_model_num_voxels(const ekg_model* m);   /* occupied voxels (inside the slab) */
int64_t ekg_model_num_layers(const ekg_model* m);   /* = Simulation::getTargetNumOfAps(), simulator.h:570 */

/* Runs the activation-time automaton on the device; the map stays resident.  delay_out (host,
 * Z*Y*X doubles, 0.0 for empty / unreached voxels) may be NULL: the map is then not copied to the
 * host at all (the ECG entry points only need it in HBM; ekg_model_get_activation and
 * ekg_model_ap_classes fetch the raster copy on demand).  sweeps_out (may be NULL) receives a work
 * count: brick visits of the frontier kernel (sweeps of the cross-check kernel). */
int  ekg_model_activation(ekg_model* m, double* delay_out, int64_t* sweeps_out);
/* Device time (ms, CUDA events) of the last ekg_model_activation call on this handle. */
double ekg_model_activation_ms(const ekg_model* m);
/* Brick visits of the last frontier automaton run (work actually done, in units of 4x4x4 bricks). */
int64_t ekg_model_activation_brick_visits(const ekg_model* m);
int  ekg_model_set_activation(ekg_model* m, const double* delay);

/* The automaton on a model sharded into z-slabs over several GPUs ("automaton on a sharded model").
 * Every rank holds the whole model and has restricted itself to its slab with ekg_model_set_slab; it relaxes only the
 * bricks that intersect the slab.  One round = relax, then exchange the planes at the slab faces with the neighbouring
 * ranks (NCCL / P2P, the caller's job) and merge what arrives by an elementwise minimum; the loop ends when no rank's
 * merge improved anything.  The result has the bits of ekg_model_activation.  ekgsim_b200/dist.py::sharded_activation
 * is the driver; planes are (z, all y, all x) slices of the zero-bordered grid, ekg_model_plane_elems() doubles each,
 * z from -1 (border) to Z.
 *   begin   times = +inf, start voxels = 1 (simulator.cpp:263); the start bricks inside the slab are queued
 *   relax   frontier relaxation from the queued bricks until the slab is at its fixed point for its current halo
 *   relax_bounded   the same, but at most ~max_brick_visits brick visits (0 = no bound); bricks_left_out = bricks still
 *           queued, they are carried into the next relax.  With a bound the wave reaches the neighbouring slabs
 *           before this slab is finished, so the ranks work side by side; the loop then ends when no merge improved
 *           anything AND no rank has bricks left.  Min-merging stale values is harmless: times only ever decrease
 *           towards the least fixed point, the bits are those of the unbounded run.
 *   export  copies planes [z_begin, z_end) into a device buffer, asynchronously on `stream` (work enqueued on the same
 *           stream afterwards -- the collective -- is ordered behind the copy)
 *   merge   time = min(time, planes); bricks of the slab that can see an improved cell are queued for the next
 *           relax; improved_out = number of improved cells (one stream synchronisation)
 *   merge_async   the same without reading anything back: the number of improved cells is ADDED to the caller's
 *           device counter *d_improved_accum (which the driver all-reduces over the ranks anyway); the next relax /
 *           end waits for the stream
 *   end     publishes the map like ekg_model_activation does (range of the times, ECG voxel list; host copy only if
 *           delay_out is not NULL) */
int     ekg_model_activation_begin(ekg_model* m);
int     ekg_model_activation_relax(ekg_model* m, int64_t* brick_visits_out);
int     ekg_model_activation_relax_bounded(ekg_model* m, int64_t max_brick_visits, int64_t* brick_visits_out, int64_t* bricks_left_out);
int64_t ekg_model_plane_elems(const ekg_model* m);
int     ekg_model_activation_export(ekg_model* m, int64_t z_begin, int64_t z_end, double* d_planes, void* stream);
int     ekg_model_activation_merge(ekg_model* m, int64_t z_begin, int64_t z_end, const double* d_planes,
                                   int64_t* improved_out, void* stream);
int     ekg_model_activation_merge_async(ekg_model* m, int64_t z_begin, int64_t z_end, const double* d_planes,
                                         uint64_t* d_improved_accum, void* stream);
int     ekg_model_activation_end(ekg_model* m, double* delay_out);

/* The same sharded automaton without host rounds ("peer-linked"): every rank maps the neighbouring ranks' grids (NVLink
 * peer memory: raw pointers inside one process, CUDA IPC between processes) and ONE kernel launch per rank does the whole
 * run -- a warp that improves a voxel on the first / last plane of its slab also writes it into the neighbour's grid and
 * queues the neighbour's bricks around it in the neighbour's work ring (system-scope atomics), idle ranks wait inside the
 * kernel, rank 0 detects global termination from all ranks' message counters.  No collective and no host synchronisation
 * on the path; the bits are those of ekg_model_activation.  Needs native atomics between the devices (NVLink);
 * ekg_model_activation_link returns EKG_E_UNSUPPORTED otherwise and the rounds above remain.
 *   link_info       fills EKG_LINK_INFO_BYTES bytes describing this handle (exchange them between the ranks, rank order)
 *   link            rank, number of ranks, all ranks' records, all ranks' slabs [n_ranks][2] (a partition of [0, Z) in rank
 *                   order; this rank's must be what ekg_model_set_slab set).  Ranks with an empty slab take part, idle.
 *   begin           ekg_model_activation_begin as above (also prepares the work ring); THEN A BARRIER OVER ALL RANKS
 *   linked_launch   starts the kernel and returns.  max_ctas > 0 caps the grid; 0 = the device's capacity, divided by the
 *                   number of ranks of this process on this device (all of them must be resident at the same time)
 *   linked_wait     waits for it; brick_visits_out = this rank's brick visits; remote_out[3] (may be NULL) = bricks queued at
 *                   other ranks, bricks other ranks queued here, cells written into other ranks' grids
 *   linked_gather   after a barrier over all ranks: copies the other ranks' slabs over the links, so that every rank
 *                   holds the whole map (optional: the ECG of a slab only needs the slab); then a barrier before the next begin
 *   end             ekg_model_activation_end as above
 *   unlink          releases the mappings (also done by ekg_model_destroy) */
#define EKG_LINK_INFO_BYTES 256
int     ekg_model_activation_link_info(ekg_model* m, void* info_out);
int     ekg_model_activation_link(ekg_model* m, int rank, int n_ranks, const void* infos, const int64_t* slabs);
int     ekg_model_activation_linked_launch(ekg_model* m, int max_ctas);
int     ekg_model_activation_linked_wait(ekg_model* m, int64_t* brick_visits_out, int64_t* remote_out);
int     ekg_model_activation_linked_gather(ekg_model* m);
int     ekg_model_activation_unlink(ekg_model* m);
int  ekg_model_get_activation(const ekg_model* m, double* delay_out);

/* (layer, delay) class table in first-seen raster order: ap_index_out[Z*Y*X] (-1 = empty);
 * returns K through n_classes_out.  Host-side bookkeeping only (EkgSim::getAps ordering). */
int  ekg_model_ap_classes(const ekg_model* m, int64_t* ap_index_out, int64_t* n_classes_out);

/* B simulations against the resident model.  Host buffers; blocking. */
int  ekg_simulate(ekg_model* m, const double* layer_k, const double* leads_zyx,
                  int64_t B, int64_t n_leads, int nbhd,
                  double t_start, double t_step, double total_time,
                  int flags, double* ecg_out);

/* Same with device pointers (on the model's device), enqueued on the caller's stream (a
 * cudaStream_t; NULL = CUDA's default stream).  Asynchronous with respect to the host: the ECGs are
 * ready when work enqueued on that stream after this call runs. */
int  ekg_simulate_device(ekg_model* m, const double* d_layer_k, const double* d_leads_zyx,
                         int64_t B, int64_t n_leads, int nbhd,
                         double t_start, double t_step, double total_time,
                         int flags, double* d_ecg_out, void* stream);

/* ekg_simulate_device for a caller who knows the batch's coefficients on the host (it uploaded them): k1_min = the
 * smallest depolarisation rate k1, decay_max = the largest of |k4 + k5| and |k5|, over all (vector, layer).  The default /
 * SEPARABLE mode needs both to decide from which sample on the sigmoid is saturated and whether the hoisted exponentials
 * stay in range; without them ekg_simulate_device reads them back from the device (one stream synchronisation in the
 * middle of the call).  With them the call is fully asynchronous -- what ekg_simulate (host buffers) does internally.
 * Hints that do not bound the batch give wrong results. */
int  ekg_simulate_device_hinted(ekg_model* m, const double* d_layer_k, const double* d_leads_zyx,
                                int64_t B, int64_t n_leads, int nbhd,
                                double t_start, double t_step, double total_time,
                                int flags, double k1_min, double decay_max, double* d_ecg_out, void* stream);

/* ekg_simulate + the reference's curve comparison on the device, so that only B x n_leads criteria
 * have to leave the GPU: criteria_out[b][l] compares ECG[b][l][0..n) with
 * targets[l][0..n), n = min(n_steps, n_target), offset 0, like calculateFitness does (sim.cpp:600-702):
 *   comparison 1 = RMS of the difference                         (vectorMath.h:322-343)
 *   comparison 2 = 1 - Pearson correlation                       (vectorMath.h:287-317)
 *   comparison 3 = deviation from linear, needs target_offsets   (vectorMath.h:372-404)
 *   comparison 4 = 1 - vector correlation                        (vectorMath.h:348-366)
 * targets is [n_leads][n_target] (host), already normalised/resampled by the caller (loadTargets);
 * target_offsets [n_leads] may be NULL unless comparison == 3; ecg_out (host) may be NULL. */
int  ekg_simulate_criteria(ekg_model* m, const double* layer_k, const double* leads_zyx,
                           int64_t B, int64_t n_leads, int nbhd,
                           double t_start, double t_step, double total_time, int flags,
                           const double* targets, int64_t n_target, const double* target_offsets,
                           int comparison, double* criteria_out, double* ecg_out);

/* Layer-AP construction for B parameter vectors.  border_k is [B][n_border][9]:
 * n_border = 2 -> (endo, epi), every inner layer interpolated between them (sim.cpp:751-821);
 * n_border = 3 -> (endo, mid, epi) with the mid AP sitting in layer index `mid` (0-based,
 * 0 < mid < n_layers-1; sim.cpp:825-916, absoluteMidPos :833).  d9 are the nine gradient offsets /
 * initial moves (sim.cpp:877 `kd`; 0 = coefficient not fitted), step_size / epsilon / iterations the
 * arguments of steepestDescend (sim.cpp:901: 0.5, 1e-3, 100).  layer_k_out is [B][n_layers][9] with
 * n_layers = ekg_model_num_layers(m), directly usable as the layer_k of ekg_simulate.  f64 on the device,
 * the reference's expression order; integer decisions (sample indices, accept/reject) make the result
 * insensitive to the last-bit differences between CUDA's and glibc's exp/log/pow. */
int  ekg_fit_layers(ekg_model* m, const double* border_k, int64_t B, int64_t n_border, int64_t mid,
                    const double* d9, double step_size, double epsilon, int64_t iterations,
                    double* layer_k_out);
/* Same with device pointers, enqueued on the caller's stream. */
int  ekg_fit_layers_device(ekg_model* m, const double* d_border_k, int64_t B, int64_t n_border, int64_t mid,
                           const double* d9, double step_size, double epsilon, int64_t iterations,
                           double* d_layer_k_out, void* stream);

/* ekg_fit_layers + ekg_simulate_criteria in one call, intermediate results stay in HBM: per vector
 * 27 (or 18) border coefficients and the lead positions go in, n_leads criteria come out.
 * layer_k_out ([B][n_layers][9]) and ecg_out ([B][n_leads][n_steps]) may be NULL; so may criteria_out
 * (targets are then ignored) as long as one output is requested. */
int  ekg_evaluate(ekg_model* m, const double* border_k, int64_t n_border, int64_t mid,
                  const double* d9, double step_size, double epsilon, int64_t iterations,
                  const double* leads_zyx, int64_t B, int64_t n_leads, int nbhd,
                  double t_start, double t_step, double total_time, int flags,
                  const double* targets, int64_t n_target, const double* target_offsets,
                  int comparison, double* criteria_out, double* layer_k_out, double* ecg_out);

/* Number of kernel launches the last ekg_simulate* / ekg_fit* / ekg_evaluate call on this handle issued. */
int64_t ekg_last_launch_count(const ekg_model* m);
/* Device time (ms) of the ECG kernel launch(es) of the last call made with EKG_FLAG_TIME_KERNEL;
 * synchronises on the recorded events.  Negative if nothing was recorded. */
double ekg_last_kernel_ms(ekg_model* m);
/* Name of the ECG kernel variant the last ekg_simulate* call used (static string). */
const char* ekg_last_kernel_name(const ekg_model* m);

#ifdef __cplusplus
}
#endif
#endif /* EKGSIM_B200_H */
