/* oracle/ekg_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's simulation hot path (synergy-twinning/ekgsim,
 * simlib/simulator.cpp + simlib/Wohlfart.h).  It is the checker for the CUDA path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * it.  The product (ekgsim_b200/) never links or imports anything from oracle/.
 *
 * Pinned: this restatement is validated against outputs of the reference itself, compiled here
 * from /root/reference into oracle/_ref (oracle/Makefile, oracle/ref_dump.cpp); the resulting
 * golden vectors live in tests/golden/ (tests/golden/make_fixtures.py generated them).
 *
 * Conventions shared with the product's C ABI (include/ekgsim_b200.h):
 *   - voxel arrays are raster z,y,x (x fastest), index (z*Y+y)*X+x        (matrix.h:166-173)
 *   - layers[] is uint16: 0 = empty, 1..n = layer, bit 0x1000 = excitation start voxel
 *     (ShapeElement::layerStartingPoint, matrix.h:90,128)
 *   - lead positions are (z,y,x) triples in voxel units                   (simulator.cpp:376-381)
 *   - layer_k is [n_layers][9] WohlfartPlus coefficients                  (Wohlfart.h:167-203)
 *   - neighbourhood ids follow sim_lib.h:133-141 (see EKG_NBHD_* below)
 */
#ifndef EKG_ORACLE_H
#define EKG_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { EKG_ORACLE_NBHD_2D4 = 0, EKG_ORACLE_NBHD_2D8 = 1, EKG_ORACLE_NBHD_3D4 = 2, EKG_ORACLE_NBHD_3D8 = 3 };

/* WohlfartPlus::operator[] (Wohlfart.h:195-203), literal. */
double ekg_oracle_wohlfart_plus(const double k[9], double t);

/* ActionPotential::init + operator() (simulator.cpp:154-170): k8 -= at, evaluate at t - at. */
double ekg_oracle_ap(const double layer_k[9], double at, double t);

/* Neighbour offsets (dz,dy,dx) in the order Neighbourhood::create produces them
 * (simulator.h:334-384).  Returns the count (<= 26); dif must hold 26*3 ints. */
int ekg_oracle_neighbourhood(int nbhd, int dif[26 * 3]);

/* Simulation::calculateExcitationSequence + exciteElement (simulator.cpp:212-286).
 * transfer is [t_rows][t_cols] row-major (row = exciting layer, column = excited layer).
 * delay_out[Z*Y*X]: activation time, 0.0 for empty / never reached voxels.
 * Returns 0, -1 no start voxel, -2 a layer is not covered by the transfer matrix. */
int ekg_oracle_activation(const uint16_t* layers, int64_t Z, int64_t Y, int64_t X,
                          const double* transfer, int64_t t_rows, int64_t t_cols,
                          double* delay_out);

/* Simulation::run (simulator.cpp:452-550), literal loop structure and summation order
 * (raster over voxels, neighbours in create() order, f64 throughout).
 * ecg_out is [n_leads][n_steps], n_steps = ceil(total_time / t_step).  Returns n_steps or <0. */
int64_t ekg_oracle_run_direct(const uint16_t* layers, const double* delay,
                              int64_t Z, int64_t Y, int64_t X,
                              const double* layer_k, int64_t n_layers,
                              const double* leads_zyx, int64_t n_leads, int nbhd,
                              double t_start, double t_step, double total_time,
                              double* ecg_out);

/* Same result up to f64 summation order (<= 1e-12 relative), 50-100x faster: AP evaluated once
 * per distinct (layer, delay) class -- the reference's own setApIndices dedup
 * (simulator.cpp:561-621) -- and lead-field coefficients folded per voxel (SURVEY 7,
 * "algebraic shortcut").  Used for the 256-vector batch and the 4x heart. */
int64_t ekg_oracle_run_factored(const uint16_t* layers, const double* delay,
                                int64_t Z, int64_t Y, int64_t X,
                                const double* layer_k, int64_t n_layers,
                                const double* leads_zyx, int64_t n_leads, int nbhd,
                                double t_start, double t_step, double total_time,
                                double* ecg_out);

/* Partial ECG of the voxels with z in [z0, z1) (see ekg_oracle.c); slabs add up to run_factored. */
int64_t ekg_oracle_run_factored_slab(const uint16_t* layers, const double* delay,
                                     int64_t Z, int64_t Y, int64_t X,
                                     const double* layer_k, int64_t n_layers,
                                     const double* leads_zyx, int64_t n_leads, int nbhd,
                                     double t_start, double t_step, double total_time,
                                     int64_t z0, int64_t z1, double* ecg_out);

/* Simulation::setApIndices (simulator.cpp:561-621): ap_index_out[Z*Y*X] in first-seen raster
 * order (-1 for empty voxels).  Returns the number of classes K. */
int64_t ekg_oracle_ap_classes(const uint16_t* layers, const double* delay,
                              int64_t n_voxels, int64_t n_layers, int64_t* ap_index_out);

/* Simulation::runApproximation (simulator.cpp:552-559). n = (size_t)(total_time / t_step). */
int64_t ekg_oracle_run_approximation(const double* layer_k, int64_t n_layers,
                                     double t_start, double t_step, double total_time,
                                     double delay, double* out);

/* Wohlfart.h:206-223 */
double ekg_oracle_apd90(const double k[9]);

/* Layer-AP construction for ONE parameter vector (sim.cpp:751-916 with WohlfartInterpolationEvaluator
 * sim.cpp:91-313 and steepestDescend nonlinearFit.h:92-168).  border_k [n_border][9] (2: endo, epi;
 * 3: endo, mid, epi with the mid AP in 0-based layer `mid`), out [n_layers][9]. */
int ekg_oracle_fit_layers(const double* border_k, int64_t n_border, int64_t n_layers, int64_t mid,
                          const double d9[9], double step, double eps, int64_t iterations, double* out);

#ifdef __cplusplus
}
#endif
#endif
