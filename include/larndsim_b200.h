/*
 * larndsim_b200.h -- C ABI of the B200-native larnd-sim charge/light readout chain.
 *
 * This is the drop-in boundary.  Every entry point replaces one
 * Numba `@cuda.jit` kernel (or one piece of CuPy glue) of the reference; the reference
 * file:line it replaces is cited on each declaration (paths relative to the reference
 * tree, e.g. larndsim/detsim.py).  The Python host layer in `larndsim_b200/` binds these
 * with ctypes and re-creates the reference's `kernel[grid, block](*arrays)` call
 * surface on top; INTEGRATION.md shows the binding a larnd-sim maintainer would add.
 *
 * Conventions
 *  - plain pointers + sizes, POD structs, no C++/torch types;
 *  - every array pointer is a DEVICE pointer unless the name ends in `_host`;
 *    arrays are C-contiguous with exactly the dtypes the reference CLI passes
 *    (cli/simulate_pixels.py:930-1087): pixel ids / radii int32, maps int64,
 *    accumulators float64, induced-current `signals` float32;
 *  - outputs are caller-allocated and caller-initialised (-1 / 0 fills); kernels only
 *    write or accumulate in place, exactly like the reference kernels;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, which is
 *    what Numba and CuPy use in the reference);
 *  - return value: 0 ok, <0 argument error, >0 cudaError_t.  Never throws.
 *    `lsb_last_error()` returns a static description of the last failure.
 *  - detector/physics/light/simulation constants are read by the host layer from
 *    `larndsim.consts.*` AT CALL TIME and passed in `lsb_consts` (the reference freezes
 *    them into the JIT: cli/simulate_pixels.py:459-464).
 */
#ifndef LARNDSIM_B200_H
#define LARNDSIM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSB_ABI_VERSION 2
#define LSB_MAX_TPC 128

/* dtype codes for record fields */
enum lsb_dtype {
    LSB_NONE = 0, LSB_F32 = 1, LSB_F64 = 2, LSB_I32 = 3, LSB_U32 = 4, LSB_I64 = 5, LSB_U64 = 6
};

/* segment-record fields the chain touches (cli/dumpTree.py:17-28) */
enum lsb_field {
    LSB_F_X = 0, LSB_F_Y, LSB_F_Z,
    LSB_F_X_START, LSB_F_Y_START, LSB_F_Z_START,
    LSB_F_X_END, LSB_F_Y_END, LSB_F_Z_END,
    LSB_F_T, LSB_F_T_START, LSB_F_T_END,
    LSB_F_T0, LSB_F_T0_START, LSB_F_T0_END,
    LSB_F_DEDX, LSB_F_DE, LSB_F_DX,
    LSB_F_N_ELECTRONS, LSB_F_N_PHOTONS,
    LSB_F_LONG_DIFF, LSB_F_TRAN_DIFF,
    LSB_F_PIXEL_PLANE,
    LSB_F_COUNT
};

/* Layout of one structured-array record, resolved by the host layer from
 * `arr.dtype.fields` at call time (tests use an all-f8 record, production a 152-byte
 * f4/u4 record: tests/testQuenching.py:16-19 vs cli/dumpTree.py:17-28). offset<0: absent. */
typedef struct lsb_track_layout {
    int32_t itemsize;
    int32_t offset[LSB_F_COUNT];
    int32_t dtype[LSB_F_COUNT];
} lsb_track_layout;

/* Flat snapshot of larndsim.consts.{physics,detector,light,sim,units} and
 * pixels_from_track.MAX_NEIGHBOR_BACKTRACK_DISTANCE. */
typedef struct lsb_consts {
    /* consts/physics.py:7-17 */
    double box_alpha, box_beta, birks_ab, birks_kb, w_ion;
    int32_t mode_box, mode_birks;
    /* consts/detector.py:19-67 */
    double e_field, lar_density, v_drift, electron_lifetime, long_diff, tran_diff;
    double time_sampling, time_padding, time_window, time_interval[2];
    double response_sampling, response_bin_size, pixel_pitch;
    int32_t n_time_ticks;          /* len(detector.TIME_TICKS) */
    int32_t n_pixels[2];
    int32_t n_tpc;                 /* TPC_BORDERS.shape[0] */
    int32_t default_plane_index;   /* 0xBEEF */
    int32_t sampled_points;
    int32_t max_neighbor_backtrack_distance; /* pixels_from_track.py:11 */
    int32_t pad0_;
    /* consts/detector.py:92-131 (front-end electronics) */
    double discrimination_threshold, adc_hold_delay, adc_busy_delay, reset_cycles, clock_cycle;
    double gain, buffer_risetime, v_cm, v_ref, v_pedestal, adc_counts;
    double reset_noise_charge, uncorrelated_noise_charge, discriminator_noise;
    /* consts/units.py */
    double unit_e, unit_mV, unit_ns, unit_mus;
    /* consts/light.py:8-61 */
    double w_ph, scint_prescale, light_tick_size, light_window[2];
    double singlet_fraction, tau_s, tau_t;
    double light_response_time, light_oscillation_period, impulse_tick_size;
    int32_t sipm_response_model, light_trig_mode, enable_lut_smearing, n_op_channel;
    /* consts/sim.py:26-39 */
    double min_step_size, mc_truth_threshold;
    int32_t max_tracks_per_pixel, mc_sample_multiplier, max_adc_values, pad1_;
    /* consts/detector.py:329-345: TPC_BORDERS[n_tpc][3][2] (x,y,z) x (lo,hi); z[0] = anode */
    double tpc_borders[LSB_MAX_TPC][3][2];
} lsb_consts;

/* Layout of the light LUT record (lightLUT.py:104-114, light_sim.py:97,116):
 * lut[nx][ny][nz][ndet_tpc] of records with float32 fields. */
typedef struct lsb_lut_layout {
    int32_t itemsize;
    int32_t off_vis, off_t0, off_t0_avg, off_time_dist;  /* <0: absent */
    int32_t n_time_dist;
    int32_t shape[4];
} lsb_lut_layout;

/* Layout of light_incidence records [S][ndet] {n_photons_det f4, t0_det f4, ...}
 * (cli/simulate_pixels.py:758-759). */
typedef struct lsb_linc_layout {
    int32_t itemsize;
    int32_t off_n_photons_det, off_t0_det;
} lsb_linc_layout;

int         lsb_abi_version(void);
const char* lsb_last_error(void);
/* number of kernel launches issued by this library since load (bench.py gpu_launches) */
int64_t     lsb_launch_count(void);

/* per-kernel device timing for bench.py: between begin and end, one CUDA event is recorded on the
 * launching stream after every kernel launch / memset of this library; lsb_profile_end writes
 * "name count total_ms" lines (sorted by total time) into `out` and returns the bytes needed. */
int         lsb_profile_begin(void* stream);
int64_t     lsb_profile_end(char* out, int64_t cap);

/* ---- RNG: numba.cuda.random state layout {s0:u8, s1:u8}, 16 bytes ------------------ */
/* numba/cuda/random.py create_xoroshiro128p_states(n, seed, subsequence_start):
 * state[i] = jump^(subsequence_start+i)(splitmix64(seed)); host-side, multithreaded. */
int lsb_rng_create_states_host(uint64_t* states_host, int64_t n, uint64_t seed, uint64_t subsequence_start);
/* the same states written straight into device memory by one kernel launch: the 2^64-step jump is a linear map over
 * GF(2), state i = J^i state 0 is reached with one 128x128 bit-matrix product per set bit of i (csrc/rng.cuh).
 * Replaces numba.cuda.random.create_xoroshiro128p_states (cli/simulate_pixels.py:96,101), bit-identical. */
int lsb_rng_create_states(uint64_t* states_dev, int64_t n, uint64_t seed, uint64_t subsequence_start, void* stream);

/* ---- per-segment kernels ----------------------------------------------------------- */
/* larndsim/quenching.py:11-44  quench(tracks, mode) */
int lsb_quench(const lsb_consts* c, const lsb_track_layout* L, void* tracks, int64_t n, int32_t mode, void* stream);
/* larndsim/drifting.py:11-58  drift(tracks) */
int lsb_drift(const lsb_consts* c, const lsb_track_layout* L, void* tracks, int64_t n, void* stream);
/* larndsim/pixels_from_track.py:43-65  max_pixels(tracks, n_max_pixels); n_max_pixels int64[1], atomic max */
int lsb_max_pixels(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t n,
                   int64_t* n_max_pixels, void* stream);
/* larndsim/pixels_from_track.py:67-109  get_pixels(tracks, active, neighboring, radius_class, n_pixels_list, radius) */
int lsb_get_pixels(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t n,
                   int32_t* active_pixels, int32_t max_active,
                   int32_t* neighboring_pixels, int32_t* neighboring_radius, int32_t max_neighbors,
                   double* n_pixels_list, int32_t radius, void* stream);
/* larndsim/detsim.py:18-40  time_intervals(track_starts, time_max, tracks) */
int lsb_time_intervals(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t n,
                       double* track_starts, int64_t* time_max, void* stream);

/* ---- glue the reference does with CuPy ---------------------------------------------- */
/* cli/simulate_pixels.py:953-956  unique_pix = cp.unique(neighboring_pixels); drop -1.
 * `unique_out` needs room for min(n_entries, max_pixel_id) ids; *n_unique is a device int64. */
int lsb_unique_pixels(const int32_t* pixels, int64_t n_entries, int64_t max_pixel_id,
                      int32_t* unique_out, int64_t* n_unique, void* workspace, int64_t workspace_bytes,
                      void* stream);
int64_t lsb_unique_pixels_workspace_bytes(int64_t max_pixel_id);
/* cli/simulate_pixels.py:1021-1025  pixel_index_map[s][p] = index of pixels[s][p] in unique_pix (else -1).
 * Uses the rank table left in `workspace` by lsb_unique_pixels. */
int lsb_pixel_index_map(const int32_t* pixels, int64_t n_entries, int64_t max_pixel_id,
                        const void* workspace, int64_t* pixel_index_map, void* stream);
/* same, but by binary search in an arbitrary sorted unique_pix (no workspace) */
int lsb_pixel_index_map_search(const int32_t* pixels, int64_t n_entries, const int32_t* unique_pix,
                               int64_t n_unique, int64_t* pixel_index_map, void* stream);

/* ---- induced current ---------------------------------------------------------------- */
/* larndsim/detsim.py:258-348  tracks_current_mc(signals, pixels, tracks, response, rng_states)
 *  signals f32[S][P][T]; pixels i32[S][P]; response [Rx][Ry][Rt] f32 (response_f64=0) or f64;
 *  rng_states {u8,u8}[>= S*P], state of (itrk,ipix) at index itrk + S*ipix (detsim.py:324).
 *  rng_mode 0 ("cloud", production): one sample cloud per (segment,pixel), drawn z,x,y per
 *    step from that stream and applied to every tick -- what a converged warp of the
 *    reference does; statistically equivalent, exact when tran_diff=long_diff=0.
 *  rng_mode 1 ("replay"): the reference's draw pattern thread for thread with ticks
 *    consumed in order (= the CUDA simulator with 1-thread blocks); sequential per pair. */
int lsb_tracks_current_mc(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t S,
                          const int32_t* pixels, int32_t P, float* signals, int32_t T,
                          const void* response, int32_t Rx, int32_t Ry, int32_t Rt, int32_t response_f64,
                          uint64_t* rng_states, int64_t n_rng, int64_t rng_stride, int32_t rng_mode,
                          void* workspace, int64_t workspace_bytes, void* stream);
/* rng_stride = ntrk of detsim.py:324 (grid size along the segment axis; <=0: S).
 * workspace: >= lsb_tracks_current_mc_workspace_bytes(S, P, n) for n sample points; if the batch needs
 * more the call splits it into segment ranges that fit (never fails for a workspace that holds the
 * samples of one segment). */
int64_t lsb_tracks_current_mc_workspace_bytes(int64_t S, int32_t P, int64_t max_steps_total);
/* sample points drawn by the last lsb_tracks_current_mc call (roofline accounting) */
int64_t lsb_tracks_current_mc_last_samples(void);
/* interior-tick strategy of tracks_current_mc for float32 tables at unit sampling ratio: 1 = grouped path
 * (default: offsets sorted per pair, aligned 4-word register windows shared by the samples of a group),
 * 0 = generic gather stream.  Same results to 1e-5; the default can also be set with LSB_MC_GROUPED=0/1. */
void lsb_mc_set_grouped(int32_t on);
int32_t lsb_mc_get_grouped(void);
/* consecutive ticks a lane owns in the grouped path: 4 (default; 8-word windows, two LDG.128 per group and 128 ticks of a warp)
 * or 8 (12-word windows, three LDG.128 per 256 ticks: fewer table words per tick but more registers -- measured slower on
 * B200); also LSB_ACC_LANE_TICKS=4/8.  Tables sampled at half the tick
 * length (RESPONSE_SAMPLING = TIME_SAMPLING / 2, ND-LAr) take the same path on a phase-split copy of the table. */
void lsb_mc_set_lane_ticks(int32_t n);
int32_t lsb_mc_get_lane_ticks(void);
/* phase-aligned variant of the grouped path: the offsets of a pair are ordered by (offset mod 4, offset / 4) and folded into one
 * 4-byte record per DISTINCT offset; for alignment class d a lane owns ticks 4 lane - d .. 4 lane - d + 3 of a 124-tick block, so
 * the table words it needs are one aligned float4 (one LDG.128 + 4 FFMA per distinct offset instead of an 8-word window per
 * group).  mode: -1 automatic (on when the table is phase-split, i.e. RESPONSE_SAMPLING = TIME_SAMPLING / 2), 0 off, 1 on;
 * also LSB_ACC_ALIGNED=0/1.  Same results to 1e-5 (summation order differs). */
void lsb_mc_set_aligned(int32_t mode);
int32_t lsb_mc_get_aligned(void);
/* larndsim/detsim.py:351-453  tracks_current(signals, pixels, tracks, response) */
int lsb_tracks_current(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t S,
                       const int32_t* pixels, int32_t P, float* signals, int32_t T,
                       const void* response, int32_t Rx, int32_t Ry, int32_t Rt, int32_t response_f64,
                       void* stream);

/* ---- per-pixel reduction ------------------------------------------------------------ */
/* larndsim/detsim.py:529-562  get_track_pixel_map(track_pixel_map, unique_pix, pixels) */
int lsb_get_track_pixel_map(int64_t* track_pixel_map, int32_t K, const int32_t* unique_pix, int64_t U,
                            const int32_t* pixels, int64_t S, int32_t P, void* stream);
/* larndsim/detsim.py:564-607  get_track_pixel_map2(track_pixel_map, unique_pix, pixels, distances, max_distance) */
int lsb_get_track_pixel_map2(int64_t* track_pixel_map, int32_t K, const int32_t* unique_pix, int64_t U,
                             const int32_t* pixels, const int32_t* distances, int64_t S, int32_t P,
                             int32_t max_distance, void* stream);
/* larndsim/detsim.py:468-527  sum_pixel_signals(pixels_signals, signals, track_starts, pixel_index_map,
 *                                              track_pixel_map, pixels_tracks_signals, overflow_flag)
 * Gather formulation: contributions to one pixel are added in ascending segment order, so
 * sums are reproducible (the reference's float64 atomics are order-free). */
int lsb_sum_pixel_signals(const lsb_consts* c, double* pixels_signals, int64_t U, int32_t Tt,
                          const float* signals, int64_t S, int32_t P, int32_t T,
                          const double* track_starts, const int64_t* pixel_index_map,
                          const int64_t* track_pixel_map, int32_t K,
                          double* pixels_tracks_signals, double* overflow_flag, void* stream);

/* ---- front-end electronics ---------------------------------------------------------- */
/* larndsim/fee.py:517-655  get_adc_values(pixels_signals, pixels_signals_tracks, time_ticks, adc_list,
 *        adc_ticks_list, time_padding, rng_states, current_fractions, pixel_thresholds) */
int lsb_get_adc_values(const lsb_consts* c, const double* pixels_signals, const double* pixels_signals_tracks,
                       int64_t U, int32_t Tt, int32_t K,
                       const double* time_ticks, int32_t n_time_ticks,
                       double* adc_list, double* adc_ticks_list, int32_t A, double time_padding,
                       uint64_t* rng_states, int64_t n_rng,
                       double* current_fractions, const double* pixel_thresholds, void* stream);
/* larndsim/fee.py:499-515  digitize(integral_list[, gain]); gain_list may be NULL (scalar gain*mV/e) */
int lsb_digitize(const lsb_consts* c, const double* integral_list, const double* gain_list, int64_t n,
                 double* adcs, void* stream);

/* ---- light ---------------------------------------------------------------------------- */
/* larndsim/lightLUT.py:65-136  calculate_light_incidence(tracks, lut, light_incidence, voxel) */
int lsb_calculate_light_incidence(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t S,
                                  const void* lut, const lsb_lut_layout* LL,
                                  void* light_incidence, const lsb_linc_layout* LI, int32_t ndet,
                                  int32_t* voxel, const double* op_channel_efficiency,
                                  const int64_t* op_channel_to_tpc, void* stream);
/* larndsim/light_sim.py:58-129  sum_light_signals(...) */
int lsb_sum_light_signals(const lsb_consts* c, const lsb_track_layout* L, const void* segments, int64_t S,
                          const int32_t* segment_voxel, const int64_t* segment_track_id,
                          const void* light_inc, const lsb_linc_layout* LI, int32_t ndet_inc,
                          const int32_t* op_channel, const void* lut, const lsb_lut_layout* LL,
                          double start_time, float* light_sample_inc, int32_t ndet, int32_t nticks,
                          int64_t* true_track_id, double* true_photons, int32_t n_true,
                          const int64_t* sorted_indices, int64_t n_sorted, double t0_profile_length,
                          void* stream);
/* larndsim/light_sim.py:148-183  calc_scintillation_effect(6 arrays) */
int lsb_calc_scintillation_effect(const lsb_consts* c, const float* light_sample_inc,
                                  const int64_t* inc_true_track_id, const double* inc_true_photons,
                                  float* light_sample_inc_scint, int64_t* scint_true_track_id,
                                  double* scint_true_photons, int32_t ndet, int32_t nticks,
                                  int32_t n_true_in, int32_t n_true_out, void* stream);
/* larndsim/light_sim.py:220-238  calc_stat_fluctuations(in, out, rng_states) */
int lsb_calc_stat_fluctuations(const lsb_consts* c, const float* light_sample_inc, float* light_sample_inc_disc,
                               int32_t ndet, int32_t nticks, uint64_t* rng_states, int64_t n_rng, void* stream);
/* larndsim/light_sim.py:303-336  calc_light_detector_response(6 arrays);
 * light_gain = light.LIGHT_GAIN (device f64[ndet...]); impulse_model_host = light.IMPULSE_MODEL (HOST f64[n_impulse]:
 * the tap weights are evaluated once per call on the host) */
int lsb_calc_light_detector_response(const lsb_consts* c, const float* light_sample_inc,
                                     const int64_t* inc_true_track_id, const double* inc_true_photons,
                                     float* light_response, int64_t* resp_true_track_id,
                                     double* resp_true_photons, int32_t ndet, int32_t nticks,
                                     int32_t n_true_in, int32_t n_true_out,
                                     const double* light_gain, const double* impulse_model_host, int32_t n_impulse,
                                     void* stream);

/* ---- fused device-resident batch driver (replaces cli/simulate_pixels.py:907-1117) ---- */
typedef struct lsb_chain lsb_chain;   /* opaque; owns its device workspace */

typedef struct lsb_chain_result {
    int64_t n_segments, n_unique_pixels, max_active, max_neighbors, n_ticks;   /* S, U, maxpix, P, T */
    int64_t n_hits;                   /* adc_list entries above pedestal */
    int64_t n_samples, n_fma;         /* MC sample points N_sp / (sample,tick) pairs that pass every test N_fma, counted on the device */
    int64_t n_pairs;                  /* (segment,pixel) pairs that hold a pixel id (S * P-bar) */
    int64_t n_groups, n_edge, n_irregular;   /* MC diagnostics: group records of the grouped gather, (sample,tick) pairs on edge ticks, samples on the exact path */
    /* device pointers valid until the next run / destroy */
    const int32_t* unique_pix;        /* [U] */
    const int64_t* track_pixel_map;   /* [U][K] */
    const double*  adc_list;          /* [U][A] integrated charge */
    const double*  adc_digit;         /* [U][A] digitize(adc_list) */
    const double*  adc_ticks_list;    /* [U][A] */
    const double*  current_fractions; /* [U][A][K] */
    const float*   signals;           /* [S][P][T] */
    const double*  pixels_signals;    /* [U][Tt] */
    float stage_ms[12];               /* per-stage CUDA-event times of the last run (if timing enabled) */
} lsb_chain_result;

lsb_chain* lsb_chain_create(const lsb_consts* c, const lsb_track_layout* L,
                            const void* response, int32_t Rx, int32_t Ry, int32_t Rt, int32_t response_f64,
                            int32_t rng_mode, int32_t enable_stage_timing);
void       lsb_chain_destroy(lsb_chain* h);
/* dense=1: materialise pixels_tracks_signals [U][Tt][K] exactly like cli/simulate_pixels.py:1053-1055 and
 * run get_adc_values on it; dense=0 (default): read the per-segment waveforms from `signals` through the
 * (pixel, slot) entry list -- same numbers, no 0.8 MB/pixel tensor. */
int        lsb_chain_set_dense(lsb_chain* h, int32_t dense);
/* current_fractions (fee.py:568-573, the per-segment share of each hit): exact=1 replays the reference's tick-major
 * summation order (bit-identical float64), exact=0 (default) evaluates the same double sum as order-free weighted
 * sums over each hit window (agrees to ~1e-15 relative; hits, timestamps and charges do not depend on it). */
int        lsb_chain_set_exact_fractions(lsb_chain* h, int32_t exact);
/* RNG policy of the handle.  fresh=0 (default): cli/simulate_pixels.py:92-104 maybe_create_rng_states -- one state array
 * per handle that evolves from batch to batch, fresh states (seed = rng_seed) appended when a batch needs more.
 * fresh=1: every batch starts from create_xoroshiro128p_states(n, seed = rng_seed); its result then depends on
 * (records, rng_seed) only, not on which handle / rank ran it or on what ran before. */
int        lsb_chain_set_rng_fresh(lsb_chain* h, int32_t fresh);
/* lsb_chain_result.signals is stored sparsely by the fused chain (per row only the ticks covered by the pair's samples are
 * written; later stages never read the rest).  Call this before reading the dense f4[S, P, T] array: it zero-fills the
 * unwritten parts of the last batch on `stream` (idempotent). */
int        lsb_chain_signals_dense(lsb_chain* h, void* stream);
/* tracks on the device, modified in place by quench/drift like the reference; quench_mode < 0: the records have been
 * quenched and drifted already (cli/simulate_pixels.py:732,742 run both over the whole file before the batch loop) */
int lsb_chain_run(lsb_chain* h, void* tracks_dev, int64_t S, int32_t quench_mode, uint64_t rng_seed,
                  int32_t n_events, lsb_chain_result* out, void* stream);
/* tracks in (pinned) host memory: H2D, chain, D2H of the packet-level outputs into host buffers
 * sized by the caller (adc_digit/adc_ticks U_cap*A doubles, unique_pix U_cap int32); the e2e path. */
int lsb_chain_run_host(lsb_chain* h, void* tracks_host, int64_t S, int32_t quench_mode, uint64_t rng_seed,
                       int32_t n_events, int32_t* unique_pix_host, double* adc_digit_host,
                       double* adc_ticks_host, int64_t U_cap, lsb_chain_result* out, void* stream);

/* Pipelined form.  The handle owns a high-priority stream (front + FEE stages) and a low-priority stream
 * (MC stage).  lsb_chain_run_async queues one batch behind the work already on `stream` and returns;
 * lsb_chain_wait blocks until that batch is complete and fills `out` (stage_ms is not filled).  Alternate two
 * handles to run the latency-bound FEE kernels of one batch under the MC kernels of the next. */
int lsb_chain_run_async(lsb_chain* h, void* tracks_dev, int64_t S, int32_t quench_mode, uint64_t rng_seed,
                        int32_t n_events, void* stream);
int lsb_chain_run_host_async(lsb_chain* h, void* tracks_host, int64_t S, int32_t quench_mode, uint64_t rng_seed,
                             int32_t n_events, int32_t* unique_pix_host, double* adc_digit_host,
                             double* adc_ticks_host, int64_t U_cap);
int lsb_chain_wait(lsb_chain* h, lsb_chain_result* out);

/* ---- light triggers and waveform digitisation ---------------------------------------------- */
/* larndsim/light_sim.py:380-477  get_triggers (threshold mode): per trigger group (channels_per_group consecutive
 * rows of `signal`) the channel sum, averaged over blocks of sample_factor ticks (zero padded), is compared with
 * group_threshold; chan_module[d] = module slot (0..n_modules-1) of signal row d, or -1; every module then searches
 * its triggers sequentially with a dead time of digit_ticks, the reference's index bookkeeping included.
 * Outputs: trig_idx[module][max_trig] (tick indices as the reference reports them), n_trig[module]. */
int lsb_light_get_triggers(const void* signal, int32_t signal_f64, int32_t ndet, int64_t nticks, int32_t channels_per_group,
                           int32_t sample_factor, const double* group_threshold, const int32_t* chan_module, int32_t n_modules,
                           int64_t digit_ticks, int32_t max_trig, int64_t* trig_idx, int32_t* n_trig, void* stream);
/* larndsim/light_sim.py:480-543 digitize_signal + the zero padding / missing-channel rows / LIGHT_NBIT rounding of
 * sim_triggers (:545-619).  The padded, channel-sorted waveform array the reference builds is described, not
 * materialised: row k of it has channel row_channel[k] and is row row_source[k] of `signal` (-1: all zeros), shifted
 * right by front_pad ticks inside a length of padded_len; array_is_f32 = that array would be float32 (no padding, no
 * added rows, float32 input).  truncate = 1 applies round(x / 2^(16-nbit)) * 2^(16-nbit).  Truth outputs must be
 * pre-filled (-1 / 0) by the caller like the reference does. */
int lsb_light_digitize(const void* signal, int32_t signal_f64, int64_t nticks, int32_t n_rows, const int64_t* row_channel,
                       const int32_t* row_source, int64_t front_pad, int64_t padded_len, int32_t array_is_f32,
                       const int64_t* true_track_id, const double* true_photons, int32_t n_truth, int64_t n_trig,
                       const int64_t* trig_channel, int32_t n_det_module, int32_t n_samples, double digit_sample_spacing,
                       double light_tick_size, double mc_truth_threshold, int32_t light_nbit, int32_t truncate,
                       double* digit_signal, int64_t* digit_true_track_id, double* digit_true_photons, int32_t n_truth_out,
                       void* stream);

/* larndsim/light_sim.py:24-42 get_nticks and :44-57 get_active_op_channel, device part: over light_incidence[n_segments][ndet]
 * entries with n_photons_det > 0: t0_minmax[0] = min t0_det, t0_minmax[1] = max t0_det (float32; +inf / -inf when there is
 * none), active[d] = 1 if any segment gives photons to channel d.  All pointers device. */
int lsb_light_extent(const void* light_incidence, const lsb_linc_layout* LI, int64_t n_segments, int32_t ndet,
                     float* t0_minmax, uint8_t* active, void* stream);

/* larndsim/light_sim.py:621-661 zero_suppress_waveform_truth: one 32-byte record {trigger_id i4, op_channel_id i4, tick i4,
 * event_id i4, segment_id i8, pe_current f8} per slot of true_track_id[n_trig][n_det][n_samples][n_truth] that is not -1,
 * in C order; op_channel[d] = channel id of column d; trigger_id = first_trigger_id + the running sum of the trigger
 * indices of the records so far (the reference accumulates, :644).  rows: device, room for every slot; *n_rows device. */
int64_t lsb_light_truth_ws_bytes(int64_t n_slots);
int lsb_light_zero_suppress_truth(const int64_t* true_track_id, const double* true_photons, int64_t n_trig, int32_t n_det,
                                  int32_t n_samples, int32_t n_truth, const int32_t* op_channel, int32_t event_id,
                                  int32_t first_trigger_id, void* rows, int64_t* n_rows, void* ws, int64_t ws_bytes, void* stream);



/* ---- active volume + batching (the callers that cut the segment array into units) ------- */
/* larndsim/active_volume.py:4-46  select_active_volume: first_tpc[i] = lowest TPC index in [tpc_lo, tpc_hi) whose open
 * box contains the start OR the end point of segment i (float64 comparisons), -1 if none; the reference's return value
 * (`nonzero(mask)[0]`) is `indices[0 .. *n_selected)` (ascending; optional: pass NULL to skip the compaction).
 * borders: DEVICE f64[n_tpc,3,2] (either order on the last axis: sorted on load, active_volume.py:24).
 * i_module >= 1 in the reference is tpc_lo = 2*(i_module-1), tpc_hi = 2*i_module. */
int64_t lsb_active_volume_ws_bytes(int64_t n);
int lsb_active_volume(const lsb_track_layout* L, const void* tracks, int64_t n, const double* borders, int32_t n_tpc,
                      int32_t tpc_lo, int32_t tpc_hi, int32_t* first_tpc, int64_t* indices, int64_t* n_selected,
                      void* ws, int64_t ws_bytes, void* stream);
/* larndsim/util/batching.py:17-67  TPCBatcher, every batch of the run in one call.  Unit u = e * n_tpc_batches + b is
 * the batch the reference's iterator yields at position u: event events_sorted[e] (= np.unique of the event field,
 * batching.py:29) and TPCs [b*tpc_batch_size, (b+1)*tpc_batch_size).  A segment belongs to the FIRST batch of its
 * event that contains it (the `_simulated` mask, batching.py:49,63), i.e. b = first_tpc / tpc_batch_size.
 * order[unit_offsets[u] .. unit_offsets[u+1]) = the rows `mask` selects for unit u, ascending (tracks[mask] order);
 * rows in no unit follow at order[unit_offsets[n_units] .. n).  event field: byte offset + lsb_dtype in the record. */
int64_t lsb_batch_units_ws_bytes(int64_t n);
int lsb_batch_units(const void* tracks, int64_t n, int32_t itemsize, int32_t event_offset, int32_t event_dtype,
                    const int64_t* events_sorted, int64_t n_events, const int32_t* first_tpc, int32_t tpc_batch_size,
                    int32_t n_tpc_batches, int64_t* order, int64_t* unit_offsets, void* ws, int64_t ws_bytes, void* stream);

/* ---- static key -> value table ----------------------------------------------------------- */
/* larndsim/util/cuda_dict.py:1-230  CudaDict lookup / contains (per-pixel thresholds and gains,
 * cli/simulate_pixels.py:1080-1100): out[i] = value of query[i] or *default_host; exists[i] = 1 if present.
 * keys_sorted: int32 ascending, unique; values: n elements of value_bytes (4 or 8) each, copied bit for bit. */
int lsb_table_lookup(const int32_t* keys_sorted, const void* values, int64_t n, int32_t value_bytes,
                     const int32_t* query, int64_t nq, const void* default_host, void* out, uint8_t* exists,
                     void* stream);

/* ---- hit compaction + LArPix packets ---------------------------------------------------- */
/* One output packet.  packet_type uses the codes of larpix.format.hdf5format (0 data, 4 timestamp, 6 sync,
 * 7 trigger); sub_type is the sync type ('S') or trigger type (0x02) byte; timestamp_s is the float
 * timestamp of a TimestampPacket [s]; parity is Packet_v2.assign_parity() of a data packet. */
typedef struct lsb_packet {
    uint8_t packet_type, io_group, io_channel, chip_id, channel_id, dataword, first_packet, parity, sub_type, pad[3];
    uint32_t receipt_timestamp;
    uint64_t timestamp;
    double timestamp_s;
} lsb_packet;
/* Readout constants of larndsim.consts.detector / light / units that fee.export_to_hdf5 reads (host arrays):
 * TILE_MAP [2][ntx][nty]; TILE_ORIENTATIONS as the sign of the x / y axis per tile id; PIXEL_CONNECTION_DICT as
 * chip * 1000 + channel per rotated in-tile pixel (-1: absent); TILE_CHIP_TO_IO as io_group * 1000 + io_channel
 * per (tile id, chip) (-1: absent); MODULE_TO_IO_GROUPS per module id; io_groups = the groups that receive the
 * per-event timestamp / sync packets (already restricted to i_mod); bad_channels = sorted keys
 * ((io_group * 1000 + io_channel) * 1000 + chip) * 64 + channel. */
typedef struct lsb_readout_tables {
    double clock_cycle, adc_pedestal, mus, s;
    int64_t clock_reset_period;
    int32_t light_trig_mode;
    int32_t n_pixels[2], n_pixels_per_tile[2], n_tiles_xy[2];
    int32_t n_tiles, n_modules, max_groups, n_io_groups, n_bad;
    const int32_t* tile_map;
    const int32_t* tile_orientation;
    const int32_t* pixel_connection;
    const int32_t* tile_chip_to_io;
    const int32_t* module_n_groups;
    const int32_t* module_io_groups;
    const int32_t* io_groups;
    const int64_t* bad_channels;
} lsb_readout_tables;
/* larndsim/fee.py:84-359  export_to_hdf5(event_id_list, adc_list, adc_ticks_list, unique_pix, current_fractions,
 * track_ids, traj_ids, filename, event_start_times, light_trigger_*, bad_channels, i_mod) without the file I/O:
 * the packets in the reference's order and the mc_packets_assn rows (n_assn = ASSOCIATION_COUNT_TO_STORE).
 * Device arrays: event_id i8[U,A], adc (digitised) f8[U,A], adc_ticks f8[U,A], unique_pix i4[U],
 * current_fractions f8[U,A,K], track_ids / traj_ids i8[U,K]; pix_t0_ticks[i] = int(event_start_times[inv[i]] /
 * CLOCK_CYCLE) and pix_t0_us[i] = event_start_times[inv[i]] with inv the rank of event_id[i,0] among the sorted
 * unique events (fee.py:136-137); light triggers (times [us], event, module).  assn_rows: cap_packets records of the
 * mc_packets_assn dtype (event_ids i8[1] | segment_ids i8[n] | fraction f8[n] | file_traj_ids i8[n] | fraction_traj f8[n],
 * 8 + 32 n bytes each).  *n_packets receives the packet
 * count; if it exceeds cap_packets nothing is written and the call fails (retry with that capacity). */
int lsb_export_packets(const lsb_readout_tables* rt, int64_t U, int32_t A, int32_t K, const int64_t* event_id,
                       const double* adc, const double* adc_ticks, const int32_t* unique_pix, const double* current_fractions,
                       const int64_t* track_ids, const int64_t* traj_ids, const int64_t* pix_t0_ticks, const double* pix_t0_us,
                       int32_t n_trig, const double* trig_times, const int64_t* trig_event, const int32_t* trig_module,
                       int64_t cap_packets, lsb_packet* packets, void* assn_rows, int32_t n_assn, int64_t* n_packets, void* stream);


/* ---- one rank's share of a spill / file: the batch loop itself ------------------------------ */
/* cli/simulate_pixels.py:864-1117 (loop over the (event, TPC group) batches of larndsim/util/batching.py:40-67) +
 * save_results :1370-1390 -> fee.export_to_hdf5 (fee.py:84-359, WRITE_BATCH_SIZE = 1: one export per batch).
 * The runner owns `depth` chain handles (RNG policy "fresh", see lsb_chain_set_rng_fresh) that are kept in flight
 * together, and per-rank output buffers the packets and mc_packets_assn rows of every unit are appended to on the device,
 * unit after unit, in the order the units are passed. */
typedef struct lsb_spill lsb_spill;
typedef struct lsb_spill_result {
    int64_t n_units, n_segments, n_packets;       /* n_packets: packets written (or required, on overflow) */
    int64_t n_hits, n_unique_pixels, n_samples, n_fma;   /* sums over the units (roofline accounting) */
    int64_t pair_ticks, pixel_ticks;              /* sum of n_pairs * T and of U * Tt over the units */
    const lsb_packet* packets;                    /* device, [n_packets] */
    const void* assn_rows;                        /* device, [n_packets] records of assn_row_bytes */
    const void* records;                          /* device, the units' records (unit after unit) as the chain left them */
    int64_t assn_row_bytes;
    int32_t overflow, pad;
} lsb_spill_result;
lsb_spill* lsb_spill_create(const lsb_consts* c, const lsb_track_layout* L, const void* response, int32_t Rx, int32_t Ry,
                            int32_t Rt, int32_t response_f64, const lsb_readout_tables* rt, int32_t n_assn, int32_t depth);
void lsb_spill_destroy(lsb_spill* sp);
/* serial=1: every stage of a unit is queued on one stream (no overlap of the current stage with front-end stages of other
 * units): what per-kernel timing with events needs */
int  lsb_spill_set_serial(lsb_spill* sp, int32_t serial);
/* tracks_dev: the selected, quenched and drifted records of the file (device).  order_dev: row permutation written by
 * lsb_batch_units (NULL: identity).  Unit i = rows order[unit_begin[i] .. unit_begin[i] + unit_count[i]) with event id
 * unit_event[i], event start time unit_t0_us[i] and RNG seed unit_seed[i] (host arrays).  seg_id / traj_id: byte offset and
 * lsb_dtype of the record fields that label the truth rows (cli/simulate_pixels.py:483-484: 'segment_id',
 * 'file_traj_id'; offset < 0: -1 is stored).  unit_packets_host[i] receives the number of packets of unit i.
 * Returns -2 if cap_packets was too small (out->n_packets = required capacity; nothing usable was written). */
int lsb_spill_run(lsb_spill* sp, const void* tracks_dev, const int64_t* order_dev, int64_t n_units,
                  const int64_t* unit_begin, const int64_t* unit_count, const int64_t* unit_event, const double* unit_t0_us,
                  const uint64_t* unit_seed, int32_t seg_id_offset, int32_t seg_id_dtype, int32_t traj_id_offset,
                  int32_t traj_id_dtype, int64_t cap_packets, int64_t* unit_packets_host, lsb_spill_result* out, void* stream);
/* tracks[indices] on the device (cli/simulate_pixels.py:670): dst[r] = src[order[r]], records of `itemsize` bytes */
int lsb_gather_records(const void* src_dev, const int64_t* order_dev, int64_t n, int32_t itemsize, void* dst_dev, void* stream);
/* gather of variable-length blocks into one buffer (all device pointers; the three arrays are host arrays): block b =
 * src_dev[b] .. + bytes[b] -> dst_dev + dst_off[b].  Puts the units received from all ranks into file order. */
int lsb_copy_blocks(int64_t n_blocks, const void* const* src_dev, const int64_t* dst_off, const int64_t* bytes, void* dst_dev,
                    void* stream);
/* Multi-rank host output on one node: the ranks share one host table (a shared-memory mapping made by the caller).  Each rank
 * registers the mapping with its CUDA context (lsb_host_register; page-locks it) and copies the blocks of its own units to their
 * file-order byte offsets (lsb_d2h_blocks: asynchronous on `stream`, src_dev / dst_off / bytes are host arrays) -- the copy the
 * reference does once per batch in fee.export_to_hdf5 (fee.py:84-359), spread over the ranks' PCIe links. */
int lsb_host_register(void* host_ptr, int64_t bytes);
int lsb_host_unregister(void* host_ptr);
int lsb_d2h_blocks(int64_t n_blocks, const void* const* src_dev, const int64_t* dst_off, const int64_t* bytes, void* dst_host,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LARNDSIM_B200_H */
