// chain.cuh -- device-resident driver for one batch of the charge readout chain,
// quench -> drift -> get_pixels -> tracks_current_mc -> sum_pixel_signals -> get_adc_values -> digitize,
// i.e. the sequence of cli/simulate_pixels.py:907-1117 with the CuPy glue (unique, index maps, fills,
// linspace) done by the kernels of glue.cuh / pixelmap.cuh.  Everything stays in HBM; the host reads
// back three scalars per batch (P, U, T) that size the buffers, exactly where the reference
// synchronises (max_pixels[0], cp.unique, max_length).
#pragma once
#include "common.cuh"
#include "segments.cuh"
#include "glue.cuh"
#include "current.cuh"
#include "pixelmap.cuh"
#include "fee.cuh"

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    int need(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return lsb_fail_cuda(e, "cudaFree"); }
        // batches of a run differ in size (event x TPC group): grow generously so that a run re-allocates a handful of
        // times, not at every new maximum (a cudaFree + cudaMalloc of GB-sized buffers costs 10-500 ms on the host)
        size_t want = bytes + (bytes < ((size_t)4 << 30) ? bytes / 2 : bytes / 8) + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e != cudaSuccess) { p = nullptr; return lsb_fail_cuda(e, "cudaMalloc (chain buffer)"); }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

enum { ST_QUENCH_DRIFT = 0, ST_GET_PIXELS, ST_UNIQUE, ST_TIME_INTERVALS, ST_TRACKS_CURRENT, ST_INDEX_MAPS, ST_SUM_PIXELS,
       ST_GET_ADC, ST_DIGITIZE, ST_COUNT };

struct ChainScalars { long long max_pixels; unsigned long long max_tran_bits; long long n_unique; long long t_max; int max_dist; int mc_overflow;
                      long long n_hits; double sum_len; long long mc_total; unsigned long long mc_nfma, mc_npairs, mc_diag[3]; };

struct lsb_chain {
    lsb_consts c;
    lsb_track_layout L;
    const void* response; int Rx, Ry, Rt, f64, rng_mode, timing;
    DevBuf response_split;    // phase-split copy of the table (RESPONSE_SAMPLING = TIME_SAMPLING / 2), see current.cuh
    DevBuf sig_ranges;        // int2 per (segment, pixel) row of `signals`: ticks that hold data (rows are stored sparsely)
    int signals_dense;        // the rows of the last batch have been zero-filled outside their ranges
    long long last_rows; int last_T;
    DevBuf tracks, scal, active, neigh, nrad, npl, uniq, uniq_ws, starts, signals, mc_ws, pim, tpm, psig, pts, oflow, tticks,
           integral, adc_digit, adc_ticks, cf, thr, rng, nhits, sx_slot, sx_counts, sx_cursor, sx_raw, sx_offs, sx_bsums, sx_sorted;
    DevBuf arena_buf; TmpArena arena;
    int exact_fractions;      // 1: current_fractions in the reference's summation order (bit-identical), 0: order-free weighted sums
    int dense;                // 1: materialise pixels_tracks_signals like the reference (parity / debugging)
    long long n_rng;
    int rng_fresh;            // 1: the states are re-created from rng_seed for every batch (result independent of the batches seen before)
    int tticks_n; long long tticks_events;
    cudaEvent_t ev[ST_COUNT + 1];
    // pipelined (asynchronous) operation: front + FEE stages on a high-priority stream, the MC stage on a
    // low-priority one, so the latency-bound FEE kernels of one batch run under the L1-bound MC kernels of the
    // next batch (issued through a second chain handle)
    cudaStream_t hp, lp;
    cudaEvent_t ev_in, ev_front, ev_mc, ev_done;
    cudaEvent_t tl[6];        // timeline of the last async batch: front begin/end, MC begin/end, FEE begin, done
    ChainScalars* hs_pinned;
    lsb_chain_result pending;
    int pending_valid;
};


// max(tran_diff) (-> neighbour radius) and the summed 3-D segment length (-> upper bound of the MC sample count:
// every (segment,pixel) takes at most len/MIN_STEP_SIZE + 1 steps, detsim.py:319)
__global__ void k_chain_max_tran(Layout L, const char* __restrict__ tracks, long long n, unsigned long long* __restrict__ out,
                                 double* __restrict__ sum_len) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    double v = 0.0, len = 0.0;
    if (i < n) {
        const char* t = tracks + i * L.itemsize;
        double td = fld_get(L, t, LSB_F_TRAN_DIFF); if (td > v) v = td;   // NaN/negatives ignored
        double dx = fld_get(L, t, LSB_F_X_END) - fld_get(L, t, LSB_F_X_START);
        double dy = fld_get(L, t, LSB_F_Y_END) - fld_get(L, t, LSB_F_Y_START);
        double dz = fld_get(L, t, LSB_F_Z_END) - fld_get(L, t, LSB_F_Z_START);
        len = sqrt(dx * dx + dy * dy + dz * dz);
        if (!(len >= 0.0) || len > 1e30) len = 1e30;                     // NaN / inf: force the synchronous path
    }
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, b, o); b = w > b ? w : b;
        len += __shfl_xor_sync(0xffffffffu, len, o);
    }
    if ((threadIdx.x & 31) == 0) { if (b) atomicMax(out, b); atomicAdd(sum_len, len); }
}
__global__ void k_fill_i32(int32_t* p, long long n, int32_t v) { long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; if (i < n) p[i] = v; }
__global__ void k_fill_i64(long long* p, long long n, long long v) { long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; if (i < n) p[i] = v; }
__global__ void k_fill_f64(double* p, long long n, double v) { long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; if (i < n) p[i] = v; }
__global__ void k_max_i32(const int32_t* __restrict__ p, long long n, int* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int v = i < n ? p[i] : -2147483647;
    for (int o = 16; o > 0; o >>= 1) { int w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    if ((threadIdx.x & 31) == 0) atomicMax(out, v);
}
__global__ void k_count_hits(const double* __restrict__ adc_digit, long long n, double pedestal_adc, long long* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int h = (i < n && adc_digit[i] > pedestal_adc) ? 1 : 0;              // fee.py:141 `if adc > digitize(0)`
    unsigned m = __ballot_sync(0xffffffffu, h);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd((unsigned long long*)out, (unsigned long long)__popc(m));
}

LSB_EXPORT lsb_chain* lsb_chain_create(const lsb_consts* c, const lsb_track_layout* L, const void* response, int32_t Rx,
                                       int32_t Ry, int32_t Rt, int32_t response_f64, int32_t rng_mode, int32_t enable_stage_timing) {
    if (!c || !L || !response || Rx <= 0 || Ry <= 0 || Rt <= 0) { lsb_fail_arg("chain_create: bad arguments"); return nullptr; }
    lsb_chain* h = new lsb_chain();
    h->c = *c; h->L = *L; h->response = response; h->Rx = Rx; h->Ry = Ry; h->Rt = Rt; h->f64 = response_f64;
    h->rng_mode = rng_mode; h->timing = enable_stage_timing; h->dense = 0; h->exact_fractions = 0; h->n_rng = 0; h->rng_fresh = 0; h->tticks_n = 0; h->tticks_events = -1;
    for (int i = 0; i <= ST_COUNT; i++) cudaEventCreate(&h->ev[i]);
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    // non-blocking: no implicit synchronisation with the legacy default stream (torch / numba / cupy work there)
    cudaStreamCreateWithPriority(&h->hp, cudaStreamNonBlocking, prio_hi);
    // ONE low-priority stream for the MC stage of every handle: MC stages of different batches must run back to
    // back, not interleaved (they are L1/L2-bound and would evict each other's LUT lines)
    static cudaStream_t s_mc_stream[64] = {nullptr};          // per device
    int dev = 0; cudaGetDevice(&dev); dev = (dev < 0 || dev >= 64) ? 0 : dev;
    if (!s_mc_stream[dev]) cudaStreamCreateWithPriority(&s_mc_stream[dev], cudaStreamNonBlocking, prio_lo);
    h->lp = s_mc_stream[dev];
    cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_front, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_mc, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming);
    for (int i = 0; i < 6; i++) cudaEventCreate(&h->tl[i]);
    h->hs_pinned = nullptr;
    cudaMallocHost((void**)&h->hs_pinned, sizeof(ChainScalars));
    h->pending_valid = 0;
    h->arena.base = nullptr; h->arena.cap = h->arena.off = h->arena.high_water = h->arena.overflow = 0;
    if (rng_mode == 0 && mc_wants_split(c, response_f64)) {
        cudaStream_t st = h->hp;
        if (h->response_split.need((size_t)Rx * Ry * Rt * 4) || mc_build_split(response, Rx, Ry, Rt, (float*)h->response_split.p, st) ||
            cudaStreamSynchronize(st) != cudaSuccess) { lsb_chain_destroy(h); return nullptr; }
    }
    return h;
}
LSB_EXPORT int lsb_chain_set_dense(lsb_chain* h, int32_t dense) {
    LSB_REQUIRE(h, "chain_set_dense: null handle");
    h->dense = dense ? 1 : 0;
    return 0;
}
LSB_EXPORT int lsb_chain_set_exact_fractions(lsb_chain* h, int32_t exact) {
    LSB_REQUIRE(h, "chain_set_exact_fractions: null handle");
    h->exact_fractions = exact ? 1 : 0;
    return 0;
}
// RNG policy.  0 (default) = the reference's maybe_create_rng_states (cli/simulate_pixels.py:92-104): the handle keeps ONE
// evolving state array, fresh states (seed = rng_seed) are appended when a batch needs more.  1 = fresh: every batch
// starts from create_xoroshiro128p_states(n, seed = rng_seed), made on the device (rng.cuh); the result of a batch then
// depends on (input, rng_seed) only -- not on the handle, rank or order that processed it (SURVEY.md 8e "RNG across
// ranks": per-unit states from (rand_seed, event, module)).
LSB_EXPORT int lsb_chain_set_rng_fresh(lsb_chain* h, int32_t fresh) {
    LSB_REQUIRE(h, "chain_set_rng_fresh: null handle");
    h->rng_fresh = fresh ? 1 : 0;
    return 0;
}
// `signals` rows are stored sparsely by the fused MC stage (only the ticks covered by a pair's samples are written; the
// rest is zero by definition and no later stage reads it).  A caller who wants to look at the dense [S, P, T] array --
// lsb_chain_result.signals -- calls this first: it zero-fills the unwritten parts of the last batch (idempotent).
LSB_EXPORT int lsb_chain_signals_dense(lsb_chain* h, void* stream) {
    LSB_REQUIRE(h, "chain_signals_dense: null handle");
    if (h->signals_dense || h->last_rows <= 0 || h->last_T <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    k_signals_dense<<<(unsigned)h->last_rows, 128, 0, st>>>((float*)h->signals.p, (const int2*)h->sig_ranges.p, h->last_rows, h->last_T);
    LSB_LAUNCH_CHECK("k_signals_dense");
    h->signals_dense = 1;
    return 0;
}
LSB_EXPORT void lsb_chain_destroy(lsb_chain* h) {
    if (!h) return;
    h->sig_ranges.release(); h->response_split.release();
    DevBuf* all[] = {&h->tracks, &h->scal, &h->active, &h->neigh, &h->nrad, &h->npl, &h->uniq, &h->uniq_ws, &h->starts, &h->signals,
                     &h->mc_ws, &h->pim, &h->tpm, &h->psig, &h->pts, &h->oflow, &h->tticks, &h->integral, &h->adc_digit,
                     &h->adc_ticks, &h->cf, &h->thr, &h->rng, &h->nhits, &h->sx_slot, &h->sx_counts, &h->sx_cursor, &h->sx_raw,
                     &h->sx_offs, &h->sx_bsums, &h->sx_sorted, &h->arena_buf};
    for (DevBuf* b : all) b->release();
    for (int i = 0; i <= ST_COUNT; i++) cudaEventDestroy(h->ev[i]);
    cudaStreamSynchronize(h->hp); cudaStreamSynchronize(h->lp);
    cudaStreamDestroy(h->hp);          // h->lp is shared by all handles
    cudaEventDestroy(h->ev_in); cudaEventDestroy(h->ev_front); cudaEventDestroy(h->ev_mc); cudaEventDestroy(h->ev_done);
    for (int i = 0; i < 6; i++) cudaEventDestroy(h->tl[i]);
    if (h->hs_pinned) cudaFreeHost(h->hs_pinned);
    delete h;
}

cudaEvent_t lsb_reference_event();

// cli/simulate_pixels.py:92-104 maybe_create_rng_states: keep evolved states, append fresh ones (made on the device)
static int chain_grow_rng(lsb_chain* h, long long n, uint64_t seed, cudaStream_t st) {
    if (n <= h->n_rng) return 0;
    if ((size_t)n * 16 > h->rng.cap) {
        DevBuf nb;
        if (nb.need((size_t)n * 16)) return -1;
        if (h->n_rng) {
            LSB_CUDA(cudaMemcpyAsync(nb.p, h->rng.p, (size_t)h->n_rng * 16, cudaMemcpyDeviceToDevice, st));
            LSB_CUDA(cudaStreamSynchronize(st));              // the old buffer is freed below
        }
        h->rng.release();
        h->rng = nb;
    }
    int rc = rng_create_states_dev((unsigned long long*)h->rng.p + 2 * h->n_rng, n - h->n_rng, seed, 0, st);
    if (rc) return rc;
    h->n_rng = n;
    return 0;
}

// LSB_CHAIN_TRACE=1: host-side time between the stage markers of one enqueue (allocation growth, host loops), on stderr
#include <chrono>
static double g_ch_trace_t0 = 0.0;
static inline void ch_trace(int i) {
    static const int on = getenv("LSB_CHAIN_TRACE") ? 1 : 0;
    if (!on) return;
    const double now = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    if (i > 0 && now - g_ch_trace_t0 > 1.0) fprintf(stderr, "[lsb chain trace] host %.1f ms before marker %d\n", now - g_ch_trace_t0, i);
    g_ch_trace_t0 = now;
}
// NVTX ranges with the reference's names (cli/simulate_pixels.py:917-1105 RangePush / RangePop), host side like there
#include <nvtx3/nvToolsExt.h>
static const char* const k_chain_range_names[ST_COUNT + 1] = {"quench+drift", "max_pixels/get_pixels", "unique_pix", "time_intervals", "tracks_current",
                                                              "pixel_index_map/track_pixel_map", "sum_pixels_signals", "get_adc_values", "digitize", nullptr};
struct ChainRange {
    bool open = false;
    void next(int i) { if (open) nvtxRangePop(); open = false; if (i >= 0 && i < ST_COUNT) { nvtxRangePushA(k_chain_range_names[i]); open = true; } }
    ~ChainRange() { if (open) nvtxRangePop(); }
};
#define CH_STAGE(i) do { ch_trace(i); nvtx_range.next(i); if (h->timing) cudaEventRecord(h->ev[i], st); } while (0)

// Enqueue one batch.  `st` = stream of the front and FEE stages, `st_mc` = stream of the MC stage (may be the
// same).  Returns with everything queued; the hit count arrives in h->hs_pinned once `st` has drained.
static int chain_enqueue(lsb_chain* h, void* tracks_dev, int64_t S, int32_t quench_mode, uint64_t rng_seed,
                         int32_t n_events, lsb_chain_result* out, cudaStream_t st, cudaStream_t st_mc) {
    LSB_REQUIRE(h && out && (tracks_dev || S == 0), "chain_run: null pointer");
    const lsb_consts* c = &h->c;
    const lsb_track_layout* L = &h->L;
    memset(out, 0, sizeof(*out));
    out->n_segments = S;
    if (S == 0) return 0;
    ChainRange nvtx_range;
    const int K = c->max_tracks_per_pixel, A = c->max_adc_values, Tt = c->n_time_ticks;
    int rc;
    if ((rc = h->scal.need(sizeof(ChainScalars)))) return rc;
    ChainScalars* d_s = (ChainScalars*)h->scal.p;
    ChainScalars hs;
    LSB_CUDA(cudaMemsetAsync(d_s, 0, sizeof(ChainScalars), st));
    CH_STAGE(0);
    if (st_mc != st) cudaEventRecord(h->tl[0], st);
    // ---- quench, drift (simulate_pixels.py:732,742) -------------------------------------
    if (quench_mode >= 0) {                                  // < 0: the caller has quenched and drifted the records already
        if ((rc = lsb_quench(c, L, tracks_dev, S, quench_mode, st))) return rc;
        if ((rc = lsb_drift(c, L, tracks_dev, S, st))) return rc;
    }
    CH_STAGE(1);
    // ---- max_radius, max_pixels (:918-928) ----------------------------------------------
    k_chain_max_tran<<<lsb_blocks(S, 256), 256, 0, st>>>(make_layout(L), (const char*)tracks_dev, S, &d_s->max_tran_bits, &d_s->sum_len);
    LSB_LAUNCH_CHECK("k_chain_max_tran");
    if ((rc = lsb_max_pixels(c, L, tracks_dev, S, (int64_t*)&d_s->max_pixels, st))) return rc;
    LSB_CUDA(cudaMemcpyAsync(&hs, d_s, sizeof(hs), cudaMemcpyDeviceToHost, st));
    LSB_CUDA(cudaStreamSynchronize(st));
    double max_tran; memcpy(&max_tran, &hs.max_tran_bits, 8);
    const double sum_len = hs.sum_len;
    long long mc_total_host = -1;
    const int radius = (int)ceil(max_tran * 5 / c->pixel_pitch);
    const long long maxpix = hs.max_pixels;
    const long long P = (2LL * radius + 1) * maxpix + (1 + 2LL * radius) * radius * 2;
    out->max_active = maxpix; out->max_neighbors = P;
    if (maxpix == 0 || P == 0) return 0;                                              // :935-941
    LSB_REQUIRE(P < 100000, "chain_run: unreasonable neighbour row length");
    // ---- get_pixels (:930-950) ----------------------------------------------------------
    if ((rc = h->active.need((size_t)S * maxpix * 4)) || (rc = h->neigh.need((size_t)S * P * 4)) ||
        (rc = h->nrad.need((size_t)S * P * 4)) || (rc = h->npl.need((size_t)S * 8))) return rc;
    k_fill_i32<<<lsb_blocks(S * maxpix, 256), 256, 0, st>>>((int32_t*)h->active.p, S * maxpix, -1); LSB_LAUNCH_CHECK("k_fill_i32");
    k_fill_i32<<<lsb_blocks(S * P, 256), 256, 0, st>>>((int32_t*)h->neigh.p, S * P, -1); LSB_LAUNCH_CHECK("k_fill_i32");
    k_fill_i32<<<lsb_blocks(S * P, 256), 256, 0, st>>>((int32_t*)h->nrad.p, S * P, -1); LSB_LAUNCH_CHECK("k_fill_i32");
    LSB_CUDA(cudaMemsetAsync(h->npl.p, 0, (size_t)S * 8, st));
    if ((rc = lsb_get_pixels(c, L, tracks_dev, S, (int32_t*)h->active.p, (int32_t)maxpix, (int32_t*)h->neigh.p, (int32_t*)h->nrad.p,
                             (int32_t)P, (double*)h->npl.p, radius, st))) return rc;
    CH_STAGE(2);
    // ---- unique pixels (:953-956), time_intervals (:1000-1002), max distance class (:1040) --
    const long long max_id = (long long)c->n_pixels[0] * c->n_pixels[1] * c->n_tpc - 1;
    const long long ws_bytes = lsb_unique_pixels_workspace_bytes(max_id);
    long long ucap = S * P < max_id + 1 ? S * P : max_id + 1;
    if ((rc = h->uniq_ws.need((size_t)ws_bytes)) || (rc = h->uniq.need((size_t)ucap * 4)) || (rc = h->starts.need((size_t)S * 8))) return rc;
    if ((rc = lsb_unique_pixels((const int32_t*)h->neigh.p, S * P, max_id, (int32_t*)h->uniq.p, (int64_t*)&d_s->n_unique, h->uniq_ws.p,
                                ws_bytes, st))) return rc;
    CH_STAGE(3);
    if ((rc = lsb_time_intervals(c, L, tracks_dev, S, (double*)h->starts.p, (int64_t*)&d_s->t_max, st))) return rc;
    k_max_i32<<<lsb_blocks(S * P, 256), 256, 0, st>>>((const int32_t*)h->nrad.p, S * P, &d_s->max_dist); LSB_LAUNCH_CHECK("k_max_i32");
    LSB_CUDA(cudaMemcpyAsync(&hs, d_s, sizeof(hs), cudaMemcpyDeviceToHost, st));
    LSB_CUDA(cudaStreamSynchronize(st));
    const long long U = hs.n_unique, T = hs.t_max;
    const int max_distance = hs.max_dist + 1;
    out->n_unique_pixels = U; out->n_ticks = T;
    if (U == 0 || T <= 0) return 0;
    LSB_REQUIRE(T < 2147483647LL, "chain_run: tick count overflow");
    CH_STAGE(4);
    // ---- tracks_current_mc (:1007-1016) -------------------------------------------------
    if ((rc = h->signals.need((size_t)S * P * T * 4))) return rc;
    if ((rc = h->sig_ranges.need((size_t)S * P * sizeof(int2)))) return rc;
    h->signals_dense = 0; h->last_rows = S * P; h->last_T = (int)T;
    // states for the MC stage (index = segment + S * pixel slot, detsim.py:324) and for get_adc_values (index = pixel,
    // fee.py:557; the reference asks for TPB * BPG = roundup128(U) of them).  Both requests are served BEFORE the MC stage
    // is queued: it advances the states on its own stream.
    long long need_rng = S * P;
    long long need_rng2 = 128LL * ((U + 127) / 128);
    if (h->rng_fresh) {
        const long long n = need_rng > need_rng2 ? need_rng : need_rng2;
        h->n_rng = 0;
        if ((rc = h->rng.need((size_t)n * 16))) return rc;
        if ((rc = rng_create_states_dev((unsigned long long*)h->rng.p, need_rng, rng_seed, 0, st))) return rc;
        if (n > need_rng && (rc = rng_create_states_dev((unsigned long long*)h->rng.p + 2 * need_rng, n - need_rng, rng_seed, 0, st))) return rc;
        h->n_rng = n;
    } else {
        if ((rc = chain_grow_rng(h, need_rng, rng_seed, st))) return rc;
        if ((rc = chain_grow_rng(h, need_rng2, rng_seed, st))) return rc;
    }
    if (st_mc != st) { LSB_CUDA(cudaEventRecord(h->ev_front, st)); LSB_CUDA(cudaStreamWaitEvent(st_mc, h->ev_front, 0)); cudaEventRecord(h->tl[1], st); cudaEventRecord(h->tl[2], st_mc); }
    if (st_mc != st) {
        // The MC kernels re-read the response table (15.8 MB) ~2000 times per batch and need it L2-resident,
        // while the FEE stage of the previous batch streams GBs through L2 at the same time: pin the table
        // with a persisting access-policy window on the MC stream.
        static size_t persist_by_dev[64]; static bool persist_init[64] = {false};
        int dev = 0; cudaGetDevice(&dev); dev = (dev < 0 || dev >= 64) ? 0 : dev;
        size_t& persist_max = persist_by_dev[dev];
        if (!persist_init[dev]) {
            persist_init[dev] = true;
            cudaDeviceProp prop; persist_max = 0;
            if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) persist_max = (size_t)prop.persistingL2CacheMaxSize;
            if (getenv("LSB_NO_L2_PERSIST")) persist_max = 0;        // tuning / A-B switch
            if (persist_max > (64u << 20)) persist_max = 64u << 20;   // the table needs 16-32 MB; leave the rest of L2 alone
            if (persist_max) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist_max);
            (void)cudaGetLastError();
        }
        const size_t lut_bytes = (size_t)h->Rx * h->Ry * h->Rt * (h->f64 ? 8 : 4);
        if (persist_max) {
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof(attr));
            attr.accessPolicyWindow.base_ptr = h->response_split.p ? h->response_split.p : const_cast<void*>(h->response);
            attr.accessPolicyWindow.num_bytes = lut_bytes;
            attr.accessPolicyWindow.hitRatio = lut_bytes <= persist_max ? 1.0f : (float)((double)persist_max / (double)lut_bytes);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(st_mc, cudaStreamAttributeAccessPolicyWindow, &attr);
            (void)cudaGetLastError();
        }
    }
    {
        cudaStream_t st = st_mc;                                   // MC stage
        // sample workspace from an upper bound of the sample count (no host round trip inside the MC stage):
        // every (segment,pixel) takes at most len/MIN_STEP_SIZE + 1 steps
        const double bound_d = (sum_len / c->min_step_size + (double)S) * (double)P * (double)c->mc_sample_multiplier * 1.001 + 1024.0;
        if (h->rng_mode == 0 && bound_d < 4e9) {
            long long wsb = lsb_tracks_current_mc_workspace_bytes(S, (int32_t)P, (long long)bound_d);
            if ((size_t)wsb > h->mc_ws.cap) { if ((rc = h->mc_ws.need((size_t)wsb))) return rc; }
            if ((rc = mc_run_nosync(c, L, tracks_dev, S, (const int32_t*)h->neigh.p, (int32_t)P, (float*)h->signals.p, (int32_t)T,
                                    h->response, h->response_split.p, h->Rx, h->Ry, h->Rt, h->f64, (uint64_t*)h->rng.p, S, h->mc_ws.p,
                                    (int64_t)h->mc_ws.cap, &d_s->mc_total, &d_s->mc_overflow, (int2*)h->sig_ranges.p, &d_s->mc_nfma, &d_s->mc_npairs, d_s->mc_diag, st))) return rc;
        } else {
            LSB_CUDA(cudaMemsetAsync(h->signals.p, 0, (size_t)S * P * T * 4, st)); LSB_MARK("memset_signals", st);
            k_ranges_full<<<lsb_blocks(S * P, 256), 256, 0, st>>>((int2*)h->sig_ranges.p, S * P, (int)T);
            LSB_LAUNCH_CHECK("k_ranges_full");
            h->signals_dense = 1;
            long long guess = S * 4000LL;
            long long wsb = lsb_tracks_current_mc_workspace_bytes(S, (int32_t)P, guess);
            if ((size_t)wsb > h->mc_ws.cap) { if ((rc = h->mc_ws.need((size_t)wsb))) return rc; }
            if ((rc = lsb_tracks_current_mc(c, L, tracks_dev, S, (const int32_t*)h->neigh.p, (int32_t)P, (float*)h->signals.p, (int32_t)T,
                                            h->response, h->Rx, h->Ry, h->Rt, h->f64, (uint64_t*)h->rng.p, h->n_rng, S, h->rng_mode,
                                            h->mc_ws.p, (int64_t)h->mc_ws.cap, st))) return rc;
            mc_total_host = lsb_tracks_current_mc_last_samples();
        }
        out->n_samples = mc_total_host;
    }
    if (st_mc != st) { LSB_CUDA(cudaEventRecord(h->ev_mc, st_mc)); cudaEventRecord(h->tl[3], st_mc); }
    else CH_STAGE(5);
    // temporaries of the remaining stages (all queued on `st`) come from this handle's arena
    struct ArenaScope {
        lsb_chain* h; TmpArena* prev;
        ArenaScope(lsb_chain* hh, size_t want_bytes) : h(hh), prev(g_lsb_arena) {
            if (h->arena.overflow) {                        // last batch did not fit: grow (synchronises, warm-up only)
                size_t want = h->arena.high_water + h->arena.overflow + (64u << 20);
                h->arena_buf.need(want);
                h->arena.overflow = 0;
            }
            if (want_bytes > h->arena_buf.cap) h->arena_buf.need(want_bytes);
            h->arena.base = (char*)h->arena_buf.p; h->arena.cap = h->arena_buf.cap; h->arena.off = 0;
            g_lsb_arena = h->arena.base ? &h->arena : nullptr;
        }
        ~ArenaScope() { g_lsb_arena = prev; }
    } arena_scope(h, fee_scratch_bytes(c, U, Tt, A, S * P) + (size_t)S * P * 24 + (size_t)U * 32 + (16u << 20));
    // ---- pixel_index_map (:1021-1025), track_pixel_map (:1031-1042) ---------------------
    if ((rc = h->pim.need((size_t)S * P * 8)) || (rc = h->tpm.need((size_t)U * K * 8))) return rc;
    if ((rc = lsb_pixel_index_map((const int32_t*)h->neigh.p, S * P, max_id, h->uniq_ws.p, (int64_t*)h->pim.p, st))) return rc;
    k_fill_i64<<<lsb_blocks(U * K, 256), 256, 0, st>>>((long long*)h->tpm.p, U * K, -1); LSB_LAUNCH_CHECK("k_fill_i64");
    if ((rc = lsb_get_track_pixel_map2((int64_t*)h->tpm.p, K, (const int32_t*)h->uniq.p, U, (const int32_t*)h->neigh.p,
                                       (const int32_t*)h->nrad.p, S, (int32_t)P, max_distance, st))) return rc;
    CH_STAGE(6);
    // ---- sum_pixel_signals (:1052-1063) -------------------------------------------------
    // default: the dense per-segment tensor pixels_tracks_signals [U][Tt][K] (0.8 MB per pixel) is NOT
    // materialised; get_adc_values reads the per-segment waveforms straight from `signals` through the
    // (pixel, slot) entry list.  dense=1 reproduces the reference's buffers.
    if ((rc = h->psig.need((size_t)U * Tt * 8)) || (rc = h->oflow.need((size_t)U * 8))) return rc;
    if (h->dense && (rc = h->pts.need((size_t)U * Tt * K * 8))) return rc;
    if (h->dense) { LSB_CUDA(cudaMemsetAsync(h->psig.p, 0, (size_t)U * Tt * 8, st)); LSB_MARK("memset_psig", st); }   // else: written fresh by the sum kernel
    if (h->dense) { LSB_CUDA(cudaMemsetAsync(h->pts.p, 0, (size_t)U * Tt * K * 8, st)); LSB_MARK("memset_pts", st); }
    LSB_CUDA(cudaMemsetAsync(h->oflow.p, 0, (size_t)U * 8, st));
    SumCtx sx;
    {
        long long ne = S * P;
        if ((rc = h->sx_slot.need((size_t)ne * 4)) || (rc = h->sx_counts.need((size_t)U * 4)) || (rc = h->sx_cursor.need((size_t)U * 4)) ||
            (rc = h->sx_raw.need((size_t)ne * 4)) || (rc = h->sx_offs.need((size_t)U * 8)) ||
            (rc = h->sx_bsums.need((size_t)(scan_num_blocks(U) + 1) * 8)) || (rc = h->sx_sorted.need((size_t)ne * sizeof(SumEntry)))) return rc;
        sx.slot_of = (int*)h->sx_slot.p; sx.counts = (int*)h->sx_counts.p; sx.cursor = (int*)h->sx_cursor.p; sx.raw = (int*)h->sx_raw.p;
        sx.offs = (long long*)h->sx_offs.p; sx.bsums = (long long*)h->sx_bsums.p; sx.sorted = (SumEntry*)h->sx_sorted.p;
        if ((rc = lsb_upload_consts(c, st))) return rc;
        if ((rc = sum_build_entries(sx, U, S, (int)P, (const double*)h->starts.p, (const long long*)h->pim.p, (const long long*)h->tpm.p, K,
                                    (double*)h->oflow.p, nullptr, (int)T, st))) return rc;
        if (st_mc != st) { LSB_CUDA(cudaStreamWaitEvent(st, h->ev_mc, 0)); cudaEventRecord(h->tl[4], st); }
        k_sum_apply_ranges<<<lsb_blocks(U, 128), 128, 0, st>>>(sx.offs, sx.counts, U, (const int2*)h->sig_ranges.p, sx.sorted);
        LSB_LAUNCH_CHECK("k_sum_apply_ranges");
        if ((rc = sum_run(sx, (double*)h->psig.p, U, Tt, (const float*)h->signals.p, (int)T, K, h->dense ? (double*)h->pts.p : nullptr, st,
                          !h->dense))) return rc;
    }
    CH_STAGE(7);
    // ---- get_adc_values (:1072-1095) ----------------------------------------------------
    if (h->tticks_n != Tt + 1 || h->tticks_events != n_events) {
        // numpy/cupy linspace(0, n_events*TIME_INTERVAL[1], Tt+1): i*step, last element = stop
        if ((rc = h->tticks.need((size_t)(Tt + 1) * 8))) return rc;
        double* tt = (double*)malloc(sizeof(double) * (size_t)(Tt + 1));
        if (!tt) return lsb_fail_arg("chain: out of host memory");
        double stop = (double)n_events * c->time_interval[1];
        double step = stop / (double)Tt;
        for (int i = 0; i <= Tt; i++) tt[i] = (double)i * step + 0.0;
        tt[Tt] = stop;
        cudaError_t e = cudaMemcpyAsync(h->tticks.p, tt, sizeof(double) * (size_t)(Tt + 1), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        free(tt);
        if (e != cudaSuccess) return lsb_fail_cuda(e, "time_ticks upload");
        h->tticks_n = Tt + 1; h->tticks_events = n_events;
    }
    if ((rc = h->integral.need((size_t)U * A * 8)) || (rc = h->adc_digit.need((size_t)U * A * 8)) ||
        (rc = h->adc_ticks.need((size_t)U * A * 8)) || (rc = h->cf.need((size_t)U * A * K * 8)) || (rc = h->thr.need((size_t)U * 8))) return rc;
    LSB_CUDA(cudaMemsetAsync(h->integral.p, 0, (size_t)U * A * 8, st));
    LSB_CUDA(cudaMemsetAsync(h->adc_ticks.p, 0, (size_t)U * A * 8, st));
    LSB_CUDA(cudaMemsetAsync(h->cf.p, 0, (size_t)U * A * K * 8, st)); LSB_MARK("memset_cf", st);
    k_fill_f64<<<lsb_blocks(U, 256), 256, 0, st>>>((double*)h->thr.p, U, c->discrimination_threshold * c->unit_e); LSB_LAUNCH_CHECK("k_fill_f64");
    {
        FeeSparse sp; sp.signals = (const float*)h->signals.p; sp.T = (int)T; sp.offs = sx.offs; sp.counts = sx.counts;
        sp.sorted = sx.sorted; sp.n_entries_cap = S * P; sp.exact = h->exact_fractions;
        LSB_REQUIRE(h->n_rng >= U, "chain_run: rng_states shorter than the number of pixels");
        if ((rc = fee_run(c, (const double*)h->psig.p, h->dense ? (const double*)h->pts.p : nullptr, h->dense ? nullptr : &sp, U, Tt, K,
                          (const double*)h->tticks.p, Tt + 1, (double*)h->integral.p, (double*)h->adc_ticks.p, A, 0.0,
                          (uint64_t*)h->rng.p, (double*)h->cf.p, (const double*)h->thr.p, st))) return rc;
    }
    CH_STAGE(8);
    // ---- digitize (:1102) ---------------------------------------------------------------
    if ((rc = lsb_digitize(c, (const double*)h->integral.p, nullptr, U * A, (double*)h->adc_digit.p, st))) return rc;
    {
        double g = c->gain * c->unit_mV / c->unit_e;
        double v = 0.0 * g + c->v_pedestal * c->unit_mV - c->v_cm * c->unit_mV; v = v > 0 ? v : 0;
        double ped = nearbyint(v * c->adc_counts / (c->v_ref * c->unit_mV - c->v_cm * c->unit_mV));
        ped = ped < c->adc_counts - 1 ? ped : c->adc_counts - 1;
        k_count_hits<<<lsb_blocks(U * A, 256), 256, 0, st>>>((const double*)h->adc_digit.p, U * A, ped, &d_s->n_hits);
        LSB_LAUNCH_CHECK("k_count_hits");
    }
    CH_STAGE(9);
    LSB_CUDA(cudaMemcpyAsync(h->hs_pinned, d_s, sizeof(ChainScalars), cudaMemcpyDeviceToHost, st));
    if (st_mc != st) cudaEventRecord(h->tl[5], st);
    out->unique_pix = (const int32_t*)h->uniq.p; out->track_pixel_map = (const int64_t*)h->tpm.p;
    out->adc_list = (const double*)h->integral.p; out->adc_digit = (const double*)h->adc_digit.p;
    out->adc_ticks_list = (const double*)h->adc_ticks.p; out->current_fractions = (const double*)h->cf.p;
    out->signals = (const float*)h->signals.p; out->pixels_signals = (const double*)h->psig.p;
    return 0;
}
// after `st` has drained: hit count, stage times
static void chain_finish(lsb_chain* h, lsb_chain_result* out, bool with_timing) {
    if (out->unique_pix) {
        out->n_hits = h->hs_pinned->n_hits;
        if (out->n_samples < 0) out->n_samples = h->hs_pinned->mc_total;
        out->n_fma = (int64_t)h->hs_pinned->mc_nfma;
        out->n_pairs = (int64_t)h->hs_pinned->mc_npairs;
        out->n_groups = (int64_t)h->hs_pinned->mc_diag[0]; out->n_edge = (int64_t)h->hs_pinned->mc_diag[1]; out->n_irregular = (int64_t)h->hs_pinned->mc_diag[2];
    }
    if (h->timing && with_timing && out->unique_pix)
        for (int i = 0; i < ST_COUNT; i++) { float ms = 0; if (cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]) == cudaSuccess) out->stage_ms[i] = ms; }
}

LSB_EXPORT int lsb_chain_run(lsb_chain* h, void* tracks_dev, int64_t S, int32_t quench_mode, uint64_t rng_seed,
                             int32_t n_events, lsb_chain_result* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = chain_enqueue(h, tracks_dev, S, quench_mode, rng_seed, n_events, out, st, st);
    if (rc) return rc;
    LSB_CUDA(cudaStreamSynchronize(st));
    chain_finish(h, out, true);
    LSB_REQUIRE(!(out->unique_pix && h->hs_pinned->mc_overflow), "chain_run: MC sample bound exceeded (internal error)");
    return 0;
}

// Pipelined form: returns as soon as the batch is queued on the handle's own streams (ordered after the work
// already queued on `stream`); lsb_chain_wait blocks until it is complete and returns the result.  Use two
// handles alternately to overlap the FEE stage of one batch with the MC stage of the next.
LSB_EXPORT int lsb_chain_run_async(lsb_chain* h, void* tracks_dev, int64_t S, int32_t quench_mode, uint64_t rng_seed,
                                   int32_t n_events, void* stream) {
    LSB_REQUIRE(h, "chain_run_async: null handle");
    LSB_REQUIRE(!h->pending_valid, "chain_run_async: previous batch not collected (call lsb_chain_wait)");
    // order this batch after the work already queued on the caller's stream (e.g. the upload of `tracks_dev`);
    // an idle caller stream needs no dependency
    if (cudaStreamQuery((cudaStream_t)stream) != cudaSuccess) {
        (void)cudaGetLastError();
        LSB_CUDA(cudaEventRecord(h->ev_in, (cudaStream_t)stream));
        LSB_CUDA(cudaStreamWaitEvent(h->hp, h->ev_in, 0));
    }
    int rc = chain_enqueue(h, tracks_dev, S, quench_mode, rng_seed, n_events, &h->pending, h->hp, h->lp);
    if (rc) return rc;
    LSB_CUDA(cudaEventRecord(h->ev_done, h->hp));
    h->pending_valid = 1;
    return 0;
}
LSB_EXPORT int lsb_chain_wait(lsb_chain* h, lsb_chain_result* out) {
    LSB_REQUIRE(h && out, "chain_wait: null pointer");
    LSB_REQUIRE(h->pending_valid, "chain_wait: nothing pending");
    LSB_CUDA(cudaEventSynchronize(h->ev_done));
    chain_finish(h, &h->pending, false);
    if (h->pending.unique_pix) {
        // timeline of this batch in ms since the library's reference event: front begin/end, MC begin/end, FEE begin, done
        // (the markers are only recorded when the MC stage ran on its own stream)
        for (int i = 0; i < 6; i++) { float ms = 0; if (cudaEventElapsedTime(&ms, lsb_reference_event(), h->tl[i]) == cudaSuccess) h->pending.stage_ms[i] = ms; }
        (void)cudaGetLastError();
    }
    *out = h->pending;
    h->pending_valid = 0;
    LSB_REQUIRE(!(out->unique_pix && h->hs_pinned->mc_overflow), "chain_wait: MC sample bound exceeded (internal error)");
    return 0;
}

LSB_EXPORT int lsb_chain_run_host(lsb_chain* h, void* tracks_host, int64_t S, int32_t quench_mode, uint64_t rng_seed,
                                  int32_t n_events, int32_t* unique_pix_host, double* adc_digit_host, double* adc_ticks_host,
                                  int64_t U_cap, lsb_chain_result* out, void* stream) {
    LSB_REQUIRE(h && out && (tracks_host || S == 0), "chain_run_host: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    size_t bytes = (size_t)S * h->L.itemsize;
    int rc;
    if ((rc = h->tracks.need(bytes ? bytes : 1))) return rc;
    if (S) LSB_CUDA(cudaMemcpyAsync(h->tracks.p, tracks_host, bytes, cudaMemcpyHostToDevice, st));
    if ((rc = lsb_chain_run(h, h->tracks.p, S, quench_mode, rng_seed, n_events, out, st))) return rc;
    // Numba semantics: the record array is copied back after quench/drift modified it in place
    if (S) LSB_CUDA(cudaMemcpyAsync(tracks_host, h->tracks.p, bytes, cudaMemcpyDeviceToHost, st));
    const long long U = out->n_unique_pixels, A = h->c.max_adc_values;
    if (U > 0 && out->unique_pix) {
        LSB_REQUIRE(U <= U_cap, "chain_run_host: U_cap too small");
        if (unique_pix_host) LSB_CUDA(cudaMemcpyAsync(unique_pix_host, out->unique_pix, (size_t)U * 4, cudaMemcpyDeviceToHost, st));
        if (adc_digit_host) LSB_CUDA(cudaMemcpyAsync(adc_digit_host, out->adc_digit, (size_t)U * A * 8, cudaMemcpyDeviceToHost, st));
        if (adc_ticks_host) LSB_CUDA(cudaMemcpyAsync(adc_ticks_host, out->adc_ticks_list, (size_t)U * A * 8, cudaMemcpyDeviceToHost, st));
    }
    LSB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// pipelined host-buffer form: H2D, chain and D2H are queued on the handle's streams; collect with lsb_chain_wait
LSB_EXPORT int lsb_chain_run_host_async(lsb_chain* h, void* tracks_host, int64_t S, int32_t quench_mode, uint64_t rng_seed,
                                        int32_t n_events, int32_t* unique_pix_host, double* adc_digit_host,
                                        double* adc_ticks_host, int64_t U_cap) {
    LSB_REQUIRE(h && (tracks_host || S == 0), "chain_run_host_async: null pointer");
    LSB_REQUIRE(!h->pending_valid, "chain_run_host_async: previous batch not collected (call lsb_chain_wait)");
    cudaStream_t st = h->hp;
    size_t bytes = (size_t)S * h->L.itemsize;
    int rc;
    if ((rc = h->tracks.need(bytes ? bytes : 1))) return rc;
    if (S) LSB_CUDA(cudaMemcpyAsync(h->tracks.p, tracks_host, bytes, cudaMemcpyHostToDevice, st));
    lsb_chain_result* out = &h->pending;
    if ((rc = chain_enqueue(h, h->tracks.p, S, quench_mode, rng_seed, n_events, out, st, h->lp))) return rc;
    if (S) LSB_CUDA(cudaMemcpyAsync(tracks_host, h->tracks.p, bytes, cudaMemcpyDeviceToHost, st));
    const long long U = out->n_unique_pixels, A = h->c.max_adc_values;
    if (U > 0 && out->unique_pix) {
        LSB_REQUIRE(U <= U_cap, "chain_run_host_async: U_cap too small");
        if (unique_pix_host) LSB_CUDA(cudaMemcpyAsync(unique_pix_host, out->unique_pix, (size_t)U * 4, cudaMemcpyDeviceToHost, st));
        if (adc_digit_host) LSB_CUDA(cudaMemcpyAsync(adc_digit_host, out->adc_digit, (size_t)U * A * 8, cudaMemcpyDeviceToHost, st));
        if (adc_ticks_host) LSB_CUDA(cudaMemcpyAsync(adc_ticks_host, out->adc_ticks_list, (size_t)U * A * 8, cudaMemcpyDeviceToHost, st));
    }
    LSB_CUDA(cudaEventRecord(h->ev_done, st));
    h->pending_valid = 1;
    return 0;
}
