// light_trigger.cuh -- light trigger search and waveform digitisation: light_sim.get_triggers (light_sim.py:380-477),
// sim_triggers (:545-619, zero-noise part) and the digitize_signal kernel (:480-543) with interp (:241-271).
//
//   k_lt_block_above   thread per (trigger group, block of `sf` ticks): the group's channel sum per tick in the
//                      signal's own precision (sequential over the channels, like a NumPy reduction over a
//                      non-contiguous axis), its float64 mean over the block with NumPy's pairwise schedule
//                      (zero padded at the end), compared with the group threshold.
//   k_lt_search        thread per module: OR of its groups' flags and the sequential trigger search with the
//                      digitisation dead time -- including the reference's index bookkeeping (:441-452: the
//                      remaining waveform is re-sliced by an absolute index), replicated, not corrected.
//   k_lt_digitize      thread per (trigger, channel, sample): linear interpolation of the (virtually zero-padded,
//                      channel-sorted) waveform at isample * spacing / tick, truth propagation, rounding to
//                      LIGHT_NBIT bits.  Padding and the rows of channels without a simulated waveform are not
//                      materialised: a row map + tick offset describe them.
#pragma once
#include "common.cuh"
#include "glue.cuh"

template <typename TS>
__device__ __forceinline__ double lt_group_sum(const TS* __restrict__ signal, long long nticks, int g, int cpt, long long t) {
    TS acc = signal[(long long)(g * cpt) * nticks + t];
    for (int c = 1; c < cpt; c++) acc = acc + signal[(long long)(g * cpt + c) * nticks + t];     // -fmad=false: plain adds
    return (double)acc;
}
#define LT_MAX_SF 136
template <typename TS>
__global__ void k_lt_block_above(const TS* __restrict__ signal, int ngrp, long long nticks, int cpt, int sf, long long nblk,
                                 const double* __restrict__ group_threshold, uint8_t* __restrict__ above) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= ngrp * nblk) return;
    const int g = (int)(i / nblk);
    const long long b = i - g * nblk;
    // numpy add.reduce over the sf contiguous values of the block (see packets.cuh: np_sum_schedule)
    double res;
    const long long t0 = b * sf;
    auto val = [&](int k) -> double { const long long t = t0 + k; return t < nticks ? lt_group_sum(signal, nticks, g, cpt, t) : 0.0; };
    if (sf < 8) {
        res = 0.0;
        for (int k = 0; k < sf; k++) res += val(k);
    } else {
        double r[8];
        for (int k = 0; k < 8; k++) r[k] = val(k);
        int k = 8;
        for (; k < sf - (sf % 8); k += 8)
            for (int j = 0; j < 8; j++) r[j] += val(k + j);
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; k < sf; k++) res += val(k);
    }
    above[i] = (res / (double)sf) < group_threshold[g] ? 1 : 0;
}
__global__ void k_lt_search(const uint8_t* __restrict__ above, int ngrp, long long nblk, long long nticks, int sf, int cpt, int ndet,
                            const int32_t* __restrict__ chan_module, int n_mod, long long digit_ticks, int max_trig,
                            long long* __restrict__ trig_idx, int32_t* __restrict__ n_trig) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_mod) return;
    auto flag = [&](long long b) -> bool {
        for (int g = 0; g < ngrp; g++) {
            if (!above[g * nblk + b]) continue;
            for (int c = 0; c < cpt; c++) { const int d = g * cpt + c; if (d < ndet && chan_module[d] == m) return true; }
        }
        return false;
    };
    long long base = 0, last = 0;
    int n = 0;
    while (base < nticks) {
        long long b = base / sf, tick = -1;
        for (; b * sf < nticks; b++)
            if (flag(b)) { tick = b * sf > base ? b * sf : base; break; }
        if (tick < 0 || tick >= nticks) break;
        const long long idx = (tick - base) + last;
        if (n < max_trig) trig_idx[(long long)m * max_trig + n] = idx;
        n++;
        base += idx + digit_ticks;               // (sic) the remaining waveform is cut at the absolute index
        last = idx + digit_ticks;
    }
    n_trig[m] = n;
}

LSB_EXPORT int lsb_light_get_triggers(const void* signal, int32_t signal_f64, int32_t ndet, int64_t nticks, int32_t channels_per_group,
                                      int32_t sample_factor, const double* group_threshold, const int32_t* chan_module, int32_t n_modules,
                                      int64_t digit_ticks, int32_t max_trig, int64_t* trig_idx, int32_t* n_trig, void* stream) {
    LSB_REQUIRE(ndet >= 0 && nticks >= 0 && channels_per_group > 0 && sample_factor > 0 && sample_factor <= LT_MAX_SF && max_trig > 0,
                "light_get_triggers: bad sizes (sample factor must be 1..136)");
    if (n_modules == 0) return 0;
    LSB_REQUIRE(trig_idx && n_trig && chan_module, "light_get_triggers: null output");
    cudaStream_t st = (cudaStream_t)stream;
    const int ngrp = ndet / channels_per_group;
    const long long nblk = (nticks + (sample_factor - nticks % sample_factor)) / sample_factor;
    TmpPool pool(st);
    uint8_t* above;
    LSB_CUDA(pool.get(&above, (long long)ngrp * nblk));
    if (ngrp > 0 && nticks > 0) {
        LSB_REQUIRE(signal && group_threshold, "light_get_triggers: null input");
        if (signal_f64)
            k_lt_block_above<double><<<lsb_blocks(ngrp * nblk, 128), 128, 0, st>>>((const double*)signal, ngrp, nticks, channels_per_group,
                                                                                  sample_factor, nblk, group_threshold, above);
        else
            k_lt_block_above<float><<<lsb_blocks(ngrp * nblk, 128), 128, 0, st>>>((const float*)signal, ngrp, nticks, channels_per_group,
                                                                                 sample_factor, nblk, group_threshold, above);
        LSB_LAUNCH_CHECK("k_lt_block_above");
    }
    k_lt_search<<<lsb_blocks(n_modules, 32), 32, 0, st>>>(above, ngrp, nblk, ngrp > 0 ? nticks : 0, sample_factor, channels_per_group, ndet,
                                                          chan_module, n_modules, digit_ticks, max_trig, (long long*)trig_idx, n_trig);
    LSB_LAUNCH_CHECK("k_lt_search");
    return 0;
}

// ---- digitisation --------------------------------------------------------------------------------------
struct LtDigit {
    long long nticks, front, L;       // simulated ticks, zero ticks in front, padded length
    int nrow, M, Mout, ndm, nsamples;
    int diff_f32;                     // the waveform array the reference interpolates is float32 (no padding, no added rows)
    double step;                      // LIGHT_DIGIT_SAMPLE_SPACING / LIGHT_TICK_SIZE is evaluated per sample as i * spacing / tick
    double spacing, tick, truth_threshold, quantum;
    int truncate;
};
template <typename TS>
__device__ __forceinline__ double lt_interp(const TS* __restrict__ row, const LtDigit& p, double idx) {
    // interp(idx, arr, 0, 0) on the padded row (row == nullptr: a channel without waveform, all zeros)
    const long long i0 = (long long)floor(idx);
    if (i0 < 0 || i0 > p.L - 1) return 0.0;
    auto at = [&](long long i) -> TS { const long long t = i - p.front; return (row && t >= 0 && t < p.nticks) ? row[t] : (TS)0; };
    if ((double)i0 == idx) return (double)at(i0);
    if (i0 > p.L - 2) return 0.0;
    const TS v0 = at(i0), v1 = at(i0 + 1);
    const double d = p.diff_f32 ? (double)(v1 - v0) : ((double)v1 - (double)v0);
    return (double)v0 + d * (idx - (double)i0);
}
template <typename TS>
__global__ void k_lt_digitize(LtDigit p, const TS* __restrict__ signal, const long long* __restrict__ row_chan, const int32_t* __restrict__ row_src,
                              const long long* __restrict__ true_id, const double* __restrict__ true_ph, long long ntrig,
                              const long long* __restrict__ trig_chan, double* __restrict__ digit, long long* __restrict__ out_id,
                              double* __restrict__ out_ph) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= ntrig * p.ndm * p.nsamples) return;
    const int s = (int)(i % p.nsamples);
    const long long im = (i / p.nsamples) % p.ndm, it = i / ((long long)p.nsamples * p.ndm);
    const long long idet = trig_chan[it * p.ndm + im];
    int isig = p.nrow - 1;                                    // (sic) the search variable keeps the last row when nothing matches
    for (int k = 0; k < p.nrow; k++) if (row_chan[k] == idet) { isig = k; break; }
    if (p.nrow <= 0) return;
    const double st = (double)s * p.spacing / p.tick;
    const int src = row_src[isig];
    double v = lt_interp<TS>(src >= 0 ? signal + (long long)src * p.nticks : nullptr, p, st);
    if (p.truncate) v = rint(v / p.quantum) * p.quantum;      // cp.round: half to even
    digit[i] = v;
    if (p.M <= 0) return;
    const long long t0 = (long long)floor(st), t1 = (long long)ceil(st);
    // truth rows of the padded, channel-sorted arrays; -1 / 0 outside the simulated ticks and for added rows
    auto tid = [&](int row, long long t, int j) -> long long {
        const int sr = row_src[row]; const long long tt = t - p.front;
        return (sr >= 0 && tt >= 0 && tt < p.nticks && t < p.L) ? true_id[((long long)sr * p.nticks + tt) * p.M + j] : -1;
    };
    auto tph = [&](long long row, long long t, int j) -> double {
        if (row < 0 || row >= p.nrow) return 0.0;             // the reference indexes this read with the channel id (see below)
        const int sr = row_src[row]; const long long tt = t - p.front;
        return (sr >= 0 && tt >= 0 && tt < p.nticks && t < p.L) ? true_ph[((long long)sr * p.nticks + tt) * p.M + j] : 0.0;
    };
    long long* oid = out_id + i * p.Mout;
    double* oph = out_ph + i * p.Mout;
    int n = 0;
    for (int j = 0; j < p.M; j++) {
        if (n >= p.Mout) break;
        const long long id0 = tid(isig, t0, j);
        if (id0 == -1) break;
        double p0 = 0.0, p1 = 0.0;
        if (id0 == oid[n] || oid[n] == -1) {
            oid[n] = id0;
            n++;
            p0 = tph(idet, t0, j);                            // (sic) light_sim.py:521 indexes with the channel id, not the row
            if (fabs(p0) < p.truth_threshold) continue;
            if (id0 == tid(isig, t1, j)) p1 = tph(isig, t1, j);
            else for (int k = 0; k < p.M; k++) if (id0 == tid(isig, t1, k)) { p1 = tph(isig, t1, k); break; }
        }
        if (n >= 1 && oid[n - 1] != -1) {
            // interp(sample_tick - itick0, (photons0, photons1), 0, 0)
            const double x = st - (double)t0;
            const long long i0 = (long long)floor(x);
            double r;
            if (i0 < 0 || i0 > 1) r = 0.0;
            else if ((double)i0 == x) r = i0 == 0 ? p0 : p1;
            else if (i0 > 0) r = 0.0;
            else r = p0 + (p1 - p0) * (x - (double)i0);
            oph[n - 1] = r;
        }
    }
}

LSB_EXPORT int lsb_light_digitize(const void* signal, int32_t signal_f64, int64_t nticks, int32_t n_rows, const int64_t* row_channel,
                                  const int32_t* row_source, int64_t front_pad, int64_t padded_len, int32_t array_is_f32,
                                  const int64_t* true_track_id, const double* true_photons, int32_t n_truth, int64_t n_trig,
                                  const int64_t* trig_channel, int32_t n_det_module, int32_t n_samples, double digit_sample_spacing,
                                  double light_tick_size, double mc_truth_threshold, int32_t light_nbit, int32_t truncate,
                                  double* digit_signal, int64_t* digit_true_track_id, double* digit_true_photons, int32_t n_truth_out,
                                  void* stream) {
    const long long n = n_trig * (long long)n_det_module * n_samples;
    if (n == 0) return 0;
    LSB_REQUIRE(n_rows >= 0 && row_channel && row_source && trig_channel && digit_signal, "light_digitize: null pointer");
    LSB_REQUIRE(n_truth == 0 || (true_track_id && true_photons && digit_true_track_id && digit_true_photons), "light_digitize: truth arrays missing");
    LtDigit p;
    p.nticks = nticks; p.front = front_pad; p.L = padded_len; p.nrow = n_rows; p.M = n_truth; p.Mout = n_truth_out; p.ndm = n_det_module;
    p.nsamples = n_samples; p.diff_f32 = (array_is_f32 && !signal_f64) ? 1 : 0; p.step = 0;
    p.spacing = digit_sample_spacing; p.tick = light_tick_size; p.truth_threshold = mc_truth_threshold;
    p.quantum = ldexp(1.0, 16 - light_nbit); p.truncate = truncate;
    cudaStream_t st = (cudaStream_t)stream;
    if (signal_f64)
        k_lt_digitize<double><<<lsb_blocks(n, 128), 128, 0, st>>>(p, (const double*)signal, (const long long*)row_channel, row_source,
                                                                 (const long long*)true_track_id, true_photons, n_trig,
                                                                 (const long long*)trig_channel, digit_signal,
                                                                 (long long*)digit_true_track_id, digit_true_photons);
    else
        k_lt_digitize<float><<<lsb_blocks(n, 128), 128, 0, st>>>(p, (const float*)signal, (const long long*)row_channel, row_source,
                                                                (const long long*)true_track_id, true_photons, n_trig,
                                                                (const long long*)trig_channel, digit_signal,
                                                                (long long*)digit_true_track_id, digit_true_photons);
    LSB_LAUNCH_CHECK("k_lt_digitize");
    return 0;
}

// ---------------------------------------------------------------------------------------
// extent of the light simulation window: light_sim.get_nticks (:24-42) and get_active_op_channel (:44-57)
// One pass over light_incidence[S][ndet]: among entries with n_photons_det > 0 the earliest / latest t0_det (float32,
// reduced through an order-preserving integer code) and, per channel, whether any segment lights it.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f32_ordered(float x) {
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_unordered(uint32_t c) {
    return __uint_as_float((c & 0x80000000u) ? (c & 0x7fffffffu) : ~c);
}
__global__ void k_lt_extent_init(uint32_t* __restrict__ code, uint8_t* __restrict__ active, int ndet) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { code[0] = 0xffffffffu; code[1] = 0u; }
    if (i < ndet) active[i] = 0;
}
__global__ void k_lt_extent(const char* __restrict__ linc, lsb_linc_layout LI, long long n, int ndet, uint32_t* __restrict__ code,
                            uint8_t* __restrict__ active) {
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const char* rec = linc + i * (long long)LI.itemsize;
        if (*(const float*)(rec + LI.off_n_photons_det) > 0.0f) {
            const uint32_t c = f32_ordered(*(const float*)(rec + LI.off_t0_det));
            lo = min(lo, c);
            hi = max(hi, c);
            active[i % ndet] = 1;                       // same value from every writer
        }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
        atomicMin(code, lo);
        atomicMax(code + 1, hi);
    }
}
__global__ void k_lt_extent_finish(uint32_t* __restrict__ code) {
    const bool any = code[0] <= code[1];
    const float lo = any ? f32_unordered(code[0]) : INFINITY, hi = any ? f32_unordered(code[1]) : -INFINITY;
    ((float*)code)[0] = lo;
    ((float*)code)[1] = hi;
}

LSB_EXPORT int lsb_light_extent(const void* light_incidence, const lsb_linc_layout* LI, int64_t n_segments, int32_t ndet,
                                float* t0_minmax, uint8_t* active, void* stream) {
    LSB_REQUIRE(LI && t0_minmax && (active || ndet == 0) && (light_incidence || n_segments * (int64_t)ndet == 0),
                "light_extent: null pointer");
    LSB_REQUIRE(n_segments >= 0 && ndet >= 0, "light_extent: negative size");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* code = (uint32_t*)t0_minmax;
    k_lt_extent_init<<<lsb_blocks(ndet > 0 ? ndet : 1, 256), 256, 0, st>>>(code, active, ndet);
    LSB_LAUNCH_CHECK("k_lt_extent_init");
    const long long n = n_segments * (long long)ndet;
    if (n > 0) {
        long long nb = (n + 255) / 256;
        if (nb > 148 * 16) nb = 148 * 16;
        k_lt_extent<<<(unsigned)nb, 256, 0, st>>>((const char*)light_incidence, *LI, n, ndet, code, active);
        LSB_LAUNCH_CHECK("k_lt_extent");
    }
    k_lt_extent_finish<<<1, 1, 0, st>>>(code);
    LSB_LAUNCH_CHECK("k_lt_extent_finish");
    return 0;
}

// ---------------------------------------------------------------------------------------
// zero suppression of the waveform truth: light_sim.zero_suppress_waveform_truth (:621-661)
// The reference enumerates the [trigger][channel][sample][slot] array in Python and appends one record per slot whose
// track id is not -1.  Here: flag -> prefix sum -> scatter, rows in the same (C-order) sequence.  The reference's running
// `i_trig = i_trig + this_trig` (it accumulates over the entries, :644) is a second prefix sum over the kept entries.
// ---------------------------------------------------------------------------------------
struct LtTruthRow { int32_t trigger_id, op_channel_id, tick, event_id; long long segment_id; double pe_current; };
static_assert(sizeof(LtTruthRow) == 32, "truth row layout");

__global__ void k_lt_truth_flags(const long long* __restrict__ ids, long long n, long long per_trigger, uint32_t* __restrict__ flag,
                                 uint32_t* __restrict__ tsum) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool keep = ids[i] != -1;
    flag[i] = keep ? 1u : 0u;
    tsum[i] = keep ? (uint32_t)(i / per_trigger) : 0u;
}
__global__ void k_lt_truth_rows(const long long* __restrict__ ids, const double* __restrict__ photons, long long n, int n_det,
                                int n_samples, int n_truth, const int32_t* __restrict__ op_channel, int event_id, int first_trigger,
                                const uint32_t* __restrict__ flag, const long long* __restrict__ pos, const long long* __restrict__ tpre,
                                LtTruthRow* __restrict__ rows) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    const long long per_det = (long long)n_samples * n_truth, per_trigger = per_det * n_det;
    const long long trig = i / per_trigger, rem = i - trig * per_trigger;
    const int det = (int)(rem / per_det), sample = (int)((rem - det * per_det) / n_truth);
    LtTruthRow r;
    r.trigger_id = (int32_t)(first_trigger + tpre[i] + trig);          // inclusive running sum of the trigger indices
    r.op_channel_id = op_channel[det];
    r.tick = sample;
    r.event_id = event_id;
    r.segment_id = ids[i];
    r.pe_current = photons[i];
    rows[pos[i]] = r;
}

LSB_EXPORT int64_t lsb_light_truth_ws_bytes(int64_t n) {
    const size_t a4 = (((size_t)n * 4 + 15) / 16) * 16, a8 = (((size_t)n * 8 + 15) / 16) * 16;
    return (int64_t)(2 * a4 + 2 * a8 + (size_t)(scan_num_blocks(n) + 1) * 8 + 32);
}

LSB_EXPORT int lsb_light_zero_suppress_truth(const int64_t* true_track_id, const double* true_photons, int64_t n_trig, int32_t n_det,
                                             int32_t n_samples, int32_t n_truth, const int32_t* op_channel, int32_t event_id,
                                             int32_t first_trigger_id, void* rows, int64_t* n_rows, void* ws, int64_t ws_bytes,
                                             void* stream) {
    LSB_REQUIRE(n_rows, "light_zero_suppress_truth: null pointer");
    LSB_REQUIRE(n_trig >= 0 && n_det >= 0 && n_samples >= 0 && n_truth >= 0, "light_zero_suppress_truth: negative size");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)n_trig * n_det * n_samples * n_truth;
    if (n == 0) { LSB_CUDA(cudaMemsetAsync(n_rows, 0, 8, st)); return 0; }
    LSB_REQUIRE(true_track_id && true_photons && op_channel && rows, "light_zero_suppress_truth: null pointer");
    LSB_REQUIRE(ws && ws_bytes >= lsb_light_truth_ws_bytes(n), "light_zero_suppress_truth: workspace too small");
    const size_t a4 = (((size_t)n * 4 + 15) / 16) * 16, a8 = (((size_t)n * 8 + 15) / 16) * 16;
    char* p = (char*)ws;
    uint32_t* flag = (uint32_t*)p; p += a4;
    uint32_t* tsum = (uint32_t*)p; p += a4;
    long long* pos = (long long*)p; p += a8;
    long long* tpre = (long long*)p; p += a8;
    long long* bs = (long long*)p; p += (size_t)scan_num_blocks(n) * 8;
    long long* scratch_total = (long long*)(((uintptr_t)p + 7) & ~(uintptr_t)7);
    const long long per_trigger = (long long)n_det * n_samples * n_truth;
    k_lt_truth_flags<<<lsb_blocks(n, 256), 256, 0, st>>>((const long long*)true_track_id, n, per_trigger, flag, tsum);
    LSB_LAUNCH_CHECK("k_lt_truth_flags");
    int rc = exclusive_scan<uint32_t, long long>(flag, n, pos, bs, (long long*)n_rows, st);
    if (rc) return rc;
    rc = exclusive_scan<uint32_t, long long>(tsum, n, tpre, bs, scratch_total, st);
    if (rc) return rc;
    k_lt_truth_rows<<<lsb_blocks(n, 256), 256, 0, st>>>((const long long*)true_track_id, true_photons, n, n_det, n_samples, n_truth,
                                                       op_channel, event_id, first_trigger_id, flag, pos, tpre, (LtTruthRow*)rows);
    LSB_LAUNCH_CHECK("k_lt_truth_rows");
    return 0;
}
