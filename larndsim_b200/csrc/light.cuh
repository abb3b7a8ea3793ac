// light.cuh -- light readout: calculate_light_incidence (lightLUT.py:65-136), sum_light_signals
// (light_sim.py:58-129), calc_scintillation_effect (:148-183), calc_stat_fluctuations (:220-238),
// calc_light_detector_response (:303-336).
//
// The waveform buffers are float32 and the reference adds into them with `+=`, i.e. one float32
// rounding per add, so every output sample is produced by one thread that performs the adds in the
// reference's order (segments in `sorted_indices` order, profile bins ascending, taps ascending).
// The two time convolutions are shared-memory staged FIRs: a CTA owns 256 consecutive ticks of one
// channel, input samples and the float64 tap weights it needs are staged through shared memory in
// chunks, and chunks whose inputs are all zero are skipped (x + w*0 == x).  The tap weights
// (scintillation model / SiPM impulse) are evaluated once per call on the host in float64.
#pragma once
#include "common.cuh"
#include "glue.cuh"
#include "batching.cuh"

// lightLUT.py:65-136, thread per (segment, channel)
__global__ void k_light_incidence(Layout L, const char* __restrict__ tracks, long long S, const char* __restrict__ lut,
                                  lsb_lut_layout LL, char* __restrict__ linc, lsb_linc_layout LI, int ndet,
                                  int32_t* __restrict__ voxel, const double* __restrict__ eff, const long long* __restrict__ ch2tpc) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= S * ndet) return;
    long long itrk = idx / ndet;
    int o = (int)(idx % ndet);
    const char* t = tracks + itrk * L.itemsize;
    long long itpc = (long long)fld_get(L, t, LSB_F_PIXEL_PLANE);
    if (itpc == d_c.default_plane_index || itpc < 0 || itpc >= d_c.n_tpc) return;
    long long imod = py_div_ll(itpc, 2);
    const double (*b)[2] = d_c.tpc_borders[itpc];
    double px = fld_get(L, t, LSB_F_X), py = fld_get(L, t, LSB_F_Y), pz = fld_get(L, t, LSB_F_Z);
    bool is_even = b[2][1] > b[2][0];
    double x_min = b[0][0] - 2e-2, x_max = b[0][1] + 2e-2, y_min = b[1][0] - 2e-2, y_max = b[1][1] + 2e-2;
    double z_min = b[2][0] - 2e-2, z_max = b[2][1] + 2e-2;
    long long i = is_even ? (long long)((px - x_min) / (x_max - x_min) * LL.shape[0]) : (long long)((x_max - px) / (x_max - x_min) * LL.shape[0]);
    long long j = (long long)((y_max - py) / (y_max - y_min) * LL.shape[1]);
    long long k = (long long)((pz - z_min) / (z_max - z_min) * LL.shape[2]);
    i = i < 0 ? 0 : i; i = i > LL.shape[0] - 1 ? LL.shape[0] - 1 : i;
    j = j < 0 ? 0 : j; j = j > LL.shape[1] - 1 ? LL.shape[1] - 1 : j;
    k = k < 0 ? 0 : k; k = k > LL.shape[2] - 1 ? LL.shape[2] - 1 : k;
    if (o == 0) { voxel[itrk * 3 + 0] = (int32_t)i; voxel[itrk * 3 + 1] = (int32_t)j; voxel[itrk * 3 + 2] = (int32_t)k; }
    const char* vox = lut + (((i * LL.shape[1] + j) * LL.shape[2] + k) * (long long)LL.shape[3]) * LL.itemsize;
    long long channel_offset = (ndet < d_c.n_op_channel) ? ndet * imod : 0;
    long long ch = o + channel_offset;
    long long li = o % LL.shape[3];
    float vis_f = *(const float*)(vox + li * LL.itemsize + LL.off_vis);
    double vis = (double)vis_f * (ch2tpc[ch] == itpc ? 1.0 : 0.0);
    double n_photons = fld_get(L, t, LSB_F_N_PHOTONS);
    char* rec = linc + (itrk * ndet + o) * (long long)LI.itemsize;
    *(float*)(rec + LI.off_n_photons_det) = __double2float_rn(eff[ch] * vis * n_photons);
    if (d_c.light_trig_mode == 0) {
        float t1f = *(const float*)(vox + li * LL.itemsize + LL.off_t0);
        double t1 = ((double)t1f * d_c.unit_ns + fld_get(L, t, LSB_F_T0) * d_c.unit_mus) / d_c.unit_mus;
        *(float*)(rec + LI.off_t0_det) = __double2float_rn(t1);
    }
}

// per-channel compaction of the contributing segments, in sorted_indices order:
// keep[idet][s] = 1 if n_photons_det > 0 (light_sim.py:84)
struct LightSeg { double t0; float nph; int pad; long long lut_off; long long track_id; };

__global__ void k_light_gather_segs(Layout L, const char* __restrict__ segments, const int32_t* __restrict__ seg_voxel,
                                    const long long* __restrict__ seg_track_id, const char* __restrict__ linc, lsb_linc_layout LI,
                                    int ndet_inc, const int32_t* __restrict__ op_channel, lsb_lut_layout LL,
                                    const long long* __restrict__ sorted_indices, long long n_sorted, int ndet,
                                    LightSeg* __restrict__ segs) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)ndet * n_sorted) return;
    int idet = (int)(idx / n_sorted);
    long long itrk = sorted_indices[idx];
    int ch = op_channel[idet];
    LightSeg r;
    r.nph = *(const float*)(linc + (itrk * ndet_inc + ch) * (long long)LI.itemsize + LI.off_n_photons_det);
    r.t0 = fld_get(L, segments + itrk * L.itemsize, LSB_F_T0);
    const int32_t* v = seg_voxel + itrk * 3;
    long long idet_lut = py_mod_ll(ch, LL.shape[3]);
    r.lut_off = ((((long long)v[0] * LL.shape[1] + v[1]) * LL.shape[2] + v[2]) * (long long)LL.shape[3] + idet_lut) * LL.itemsize;
    r.track_id = seg_track_id ? seg_track_id[itrk] : -1;
    r.pad = 0;
    segs[idx] = r;
}

#define LT_TPB 128
#define LT_CHUNK 128
// one segment's contribution to one tick (light_sim.py:84-129), reference add order and roundings
__device__ __forceinline__ void light_add_segment(const LightSeg& g, const char* __restrict__ lut, const lsb_lut_layout& LL,
                                                  double start_tick_time, double end_tick_time, double prof_len, float& acc,
                                                  long long* __restrict__ true_id, double* __restrict__ true_ph, long long tbase, int n_true) {
    if (!(g.nph > 0.f)) return;
    const double track_time = g.t0;
    const double track_end_time = track_time + prof_len;
    if (track_end_time < start_tick_time || track_time > end_tick_time) return;
    const char* lrec = lut + g.lut_off;
    if (d_c.enable_lut_smearing) {
        // profile_time is non-decreasing in ip: only the bins around (tick start - t0) / 1 ns can pass the strict test below; the
        // window is taken with a margin of two bins on either side and every bin in it is tested exactly, in ascending order
        const double step = d_c.unit_ns / d_c.unit_mus;
        const double first = floor((start_tick_time - track_time) / step) - 2.0;
        const double last = ceil((end_tick_time - track_time) / step) + 2.0;
        const int ip_lo = first > 0.0 ? (first < (double)LL.n_time_dist ? (int)first : LL.n_time_dist) : 0;
        const int ip_hi = last < (double)(LL.n_time_dist - 1) ? (last >= 0.0 ? (int)last : -1) : LL.n_time_dist - 1;
        for (int ip = ip_lo; ip <= ip_hi; ip++) {
            double profile_time = track_time + (double)ip * d_c.unit_ns / d_c.unit_mus;
            if (profile_time < end_tick_time && profile_time > start_tick_time) {
                float tp = *(const float*)(lrec + LL.off_time_dist + 4 * ip);
                double photons = (double)__fmul_rn(g.nph, tp) / d_c.light_tick_size;
                acc = __double2float_rn((double)acc + photons);
                if (photons > d_c.mc_truth_threshold) {
                    for (int q = 0; q < n_true; q++) {
                        long long* tid = true_id + tbase + q;
                        if (*tid == -1 || *tid == g.track_id) { *tid = g.track_id; true_ph[tbase + q] += photons; break; }
                    }
                }
            }
        }
    } else {
        float ta = *(const float*)(lrec + LL.off_t0_avg);
        double t0_avg = (double)ta * d_c.unit_ns / d_c.unit_mus;
        double profile_time = track_time + t0_avg;
        if (profile_time < end_tick_time && profile_time > start_tick_time) {
            double photons = (double)g.nph / d_c.light_tick_size;
            acc = __double2float_rn((double)acc + photons);
            if (photons > d_c.mc_truth_threshold) {
                for (int q = 0; q < n_true; q++) {
                    long long* tid = true_id + tbase + q;
                    if (*tid == -1 || *tid == g.track_id) { *tid = g.track_id; true_ph[tbase + q] += photons; break; }
                }
            }
        }
    }
}

// brute force (the reference's formulation: every tick walks every segment; CTA-wide early-out): kept for A/B (LSB_LIGHT_BINNED=0)
__global__ void __launch_bounds__(LT_TPB) k_sum_light_signals(const LightSeg* __restrict__ segs, long long n_sorted,
                                                              const char* __restrict__ lut, lsb_lut_layout LL, double start_time,
                                                              float* __restrict__ lsi, int ndet, int nticks,
                                                              long long* __restrict__ true_id, double* __restrict__ true_ph, int n_true,
                                                              double t0_profile_length) {
    __shared__ LightSeg s_seg[LT_CHUNK];
    const int idet = blockIdx.y;
    const int itick = blockIdx.x * LT_TPB + threadIdx.x;
    const bool active = itick < nticks;
    const double start_tick_time = (double)itick * d_c.light_tick_size + start_time;
    const double end_tick_time = start_tick_time + d_c.light_tick_size;
    const double prof_len = t0_profile_length * d_c.unit_ns / d_c.unit_mus;
    // time span of this CTA's ticks, to discard whole segments for all threads at once
    const double cta_lo = (double)(blockIdx.x * LT_TPB) * d_c.light_tick_size + start_time;
    const double cta_hi = (double)(blockIdx.x * LT_TPB + LT_TPB) * d_c.light_tick_size + start_time + d_c.light_tick_size;
    float acc = active ? lsi[(long long)idet * nticks + itick] : 0.f;
    const long long tbase = ((long long)idet * nticks + itick) * n_true;
    for (long long c0 = 0; c0 < n_sorted; c0 += LT_CHUNK) {
        int nc = (int)(n_sorted - c0 < LT_CHUNK ? n_sorted - c0 : LT_CHUNK);
        __syncthreads();
        if ((int)threadIdx.x < nc) s_seg[threadIdx.x] = segs[(long long)idet * n_sorted + c0 + threadIdx.x];
        __syncthreads();
        if (!active) continue;
        for (int s = 0; s < nc; s++) {
            const LightSeg g = s_seg[s];
            if (!(g.nph > 0.f)) continue;
            if (g.t0 + prof_len < cta_lo || g.t0 > cta_hi) continue;          // uniform early-out (superset of the per-tick test)
            light_add_segment(g, lut, LL, start_tick_time, end_tick_time, prof_len, acc, true_id, true_ph, tbase, n_true);
        }
    }
    if (active) lsi[(long long)idet * nticks + itick] = acc;
}

// ---- binned form ---------------------------------------------------------------------------------------------------
// A segment lights a channel for t0_profile_length ns: a few dozen of the 16 000 ticks.  Every (channel, segment) entry with
// photons is expanded to the blocks of LT_TPB ticks it can touch (a superset computed with a margin; the exact per-tick test
// stays in the kernel), the (block, entry) pairs are sorted by block with the package's STABLE radix sort, so each block's list is
// in `sorted_indices` order -- the order in which the reference adds, which the float32 accumulator depends on -- and a CTA walks
// only its own list: O(entries) instead of O(ndet * nticks * S).
__global__ void k_light_block_ranges(const LightSeg* __restrict__ segs, long long n_entries, double start_time, double prof_len,
                                     int nticks, int2* __restrict__ range, uint32_t* __restrict__ count) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    const LightSeg g = segs[e];
    int2 r = make_int2(0, -1);
    if (g.nph > 0.f) {
        const double lo = floor((g.t0 - start_time) / d_c.light_tick_size) - 2.0;
        const double hi = ceil((g.t0 + prof_len - start_time) / d_c.light_tick_size) + 2.0;
        if (hi >= 0.0 && lo <= (double)(nticks - 1)) {
            const int il = lo < 0.0 ? 0 : (int)lo, ih = hi > (double)(nticks - 1) ? nticks - 1 : (int)hi;
            r = make_int2(il / LT_TPB, ih / LT_TPB);
        }
    }
    range[e] = r;
    count[e] = r.y >= r.x ? (uint32_t)(r.y - r.x + 1) : 0u;
}
__global__ void k_light_expand(const int2* __restrict__ range, const long long* __restrict__ offs, long long n_entries, long long n_sorted,
                               int n_blocks, uint32_t* __restrict__ keys, int32_t* __restrict__ idx) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    const int2 r = range[e];
    const long long idet = e / n_sorted;
    long long o = offs[e];
    for (int b = r.x; b <= r.y; b++, o++) { keys[o] = (uint32_t)(idet * n_blocks + b); idx[o] = (int32_t)e; }
}
__global__ void __launch_bounds__(LT_TPB) k_sum_light_signals_binned(const LightSeg* __restrict__ segs, const int32_t* __restrict__ idx,
                                                                     const long long* __restrict__ bucket, const char* __restrict__ lut,
                                                                     lsb_lut_layout LL, double start_time, float* __restrict__ lsi, int ndet,
                                                                     int nticks, long long* __restrict__ true_id, double* __restrict__ true_ph,
                                                                     int n_true, double t0_profile_length) {
    __shared__ LightSeg s_seg[LT_CHUNK];
    const int idet = blockIdx.y;
    const long long key = (long long)idet * gridDim.x + blockIdx.x;
    const long long e0 = bucket[key], e1 = bucket[key + 1];
    if (e0 >= e1) return;                                   // nothing reaches this block: the caller's values stay
    const int itick = blockIdx.x * LT_TPB + threadIdx.x;
    const bool active = itick < nticks;
    const double start_tick_time = (double)itick * d_c.light_tick_size + start_time;
    const double end_tick_time = start_tick_time + d_c.light_tick_size;
    const double prof_len = t0_profile_length * d_c.unit_ns / d_c.unit_mus;
    float acc = active ? lsi[(long long)idet * nticks + itick] : 0.f;
    const long long tbase = ((long long)idet * nticks + itick) * n_true;
    for (long long c0 = e0; c0 < e1; c0 += LT_CHUNK) {
        const int nc = (int)(e1 - c0 < LT_CHUNK ? e1 - c0 : LT_CHUNK);
        __syncthreads();
        if ((int)threadIdx.x < nc) s_seg[threadIdx.x] = segs[idx[c0 + threadIdx.x]];
        __syncthreads();
        if (!active) continue;
        for (int s = 0; s < nc; s++) light_add_segment(s_seg[s], lut, LL, start_tick_time, end_tick_time, prof_len, acc, true_id, true_ph, tbase, n_true);
    }
    if (active) lsi[(long long)idet * nticks + itick] = acc;
}

// ---------------------------------------------------------------------------------------
// shared-memory staged causal FIR:  out[idet][i] += sum_{j=max(i-conv,0)}^{i} g * w[i-j] * in[idet][j]
// (float32 accumulate, one rounding per add, taps in ascending j).  mode 0: scintillation (truth test
// `w*ph < thr`), mode 1: detector response (gain, |w*ph| < thr and the reference's itick-for-jtick
// indexing of the truth ids, light_sim.py:333-335).
// ---------------------------------------------------------------------------------------
#define FIR_TPB 256
template <int MODE>
__global__ void __launch_bounds__(FIR_TPB) k_light_fir(const float* __restrict__ in, const long long* __restrict__ in_id,
                                                       const double* __restrict__ in_ph, float* __restrict__ out,
                                                       long long* __restrict__ out_id, double* __restrict__ out_ph, int ndet,
                                                       int nticks, int n_in, int n_out, const double* __restrict__ w,
                                                       long long conv_ticks, const double* __restrict__ gain) {
    __shared__ float s_in[FIR_TPB];
    __shared__ double s_w[2 * FIR_TPB];
    __shared__ int s_any;
    const int idet = blockIdx.y;
    const int i0 = blockIdx.x * FIR_TPB;
    const int itick = i0 + threadIdx.x;
    const bool active = itick < nticks;
    const float* row = in + (long long)idet * nticks;
    float acc = active ? out[(long long)idet * nticks + itick] : 0.f;
    const double g = MODE == 1 ? gain[idet] : 1.0;
    long long jlo = (long long)i0 - conv_ticks; if (jlo < 0) jlo = 0;
    const long long jhi = (long long)i0 + FIR_TPB - 1 < nticks - 1 ? (long long)i0 + FIR_TPB - 1 : nticks - 1;
    for (long long j0 = jlo; j0 <= jhi; j0 += FIR_TPB) {
        __syncthreads();
        if (threadIdx.x == 0) s_any = 0;
        __syncthreads();
        long long j = j0 + threadIdx.x;
        float v = (j <= jhi) ? row[j] : 0.f;
        s_in[threadIdx.x] = v;
        if (v != 0.f) s_any = 1;
        // weights for d = i - j, i in [i0, i0+TPB), j in [j0, j0+TPB): d in [i0-j0-TPB+1, i0-j0+TPB-1]
        long long dbase = (long long)i0 - j0 - (FIR_TPB - 1);
        for (int q = threadIdx.x; q < 2 * FIR_TPB; q += FIR_TPB) {
            long long d = dbase + q;
            s_w[q] = (d >= 0 && d <= conv_ticks) ? w[d] : 0.0;
        }
        __syncthreads();
        if (!active || (!s_any && (MODE == 0 || n_in == 0))) continue;
        int nj = (int)(jhi - j0 + 1 < FIR_TPB ? jhi - j0 + 1 : FIR_TPB);
        for (int jj = 0; jj < nj; jj++) {
            long long jt = j0 + jj;
            long long d = (long long)itick - jt;
            if (d < 0 || d > conv_ticks) continue;
            float vin = s_in[jj];
            if (vin == 0.f && (MODE == 0 || n_in == 0)) continue;        // scint: reference skips; response: adds w*0
            double tw = s_w[(int)(d - dbase)];
            acc = __double2float_rn((double)acc + g * tw * (double)vin);
            if (n_in > 0) {
                const long long bj = ((long long)idet * nticks + jt) * n_in, bi = ((long long)idet * nticks + itick) * n_in;
                const long long bo = ((long long)idet * nticks + itick) * n_out;
                for (int it = 0; it < n_in; it++) {
                    if (in_id[bj + it] == -1) break;
                    double ph = in_ph[bj + it];
                    if (MODE == 0) {
                        if (tw * ph < d_c.mc_truth_threshold) continue;
                        for (int q = 0; q < n_out; q++)
                            if (out_id[bo + q] == in_id[bj + it] || out_id[bo + q] == -1) {
                                out_id[bo + q] = in_id[bj + it]; out_ph[bo + q] += tw * ph; break;
                            }
                    } else {
                        if (fabs(tw * ph) < d_c.mc_truth_threshold) continue;
                        for (int q = 0; q < n_out; q++)
                            if (in_id[bi + q] == in_id[bi + it] || in_id[bi + q] == -1) {
                                out_id[bo + q] = in_id[bi + it]; out_ph[bo + q] += tw * ph; break;
                            }
                    }
                }
            }
        }
    }
    if (active) out[(long long)idet * nticks + itick] = acc;
}

// light_sim.py:186-238
__device__ __forceinline__ int poisson_i32(double mean, Rng& r) {
    if (mean <= 0) return 0;
    if (mean < 30) {
        float u = rng_uniform_f32(r);
        int x = 0;
        double p = exp(-mean), s = p, prev_s = s;
        while ((double)u > s) {
            x += 1;
            p = p * mean / x;
            prev_s = s;
            s = s + p;
            if (s == prev_s) break;
        }
        return x;
    }
    double v = (double)rng_normal_f32(r) * sqrt(mean) + mean;
    long long iv = (long long)v;
    return (int)(iv > 0 ? iv : 0);
}
__global__ void k_stat_fluctuations(const float* __restrict__ in, float* __restrict__ out, long long n,
                                    unsigned long long* __restrict__ rng_states) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= n) return;
    float v = in[idx];
    if (v > 0) {
        Rng r; r.s0 = rng_states[2 * idx]; r.s1 = rng_states[2 * idx + 1];
        double mean = (double)v * d_c.light_tick_size;
        out[idx] = __double2float_rn(1. / d_c.light_tick_size * (double)poisson_i32(mean, r));
        rng_states[2 * idx] = r.s0; rng_states[2 * idx + 1] = r.s1;
    } else out[idx] = 0.f;
}

// ---------------------------------------------------------------------------------------
static inline bool lut_ok(const lsb_lut_layout* LL) { return LL && LL->itemsize > 0 && LL->shape[0] > 0 && LL->shape[1] > 0 && LL->shape[2] > 0 && LL->shape[3] > 0; }

LSB_EXPORT int lsb_calculate_light_incidence(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t S,
                                             const void* lut, const lsb_lut_layout* LL, void* light_incidence,
                                             const lsb_linc_layout* LI, int32_t ndet, int32_t* voxel,
                                             const double* op_channel_efficiency, const int64_t* op_channel_to_tpc, void* stream) {
    LSB_REQUIRE(c && L && LI && lut_ok(LL), "calculate_light_incidence: null consts/layout");
    if (S == 0 || ndet == 0) return 0;
    LSB_REQUIRE(tracks && lut && light_incidence && voxel && op_channel_efficiency && op_channel_to_tpc,
                "calculate_light_incidence: null pointer");
    LSB_REQUIRE(LL->off_vis >= 0 && (c->light_trig_mode != 0 || LL->off_t0 >= 0), "calculate_light_incidence: LUT lacks vis/t0");
    static const int need[] = {LSB_F_X, LSB_F_Y, LSB_F_Z, LSB_F_N_PHOTONS, LSB_F_PIXEL_PLANE, LSB_F_T0};
    for (int f : need) LSB_REQUIRE(layout_has(L, f), "calculate_light_incidence: tracks lacks a required field");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = lsb_upload_consts(c, st); if (rc) return rc;
    k_light_incidence<<<lsb_blocks(S * ndet, 256), 256, 0, st>>>(make_layout(L), (const char*)tracks, S, (const char*)lut, *LL,
                                                                (char*)light_incidence, *LI, ndet, voxel, op_channel_efficiency,
                                                                (const long long*)op_channel_to_tpc);
    LSB_LAUNCH_CHECK("k_light_incidence");
    return 0;
}

LSB_EXPORT int lsb_sum_light_signals(const lsb_consts* c, const lsb_track_layout* L, const void* segments, int64_t S,
                                     const int32_t* segment_voxel, const int64_t* segment_track_id, const void* light_inc,
                                     const lsb_linc_layout* LI, int32_t ndet_inc, const int32_t* op_channel, const void* lut,
                                     const lsb_lut_layout* LL, double start_time, float* light_sample_inc, int32_t ndet,
                                     int32_t nticks, int64_t* true_track_id, double* true_photons, int32_t n_true,
                                     const int64_t* sorted_indices, int64_t n_sorted, double t0_profile_length, void* stream) {
    LSB_REQUIRE(c && L && LI && lut_ok(LL), "sum_light_signals: null consts/layout");
    if (ndet == 0 || nticks == 0 || n_sorted == 0) return 0;
    LSB_REQUIRE(segments && segment_voxel && light_inc && op_channel && lut && light_sample_inc && sorted_indices,
                "sum_light_signals: null pointer");
    LSB_REQUIRE(n_true == 0 || (true_track_id && true_photons && segment_track_id), "sum_light_signals: null truth arrays");
    LSB_REQUIRE(layout_has(L, LSB_F_T0), "sum_light_signals: segments lacks t0");
    LSB_REQUIRE(c->enable_lut_smearing ? LL->off_time_dist >= 0 : LL->off_t0_avg >= 0, "sum_light_signals: LUT lacks time_dist/t0_avg");
    (void)S;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = lsb_upload_consts(c, st); if (rc) return rc;
    TmpPool tp(st);
    LightSeg* segs;
    LSB_CUDA(tp.get(&segs, (long long)ndet * n_sorted));
    k_light_gather_segs<<<lsb_blocks((long long)ndet * n_sorted, 256), 256, 0, st>>>(
        make_layout(L), (const char*)segments, segment_voxel, (const long long*)segment_track_id, (const char*)light_inc, *LI,
        ndet_inc, op_channel, *LL, (const long long*)sorted_indices, n_sorted, ndet, segs);
    LSB_LAUNCH_CHECK("k_light_gather_segs");
    dim3 grid((unsigned)((nticks + LT_TPB - 1) / LT_TPB), (unsigned)ndet);
    static int binned = -1;
    if (binned < 0) { const char* e = getenv("LSB_LIGHT_BINNED"); binned = (e && e[0] == '0') ? 0 : 1; }
    const long long n_entries = (long long)ndet * n_sorted, n_buckets = (long long)ndet * grid.x;
    if (!binned || n_entries >= (1ll << 31) || n_buckets >= (1ll << 31)) {
        k_sum_light_signals<<<grid, LT_TPB, 0, st>>>(segs, n_sorted, (const char*)lut, *LL, start_time, light_sample_inc, ndet, nticks,
                                                    (long long*)true_track_id, true_photons, n_true, t0_profile_length);
        LSB_LAUNCH_CHECK("k_sum_light_signals");
        return 0;
    }
    int2* range; uint32_t* count; long long* offs; long long* bs; long long* total;
    LSB_CUDA(tp.get(&range, n_entries)); LSB_CUDA(tp.get(&count, n_entries)); LSB_CUDA(tp.get(&offs, n_entries));
    LSB_CUDA(tp.get(&bs, scan_num_blocks(n_entries) + 1)); LSB_CUDA(tp.get(&total, 1));
    k_light_block_ranges<<<lsb_blocks(n_entries, 256), 256, 0, st>>>(segs, n_entries, start_time, t0_profile_length * c->unit_ns / c->unit_mus,
                                                                    nticks, range, count);
    LSB_LAUNCH_CHECK("k_light_block_ranges");
    if ((rc = exclusive_scan<uint32_t, long long>(count, n_entries, offs, bs, total, st))) return rc;
    long long E = 0;
    LSB_CUDA(cudaMemcpyAsync(&E, total, 8, cudaMemcpyDeviceToHost, st));
    LSB_CUDA(cudaStreamSynchronize(st));
    if (E == 0) return 0;
    LSB_REQUIRE(E < (1ll << 31), "sum_light_signals: more than 2^31 (block, segment) pairs");
    uint32_t *keyA, *keyB; int32_t *idxA, *idxB; long long* bucket;
    LSB_CUDA(tp.get(&keyA, E)); LSB_CUDA(tp.get(&keyB, E)); LSB_CUDA(tp.get(&idxA, E)); LSB_CUDA(tp.get(&idxB, E));
    LSB_CUDA(tp.get(&bucket, n_buckets + 1));
    k_light_expand<<<lsb_blocks(n_entries, 256), 256, 0, st>>>(range, offs, n_entries, n_sorted, (int)grid.x, keyA, idxA);
    LSB_LAUNCH_CHECK("k_light_expand");
    int bits = 0;
    while ((n_buckets >> bits) != 0) bits++;
    if ((rc = rs_sort_pairs(&keyA, &keyB, &idxA, &idxB, E, bits, tp, st))) return rc;
    k_unit_offsets<<<lsb_blocks(n_buckets + 1, 256), 256, 0, st>>>(keyA, E, n_buckets, bucket);
    LSB_LAUNCH_CHECK("k_unit_offsets");
    k_sum_light_signals_binned<<<grid, LT_TPB, 0, st>>>(segs, idxA, bucket, (const char*)lut, *LL, start_time, light_sample_inc, ndet, nticks,
                                                       (long long*)true_track_id, true_photons, n_true, t0_profile_length);
    LSB_LAUNCH_CHECK("k_sum_light_signals_binned");
    return 0;
}

static int light_fir(int mode, const lsb_consts* c, const float* in, const int64_t* in_id, const double* in_ph, float* out,
                     int64_t* out_id, double* out_ph, int ndet, int nticks, int n_in, int n_out, const double* gain,
                     const double* impulse, int n_imp, cudaStream_t st) {
    long long conv_ticks = (long long)ceil((c->light_window[1] - c->light_window[0]) / c->light_tick_size);
    if (conv_ticks < 0) return 0;                       // empty range(max(itick-conv,0), itick+1)
    long long nw = (conv_ticks < nticks - 1 ? conv_ticks : nticks - 1) + 1;
    double* wh = (double*)malloc(sizeof(double) * (size_t)nw);
    if (!wh) return lsb_fail_arg("light FIR: out of host memory");
    for (long long tt = 0; tt < nw; tt++) {
        if (mode == 0) {                                // scintillation_model light_sim.py:131-145
            double p1 = c->singlet_fraction * exp(-tt * c->light_tick_size / c->tau_s) * (1 - exp(-c->light_tick_size / c->tau_s));
            double p3 = (1 - c->singlet_fraction) * exp(-tt * c->light_tick_size / c->tau_t) * (1 - exp(-c->light_tick_size / c->tau_t));
            wh[tt] = (p1 + p3) * 1.0;
        } else if (c->sipm_response_model == 0) {       // sipm_response_model light_sim.py:286-292
            double t = tt * c->light_tick_size;
            double imp = 1.0 * exp(-t / c->light_response_time) * sin(t / c->light_oscillation_period);
            imp /= c->light_oscillation_period * (c->light_response_time * c->light_response_time);
            imp *= c->light_oscillation_period * c->light_oscillation_period + c->light_response_time * c->light_response_time;
            wh[tt] = imp * c->light_tick_size;
        } else {                                        // interp light_sim.py:241-271, :295-299
            double idx = tt * c->light_tick_size / c->impulse_tick_size;
            long long i0 = (long long)floor(idx);
            double imp;
            if (i0 < 0) imp = 0; else if (i0 > n_imp - 1) imp = 0; else if ((double)i0 == idx) imp = impulse[i0];
            else if (i0 > n_imp - 2) imp = 0; else imp = impulse[i0] + (impulse[i0 + 1] - impulse[i0]) * (idx - i0);
            imp /= c->impulse_tick_size / c->light_tick_size;
            wh[tt] = imp;
        }
    }
    // taps beyond the last non-zero weight add w * in = 0 to the accumulator (and never pass the truth threshold): the SiPM impulse is
    // 256 samples long while the reference loops over the whole light window
    while (nw > 1 && wh[nw - 1] == 0.0) nw--;
    TmpPool tp(st);
    double* wd;
    cudaError_t e = tp.get(&wd, nw);
    if (e == cudaSuccess) e = cudaMemcpyAsync(wd, wh, sizeof(double) * (size_t)nw, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);       // wh is pageable and freed below
    free(wh);
    if (e != cudaSuccess) return lsb_fail_cuda(e, "light FIR weights");
    dim3 grid((unsigned)((nticks + FIR_TPB - 1) / FIR_TPB), (unsigned)ndet);
    if (mode == 0) k_light_fir<0><<<grid, FIR_TPB, 0, st>>>(in, (const long long*)in_id, in_ph, out, (long long*)out_id, out_ph, ndet, nticks, n_in, n_out, wd, nw - 1, gain);
    else k_light_fir<1><<<grid, FIR_TPB, 0, st>>>(in, (const long long*)in_id, in_ph, out, (long long*)out_id, out_ph, ndet, nticks, n_in, n_out, wd, nw - 1, gain);
    LSB_LAUNCH_CHECK("k_light_fir");
    return 0;
}

LSB_EXPORT int lsb_calc_scintillation_effect(const lsb_consts* c, const float* light_sample_inc, const int64_t* inc_true_track_id,
                                             const double* inc_true_photons, float* light_sample_inc_scint,
                                             int64_t* scint_true_track_id, double* scint_true_photons, int32_t ndet,
                                             int32_t nticks, int32_t n_true_in, int32_t n_true_out, void* stream) {
    LSB_REQUIRE(c, "calc_scintillation_effect: null consts");
    if (ndet == 0 || nticks == 0) return 0;
    LSB_REQUIRE(light_sample_inc && light_sample_inc_scint, "calc_scintillation_effect: null pointer");
    LSB_REQUIRE(n_true_in == 0 || (inc_true_track_id && inc_true_photons && (n_true_out == 0 || (scint_true_track_id && scint_true_photons))),
                "calc_scintillation_effect: null truth arrays");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = lsb_upload_consts(c, st); if (rc) return rc;
    return light_fir(0, c, light_sample_inc, inc_true_track_id, inc_true_photons, light_sample_inc_scint, scint_true_track_id,
                     scint_true_photons, ndet, nticks, n_true_in, n_true_out, nullptr, nullptr, 0, st);
}

LSB_EXPORT int lsb_calc_stat_fluctuations(const lsb_consts* c, const float* light_sample_inc, float* light_sample_inc_disc,
                                          int32_t ndet, int32_t nticks, uint64_t* rng_states, int64_t n_rng, void* stream) {
    LSB_REQUIRE(c, "calc_stat_fluctuations: null consts");
    long long n = (long long)ndet * nticks;
    if (n == 0) return 0;
    LSB_REQUIRE(light_sample_inc && light_sample_inc_disc && rng_states, "calc_stat_fluctuations: null pointer");
    LSB_REQUIRE(n_rng >= n, "calc_stat_fluctuations: rng_states shorter than ndet*nticks");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = lsb_upload_consts(c, st); if (rc) return rc;
    k_stat_fluctuations<<<lsb_blocks(n, 256), 256, 0, st>>>(light_sample_inc, light_sample_inc_disc, n, (unsigned long long*)rng_states);
    LSB_LAUNCH_CHECK("k_stat_fluctuations");
    return 0;
}

LSB_EXPORT int lsb_calc_light_detector_response(const lsb_consts* c, const float* light_sample_inc, const int64_t* inc_true_track_id,
                                                const double* inc_true_photons, float* light_response,
                                                int64_t* resp_true_track_id, double* resp_true_photons, int32_t ndet,
                                                int32_t nticks, int32_t n_true_in, int32_t n_true_out, const double* light_gain,
                                                const double* impulse_model, int32_t n_impulse, void* stream) {
    LSB_REQUIRE(c, "calc_light_detector_response: null consts");
    if (ndet == 0 || nticks == 0) return 0;
    LSB_REQUIRE(light_sample_inc && light_response && light_gain, "calc_light_detector_response: null pointer");
    LSB_REQUIRE(c->sipm_response_model == 0 || (c->sipm_response_model == 1 && impulse_model && n_impulse > 0),
                "calc_light_detector_response: SIPM_RESPONSE_MODEL 1 needs light.IMPULSE_MODEL");
    LSB_REQUIRE(n_true_in == 0 || (inc_true_track_id && inc_true_photons && (n_true_out == 0 || (resp_true_track_id && resp_true_photons))),
                "calc_light_detector_response: null truth arrays");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = lsb_upload_consts(c, st); if (rc) return rc;
    return light_fir(1, c, light_sample_inc, inc_true_track_id, inc_true_photons, light_response, resp_true_track_id,
                     resp_true_photons, ndet, nticks, n_true_in, n_true_out, light_gain, impulse_model, n_impulse, st);
}
