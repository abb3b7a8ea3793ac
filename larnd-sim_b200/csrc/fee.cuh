// fee.cuh -- front-end electronics: get_adc_values (fee.py:517-655).
//
// The reference runs one thread per pixel through the whole waveform and, at every tick, also
// convolves all K=50 per-segment waveforms (11 taps each) to keep `current_fractions` up to
// date.  Here the work is split so each part maps onto the machine:
//   k_fee_trigger    thread per pixel: the sequential self-trigger state machine on the summed
//                    waveform only (CSA rise-time FIR, noise draws from the pixel's xoroshiro128+
//                    stream in the reference's order, threshold / hold / reset / busy logic).  It
//                    emits the hits and, per hit, the tick window [ic0, ic1] over which the
//                    reference accumulated `current_fractions` since the last reset.
//   k_fee_fractions  warp per pixel, lane per segment slot: replays exactly those windows on the
//                    dense per-segment waveforms with the reference's summation order
//                    (tick-major, taps ascending), coalesced 8*K-byte rows, all-zero rows skipped
//                    (adding 0.0 is exact).  HBM-bound: one pass over pixels_signals_tracks.
// FIR weights exp((jc-ic)*dt/tau)*(1-exp(-dt/tau)) are evaluated once on the host in float64.
#pragma once
#include "common.cuh"
#include "pixelmap.cuh"

#define FEE_MAX_TAPS 256

struct FeeParams {
    double TS, BR, e;
    double back;                 // 10*BUFFER_RISETIME/TIME_SAMPLING (fee.py:567)
    double reset_noise, unc_noise, disc_noise;    // already in electrons
    long long interval;          // round((3+ADC_HOLD_DELAY)*CLOCK_CYCLE/dt)
    long long reset_ticks;       // round(RESET_CYCLES*CLOCK_CYCLE/dt)
    long long busy_ticks;        // round(ADC_BUSY_DELAY*CLOCK_CYCLE/dt)
    int max_adc, n_w;
    int n_taps;                  // ceil(back)+1 taps jc = floor(ic-back) .. ic
};
__constant__ double d_fee_w[FEE_MAX_TAPS];

struct FeeWindow {
    int ic0, ic1;                // FIR evaluated at every ic in [ic0, ic1]; last_reset == ic0
    int flags;                   // bit0: normalise by true_q; bit1: start from zero (a failed trigger cleared the row)
    int pad;
    double true_q;
};

// one FIR evaluation (fee.py:566-578) on the summed waveform
__device__ __forceinline__ double fee_fir(const double* __restrict__ curre, long long ic, long long last_reset, int Tt,
                                          const FeeParams& fp) {
    double q = 0.0;
    if (fp.BR > 0) {
        long long conv_start = (long long)floor((double)ic - fp.back);
        if (last_reset > conv_start) conv_start = last_reset;
        long long jend = ic + 1 < Tt ? ic + 1 : Tt;
        for (long long jc = conv_start; jc < jend; jc++) {
            double c = curre[jc];
            if (c == 0.0) continue;                      // q + 0*w == q
            q += c * fp.TS * d_fee_w[ic - jc];
        }
    } else if (ic < Tt) {
        q += curre[ic] * fp.TS;
    }
    return q;
}
// normal * sigma with the reference's promotion ((float32 -> float64) * float64 * float64); the
// stream always advances by one normal (two uniforms); transcendental work is skipped when the
// noise constant is exactly 0 (the product is then +-0 unless the normal itself is non-finite,
// probability 2^-53 per draw).
__device__ __forceinline__ double fee_noise(Rng& r, double sigma, double e) {
    if (sigma == 0.0) { rng_next(r); rng_next(r); return 0.0; }
    return (double)rng_normal_f32(r) * sigma * e;
}

__global__ void __launch_bounds__(128) k_fee_trigger(FeeParams fp, const double* __restrict__ pixels_signals, long long U, int Tt,
                                                      const double* __restrict__ time_ticks, int n_tt,
                                                      double* __restrict__ adc_list, double* __restrict__ adc_ticks_list, int A,
                                                      double time_padding, unsigned long long* __restrict__ rng_states,
                                                      const double* __restrict__ thresholds, FeeWindow* __restrict__ windows,
                                                      int* __restrict__ n_windows) {
    long long ip = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (ip >= U) return;
    const double* curre = pixels_signals + ip * Tt;
    Rng rng; rng.s0 = rng_states[2 * ip]; rng.s1 = rng_states[2 * ip + 1];
    const double thr = thresholds[ip];
    long long ic = 0, adc_busy = 0, last_reset = 0;
    int iadc = 0, cleared = 0;
    double true_q = 0.0;
    double q_sum = fee_noise(rng, fp.reset_noise, 1.0) * fp.e;
    FeeWindow* win = windows + ip * (A + 1);
    bool broke = false;
    while (ic < Tt || adc_busy > 0) {
        if (iadc >= fp.max_adc || iadc >= A) { broke = true; break; }
        double q = fee_fir(curre, ic, last_reset, Tt, fp);
        q_sum += q; true_q += q;
        double q_noise = fee_noise(rng, fp.unc_noise, 1.0) * fp.e;
        double disc_noise = fee_noise(rng, fp.disc_noise, 1.0) * fp.e;
        if (adc_busy > 0) adc_busy--;
        if (q_sum + q_noise >= thr + disc_noise && adc_busy == 0) {
            long long integrate_end = ic + fp.interval;
            ic++;
            while (ic <= integrate_end) {
                q = fee_fir(curre, ic, last_reset, Tt, fp);
                q_sum += q; true_q += q; ic++;
            }
            double adc = q_sum + fee_noise(rng, fp.unc_noise, 1.0) * fp.e;
            disc_noise = fee_noise(rng, fp.disc_noise, 1.0) * fp.e;
            if (adc < thr + disc_noise) {
                ic += fp.reset_ticks;
                q_sum = fee_noise(rng, fp.reset_noise, 1.0) * fp.e;
                true_q = 0.0;
                cleared = 1;
                last_reset = ic;
                continue;
            }
            FeeWindow w; w.ic0 = (int)last_reset; w.ic1 = (int)(ic - 1); w.flags = (true_q > 0 ? 1 : 0) | (cleared ? 2 : 0);
            w.pad = 0; w.true_q = true_q;
            win[iadc] = w;
            adc_list[ip * A + iadc] = adc;
            long long crossing = ic < n_tt - 1 ? ic : n_tt - 1;
            long long post = ic - crossing > 0 ? ic - crossing : 0;
            adc_ticks_list[ip * A + iadc] = time_ticks[crossing] + time_padding - 2 + (double)post;
            ic += fp.reset_ticks;
            last_reset = ic;
            adc_busy = fp.busy_ticks;
            q_sum = fee_noise(rng, fp.reset_noise, 1.0) * fp.e;
            true_q = 0.0;
            cleared = 0;
            iadc++;
            continue;
        }
        ic++;
    }
    int nw = iadc;
    if (!broke && iadc < A && (ic - 1 >= last_reset || cleared)) {
        // trailing window: accumulated but never normalised (the reference leaves it in the row)
        FeeWindow w; w.ic0 = (int)last_reset; w.ic1 = (int)(ic - 1); w.flags = (cleared ? 2 : 0); w.pad = 0; w.true_q = 0.0;
        win[iadc] = w;
        nw = iadc + 1;
    }
    n_windows[ip] = nw;
    rng_states[2 * ip] = rng.s0; rng_states[2 * ip + 1] = rng.s1;
}

// warp per pixel; lane = segment slot (K > 32: several passes)
__global__ void __launch_bounds__(128) k_fee_fractions(FeeParams fp, const double* __restrict__ pst, long long U, int Tt, int K,
                                                       const FeeWindow* __restrict__ windows, const int* __restrict__ n_windows,
                                                       int A, double* __restrict__ cf) {
    const int lane = threadIdx.x & 31;
    long long ip = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (ip >= U) return;
    const int nw = n_windows[ip];
    const double* base = pst + ip * (long long)Tt * K;
    for (int k0 = 0; k0 < K; k0 += 32) {
        const int k = k0 + lane;
        const bool kok = k < K;
        for (int iw = 0; iw < nw; iw++) {
            const FeeWindow w = windows[ip * (A + 1) + iw];
            double* out = cf + (ip * A + iw) * (long long)K + k;
            double acc = (kok && !(w.flags & 2)) ? *out : 0.0;
            const long long last_reset = w.ic0;
            if (fp.BR > 0) {
                // rows jc >= Tt never contribute; rows are visited tick-major, taps ascending
                for (long long ic = w.ic0; ic <= w.ic1; ic++) {
                    long long conv_start = (long long)floor((double)ic - fp.back);
                    if (last_reset > conv_start) conv_start = last_reset;
                    long long jend = ic + 1 < Tt ? ic + 1 : Tt;
                    for (long long jc = conv_start; jc < jend; jc++) {
                        double v = kok ? __ldg(base + jc * K + k) : 0.0;
                        if (v != 0.0) acc += v * fp.TS * d_fee_w[ic - jc];
                    }
                }
            } else {
                long long hi = w.ic1 < Tt - 1 ? w.ic1 : Tt - 1;
                for (long long ic = w.ic0; ic <= hi; ic++) {
                    double v = kok ? __ldg(base + ic * K + k) : 0.0;
                    if (v != 0.0) acc += v * fp.TS;
                }
            }
            if (w.flags & 1) acc /= w.true_q;
            if (kok) *out = acc;
        }
    }
}

// Sparse variant for the fused chain: the dense [U][Tt][K] tensor is never materialised.  One thread per
// (pixel, slot) pair -- the entry list built for sum_pixel_signals says which signals[segment][pixel] row
// feeds slot k of pixel p -- walks the pixel's windows with an 11-deep (FEE_RING) register ring of
// x = I*dt, adding x*w in the reference's order.  Taps the reference does not visit (before the last
// reset, beyond the waveform) enter as +0.0, which leaves a float64 sum unchanged.  Values are the
// float32 samples the dense tensor would hold, so the fractions are bit-identical to the dense path.
#define FEE_RING 11
__global__ void __launch_bounds__(128) k_fee_fractions_sparse(FeeParams fp, const float* __restrict__ signals, int T, long long U,
                                                              int Tt, int K, const long long* __restrict__ offs,
                                                              const int* __restrict__ counts, const SumEntry* __restrict__ sorted,
                                                              long long n_sorted_total, const int* __restrict__ entry_pixel,
                                                              const FeeWindow* __restrict__ windows,
                                                              const int* __restrict__ n_windows, int A, double* __restrict__ cf) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_sorted_total) return;
    const int p = entry_pixel[i];
    if (p < 0) return;                                                   // unused tail of the entry array
    const SumEntry* L = sorted + offs[p];
    const int n = counts[p];
    const int me = (int)(i - offs[p]);
    const int slot = L[me].slot;
    for (int q = 0; q < me; q++) if (L[q].slot == slot) return;          // an earlier entry owns this (pixel, slot)
    int n_same = 0;
    for (int q = me + 1; q < n; q++) if (L[q].slot == slot) n_same++;
    const float* row = signals + (long long)L[me].e * T;
    const long long start = L[me].start_tick;
    const int nw = n_windows[p];
    const FeeWindow* W = windows + (long long)p * (A + 1);
    for (int iw = 0; iw < nw; iw++) {
        const FeeWindow w = W[iw];
        double* out = cf + ((long long)p * A + iw) * K + slot;
        double acc = (w.flags & 2) ? 0.0 : *out;
        if (fp.BR > 0) {
            double ring[FEE_RING];
#pragma unroll
            for (int r = 0; r < FEE_RING; r++) ring[r] = 0.0;
            for (long long ic0 = w.ic0; ic0 <= w.ic1; ic0 += FEE_RING) {
                // the FEE_RING loads of this block are independent: issue them before the dependent adds
                double x[FEE_RING];
#pragma unroll
                for (int r = 0; r < FEE_RING; r++) {
                    const long long ic = ic0 + r;
                    double v = 0.0;
                    if (ic <= w.ic1 && ic < Tt) {
                        long long it = ic - start;
                        if (it >= 0 && it < T) v = (double)__ldg(row + it);
                        if (n_same) {
                            for (int q = me + 1; q < n; q++)
                                if (L[q].slot == slot) {
                                    long long it2 = ic - L[q].start_tick;
                                    if (it2 >= 0 && it2 < T) v += (double)__ldg(signals + (long long)L[q].e * T + it2);
                                }
                        }
                    }
                    x[r] = v * fp.TS;
                }
#pragma unroll
                for (int r = 0; r < FEE_RING; r++) {
                    if (ic0 + r <= w.ic1) {
                        ring[r] = x[r];
                        // taps jc = ic-10 .. ic (ascending): ring slot of jc = (r - (ic - jc)) mod FEE_RING
#pragma unroll
                        for (int d = FEE_RING - 1; d >= 0; d--) {
                            if (d < fp.n_taps) {
                                const double xv = ring[(r - d + 2 * FEE_RING) % FEE_RING];
                                acc += xv * d_fee_w[d];
                            }
                        }
                    }
                }
            }
        } else {
            long long hi = w.ic1 < Tt - 1 ? w.ic1 : Tt - 1;
            for (long long ic = w.ic0; ic <= hi; ic++) {
                double v = 0.0;
                long long it = ic - start;
                if (it >= 0 && it < T) v = (double)__ldg(row + it);
                for (int q = me + 1; n_same && q < n; q++)
                    if (L[q].slot == slot) {
                        long long it2 = ic - L[q].start_tick;
                        if (it2 >= 0 && it2 < T) v += (double)__ldg(signals + (long long)L[q].e * T + it2);
                    }
                acc += v * fp.TS;
            }
        }
        if (w.flags & 1) acc /= w.true_q;
        *out = acc;
    }
}
__global__ void k_entry_pixel(const long long* __restrict__ offs, const int* __restrict__ counts, long long U, int* __restrict__ entry_pixel) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= U) return;
    long long o = offs[p];
    for (int q = 0; q < counts[p]; q++) entry_pixel[o + q] = (int)p;
}

static int fee_params(const lsb_consts* c, FeeParams& fp, double* w_host) {
    fp.TS = c->time_sampling; fp.BR = c->buffer_risetime; fp.e = c->unit_e;
    fp.back = fp.BR > 0 ? 10 * fp.BR / fp.TS : 0.0;
    fp.reset_noise = c->reset_noise_charge; fp.unc_noise = c->uncorrelated_noise_charge; fp.disc_noise = c->discriminator_noise;
    fp.interval = llrint((3 * c->clock_cycle + c->adc_hold_delay * c->clock_cycle) / fp.TS);
    fp.reset_ticks = llrint(c->reset_cycles * c->clock_cycle / fp.TS);
    fp.busy_ticks = llrint(c->adc_busy_delay * c->clock_cycle / fp.TS);
    fp.max_adc = c->max_adc_values;
    fp.n_w = 0; fp.n_taps = 1;
    if (fp.BR > 0) {
        long long n = (long long)ceil(fp.back) + 2;
        if (n > FEE_MAX_TAPS) return lsb_fail_arg("get_adc_values: 10*BUFFER_RISETIME/TIME_SAMPLING too large (max 254 taps)");
        fp.n_w = (int)n; fp.n_taps = (int)ceil(fp.back) + 1;
        for (long long d = 0; d < n; d++) w_host[d] = exp((double)(-d) * fp.TS / fp.BR) * (1 - exp(-fp.TS / fp.BR));
    }
    return 0;
}

// sparse context: the (pixel, slot) entries of sum_pixel_signals and the per-segment waveforms
struct FeeSparse { const float* signals; int T; const long long* offs; const int* counts; const SumEntry* sorted; long long n_entries_cap; };

static int fee_run(const lsb_consts* c, const double* pixels_signals, const double* pst, const FeeSparse* sp, long long U, int Tt,
                   int K, const double* time_ticks, int n_time_ticks, double* adc_list, double* adc_ticks_list, int A,
                   double time_padding, uint64_t* rng_states, double* current_fractions, const double* pixel_thresholds,
                   cudaStream_t st) {
    FeeParams fp;
    double w_host[FEE_MAX_TAPS];
    if (fee_params(c, fp, w_host)) return -1;
    if (fp.n_w > 0) LSB_CUDA(cudaMemcpyToSymbolAsync(d_fee_w, w_host, sizeof(double) * fp.n_w, 0, cudaMemcpyHostToDevice, st));
    TmpPool tp(st);
    FeeWindow* windows; int* n_windows;
    LSB_CUDA(tp.get(&windows, U * (long long)(A + 1)));
    LSB_CUDA(tp.get(&n_windows, U));
    k_fee_trigger<<<lsb_blocks(U, 128), 128, 0, st>>>(fp, pixels_signals, U, Tt, time_ticks, n_time_ticks, adc_list, adc_ticks_list, A,
                                                     time_padding, (unsigned long long*)rng_states, pixel_thresholds, windows, n_windows);
    LSB_LAUNCH_CHECK("k_fee_trigger");
    if (K > 0 && A > 0) {
        const bool ring_ok = fp.BR <= 0 || fp.n_taps <= FEE_RING;
        if (sp && ring_ok) {
            int* entry_pixel;
            LSB_CUDA(tp.get(&entry_pixel, sp->n_entries_cap));
            LSB_CUDA(cudaMemsetAsync(entry_pixel, 0xff, sp->n_entries_cap * 4, st));
            k_entry_pixel<<<lsb_blocks(U, 256), 256, 0, st>>>(sp->offs, sp->counts, U, entry_pixel);
            LSB_LAUNCH_CHECK("k_entry_pixel");
            // entries are packed at the front of `sorted` (exclusive scan of the bucket sizes); unused tail has pixel -1
            k_fee_fractions_sparse<<<lsb_blocks(sp->n_entries_cap, 128), 128, 0, st>>>(
                fp, sp->signals, sp->T, U, Tt, K, sp->offs, sp->counts, sp->sorted, sp->n_entries_cap, entry_pixel, windows,
                n_windows, A, current_fractions);
            LSB_LAUNCH_CHECK("k_fee_fractions_sparse");
        } else {
            if (!pst) return lsb_fail_arg("get_adc_values: dense per-segment waveforms required for this rise time");
            k_fee_fractions<<<lsb_blocks(U * 32, 128), 128, 0, st>>>(fp, pst, U, Tt, K, windows, n_windows, A, current_fractions);
            LSB_LAUNCH_CHECK("k_fee_fractions");
        }
    }
    return 0;
}

LSB_EXPORT int lsb_get_adc_values(const lsb_consts* c, const double* pixels_signals, const double* pixels_signals_tracks,
                                  int64_t U, int32_t Tt, int32_t K, const double* time_ticks, int32_t n_time_ticks,
                                  double* adc_list, double* adc_ticks_list, int32_t A, double time_padding,
                                  uint64_t* rng_states, int64_t n_rng, double* current_fractions,
                                  const double* pixel_thresholds, void* stream) {
    LSB_REQUIRE(c, "get_adc_values: null consts");
    if (U == 0) return 0;
    LSB_REQUIRE(pixels_signals && time_ticks && adc_list && adc_ticks_list && rng_states && pixel_thresholds,
                "get_adc_values: null pointer");
    LSB_REQUIRE(K == 0 || (pixels_signals_tracks && current_fractions), "get_adc_values: null per-segment arrays");
    LSB_REQUIRE(n_rng >= U, "get_adc_values: rng_states shorter than the number of pixels");
    LSB_REQUIRE(n_time_ticks >= 1 && Tt >= 0 && A >= 0, "get_adc_values: bad sizes");
    return fee_run(c, pixels_signals, pixels_signals_tracks, nullptr, U, Tt, K, time_ticks, n_time_ticks, adc_list, adc_ticks_list,
                   A, time_padding, rng_states, current_fractions, pixel_thresholds, (cudaStream_t)stream);
}
