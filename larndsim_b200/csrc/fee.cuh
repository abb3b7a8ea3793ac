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
#include "rng.cuh"

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
__constant__ double d_fee_wp[FEE_MAX_TAPS];      // d_fee_wp[m] = w[0] + ... + w[m]

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

// ---- pre-computed inputs of the state machine (parallel, off the serial path) -------------------
// q_pre[ip][ic] = the CSA FIR of fee.py:566-573 with conv_start = max(0, floor(ic - back)), i.e. the value
// the state machine needs whenever no reset lies inside the tap window (same taps, same order).
__global__ void k_fee_fir_pre(FeeParams fp, const double* __restrict__ pixels_signals, long long U, int Tt, int Tq,
                              double* __restrict__ q_pre) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= U * Tq) return;
    long long ip = idx / Tq;
    long long ic = idx - ip * Tq;
    q_pre[idx] = fee_fir(pixels_signals + ip * Tt, ic, 0, Tt, fp);
}
// xoroshiro128+ is sequential per pixel but cheap; the Box-Muller transcendentals are not.  Step every
// pixel's stream NMAX normals ahead, store the float32 uniform pairs ([i][pixel]: coalesced) and a state
// snapshot every FEE_SNAP normals, then turn the pairs into normals with one thread per value.
#define FEE_SNAP (RNG_STEP_UNIT / 2)
__global__ void k_fee_rng_uniforms(const unsigned long long* __restrict__ rng_states, long long U, int NMAX,
                                   float2* __restrict__ uu, ulonglong2* __restrict__ snaps) {
    long long ip = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (ip >= U) return;
    Rng r; r.s0 = rng_states[2 * ip]; r.s1 = rng_states[2 * ip + 1];
    for (int i = 0; i < NMAX; i++) {
        if ((i % FEE_SNAP) == 0) snaps[(long long)(i / FEE_SNAP) * U + ip] = make_ulonglong2(r.s0, r.s1);
        float u1 = rng_uniform_f32(r);
        float u2 = rng_uniform_f32(r);
        uu[(long long)i * U + ip] = make_float2(u1, u2);
    }
    snaps[(long long)(NMAX / FEE_SNAP) * U + ip] = make_ulonglong2(r.s0, r.s1);
}
__global__ void k_fee_rng_normals(const float2* __restrict__ uu, float* __restrict__ nrm, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 u = uu[i];
    float a = sqrtf(__fmul_rn(-2.0f, logf(u.x)));              // == rng_normal_f32
    float b = cosf(__fmul_rn(6.283185307179586f, u.y));
    nrm[i] = __fmul_rn(a, b);
}

// The same normals and snapshots in ONE pass with (pixel, chunk) parallelism: thread (g, pixel) jumps the pixel's stream to draw
// RNG_STEP_UNIT * g (rng.cuh: one GF(2) matrix product per set bit of g; the matrices are shared by the warp, g is warp-uniform),
// records the snapshot, and turns its FEE_SNAP uniform pairs into normals.  Pixels are the fast index, so stores are coalesced; the
// 8-byte uniform pairs never reach HBM.  (k_fee_rng_uniforms walked each stream sequentially with one thread per pixel --
// 15 000 threads on a 300 000-thread machine -- and k_fee_rng_normals re-read its 0.8 GB of output.)
static_assert(RNG_STEP_UNIT == 2 * FEE_SNAP, "one chunk = FEE_SNAP normals");
__global__ void __launch_bounds__(128) k_fee_rng_chunks(const RngStepTable* __restrict__ tab, const unsigned long long* __restrict__ rng_states,
                                                        long long U, int n_chunks, float* __restrict__ nrm, ulonglong2* __restrict__ snaps) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= U * n_chunks) return;
    const int g = (int)(idx / U);
    const long long ip = idx - (long long)g * U;
    unsigned long long s0 = rng_states[2 * ip], s1 = rng_states[2 * ip + 1];
    for (int k = 0, bits = g; bits; k++, bits >>= 1)
        if (bits & 1) rng_matvec_dev(reinterpret_cast<const ulonglong2*>(tab->col[k]), s0, s1);
    snaps[(long long)g * U + ip] = make_ulonglong2(s0, s1);
    Rng r; r.s0 = s0; r.s1 = s1;
    float* out = nrm + (long long)g * FEE_SNAP * U + ip;
#pragma unroll 4
    for (int i = 0; i < FEE_SNAP; i++) out[(long long)i * U] = rng_normal_f32(r);
    if (g == n_chunks - 1) snaps[(long long)n_chunks * U + ip] = make_ulonglong2(r.s0, r.s1);
}

struct FeePre { const double* q_pre; int Tq; const float* nrm; const ulonglong2* snaps; int NMAX; };

// (Measured and dropped in round 2: a fast path for blocks of ticks without charge and a parallel pre-check of the ticks before the
// first current -- the response table is non-zero for ~186 us before the arrival of the charge, so a pixel's waveform is non-zero
// from its first sample window on and there are no quiet stretches worth skipping; profiles/r02_spill_pipeline.md.)
// The state machine is a serial, data-dependent loop: a global load per draw would expose the full memory
// latency every tick.  Each thread therefore keeps small windows of its inputs in shared memory
// ([slot][thread], conflict-free) and refills them with FEE_NBUF / FEE_QBUF independent loads at a time.
#define FEE_TRIG_TPB 64
// The state machine is latency-bound (a serial chain per pixel) and a batch has far fewer pixels than the GPU has
// thread slots: only the first FEE_PPW lanes of each warp carry a pixel, so the pixels spread over 32/FEE_PPW times
// more warps and every scheduler has several chains to interleave (and fewer pixels in different states per warp).
#ifndef FEE_PPW
#define FEE_PPW 8
#endif
#define FEE_TRIG_PIX (FEE_TRIG_TPB / 32 * FEE_PPW)          // pixels per block
#define FEE_NBUF 64         // staged normals per pixel (two per tick); 256 / 128 was measured: 2.55 against 1.69 ms per module0 batch
#define FEE_QBUF 32         // staged FIR values per pixel
#ifndef FEE_WBLK
#define FEE_WBLK 8          // ticks per block of the watching loop
#endif

// fee.py:548-655 as ONE flat loop: every iteration evaluates the CSA FIR at the pixel's current tick and
// then does the work of the state it is in (watching for a threshold crossing, or integrating after one),
// so the lanes of a warp -- pixels that trigger at different times -- run the same instructions.  Tick and
// draw counters are 32-bit.  The order of FIR evaluations and of noise draws is the reference's.
template <bool PRE>
__global__ void __launch_bounds__(FEE_TRIG_TPB) k_fee_trigger(FeeParams fp, FeePre pre, const double* __restrict__ pixels_signals, long long U, int Tt,
                                                      const double* __restrict__ time_ticks, int n_tt,
                                                      double* __restrict__ adc_list, double* __restrict__ adc_ticks_list, int A,
                                                      double time_padding, unsigned long long* __restrict__ rng_states,
                                                      const double* __restrict__ thresholds, FeeWindow* __restrict__ windows,
                                                      int* __restrict__ n_windows) {
    __shared__ float s_n[PRE ? FEE_NBUF * FEE_TRIG_PIX : 1];
    __shared__ double s_q[PRE ? FEE_QBUF * FEE_TRIG_PIX : 1];
    // lanes 0..FEE_PPW-1 of a warp each run one pixel ("main" lanes); the other lanes only help with the window refills
    const int lane = threadIdx.x & 31, wslot0 = (threadIdx.x >> 5) * FEE_PPW;
    const int ps = lane % FEE_PPW, part = lane / FEE_PPW;                      // pixel slot this lane serves, its share of a refill
    constexpr int NPART = 32 / FEE_PPW;
    const int slot = wslot0 + ps;
    const long long ip_s = blockIdx.x * (long long)FEE_TRIG_PIX + slot;         // pixel served (valid if < U)
    const bool is_main = lane < FEE_PPW && ip_s < U;
    if (!PRE && !is_main) return;                                              // no staged windows without the pre-computed inputs
    const long long ip = is_main ? ip_s : 0;
    const double* curre = pixels_signals + ip * Tt;
    const double* qrow = PRE ? pre.q_pre + ip * pre.Tq : nullptr;
    const float* ncol = PRE ? pre.nrm + ip : nullptr;
    float* nbuf = s_n + (PRE ? slot : 0);
    double* qbuf = s_q + (PRE ? slot : 0);
    const int cs_back = (int)ceil(fp.back);          // floor(ic - back) = ic - ceil(back) for integer ic
    const int NMAX = pre.NMAX, Tq = pre.Tq;
    const int interval = (int)fp.interval, reset_ticks = (int)fp.reset_ticks, busy_ticks = (int)fp.busy_ticks;
    const int max_hits = fp.max_adc < A ? fp.max_adc : A;
    Rng rng;
    bool inl = !PRE;
    if (!PRE) { rng.s0 = rng_states[2 * ip]; rng.s1 = rng_states[2 * ip + 1]; }
    int idx = 0, nbase = -FEE_NBUF - 1, qbase = -FEE_QBUF - 1;

    auto refill_n = [&](int i) {
        nbase = i;
#pragma unroll
        for (int k = 0; k < FEE_NBUF; k++) nbuf[k * FEE_TRIG_PIX] = (i + k < NMAX) ? __ldg(ncol + (long long)(i + k) * U) : 0.f;
    };
    auto refill_q = [&](int ic) {
        qbase = ic;
#pragma unroll
        for (int k = 0; k < FEE_QBUF; k++) qbuf[k * FEE_TRIG_PIX] = (ic + k < Tq) ? __ldg(qrow + ic + k) : 0.0;
    };
    auto draw = [&](double sigma) -> double {
        if (PRE && !inl && idx >= NMAX) {            // more draws than provisioned: continue inline from the last snapshot
            ulonglong2 sn = pre.snaps[(long long)(NMAX / FEE_SNAP) * U + ip];
            rng.s0 = sn.x; rng.s1 = sn.y; inl = true;
        }
        if (inl) { idx++; return fee_noise(rng, sigma, 1.0); }
        const int i = idx++;
        if (sigma == 0.0) return 0.0;
        if (i >= nbase + FEE_NBUF) refill_n(i);
        return (double)nbuf[(i - nbase) * FEE_TRIG_PIX] * sigma;
    };
    auto fir = [&](int ic, int last_reset) -> double {
        if (PRE && fp.BR > 0) {
            int cs = ic - cs_back;
            if (cs < 0) cs = 0;
            if (last_reset <= cs) {
                if (ic >= Tq) return 0.0;
                if (ic < qbase || ic >= qbase + FEE_QBUF) refill_q(ic);
                return qbuf[(ic - qbase) * FEE_TRIG_PIX];
            }
        }
        return fee_fir(curre, ic, last_reset, Tt, fp);
    };

    const double thr = thresholds[ip];
    int ic = 0, adc_busy = 0, last_reset = 0, iadc = 0, cleared = 0, integrate_end = 0;
    bool integrating = false, broke = false;
    double true_q = 0.0;
    bool done = !is_main;
    double q_sum = is_main ? draw(fp.reset_noise) * fp.e : 0.0;         // fee.py:557
    FeeWindow* win = windows + ip * (A + 1);
    for (;;) {
        if (PRE) {
            if (!__any_sync(0xffffffffu, !done)) break;
            // refill the staging windows for the whole warp at once (lanes drift apart by a tick or two per hit, and per-lane
            // refills would expose one memory latency per lane instead of one per warp); all 32 lanes share the loads:
            // lane = (pixel slot, part), part p fetches entries p, p + NPART, ... of that pixel's windows
            const bool mine = !done && !inl;
            const bool need = mine && ((idx + 6 > nbase + FEE_NBUF) || (fp.BR > 0 && (ic < qbase || ic >= qbase + FEE_QBUF)));
            if (__any_sync(0xffffffffu, need)) {
                const int idx_s = __shfl_sync(0xffffffffu, idx, ps), ic_s = __shfl_sync(0xffffffffu, ic, ps);
                const bool act_s = __shfl_sync(0xffffffffu, mine ? 1 : 0, ps) != 0;
                if (act_s) {
                    const float* src = pre.nrm + (long long)idx_s * U + ip_s;
#pragma unroll
                    for (int k = part; k < FEE_NBUF; k += NPART) nbuf[k * FEE_TRIG_PIX] = (idx_s + k < NMAX) ? __ldg(src + (long long)k * U) : 0.f;
                    if (fp.BR > 0) {
                        const double* qs = pre.q_pre + ip_s * pre.Tq + ic_s;
#pragma unroll
                        for (int k = part; k < FEE_QBUF; k += NPART) qbuf[k * FEE_TRIG_PIX] = (ic_s + k < Tq) ? __ldg(qs + k) : 0.0;
                    }
                }
                __syncwarp();
                if (mine) { nbase = idx; if (fp.BR > 0) qbase = ic; }
            }
            if (done) continue;
        }
        if (PRE && !inl && fp.BR > 0 && interval >= 1 && ic >= qbase && idx >= nbase) {
            // Fast paths.  While no reset lies inside the tap window the FIR values are the pre-computed ones and a tick
            // of either state is a handful of instructions on the staged windows -- same operations, same order as the
            // general step below, without its bookkeeping.
            const int cs = ic - cs_back;
            if (last_reset <= (cs < 0 ? 0 : cs)) {
                const double* qb = qbuf + (ic - qbase) * FEE_TRIG_PIX;
                int nq = qbase + FEE_QBUF - ic;
                if (Tq - ic < nq) nq = Tq - ic;
                if (!integrating) {
                    // watching for a threshold crossing (:559-593): two draws per tick
                    int nf = (nbase + FEE_NBUF - idx) >> 1;
                    if (((NMAX - idx) >> 1) < nf) nf = (NMAX - idx) >> 1;
                    if (nq < nf) nf = nq;
                    if (Tt - ic < nf) nf = Tt - ic;
                    if (iadc < max_hits && nf > 0) {
                        const float* nb = nbuf + (idx - nbase) * FEE_TRIG_PIX;
                        const double su = fp.unc_noise, sd = fp.disc_noise;
                        bool trig = false;
                        int k = 0;
                        // blocks of FEE_WBLK ticks: the noise terms and right-hand sides of the block do not depend on the
                        // running sum, so they are evaluated first (independent instructions, issued back to back); the
                        // serial part per tick is then two adds and a compare
                        while (!trig && k + FEE_WBLK <= nf) {
                            double qv[FEE_WBLK], qn[FEE_WBLK], rhs[FEE_WBLK];
#pragma unroll
                            for (int u = 0; u < FEE_WBLK; u++) {
                                qv[u] = qb[(k + u) * FEE_TRIG_PIX];
                                qn[u] = (su == 0.0 ? 0.0 : (double)nb[(2 * (k + u)) * FEE_TRIG_PIX] * su) * fp.e;
                                const double disc_noise = (sd == 0.0 ? 0.0 : (double)nb[(2 * (k + u) + 1) * FEE_TRIG_PIX] * sd) * fp.e;
                                rhs[u] = thr + disc_noise;
                            }
#pragma unroll
                            for (int u = 0; u < FEE_WBLK; u++) {
                                if (!trig) {
                                    q_sum += qv[u]; true_q += qv[u];
                                    if (adc_busy > 0) adc_busy--;
                                    k++;
                                    if (q_sum + qn[u] >= rhs[u] && adc_busy == 0) trig = true;
                                }
                            }
                        }
                        while (!trig && k < nf) {
                            const double q = qb[k * FEE_TRIG_PIX];
                            q_sum += q; true_q += q;
                            const double q_noise = (su == 0.0 ? 0.0 : (double)nb[(2 * k) * FEE_TRIG_PIX] * su) * fp.e;
                            const double disc_noise = (sd == 0.0 ? 0.0 : (double)nb[(2 * k + 1) * FEE_TRIG_PIX] * sd) * fp.e;
                            if (adc_busy > 0) adc_busy--;
                            k++;
                            if (q_sum + q_noise >= thr + disc_noise && adc_busy == 0) { trig = true; break; }
                        }
                        ic += k; idx += 2 * k;
                        if (trig) { integrate_end = ic - 1 + interval; integrating = true; }
                        continue;
                    }
                } else {
                    // integrating after a crossing (:595-613): all but the last tick of the window, which the general
                    // step takes together with the read-out
                    int nf = integrate_end - ic;
                    if (nq < nf) nf = nq;
                    if (nf > 0) {
                        for (int k = 0; k < nf; k++) { const double q = qb[k * FEE_TRIG_PIX]; q_sum += q; true_q += q; }
                        ic += nf;
                        continue;
                    }
                }
            }
        }
        if (!integrating) {
            if (!(ic < Tt || adc_busy > 0)) { if (PRE) { done = true; continue; } break; }                 // :559
            if (iadc >= max_hits) { broke = true; if (PRE) { done = true; continue; } break; }             // :561-563 (and the rows of the outputs)
        }
        const double q = fir(ic, last_reset);                            // :565-578 / :597-610
        q_sum += q; true_q += q;
        if (!integrating) {
            const double q_noise = draw(fp.unc_noise) * fp.e;            // :583-584
            const double disc_noise = draw(fp.disc_noise) * fp.e;
            if (adc_busy > 0) adc_busy--;
            if (q_sum + q_noise >= thr + disc_noise && adc_busy == 0) {  // :589
                integrate_end = ic + interval;
                integrating = true;
            }
            ic++;
            if (!integrating || ic <= integrate_end) continue;           // interval == 0: fall through to the read-out
        } else {
            ic++;
            if (ic <= integrate_end) continue;
        }
        // ---- end of the integration window (:615-652) ----
        integrating = false;
        const double adc = q_sum + draw(fp.unc_noise) * fp.e;
        const double disc_noise = draw(fp.disc_noise) * fp.e;
        if (adc < thr + disc_noise) {                                    // :619-627
            ic += reset_ticks;
            q_sum = draw(fp.reset_noise) * fp.e;
            true_q = 0.0;
            cleared = 1;
            last_reset = ic;
            continue;
        }
        FeeWindow w; w.ic0 = last_reset; w.ic1 = ic - 1; w.flags = (true_q > 0 ? 1 : 0) | (cleared ? 2 : 0);
        w.pad = 0; w.true_q = true_q;
        win[iadc] = w;
        adc_list[ip * A + iadc] = adc;
        const int crossing = ic < n_tt - 1 ? ic : n_tt - 1;              // :639-643
        const int post = ic - crossing > 0 ? ic - crossing : 0;
        adc_ticks_list[ip * A + iadc] = time_ticks[crossing] + time_padding - 2 + (double)post;
        ic += reset_ticks;
        last_reset = ic;
        adc_busy = busy_ticks;
        q_sum = draw(fp.reset_noise) * fp.e;
        true_q = 0.0;
        cleared = 0;
        iadc++;
    }
    if (!is_main) return;
    int nw = iadc;
    if (!broke && iadc < A && (ic - 1 >= last_reset || cleared)) {
        // trailing window: accumulated but never normalised (the reference leaves it in the row)
        FeeWindow w; w.ic0 = last_reset; w.ic1 = ic - 1; w.flags = (cleared ? 2 : 0); w.pad = 0; w.true_q = 0.0;
        win[iadc] = w;
        nw = iadc + 1;
    }
    n_windows[ip] = nw;
    if (PRE && !inl) {
        // the stream position after idx normals: nearest snapshot + the remaining steps
        const int snap = idx / FEE_SNAP;
        ulonglong2 sn = pre.snaps[(long long)snap * U + ip];
        rng.s0 = sn.x; rng.s1 = sn.y;
        for (int k = snap * FEE_SNAP; k < idx; k++) { rng_next(rng); rng_next(rng); }
    }
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
template <int NTAPS>
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
    if (nw == 0) return;
    // value of this (pixel, slot) waveform at pixel tick ic, as the dense tensor would hold it
    const int start32 = (int)(start < -2000000000LL ? -2000000000LL : (start > 2000000000LL ? 2000000000LL : start));
    const int my_lo = L[me].lo, my_hi = L[me].hi;                       // ticks of the row that hold data
    const int my_lo_safe = my_lo >= 0 && my_lo < T ? my_lo : 0;
    auto sample = [&](int ic) -> double {
        // branch-free so that the FEE_RING loads of a block are issued back to back
        const int it = ic - start32;
        const bool ok = (ic < Tt) && it >= my_lo && it <= my_hi;
        const float f = __ldg(row + (ok ? it : my_lo_safe));
        double v = ok ? (double)f : 0.0;
        if (n_same && ic < Tt) {                     // several row entries of one segment on this pixel (not produced by get_pixels)
            for (int q = me + 1; q < n; q++)
                if (L[q].slot == slot) {
                    long long it2 = ic - L[q].start_tick;
                    if (it2 >= L[q].lo && it2 <= L[q].hi) v += (double)__ldg(signals + (long long)L[q].e * T + it2);
                }
        }
        return v;
    };
    double* out = cf + ((long long)p * A) * K + slot;
    int iw = 0;
    FeeWindow w = W[0];
    double acc = 0.0;
    if (fp.BR > 0) {
        // one tick loop for the whole waveform, identical for every lane of the warp (ring position = ic mod
        // FEE_RING); window starts / ends are rare per-lane events.  NTAPS = taps of the FIR (<= FEE_RING).
        double ring[FEE_RING];
#pragma unroll
        for (int r = 0; r < FEE_RING; r++) ring[r] = 0.0;
        const int last = W[nw - 1].ic1;
        int w0 = w.ic0, w1 = w.ic1;
        for (int ic0 = 0; ic0 <= last && iw < nw; ic0 += FEE_RING) {
            double x[FEE_RING];                      // independent loads first
#pragma unroll
            for (int r = 0; r < FEE_RING; r++) x[r] = sample(ic0 + r) * fp.TS;
#pragma unroll
            for (int r = 0; r < FEE_RING; r++) {
                const int ic = ic0 + r;
                if (ic == w0) {
#pragma unroll
                    for (int k = 0; k < FEE_RING; k++) ring[k] = 0.0;
                    acc = (w.flags & 2) ? 0.0 : *out;
                }
                ring[r] = x[r];
                if (ic >= w0 && ic <= w1) {
                    // taps jc = ic-(NTAPS-1) .. ic (ascending): ring slot of jc = (r - (ic - jc)) mod FEE_RING
#pragma unroll
                    for (int d = NTAPS - 1; d >= 0; d--) acc += ring[(r - d + 2 * FEE_RING) % FEE_RING] * d_fee_w[d];
                    if (ic == w1) {
                        if (w.flags & 1) acc /= w.true_q;
                        *out = acc;
                        iw++; out += K;
                        if (iw < nw) { w = W[iw]; w0 = w.ic0; w1 = w.ic1; } else { w0 = 2147483647; w1 = -1; }
                    }
                }
            }
        }
    } else {
        for (; iw < nw; iw++, out += K) {
            w = W[iw];
            acc = (w.flags & 2) ? 0.0 : *out;
            long long hi = w.ic1 < Tt - 1 ? w.ic1 : Tt - 1;
            for (long long ic = w.ic0; ic <= hi; ic++) acc += sample((int)ic) * fp.TS;
            if (w.flags & 1) acc /= w.true_q;
            *out = acc;
        }
    }
    // windows without a single FIR evaluation (a failed trigger cleared the row at the very end)
    for (; iw < nw; iw++, out += K) {
        w = W[iw];
        if (w.ic1 >= w.ic0) continue;               // cannot happen: evaluated windows are closed in the loop above
        if (w.flags & 2) *out = 0.0;
    }
}
// Order-free variant (default of the fused chain): the double sum of fee.py:566-573 over a window [ic0, ic1],
//     sum_{ic=ic0}^{ic1} sum_{jc=max(ic0,ic-(n-1))}^{min(ic,Tt-1)} I[jc]*dt*w[ic-jc]
//   = dt * sum_{jc=ic0}^{min(ic1,Tt-1)} I[jc] * Wp[min(n-1, ic1-jc)],        Wp = prefix sums of w,
// is a weighted sum of the waveform, evaluated by one warp per (pixel, slot) with coalesced row reads and a
// shuffle reduction.  Mathematically identical to the replay above; the float64 rounding differs (1e-15).
__global__ void __launch_bounds__(128) k_fee_fractions_fast(FeeParams fp, const float* __restrict__ signals, int T, long long U,
                                                            int Tt, int K, const long long* __restrict__ offs,
                                                            const int* __restrict__ counts, const SumEntry* __restrict__ sorted,
                                                            long long n_sorted_total, const int* __restrict__ entry_pixel,
                                                            const FeeWindow* __restrict__ windows,
                                                            const int* __restrict__ n_windows, int A, double* __restrict__ cf) {
    const int lane = threadIdx.x & 31;
    const long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;      // one warp per sorted entry
    if (i >= n_sorted_total) return;
    const int p = entry_pixel[i];
    if (p < 0) return;
    const SumEntry* L = sorted + offs[p];
    const int n = counts[p];
    const int me = (int)(i - offs[p]);
    const int slot = L[me].slot;
    for (int q = 0; q < me; q++) if (L[q].slot == slot) return;          // an earlier entry owns this (pixel, slot)
    const int nw = n_windows[p];
    const FeeWindow* W = windows + (long long)p * (A + 1);
    const int ntap = fp.BR > 0 ? fp.n_taps : 1;
    for (int iw = 0; iw < nw; iw++) {
        const FeeWindow w = W[iw];
        double acc = 0.0;
        const int hi = w.ic1 < Tt - 1 ? w.ic1 : Tt - 1;
        for (int q = me; q < n; q++) {                                    // normally exactly one entry per (pixel, slot)
            if (L[q].slot != slot) continue;
            const float* row = signals + (long long)L[q].e * T;
            const long long start = L[q].start_tick;
            long long lo = w.ic0 > start + L[q].lo ? w.ic0 : start + L[q].lo;      // ticks where this row has samples
            long long up = hi < start + L[q].hi ? hi : start + L[q].hi;
            for (long long jc = lo + lane; jc <= up; jc += 32) {
                const double v = (double)__ldg(row + (jc - start));
                const int m = (int)(w.ic1 - jc);
                const double wt = fp.BR > 0 ? d_fee_wp[m < ntap - 1 ? m : ntap - 1] : 1.0;
                acc += v * wt;
            }
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            double* out = cf + ((long long)p * A + iw) * K + slot;
            double r = acc * fp.TS + ((w.flags & 2) ? 0.0 : *out);
            if (w.flags & 1) r /= w.true_q;
            *out = r;
        }
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

// provision of normals per pixel for one call (see k_fee_rng_uniforms) and the bytes of temporaries fee_run needs
static long long fee_nmax(const FeeParams& fp, int Tt, int A) {
    const long long iters = (long long)Tt + fp.busy_ticks + 2;
    long long nmax = 1 + 2 * iters + 3 * (iters / (fp.interval + 1) + A + 2) + 8;
    return (nmax + FEE_SNAP - 1) / FEE_SNAP * FEE_SNAP;
}
static size_t fee_scratch_bytes(const lsb_consts* c, long long U, int Tt, int A, long long n_entries) {
    FeeParams fp; double w[FEE_MAX_TAPS];
    if (fee_params(c, fp, w)) return 0;
    const long long nmax = fee_nmax(fp, Tt, A);
    return (size_t)U * ((size_t)nmax * 12 + (size_t)(Tt + fp.n_taps) * 8 + (size_t)(A + 1) * sizeof(FeeWindow) + 4 +
                        16 * (size_t)(nmax / FEE_SNAP + 1)) + (size_t)n_entries * 4 + 16 * 256;
}

// sparse context: the (pixel, slot) entries of sum_pixel_signals and the per-segment waveforms
struct FeeSparse { const float* signals; int T; const long long* offs; const int* counts; const SumEntry* sorted; long long n_entries_cap;
                   int exact; /* 1: replay the reference's summation order (bit-identical fractions) */ };

static int fee_run(const lsb_consts* c, const double* pixels_signals, const double* pst, const FeeSparse* sp, long long U, int Tt,
                   int K, const double* time_ticks, int n_time_ticks, double* adc_list, double* adc_ticks_list, int A,
                   double time_padding, uint64_t* rng_states, double* current_fractions, const double* pixel_thresholds,
                   cudaStream_t st) {
    FeeParams fp;
    double w_host[FEE_MAX_TAPS];
    if (fee_params(c, fp, w_host)) return -1;
    if (fp.n_w > 0) {
        LSB_CUDA(cudaMemcpyToSymbolAsync(d_fee_w, w_host, sizeof(double) * fp.n_w, 0, cudaMemcpyHostToDevice, st));
        double wp_host[FEE_MAX_TAPS];
        double run = 0.0;
        for (int d = 0; d < fp.n_w; d++) { run += w_host[d]; wp_host[d] = run; }
        LSB_CUDA(cudaMemcpyToSymbolAsync(d_fee_wp, wp_host, sizeof(double) * fp.n_w, 0, cudaMemcpyHostToDevice, st));
    }
    TmpPool tp(st);
    FeeWindow* windows; int* n_windows;
    LSB_CUDA(tp.get(&windows, U * (long long)(A + 1)));
    LSB_CUDA(tp.get(&n_windows, U));
    {
        // provision of pre-computed noise / FIR values per pixel (see k_fee_rng_uniforms)
        const long long nmax = fee_nmax(fp, Tt, A);
        const int Tq = Tt + fp.n_taps;
        const double pre_bytes = (double)U * ((double)nmax * 12.0 + (double)Tq * 8.0);
        FeePre pre; pre.q_pre = nullptr; pre.Tq = Tq; pre.nrm = nullptr; pre.snaps = nullptr; pre.NMAX = (int)nmax;
        if (pre_bytes < 12e9 && nmax < 2000000) {
            double* q_pre; float* nrm; ulonglong2* snaps;
            LSB_CUDA(tp.get(&q_pre, U * (long long)Tq));
            LSB_CUDA(tp.get(&nrm, U * nmax));
            LSB_CUDA(tp.get(&snaps, U * (nmax / FEE_SNAP + 1)));
            k_fee_fir_pre<<<lsb_blocks(U * (long long)Tq, 256), 256, 0, st>>>(fp, pixels_signals, U, Tt, Tq, q_pre);
            LSB_LAUNCH_CHECK("k_fee_fir_pre");
            const int n_chunks = (int)(nmax / FEE_SNAP);
            const RngStepTable* step_tab = (fp.reset_noise != 0.0 || fp.unc_noise != 0.0 || fp.disc_noise != 0.0) && n_chunks < (1 << RNG_STEP_LEVELS)
                                               ? rng_step_table_dev() : nullptr;
            static int chunked = -1;
            if (chunked < 0) { const char* e = getenv("LSB_FEE_RNG_CHUNKED"); chunked = (e && e[0] == '0') ? 0 : 1; }
            if (step_tab && chunked) {
                k_fee_rng_chunks<<<lsb_blocks(U * (long long)n_chunks, 128), 128, 0, st>>>(step_tab, (const unsigned long long*)rng_states, U, n_chunks, nrm, snaps);
                LSB_LAUNCH_CHECK("k_fee_rng_chunks");
            } else {
                float2* uu;
                LSB_CUDA(tp.get(&uu, U * nmax));
                k_fee_rng_uniforms<<<lsb_blocks(U, 64), 64, 0, st>>>((const unsigned long long*)rng_states, U, (int)nmax, uu, snaps);
                LSB_LAUNCH_CHECK("k_fee_rng_uniforms");
                k_fee_rng_normals<<<lsb_blocks(U * nmax, 256), 256, 0, st>>>(uu, nrm, U * nmax);
                LSB_LAUNCH_CHECK("k_fee_rng_normals");
            }
            pre.q_pre = q_pre; pre.nrm = nrm; pre.snaps = snaps;
            k_fee_trigger<true><<<lsb_blocks(U, FEE_TRIG_PIX), FEE_TRIG_TPB, 0, st>>>(fp, pre, pixels_signals, U, Tt, time_ticks, n_time_ticks, adc_list,
                                                                 adc_ticks_list, A, time_padding, (unsigned long long*)rng_states,
                                                                 pixel_thresholds, windows, n_windows);
        } else {
            k_fee_trigger<false><<<lsb_blocks(U, FEE_TRIG_PIX), FEE_TRIG_TPB, 0, st>>>(fp, pre, pixels_signals, U, Tt, time_ticks, n_time_ticks, adc_list,
                                                                    adc_ticks_list, A, time_padding, (unsigned long long*)rng_states,
                                                                    pixel_thresholds, windows, n_windows);
        }
        LSB_LAUNCH_CHECK("k_fee_trigger");
    }
    if (K > 0 && A > 0) {
        const bool ring_ok = fp.BR <= 0 || fp.n_taps <= FEE_RING;
        if (sp && (ring_ok || !sp->exact)) {
            int* entry_pixel;
            LSB_CUDA(tp.get(&entry_pixel, sp->n_entries_cap));
            LSB_CUDA(cudaMemsetAsync(entry_pixel, 0xff, sp->n_entries_cap * 4, st));
            k_entry_pixel<<<lsb_blocks(U, 256), 256, 0, st>>>(sp->offs, sp->counts, U, entry_pixel);
            LSB_LAUNCH_CHECK("k_entry_pixel");
            // entries are packed at the front of `sorted` (exclusive scan of the bucket sizes); unused tail has pixel -1
            if (!sp->exact) {
                k_fee_fractions_fast<<<lsb_blocks(sp->n_entries_cap * 32, 128), 128, 0, st>>>(
                    fp, sp->signals, sp->T, U, Tt, K, sp->offs, sp->counts, sp->sorted, sp->n_entries_cap, entry_pixel, windows,
                    n_windows, A, current_fractions);
                LSB_LAUNCH_CHECK("k_fee_fractions_fast");
                return 0;
            }
#define FEE_FRAC_LAUNCH(N) k_fee_fractions_sparse<N><<<lsb_blocks(sp->n_entries_cap, 128), 128, 0, st>>>(                       \
                fp, sp->signals, sp->T, U, Tt, K, sp->offs, sp->counts, sp->sorted, sp->n_entries_cap, entry_pixel, windows,       \
                n_windows, A, current_fractions)
            switch (fp.BR > 0 ? fp.n_taps : 1) {
                case 1: FEE_FRAC_LAUNCH(1); break;   case 2: FEE_FRAC_LAUNCH(2); break;   case 3: FEE_FRAC_LAUNCH(3); break;
                case 4: FEE_FRAC_LAUNCH(4); break;   case 5: FEE_FRAC_LAUNCH(5); break;   case 6: FEE_FRAC_LAUNCH(6); break;
                case 7: FEE_FRAC_LAUNCH(7); break;   case 8: FEE_FRAC_LAUNCH(8); break;   case 9: FEE_FRAC_LAUNCH(9); break;
                case 10: FEE_FRAC_LAUNCH(10); break; default: FEE_FRAC_LAUNCH(11); break;
            }
#undef FEE_FRAC_LAUNCH
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
