// current.cuh -- induced current on the pixels: tracks_current_mc (detsim.py:258-348) and the
// deterministic tracks_current (detsim.py:351-453).
//
// tracks_current_mc, production ("cloud") mode, is three kernels:
//   k_mc_pairs      thread per (segment,pixel): tick-independent geometry (detsim.py:275-322), FP64
//                   where Numba promotes, float32-rounded where Numba keeps float32.
//   k_mc_sampler    thread per pair: consumes that pair's xoroshiro128+ stream (z,x,y normals per
//                   step), resolves every sample to {LUT row offset, tick shift, valid tick range}
//                   with the reference's FP64 round()/window expressions evaluated exactly.
//   k_mc_sort       warp per pair: the pair's flat LUT offsets (row offset + tick shift) sorted in
//                   registers (bitonic network, 512 keys per pass).
//   k_mc_accumulate CTA per pair: signal[tick] = charge * sum_samples LUT[row][stride*tick+shift].
//                   Ticks inside the intersection of all sample windows ("interior", ~98% of the
//                   work): samples whose offsets fall into the same aligned 4-word group share one
//                   register window of the LUT -- a thread owns 4 consecutive ticks, fetches the
//                   window with aligned LDG.128 (reusing the upper half when the next group is
//                   adjacent) and applies every sample of the group as a count-weighted FFMA.  About
//                   1 LDG.128 per 8 FFMA instead of 1 LDG.32 per FADD (the generic path, kept for
//                   float64 tables and non-unit sampling ratios).
//                   Phase-aligned variant (default for phase-split tables, where few offsets share an aligned
//                   block): the offsets are ordered by (offset mod 4, offset / 4) and folded into one 4-byte
//                   record per DISTINCT offset; for alignment class d a lane owns ticks 4 lane - d .. + 3 of a
//                   124-tick block, so the words it needs are ONE aligned float4 -- 1 LDG.128 + 4 FFMA per
//                   distinct offset, half the instructions, bound by L2 -> L1 bandwidth instead of the L1 data
//                   pipe (profiles/r02_acc_aligned.md).
//                   The few edge ticks and irregular samples take an exact predicated path.
// Replay mode (k_mc_replay) is the reference's thread-for-thread draw pattern, sequential per pair.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "glue.cuh"

#ifndef ACC_GW
#define ACC_GW 4                         // table words per group of the grouped accumulate path: 4 (lane = 4 ticks, LDG.128
#endif                                   // windows; default) or 8 (lane = 8 ticks, 256-bit loads: measured slower on B200,
                                         // 5.05 + 0.75 ms against 4.62 + 0.56 ms, profiles/r01_mc_grouped.md)

struct PairRec {
    double x_p, y_p, t_start, z_anode;
    double sub_start[3];
    double dir[3];
    double step, charge, sig_t, sig_l;
    long long nstep;          // steps * MC_SAMPLE_MULTIPLIER handled in the loops
    long long sample_off;     // first SampleRec of this pair
    int valid, s32;
    int it_first;             // first tick with time_tick >= 0 (T if none)
    int n_live;               // live samples written by the sampler
    int n_irregular;          // live samples that need the exact per-tick path
    int int_lo, int_hi;       // intersection of the live regular samples' tick ranges
    int uni_lo, uni_hi;       // union of all live samples' tick ranges
    int n_groups;             // group records written by k_mc_sort (grouped accumulate path)
};

struct SampleRec {
    double t0;
    int rowoff;               // (i*Ry + j)*Rt
    int shift;                // k = stride*tick + shift ; INT_MIN: irregular -> exact per-tick path
    int lo, hi;               // valid ticks [lo, hi]
};
#define SHIFT_IRREGULAR (-2147483647 - 1)
#define OFF_IRREGULAR (-2147483647 - 1)

struct McParams {
    long long S;              // segments in this launch
    long long seg0;           // global index of the first segment (multi-pass)
    long long rng_stride;     // ntrk of detsim.py:324
    int P, T, Rx, Ry, Rt;
    int stride;               // TIME_SAMPLING / RESPONSE_SAMPLING if integral, else 0
    int split, split_len;     // split == 2: the accumulate kernel reads a PHASE-SPLIT copy of the table (every row stored as
                              // [even samples | odd samples], split_len = ceil(Rt / 2) = start of the odd half), which turns the
                              // stride-2 gather LUT[row + 2 tick + shift] into the unit-stride gather
                              // split[row + (shift & 1) * split_len + (shift >> 1) + tick] of the grouped path
    // sync-free operation (fused chain): the sample total stays on the device; every kernel after the scan
    // returns at once if it exceeds the capacity the caller provisioned (and *overflow is raised)
    const long long* total_dev; long long cap; int* overflow;
    unsigned int* lohi;         // per live sample: valid ticks lo | hi << 16 (T < 65536); with offs32 it is all the edge path needs, so the
                                // 24-byte SampleRec is written (and read) for IRREGULAR samples only
    unsigned long long* diag;   // optional device counters {group records, edge (sample, tick) pairs, irregular samples}
    unsigned long long* npairs; // device counter of (segment, pixel) pairs that hold a pixel id (S * P-bar of SURVEY 8d), or null
    unsigned long long* nfma; // device counter of (sample, tick) pairs that pass every test of detsim.py:299,333,341-344 (N_fma of SURVEY 8d), or null
    int2* ranges;             // fused chain: rows of `signals` are stored sparsely -- only the ticks [lo, hi] covered by the pair's
                              // samples are written, and (lo, hi) is recorded here (empty: lo > hi); everything else is zero by
                              // definition and is neither written nor read (lsb_chain_signals_dense fills it in on request)
};
#define MC_GUARD(p) do { if ((p).total_dev && *(p).total_dev > (p).cap) return; } while (0)

__device__ __forceinline__ double tick_time(double t_start, int it) { return t_start + (double)it * d_c.time_sampling; }

// overlapping_segment (detsim.py:220-256)
__device__ void overlapping_segment(double x, double y, const double* start, const double* end, double radius, bool c32,
                                    double* ns, double* ne) {
    double dxy0 = x - start[0], dxy1 = y - start[1];
    double v0 = R32(end[0] - start[0], c32), v1 = R32(end[1] - start[1], c32);
    double l = R32(sqrt(R32(R32(v0 * v0, c32) + R32(v1 * v1, c32), c32)), c32);
    v0 = R32(v0 / l, c32); v1 = R32(v1 / l, c32);
    double s = (dxy0 * v0 + dxy1 * v1) / l;
    double a = dxy0 - v0 * s * l, b = dxy1 - v1 * s * l;
    double r = sqrt(a * a + b * b);
    if (r > radius) { for (int k = 0; k < 3; k++) { ns[k] = start[k]; ne[k] = start[k]; } return; }
    double s_plus = s + sqrt(radius * radius - r * r) / l;
    double s_minus = s - sqrt(radius * radius - r * r) / l;
    if (s_plus > 1) s_plus = 1; else if (s_plus < 0) s_plus = 0;
    if (s_minus > 1) s_minus = 1; else if (s_minus < 0) s_minus = 0;
    for (int k = 0; k < 3; k++) {
        ns[k] = start[k] * (1 - s_minus) + end[k] * s_minus;
        ne[k] = start[k] * (1 - s_plus) + end[k] * s_plus;
    }
}

// ordered segment end points (detsim.py:289-294) and float32-typed direction (:302-305)
__device__ __forceinline__ void load_endpoints(const Layout& L, const char* t, double* start, double* end, bool& c32) {
    c32 = fld_f32(L, LSB_F_X_START) && fld_f32(L, LSB_F_Y_START) && fld_f32(L, LSB_F_Z_START) &&
          fld_f32(L, LSB_F_X_END) && fld_f32(L, LSB_F_Y_END) && fld_f32(L, LSB_F_Z_END);
    double zs = fld_get(L, t, LSB_F_Z_START), ze = fld_get(L, t, LSB_F_Z_END);
    if (zs < ze) {
        start[0] = fld_get(L, t, LSB_F_X_START); start[1] = fld_get(L, t, LSB_F_Y_START); start[2] = zs;
        end[0] = fld_get(L, t, LSB_F_X_END); end[1] = fld_get(L, t, LSB_F_Y_END); end[2] = ze;
    } else {
        end[0] = fld_get(L, t, LSB_F_X_START); end[1] = fld_get(L, t, LSB_F_Y_START); end[2] = zs;
        start[0] = fld_get(L, t, LSB_F_X_END); start[1] = fld_get(L, t, LSB_F_Y_END); start[2] = ze;
    }
}
__device__ __forceinline__ double seg_length(const double* seg, bool c32) {
    return R32(sqrt(R32(R32(R32(seg[0] * seg[0], c32) + R32(seg[1] * seg[1], c32), c32) + R32(seg[2] * seg[2], c32), c32)), c32);
}
// pixel centre (get_pixel_coordinates detsim.py:180-191 + :285-288); pID=-1 wraps like Python
__device__ __forceinline__ bool pixel_center(int pID, double& x_p, double& y_p) {
    long long px, py, pl;
    id2pixel(pID, px, py, pl);
    if (pl < 0) pl += d_c.n_tpc;                       // negative index wraps to the last TPC
    if (pl < 0 || pl >= d_c.n_tpc) return false;
    const double (*b)[2] = d_c.tpc_borders[pl];
    x_p = (double)px * d_c.pixel_pitch + b[0][0] + d_c.pixel_pitch / 2;
    y_p = (double)py * d_c.pixel_pitch + b[1][0] + d_c.pixel_pitch / 2;
    return true;
}

__device__ void mc_pair_geometry(const Layout& L, const char* t, int pID, int Rx, int Ry, int T, PairRec& g) {
    g.valid = 0; g.nstep = 0; g.n_live = 0; g.n_irregular = 0;
    if (!pixel_center(pID, g.x_p, g.y_p)) return;
    bool c32;
    double start[3], end[3];
    load_endpoints(L, t, start, end, c32);
    g.s32 = fld_f32(L, LSB_F_TRAN_DIFF) && fld_f32(L, LSB_F_LONG_DIFF);
    g.t_start = (double)__double2ll_rn((fld_get(L, t, LSB_F_T_START) - fld_get(L, t, LSB_F_T0_START) - d_c.time_padding) /
                                       d_c.time_sampling) * d_c.time_sampling;
    double seg[3];
    for (int k = 0; k < 3; k++) seg[k] = R32(end[k] - start[k], c32);
    double length = seg_length(seg, c32);
    for (int k = 0; k < 3; k++) g.dir[k] = R32(seg[k] / length, c32);
    g.sig_t = fld_get(L, t, LSB_F_TRAN_DIFF);
    g.sig_l = fld_get(L, t, LSB_F_LONG_DIFF);
    double impact = sqrt((double)((long long)Rx * Rx + (long long)Ry * Ry)) * d_c.response_bin_size;
    double ss[3], se[3];
    overlapping_segment(g.x_p, g.y_p, start, end, impact, c32, ss, se);
    double sub[3] = {se[0] - ss[0], se[1] - ss[1], se[2] - ss[2]};
    double sub_len = sqrt(sub[0] * sub[0] + sub[1] * sub[1] + sub[2] * sub[2]);
    if (sub_len == 0) return;
    long long ns = __double2ll_rn(sub_len / d_c.min_step_size);
    g.nstep = ns > 1 ? ns : 1;
    g.step = sub_len / (double)g.nstep;
    g.charge = fld_get(L, t, LSB_F_N_ELECTRONS) * (sub_len / length) / (double)(g.nstep * d_c.mc_sample_multiplier);
    for (int k = 0; k < 3; k++) g.sub_start[k] = ss[k];
    long long plane = (long long)fld_get(L, t, LSB_F_PIXEL_PLANE);
    if (plane < 0 || plane >= d_c.n_tpc) { g.nstep = 0; return; }
    g.z_anode = d_c.tpc_borders[plane][2][0];
    // first tick with time_tick >= 0 (monotone in it)
    int f = 0;
    if (tick_time(g.t_start, 0) < 0) {
        double e = ceil(-g.t_start / d_c.time_sampling);
        f = e > (double)T ? T : (int)e;
        while (f > 0 && tick_time(g.t_start, f - 1) >= 0) f--;
        while (f < T && tick_time(g.t_start, f) < 0) f++;
    }
    g.it_first = f;
    g.valid = 1;
}

// ---------------------------------------------------------------------------------------
__global__ void k_mc_pairs(Layout L, const char* __restrict__ tracks, const int32_t* __restrict__ pixels, McParams p,
                           PairRec* __restrict__ pairs, uint32_t* __restrict__ nsamp) {
    long long pr = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (pr >= p.S * p.P) return;
    long long itrk = pr / p.P;
    PairRec g;
    mc_pair_geometry(L, tracks + (p.seg0 + itrk) * L.itemsize, pixels[(p.seg0 + itrk) * p.P + (pr % p.P)], p.Rx, p.Ry, p.T, g);
    long long n = g.valid ? g.nstep * d_c.mc_sample_multiplier : 0;
    if (n > 0xffffffffLL) n = 0xffffffffLL;
    nsamp[pr] = (uint32_t)n;
    pairs[pr] = g;
    if (p.npairs) {
        const unsigned m = __ballot_sync(__activemask(), pixels[(p.seg0 + itrk) * p.P + (pr % p.P)] >= 0);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(p.npairs, (unsigned long long)__popc(m));
    }
}

__global__ void k_mc_set_offsets(McParams p, PairRec* __restrict__ pairs, const long long* __restrict__ offs, long long n) {
    long long pr = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p.total_dev && *p.total_dev > p.cap) { if (pr == 0 && p.overflow) *p.overflow = 1; return; }
    if (pr < n) pairs[pr].sample_off = offs[pr];
}

struct SampleGeom { bool live; double t0; int rowoff; };
// round(a / b - sub) (round half to even, as Python's round on the float64 quotient) without the FP64 division when the answer
// cannot depend on its last bits: a * (1 / b) is within a few ulp of a / b, so unless the value lies within 1e-9 of a tie the
// rounded integer is the same; otherwise the exact expression is evaluated.  (A float64 division is ~30 instructions on the
// FP64 pipe; the sampler used nine per sample.)
__device__ __forceinline__ long long round_quotient(double a, double b, double inv_b, double sub) {
    const double v = a * inv_b - sub;
    const double r = rint(v);
    if (fabs(fabs(v - r) - 0.5) < 1e-9 || !(fabs(v) < 4.0e9)) return __double2ll_rn(a / b - sub);
    return (long long)r;
}
// one MC sample given its three normals (detsim.py:326-346 without the tick dependence)
__device__ __forceinline__ SampleGeom mc_sample(const PairRec& g, long long istep, float nz, float nx, float ny,
                                                int Rx, int Ry, int Rt, double inv_bin) {
    SampleGeom s;
    double f = (double)istep + 0.5;
    double x = g.sub_start[0] + g.step * f * g.dir[0];
    double y = g.sub_start[1] + g.step * f * g.dir[1];
    double z = g.sub_start[2] + g.step * f * g.dir[2];
    z += R32((double)nz * g.sig_l, g.s32);
    s.t0 = fabs(z - g.z_anode) / d_c.v_drift - d_c.time_window;
    x += R32((double)nx * g.sig_t, g.s32);
    y += R32((double)ny * g.sig_t, g.s32);
    double x_dist = fabs(g.x_p - x), y_dist = fabs(g.y_p - y);
    s.live = true;
    if (x_dist > d_c.response_bin_size * Rx) s.live = false;
    if (y_dist > d_c.response_bin_size * Ry) s.live = false;
    long long i = round_quotient(x_dist, d_c.response_bin_size, inv_bin, 0.5);
    long long j = round_quotient(y_dist, d_c.response_bin_size, inv_bin, 0.5);
    if (!(0 <= i && i < Rx && 0 <= j && j < Ry)) s.live = false;
    s.rowoff = s.live ? (int)((i * Ry + j) * (long long)Rt) : -1;
    return s;
}
__device__ __forceinline__ long long resp_k(double tick, double t0) { return __double2ll_rn((tick - t0) / d_c.response_sampling); }

// The sampler is split so that the only sequential part -- stepping each pair's xoroshiro128+ stream -- is a
// short integer chain, and the expensive part (3 Box-Muller normals, FP64 geometry with exact divisions, the
// tick-range search) runs one SAMPLE per lane with balanced warps:
//   k_mc_uniforms  thread per (segment,pixel): 6 float32 uniforms per sample -> uu[sample][6]; state written back
//   k_mc_sampler   warp per (segment,pixel): lane = sample; live samples are compacted in order (ballot prefix)
struct SampleU { float u[6]; };

// k_mc_uniforms walks a pair's stream sequentially, so a warp is as slow as its longest pair.  Sample counts range from 0
// (no pixel / no overlap) to several hundred: the pairs are therefore handed out in descending order of their count
// (counting sort into MC_NBUCKET buckets of 16 samples; the order inside a bucket is arbitrary and does not matter --
// every pair writes its own slots), which makes the lanes of a warp finish together.
#define MC_NBUCKET 64
__device__ __forceinline__ int mc_bucket_of(uint32_t n) { const uint32_t b = n >> 4; return MC_NBUCKET - 1 - (int)(b < MC_NBUCKET - 1 ? b : MC_NBUCKET - 1); }
__global__ void k_mc_bucket_count(const uint32_t* __restrict__ nsamp, long long npair, int* __restrict__ bucket) {
    __shared__ int s_cnt[MC_NBUCKET];
    if (threadIdx.x < MC_NBUCKET) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const long long pr = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (pr < npair) atomicAdd(&s_cnt[mc_bucket_of(nsamp[pr])], 1);
    __syncthreads();
    if (threadIdx.x < MC_NBUCKET && s_cnt[threadIdx.x]) atomicAdd(&bucket[threadIdx.x], s_cnt[threadIdx.x]);
}
__global__ void k_mc_bucket_scan(int* __restrict__ bucket) {          // counts [0, NB) -> cursors [NB, 2NB)
    if (threadIdx.x == 0) { int run = 0; for (int b = 0; b < MC_NBUCKET; b++) { bucket[MC_NBUCKET + b] = run; run += bucket[b]; } }
}
__global__ void k_mc_bucket_scatter(const uint32_t* __restrict__ nsamp, long long npair, int* __restrict__ bucket, int* __restrict__ perm) {
    __shared__ int s_cnt[MC_NBUCKET], s_base[MC_NBUCKET];
    if (threadIdx.x < MC_NBUCKET) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const long long pr = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int b = 0, local = 0;
    if (pr < npair) { b = mc_bucket_of(nsamp[pr]); local = atomicAdd(&s_cnt[b], 1); }
    __syncthreads();
    if (threadIdx.x < MC_NBUCKET && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&bucket[MC_NBUCKET + threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
    if (pr < npair) perm[s_base[b] + local] = (int)pr;
}

#define UNI_TPB 64
#define UNI_CHUNK 16         // samples generated per thread between two cooperative write-outs (16*6 floats = 3 warp rows)
__global__ void __launch_bounds__(UNI_TPB) k_mc_uniforms(McParams p, const PairRec* __restrict__ pairs, const int* __restrict__ perm,
                                                        SampleU* __restrict__ uu, unsigned long long* __restrict__ rng_states) {
    // thread = (segment,pixel).  Each thread's samples are contiguous in `uu`, so direct stores would scatter
    // 24-byte pieces over 32 different places per warp; the values go through shared memory and leave as
    // contiguous 192-byte runs per pair instead.
    MC_GUARD(p);
    __shared__ float s_u[UNI_TPB][UNI_CHUNK * 6 + 1];
    __shared__ long long s_base[UNI_TPB];
    __shared__ long long s_n[UNI_TPB];
    __shared__ long long s_max;
    const int tid = threadIdx.x;
    const long long slot = blockIdx.x * (long long)UNI_TPB + tid;
    const long long pr = slot < p.S * p.P ? perm[slot] : 0;
    const bool valid = slot < p.S * p.P && pairs[pr].valid;
    const long long n = valid ? pairs[pr].nstep * d_c.mc_sample_multiplier : 0;
    Rng rng; rng.s0 = 0; rng.s1 = 0;
    unsigned long long* sp = nullptr;
    if (valid) {
        long long itrk = p.seg0 + pr / p.P;
        int ipix = (int)(pr % p.P);
        sp = rng_states + 2 * (itrk + p.rng_stride * ipix);
        rng.s0 = sp[0]; rng.s1 = sp[1];
    }
    s_base[tid] = valid ? pairs[pr].sample_off : 0;
    s_n[tid] = n;
    if (tid == 0) s_max = 0;
    __syncthreads();
    if (n > 0) atomicMax((unsigned long long*)&s_max, (unsigned long long)n);
    __syncthreads();
    const long long nmax = s_max;
    float* dst = reinterpret_cast<float*>(uu);
    for (long long i0 = 0; i0 < nmax; i0 += UNI_CHUNK) {
#pragma unroll
        for (int j = 0; j < UNI_CHUNK; j++) {
            if (i0 + j < n) {
#pragma unroll
                for (int k = 0; k < 6; k++) s_u[tid][j * 6 + k] = rng_uniform_f32(rng);
            }
        }
        __syncthreads();
        // warp w writes the runs of pairs w, w + nwarps, ...: 96 contiguous floats each = 3 coalesced rows
        for (int t = tid >> 5; t < UNI_TPB; t += UNI_TPB / 32) {
            const long long left = (s_n[t] - i0) * 6;              // floats still owed to this pair
            if (left <= 0) continue;
            float* row = dst + (s_base[t] + i0) * 6;
#pragma unroll
            for (int j = 0; j < UNI_CHUNK * 6 / 32; j++) {
                const int o = (tid & 31) + 32 * j;
                if (o < left) row[o] = s_u[t][o];
            }
        }
        __syncthreads();
    }
    if (valid) { sp[0] = rng.s0; sp[1] = rng.s1; }
}
__device__ __forceinline__ float normal_from_uniforms(float u1, float u2) {     // == rng_normal_f32
    float a = sqrtf(__fmul_rn(-2.0f, logf(u1)));
    float b = cosf(__fmul_rn(6.283185307179586f, u2));
    return __fmul_rn(a, b);
}

#ifndef SMP_WARPS
#define SMP_WARPS 4
#endif
#ifndef SMP_MINB
#define SMP_MINB 8
#endif
__global__ void __launch_bounds__(32 * SMP_WARPS, SMP_MINB) k_mc_sampler(McParams p, PairRec* __restrict__ pairs,
                                                               const SampleU* __restrict__ uu, SampleRec* __restrict__ samples,
                                                               int* __restrict__ offs32) {
    MC_GUARD(p);
    const int lane = threadIdx.x & 31;
    const long long pr = blockIdx.x * (long long)SMP_WARPS + (threadIdx.x >> 5);
    if (pr >= p.S * p.P) return;
    const PairRec g = pairs[pr];
    if (!g.valid) return;
    SampleRec* out = samples + g.sample_off;
    int* out32 = offs32 + g.sample_off;
    const SampleU* in = uu + g.sample_off;
    int n_live = 0, n_irr = 0, int_lo = 0, int_hi = p.T - 1, uni_lo = p.T, uni_hi = -1;
    unsigned long long n_fma = 0;                                    // sum of the live samples' tick counts
    const int M = d_c.mc_sample_multiplier;
    const double W = d_c.time_window, TS = d_c.time_sampling;
    const long long n = g.nstep * M;
    const double inv_TS = 1.0 / TS, inv_bin = 1.0 / d_c.response_bin_size, inv_rs = 1.0 / d_c.response_sampling;
    for (long long i0 = 0; i0 < n; i0 += 32) {
        const long long i = i0 + lane;
        bool keep = false;
        SampleRec r; r.t0 = 0; r.rowoff = 0; r.shift = SHIFT_IRREGULAR; r.lo = 0; r.hi = -1;
        if (i < n) {
            const SampleU v = in[i];
            const float nz = normal_from_uniforms(v.u[0], v.u[1]);
            const float nx = normal_from_uniforms(v.u[2], v.u[3]);
            const float ny = normal_from_uniforms(v.u[4], v.u[5]);
            const long long istep = i / M;
            SampleGeom s = mc_sample(g, istep, nz, nx, ny, p.Rx, p.Ry, p.Rt, inv_bin);
            if (s.live) {
                double t0 = s.t0, t0W = t0 + W;
                // lower bound: first tick with tick>=0, tick>t0 (tick > t0 already implies k = round((tick-t0)/rs) >= 0)
                double tl = t0 > 0 ? t0 : 0;
                double e = floor((tl - g.t_start) * inv_TS);          // an estimate: the two loops below settle the exact tick
                int lo = e < 0 ? 0 : (e > (double)p.T ? p.T : (int)e);
#define LOW_OK(it) (tick_time(g.t_start, (it)) >= 0 && t0 < tick_time(g.t_start, (it)))
                while (lo > 0 && LOW_OK(lo - 1)) lo--;
                while (lo < p.T && !LOW_OK(lo)) lo++;
#undef LOW_OK
                // upper bound: last tick with tick < t0+W and k < Rt.  The window test needs no division; k < Rt only
                // bites when the table is shorter than the window, and k is monotone in the tick index.
                e = ceil((t0W - g.t_start) * inv_TS);                 // (estimate, settled below)
                int hi = e < -1 ? -1 : (e > (double)(p.T - 1) ? p.T - 1 : (int)e);
#define HIGH_T(it) (tick_time(g.t_start, (it)) < t0W)
#define K_OK(it) (round_quotient(tick_time(g.t_start, (it)) - t0, d_c.response_sampling, inv_rs, 0.0) < p.Rt)
                while (hi < p.T - 1 && HIGH_T(hi + 1)) hi++;
                while (hi >= 0 && !HIGH_T(hi)) hi--;
                if (hi >= 0 && !K_OK(hi)) {
                    double e2 = floor((t0 + ((double)p.Rt - 0.5) * d_c.response_sampling - g.t_start) / TS) + 1.0;
                    int h2 = e2 < -1 ? -1 : (e2 > (double)hi ? hi : (int)e2);
                    while (h2 < hi && K_OK(h2 + 1)) h2++;          // safety: the estimate must not undershoot
                    hi = h2;
                    while (hi >= 0 && !K_OK(hi)) hi--;
                }
#undef HIGH_T
#undef K_OK
                if (lo <= hi) {
                    int shift = SHIFT_IRREGULAR;
                    if (p.stride > 0) {
                        const long long klo = round_quotient(tick_time(g.t_start, lo) - t0, d_c.response_sampling, inv_rs, 0.0);
                        const long long khi = round_quotient(tick_time(g.t_start, hi) - t0, d_c.response_sampling, inv_rs, 0.0);
                        long long sh = klo - (long long)p.stride * lo;
                        long long last = khi;                                   // table index (within the row) read at tick hi
                        if (khi == (long long)p.stride * hi + sh) {
                            if (p.split == 2) {                                  // phase-split table: unit stride from here on
                                sh = (sh & 1) * (long long)p.split_len + (sh >> 1);
                                last = sh + hi;
                            }
                            shift = (int)sh;
                        }
                        // the grouped accumulate path reads the table in aligned 4-word blocks: a sample that reaches
                        // the last complete block of the table (or the partial one after it) takes the exact path
                        const long long L4 = (((long long)p.Rx * p.Ry * p.Rt) & ~(long long)(ACC_GW - 1)) - 2 * ACC_GW;   // 12-word windows (8-tick lanes)
                        if ((long long)s.rowoff + last >= L4) shift = SHIFT_IRREGULAR;
                    }
                    r.t0 = t0; r.rowoff = s.rowoff; r.shift = shift; r.lo = lo; r.hi = hi;
                    keep = true;
                }
            }
        }
        // ordered compaction of the live samples of this chunk
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int pos = n_live + __popc(m & ((1u << lane) - 1));
            if (r.shift == SHIFT_IRREGULAR) out[pos] = r;
            out32[pos] = (r.shift == SHIFT_IRREGULAR) ? OFF_IRREGULAR : r.rowoff + r.shift;
            p.lohi[g.sample_off + pos] = (unsigned int)r.lo | ((unsigned int)r.hi << 16);
            if (r.shift != SHIFT_IRREGULAR) { int_lo = r.lo > int_lo ? r.lo : int_lo; int_hi = r.hi < int_hi ? r.hi : int_hi; }
            else n_irr++;
            uni_lo = r.lo < uni_lo ? r.lo : uni_lo; uni_hi = r.hi > uni_hi ? r.hi : uni_hi;
            n_fma += (unsigned long long)(r.hi - r.lo + 1);
        }
        n_live += __popc(m);
    }
    // per-pair reductions over the lanes
    for (int o = 16; o > 0; o >>= 1) {
        int v;
        v = __shfl_xor_sync(0xffffffffu, int_lo, o); int_lo = v > int_lo ? v : int_lo;
        v = __shfl_xor_sync(0xffffffffu, int_hi, o); int_hi = v < int_hi ? v : int_hi;
        v = __shfl_xor_sync(0xffffffffu, uni_lo, o); uni_lo = v < uni_lo ? v : uni_lo;
        v = __shfl_xor_sync(0xffffffffu, uni_hi, o); uni_hi = v > uni_hi ? v : uni_hi;
        n_irr += __shfl_xor_sync(0xffffffffu, n_irr, o);
        n_fma += __shfl_xor_sync(0xffffffffu, n_fma, o);
    }
    if (lane == 0) {
        if (p.nfma && n_fma) atomicAdd(p.nfma, n_fma);
        PairRec* gp = pairs + pr;
        gp->n_live = n_live; gp->n_irregular = n_irr; gp->int_lo = int_lo; gp->int_hi = int_hi; gp->uni_lo = uni_lo; gp->uni_hi = uni_hi;
    }
}


// ---------------------------------------------------------------------------------------
// k_mc_sort: warp per pair.  The pair's live offsets are sorted (bitonic network on 32*NR keys held NR per
// lane, "blocked": position = lane*NR + r, so compare distances below NR stay in registers and only the rest
// goes through SHFL; chunks of 512 keys are sorted independently) and folded into GROUP records for
// k_mc_accumulate: one record per run of sorted offsets that fall into the same aligned 4-word block of the
// table, off = 4q + d, carrying the number of samples at each d.  Offsets are taken relative to the pair's
// first interior tick (key = off + int_lo), so key + (tick - int_lo) is the table index and never negative.
struct __align__(16) GroupRec { int q; __half2 c01, c23; int pad; };      // counts <= 512: exact in binary16
// 8-word variant (ACC_GW == 8): off = 8q + d, d = 0..7; 24 bytes (the uniforms buffer the records reuse has 24 bytes per sample)
struct __align__(8) GroupRec8 { int q; int pad; __half2 c[4]; };
#if ACC_GW == 8
typedef GroupRec8 GroupRecT;
#define ACC_GSH 3
#else
typedef GroupRec GroupRecT;
#define ACC_GSH 2
#endif
__device__ __forceinline__ void group_store(GroupRecT* dst, int q, const int (&c)[ACC_GW]) {
    GroupRecT g;
    g.q = q; g.pad = 0;
#if ACC_GW == 8
#pragma unroll
    for (int k = 0; k < 4; k++) g.c[k] = __floats2half2_rn((float)c[2 * k], (float)c[2 * k + 1]);
#else
    g.c01 = __floats2half2_rn((float)c[0], (float)c[1]); g.c23 = __floats2half2_rn((float)c[2], (float)c[3]);
#endif
    *dst = g;
}
#ifndef SORT_WARPS
#define SORT_WARPS 4
#endif
#define SORT_MAXKEYS 512
template <int NR>
__device__ __forceinline__ void warp_bitonic_to_smem(int (&v)[NR], int* s_buf, int lane) {
#pragma unroll
    for (int k = 2; k <= 32 * NR; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j < NR) {
#pragma unroll
                for (int r = 0; r < NR; r++) {
                    if ((r & j) == 0) {
                        const bool up = (k < NR) ? ((r & k) == 0) : (((lane * NR) & k) == 0);
                        const int a = v[r], b = v[r | j];
                        const int lo = a < b ? a : b, hi = a < b ? b : a;
                        v[r] = up ? lo : hi; v[r | j] = up ? hi : lo;
                    }
                }
            } else {
                const int lj = j / NR;
                const bool take_min = ((lane & lj) == 0) == (((lane * NR) & k) == 0);
#pragma unroll
                for (int r = 0; r < NR; r++) {
                    const int o = __shfl_xor_sync(0xffffffffu, v[r], lj);
                    v[r] = take_min ? (v[r] < o ? v[r] : o) : (v[r] > o ? v[r] : o);
                }
            }
        }
    }
    // blocked -> position order in shared memory (one pad word per 32)
#pragma unroll
    for (int r = 0; r < NR; r++) { const int pos = lane * NR + r; s_buf[pos + (pos >> 5)] = v[r]; }
    __syncwarp();
}
template <int NR>
__device__ __forceinline__ int warp_sort_group(const int* __restrict__ keys, int n, int key_add, int* s_buf, GroupRecT* __restrict__ out,
                                               int lane) {
    int v[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) {
        const int g = r * 32 + lane;
        int k = g < n ? keys[g] : 2147483647;
        if (k != OFF_IRREGULAR && k != 2147483647) k += key_add;
        v[r] = k;
    }
    warp_bitonic_to_smem<NR>(v, s_buf, lane);
    // group heads: position whose aligned ACC_GW-word block differs from its predecessor's.  Runs are measured with ballots
    // (32 positions at a time); the last run of a row stays "pending" (warp-uniform registers) because it may continue in
    // the next row.
    int ng = 0;
    int pq = 0, pc[ACC_GW];
#pragma unroll
    for (int k = 0; k < ACC_GW; k++) pc[k] = 0;
    bool pending = false;
#pragma unroll
    for (int r = 0; r < NR; r++) {
        const int pos = r * 32 + lane;
        const int key = s_buf[pos + (pos >> 5)];                       // padding keys (INT_MAX) sort last
        const int prev = pos > 0 ? s_buf[pos - 1 + ((pos - 1) >> 5)] : OFF_IRREGULAR;
        const bool valid = pos < n && key != OFF_IRREGULAR;
        const bool head = valid && (prev == OFF_IRREGULAR || (prev >> ACC_GSH) != (key >> ACC_GSH));
        const unsigned m = __ballot_sync(0xffffffffu, head);
        const int d = key & (ACC_GW - 1);
        unsigned md[ACC_GW];
#pragma unroll
        for (int k = 0; k < ACC_GW; k++) md[k] = __ballot_sync(0xffffffffu, valid && d == k);
        const unsigned lead = m ? ((1u << (__ffs(m) - 1)) - 1u) : 0xffffffffu;     // positions continuing the pending run
        if (pending) {
#pragma unroll
            for (int k = 0; k < ACC_GW; k++) pc[k] += __popc(md[k] & lead);
        }
        if (m) {
            if (pending) {
                if (lane == 0) group_store(out + ng, pq, pc);
                ng++;
            }
            const unsigned below = (1u << lane) - 1u;
            const unsigned higher = lane == 31 ? 0u : (m & ~((2u << lane) - 1u));
            const unsigned upto = higher ? ((1u << (__ffs(higher) - 1)) - 1u) : 0xffffffffu;
            const unsigned run = upto & ~below;
            int c[ACC_GW];
#pragma unroll
            for (int k = 0; k < ACC_GW; k++) c[k] = __popc(md[k] & run);
            if (head && higher) group_store(out + ng + __popc(m & below), key >> ACC_GSH, c);
            const int src = 31 - __clz(m);                                         // last head of the row: the new pending run
            pq = __shfl_sync(0xffffffffu, key >> ACC_GSH, src);
#pragma unroll
            for (int k = 0; k < ACC_GW; k++) pc[k] = __shfl_sync(0xffffffffu, c[k], src);
            pending = true;
            ng += __popc(m) - 1;
        }
    }
    if (pending) {
        if (lane == 0) group_store(out + ng, pq, pc);
        ng++;
    }
    __syncwarp();
    return ng;
}


// ---- run records of the phase-aligned accumulate path (LSB_ACC_ALIGNED, FAST == 4) ---------------------------------------------------
// The keys are ordered by (key & 3, key >> 2): all offsets of one alignment class d = off mod 4 are neighbours, ascending in the
// float4 index q.  One 4-byte record q << 10 | count per DISTINCT offset (count <= 512 = one chunk); a chunk is
//   [header: c1 | c2 << 10 | c3 << 20, n_rec] [n_rec records], class d = records [c_d, c_(d+1)), c_0 = 0, c_4 = n_rec.
// The words of a pair start at the first 16-byte boundary of its slice of the uniforms buffer (24 bytes per sample):
// 2 (alignment) + 2 (header) + n_rec <= 6 n words.
#define RUN_CLASS_SHIFT 28
#define RUN_COUNT_BITS 10
#define RUN_MAX_WORDS (1LL << (32 - RUN_COUNT_BITS + 2))           // offsets + ticks representable in a record
__device__ __forceinline__ const int* run_stream(const void* groups, long long sample_off) {
    return reinterpret_cast<const int*>((reinterpret_cast<uintptr_t>(groups) + (uintptr_t)sample_off * 24u + 15u) & ~(uintptr_t)15u);
}
template <int NR>
__device__ __forceinline__ int warp_sort_runs(const int* __restrict__ keys, int n, int key_add, int* s_buf, int* __restrict__ out, int lane) {
    int v[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) {
        const int g = r * 32 + lane;
        int k = g < n ? keys[g] : 2147483647;
        if (k != OFF_IRREGULAR && k != 2147483647) { k += key_add; k = ((k & 3) << RUN_CLASS_SHIFT) | (k >> 2); }
        v[r] = k;
    }
    warp_bitonic_to_smem<NR>(v, s_buf, lane);
    int key[NR];
    unsigned headbits = 0;
#pragma unroll
    for (int r = 0; r < NR; r++) {
        const int pos = r * 32 + lane;
        key[r] = s_buf[pos + (pos >> 5)];
        const int prev = pos > 0 ? s_buf[pos - 1 + ((pos - 1) >> 5)] : OFF_IRREGULAR;
        if (pos < n && key[r] != OFF_IRREGULAR && key[r] != prev) headbits |= 1u << r;
    }
    __syncwarp();                                                      // every key is in registers: s_buf becomes {head position, key}
    int nrec = 0, c1 = 0, c2 = 0, c3 = 0;
    const unsigned below = (1u << lane) - 1u;
    int* s_pos = s_buf;                                                // [0, 256]: packed two 16-bit positions would do; positions <= 512
    // positions and keys of the heads: positions in the low half-words of s_buf[0..], keys go straight to the output
#pragma unroll
    for (int r = 0; r < NR; r++) {
        const bool head = (headbits >> r) & 1u;
        const unsigned m = __ballot_sync(0xffffffffu, head);
        const int cls = key[r] >> RUN_CLASS_SHIFT;
        c1 += __popc(__ballot_sync(0xffffffffu, head && cls < 1));
        c2 += __popc(__ballot_sync(0xffffffffu, head && cls < 2));
        c3 += __popc(__ballot_sync(0xffffffffu, head && cls < 3));
        if (head) {
            const int idx = nrec + __popc(m & below);
            s_pos[idx] = r * 32 + lane;
            out[2 + idx] = (key[r] & ((1 << RUN_CLASS_SHIFT) - 1)) << RUN_COUNT_BITS;
        }
        nrec += __popc(m);
    }
    if (lane == 0) { s_pos[nrec] = n; out[0] = c1 | (c2 << 10) | (c3 << 20); out[1] = nrec; }   // irregular keys sort first: runs end at n
    __syncwarp();
    for (int i = lane; i < nrec; i += 32) out[2 + i] |= s_pos[i + 1] - s_pos[i];
    __syncwarp();
    return 2 + nrec + (nrec & 1);                                      // the next chunk header stays 8-byte aligned
}

#ifndef SORT_MINB
#define SORT_MINB 12
#endif
#ifndef SORT_RUNS_MINB
#define SORT_RUNS_MINB 10
#endif
template <bool RUNS>
__global__ void __launch_bounds__(32 * SORT_WARPS, RUNS ? SORT_RUNS_MINB : SORT_MINB) k_mc_sort(McParams p, PairRec* __restrict__ pairs, const int* __restrict__ offs32,
                                                             GroupRecT* __restrict__ groups) {
    MC_GUARD(p);
    __shared__ int s_buf[SORT_WARPS][SORT_MAXKEYS + SORT_MAXKEYS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long pr = blockIdx.x * (long long)SORT_WARPS + warp;
    if (pr >= p.S * p.P) return;
    PairRec* gp = pairs + pr;
    if (!gp->valid) return;
    const int n_live = gp->n_live, n_reg = n_live - gp->n_irregular;
    const int key_add = gp->int_lo;
    int ng = 0;
    if (n_reg > 0 && gp->int_lo <= gp->int_hi) {
        const int* keys = offs32 + gp->sample_off;
        if constexpr (RUNS) {
            int* out = const_cast<int*>(run_stream(groups, gp->sample_off));
            for (int c0 = 0; c0 < n_live; c0 += SORT_MAXKEYS) {
                const int n = n_live - c0 < SORT_MAXKEYS ? n_live - c0 : SORT_MAXKEYS;
                if (n <= 32) ng += warp_sort_runs<1>(keys + c0, n, key_add, s_buf[warp], out + ng, lane);
                else if (n <= 64) ng += warp_sort_runs<2>(keys + c0, n, key_add, s_buf[warp], out + ng, lane);
                else if (n <= 128) ng += warp_sort_runs<4>(keys + c0, n, key_add, s_buf[warp], out + ng, lane);
                else if (n <= 256) ng += warp_sort_runs<8>(keys + c0, n, key_add, s_buf[warp], out + ng, lane);
                else ng += warp_sort_runs<16>(keys + c0, n, key_add, s_buf[warp], out + ng, lane);
            }
        } else {
        GroupRecT* out = groups + gp->sample_off;         // <= one record per sample
        for (int c0 = 0; c0 < n_live; c0 += SORT_MAXKEYS) {
            const int n = n_live - c0 < SORT_MAXKEYS ? n_live - c0 : SORT_MAXKEYS;
            if (n <= 32) ng += warp_sort_group<1>(keys + c0, n, key_add, s_buf[warp], out + ng, lane);
            else if (n <= 64) ng += warp_sort_group<2>(keys + c0, n, key_add, s_buf[warp], out + ng, lane);
            else if (n <= 128) ng += warp_sort_group<4>(keys + c0, n, key_add, s_buf[warp], out + ng, lane);
            else if (n <= 256) ng += warp_sort_group<8>(keys + c0, n, key_add, s_buf[warp], out + ng, lane);
            else ng += warp_sort_group<16>(keys + c0, n, key_add, s_buf[warp], out + ng, lane);
        }
        }
    }
    if (lane == 0) {
        gp->n_groups = ng;
        if (p.diag) {
            atomicAdd(p.diag, (unsigned long long)ng);
            const int interior = gp->int_lo <= gp->int_hi && n_reg > 0 ? gp->int_hi - gp->int_lo + 1 : 0;
            const int uni = gp->uni_hi >= gp->uni_lo ? gp->uni_hi - gp->uni_lo + 1 : 0;
            atomicAdd(p.diag + 1, (unsigned long long)(uni - interior) * (unsigned long long)n_live);
            atomicAdd(p.diag + 2, (unsigned long long)gp->n_irregular);
        }
    }
}

// ---------------------------------------------------------------------------------------
#ifndef ACC_TPB
#define ACC_TPB 128
#endif
#define ACC_RMAX 8
#ifndef ACC_CHUNK
#define ACC_CHUNK 256      // samples staged per smem chunk
#endif

// Samples are summed in float32 in groups of ACC_GROUP and the group sums are added in float64: samples
// that hit the same LUT row with the same tick shift contribute identical values, so a plain float32
// running sum would round in the same direction every time (error growing linearly with the count).
// The loop is written for memory-level parallelism: the offsets of 4 samples are fetched with one uniform
// LDS.128, their 4 x NFULL independent LDGs are issued back to back, and only then added.
#define ACC_GROUP 8
template <typename TL, int STRIDE, int NFULL, bool CHECK>
__device__ __forceinline__ void acc_interior(const TL* __restrict__ lut, const int* s_off, int ns, int tick0, bool rem_ok,
                                             double (&dacc)[ACC_RMAX]) {
    // signal[tick0 + TPB*r] += LUT[off + STRIDE*(tick0 + TPB*r)], tick0 = base_tick + tid
    constexpr int NL = NFULL < ACC_RMAX ? NFULL + 1 : NFULL;      // loads per sample incl. the partial row
    const int t0 = STRIDE * tick0;
    int s0 = 0;
    for (; s0 + ACC_GROUP <= ns; s0 += ACC_GROUP) {
        float acc[ACC_RMAX];
#pragma unroll
        for (int r = 0; r < ACC_RMAX; r++) acc[r] = 0.f;
#pragma unroll
        for (int h = 0; h < ACC_GROUP; h += 4) {
            const int4 o = *reinterpret_cast<const int4*>(s_off + s0 + h);
            const int off[4] = {o.x, o.y, o.z, o.w};
            float v[4][NL];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const bool ok = !CHECK || off[u] != OFF_IRREGULAR;
                const TL* p = lut + (ok ? off[u] + t0 : 0);
#pragma unroll
                for (int r = 0; r < NL; r++) {
                    const bool ld = ok && (r < NFULL || rem_ok);
                    v[u][r] = ld ? (float)__ldg(p + r * ACC_TPB * STRIDE) : 0.f;
                }
            }
#pragma unroll
            for (int r = 0; r < NL; r++) acc[r] += (v[0][r] + v[1][r]) + (v[2][r] + v[3][r]);
        }
#pragma unroll
        for (int r = 0; r < NL; r++) dacc[r] += (double)acc[r];
    }
    if (s0 < ns) {
        float acc[ACC_RMAX];
#pragma unroll
        for (int r = 0; r < ACC_RMAX; r++) acc[r] = 0.f;
        for (int s = s0; s < ns; s++) {
            const int off = s_off[s];
            if (CHECK && off == OFF_IRREGULAR) continue;
            const TL* p = lut + off + t0;
#pragma unroll
            for (int r = 0; r < NL; r++)
                if (r < NFULL || rem_ok) acc[r] += (float)__ldg(p + r * ACC_TPB * STRIDE);
        }
#pragma unroll
        for (int r = 0; r < NL; r++) dacc[r] += (double)acc[r];
    }
}


// ---- grouped interior (float table, unit sampling ratio, 16-byte aligned table) --------------------
// A warp owns up to ACC_FB blocks of 128 ticks (lane: 4 consecutive ticks of each) and walks the pair's group
// records: the 8-word window [4q, 4q+8) of the table, shifted by the lane's ticks, is fetched with two aligned
// LDG.128 -- one if the previous group was q-1, the upper half is kept -- and every sample of the group is a
// count-weighted FFMA on it.  No shared memory, no CTA barrier.
#ifndef ACC_FB
#define ACC_FB 1                         // 128-tick blocks a warp register-blocks
#endif
#ifndef ACC_MINB
#define ACC_MINB 10
#endif
#define ACC_FLUSH 2                      // groups between two float32 -> float64 folds (<= 8 terms, like the generic path)
template <int R>
__device__ __forceinline__ void acc_apply(const int4& rec, const float4 (&lo)[R], const float4 (&hi)[R], float (&acc)[R][4]) {
    const float2 c01 = __half22float2(*reinterpret_cast<const __half2*>(&rec.y));
    const float2 c23 = __half22float2(*reinterpret_cast<const __half2*>(&rec.z));
    if (c01.x != 0.f) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            acc[r][0] = __fmaf_rn(c01.x, lo[r].x, acc[r][0]); acc[r][1] = __fmaf_rn(c01.x, lo[r].y, acc[r][1]);
            acc[r][2] = __fmaf_rn(c01.x, lo[r].z, acc[r][2]); acc[r][3] = __fmaf_rn(c01.x, lo[r].w, acc[r][3]);
        }
    }
    if (c01.y != 0.f) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            acc[r][0] = __fmaf_rn(c01.y, lo[r].y, acc[r][0]); acc[r][1] = __fmaf_rn(c01.y, lo[r].z, acc[r][1]);
            acc[r][2] = __fmaf_rn(c01.y, lo[r].w, acc[r][2]); acc[r][3] = __fmaf_rn(c01.y, hi[r].x, acc[r][3]);
        }
    }
    if (c23.x != 0.f) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            acc[r][0] = __fmaf_rn(c23.x, lo[r].z, acc[r][0]); acc[r][1] = __fmaf_rn(c23.x, lo[r].w, acc[r][1]);
            acc[r][2] = __fmaf_rn(c23.x, hi[r].x, acc[r][2]); acc[r][3] = __fmaf_rn(c23.x, hi[r].y, acc[r][3]);
        }
    }
    if (c23.y != 0.f) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            acc[r][0] = __fmaf_rn(c23.y, lo[r].w, acc[r][0]); acc[r][1] = __fmaf_rn(c23.y, hi[r].x, acc[r][1]);
            acc[r][2] = __fmaf_rn(c23.y, hi[r].y, acc[r][2]); acc[r][3] = __fmaf_rn(c23.y, hi[r].z, acc[r][3]);
        }
    }
}
template <int R>
__device__ __forceinline__ void acc_window(const float4* __restrict__ lut4, int n4m2, int q, const int (&Qb)[ACC_FB], float4 (&lo)[R],
                                           float4 (&hi)[R]) {
    // window [4q, 4q+8) of the table at the lane's ticks.  Blocks past the end of the table are clamped into it: they only
    // feed ticks that are not stored (the sampler routes samples that reach the last two blocks to the exact path).
#pragma unroll
    for (int r = 0; r < R; r++) {
        const float4* w = lut4 + min(Qb[r] + q, n4m2);
        lo[r] = __ldg(w);
        hi[r] = __ldg(w + 1);       // (skipping this load for groups whose samples all sit at 4q was measured: the warp-uniform test in
                                    // front of the loads costs more than the ~8 % of wavefronts it saves -- 1374 against 1202 ms per spill)
    }
}
template <int R>
__device__ __forceinline__ void acc_fold(float (&acc)[R][4], double (&dacc)[ACC_FB][4]) {
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int j = 0; j < 4; j++) { dacc[r][j] += (double)acc[r][j]; acc[r][j] = 0.f; }
}
template <int R>
__device__ __forceinline__ void acc_gather(const float4* __restrict__ lut4, int n4m2, const GroupRec* __restrict__ grp, int ng,
                                           const int (&Qb)[ACC_FB], double (&dacc)[ACC_FB][4]) {
    // two groups per trip, ping-pong windows: while group g is applied, the window of g+1 and the records of g+2, g+3
    // are in flight; the float32 partial sums (<= 8 terms) are folded into float64 once per trip
    const int4* recs = reinterpret_cast<const int4*>(grp);
    const int4 none = make_int4(0, 0, 0, 0);                          // zero counts: applies nothing
    float acc[R][4];
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[r][j] = 0.f;
    float4 loA[R], hiA[R], loB[R], hiB[R];
    int4 recA = __ldg(recs);
    int4 recB = ng > 1 ? __ldg(recs + 1) : none;
    acc_window<R>(lut4, n4m2, recA.x, Qb, loA, hiA);
    for (int g = 0; g < ng; g += 2) {
        acc_window<R>(lut4, n4m2, recB.x, Qb, loB, hiB);
        const int4 recC = g + 2 < ng ? __ldg(recs + g + 2) : none;
        const int4 recD = g + 3 < ng ? __ldg(recs + g + 3) : none;
        acc_apply<R>(recA, loA, hiA, acc);
        acc_window<R>(lut4, n4m2, recC.x, Qb, loA, hiA);
        acc_apply<R>(recB, loB, hiB, acc);
        acc_fold<R>(acc, dacc);
        recA = recC; recB = recD;
    }
}

// ---- TMA variant of the grouped interior path (FAST == 3) ---------------------------------------------------------------------------
// The 32 lanes of a warp read, for one group record and one 128-tick block, the CONTIGUOUS table words [4(Qb0 + q), + 132): one
// `cp.async.bulk` (TMA, 1-D; UBLKCP in SASS) per group brings them into a per-warp ring of shared-memory stages, completion is
// signalled on an mbarrier per stage, and every lane then takes its 8-word window with two LDS.128.  A/B against the LDG.128 path:
// profiles/r02_spill_pipeline.md (LSB_ACC_TMA=1 selects it).
#define ACC_TMA_STAGES 4
#define ACC_TMA_WORDS 136                       // 132 used; stage stride 544 B (16-byte aligned)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// ---- lane = 8 consecutive ticks on the 4-word groups (FAST == 2) -----------------------------------------------------------
// The window of a group grows from 8 to 12 table words per lane (three aligned LDG.128) but serves twice the ticks: 1.5 words
// per tick instead of 2, and the per-group overhead (record, count unpacking, branches) is paid once per 256 ticks of a warp.
__device__ __forceinline__ void acc_apply_t8(const int4& rec, const float4& a, const float4& b, const float4& c, float (&acc)[8]) {
    const float w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    const float2 c01 = __half22float2(*reinterpret_cast<const __half2*>(&rec.y));
    const float2 c23 = __half22float2(*reinterpret_cast<const __half2*>(&rec.z));
    if (c01.x != 0.f) {
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = __fmaf_rn(c01.x, w[j], acc[j]);
    }
    if (c01.y != 0.f) {
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = __fmaf_rn(c01.y, w[j + 1], acc[j]);
    }
    if (c23.x != 0.f) {
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = __fmaf_rn(c23.x, w[j + 2], acc[j]);
    }
    if (c23.y != 0.f) {
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = __fmaf_rn(c23.y, w[j + 3], acc[j]);
    }
}
__device__ __forceinline__ void acc_gather_t8(const float4* __restrict__ lut4, int n4m3, const GroupRec* __restrict__ grp, int ng, int Qb,
                                              double (&dacc)[8]) {
    // record g+1 and its window are in flight while group g is applied; float32 partial sums (<= 8 terms) are folded into
    // float64 every second group, as in the 4-tick path
    const int4* recs = reinterpret_cast<const int4*>(grp);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
    int4 rec = __ldg(recs);
    const float4* w = lut4 + min(Qb + rec.x, n4m3);
    float4 a = __ldg(w), b = __ldg(w + 1), c = __ldg(w + 2);
    for (int g = 0; g < ng; g++) {
        int4 nrec = rec;
        float4 na = a, nb = b, nc = c;
        if (g + 1 < ng) {
            nrec = __ldg(recs + g + 1);
            const float4* w1 = lut4 + min(Qb + nrec.x, n4m3);
            na = __ldg(w1); nb = __ldg(w1 + 1); nc = __ldg(w1 + 2);
        }
        acc_apply_t8(rec, a, b, c, acc);
        if (g & 1) {
#pragma unroll
            for (int j = 0; j < 8; j++) { dacc[j] += (double)acc[j]; acc[j] = 0.f; }
        }
        rec = nrec; a = na; b = nb; c = nc;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) dacc[j] += (double)acc[j];
}

// ---- 8-word groups: lane = 8 consecutive ticks, 16-word window fetched with two 256-bit loads ------------------------
struct __align__(32) float8 { float v[8]; };
__device__ __forceinline__ float8 ldg256(const float* p) {
    float8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void acc_apply8(const int2& c0123, const int2& c4567, const float8& lo, const float8& hi, float (&acc)[8]) {
    const float w[16] = {lo.v[0], lo.v[1], lo.v[2], lo.v[3], lo.v[4], lo.v[5], lo.v[6], lo.v[7],
                         hi.v[0], hi.v[1], hi.v[2], hi.v[3], hi.v[4], hi.v[5], hi.v[6], hi.v[7]};
    const int packed[4] = {c0123.x, c0123.y, c4567.x, c4567.y};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (packed[k] == 0) continue;                                   // both counts of the pair are zero
        const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&packed[k]));
        if (c.x != 0.f) {
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] = __fmaf_rn(c.x, w[2 * k + j], acc[j]);
        }
        if (c.y != 0.f) {
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] = __fmaf_rn(c.y, w[2 * k + 1 + j], acc[j]);
        }
    }
}
// one warp, one block of 256 ticks: walks the pair's group records; record g+1 and its window are in flight while group g
// is applied; float32 partial sums are folded into float64 every ACC_FLUSH groups
__device__ __forceinline__ void acc_gather8(const float* __restrict__ lut, int n8m2, const GroupRec8* __restrict__ grp, int ng, int Qb,
                                            double (&dacc)[8]) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
    const int2* recs = reinterpret_cast<const int2*>(grp);              // 3 x int2 per record: {q, pad} {c01, c23} {c45, c67}
    int2 h = __ldg(recs), ca = __ldg(recs + 1), cb = __ldg(recs + 2);
    const float* w0 = lut + 8LL * min(Qb + h.x, n8m2);
    float8 lo = ldg256(w0), hi = ldg256(w0 + 8);
    for (int g = 0; g < ng; g++) {
        int2 nh = h, nca = ca, ncb = cb;
        float8 nlo = lo, nhi = hi;
        if (g + 1 < ng) {
            nh = __ldg(recs + 3 * (g + 1)); nca = __ldg(recs + 3 * (g + 1) + 1); ncb = __ldg(recs + 3 * (g + 1) + 2);
            // blocks past the end of the table are clamped into it: they only feed ticks that are not stored
            const float* w1 = lut + 8LL * min(Qb + nh.x, n8m2);
            nlo = ldg256(w1); nhi = ldg256(w1 + 8);
        }
        acc_apply8(ca, cb, lo, hi, acc);
        if ((g & (ACC_FLUSH - 1)) == ACC_FLUSH - 1) {
#pragma unroll
            for (int j = 0; j < 8; j++) { dacc[j] += (double)acc[j]; acc[j] = 0.f; }
        }
        h = nh; ca = nca; cb = ncb; lo = nlo; hi = nhi;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) dacc[j] += (double)acc[j];
}


// ---- phase-aligned interior path (FAST == 4) --------------------------------------------------------------------------------------
// The grouped path above fetches an 8-word window per lane and group because an offset 4q + d with d != 0 straddles two aligned
// float4s: 2 table words travel from L1 to the registers for every tick of a group, used or not (ncu: the L1 data pipe is the
// limit, profiles/r02_top_kernels_ndlar_unit.md).  Here the LANES move instead of the window: for the offsets of alignment class d
// a lane accumulates the four ticks 4 lane - d .. 4 lane - d + 3 of the block, so the words it needs, LUT[4q + d + tick], are the
// ALIGNED float4 q + lane -- one LDG.128 and four FFMA per distinct offset, 1 word per tick.  A warp's 32 lanes cover the ticks
// [-d, 128 - d) of its block; blocks are 124 ticks long, so every class covers the block whatever d is (3 % of the lanes' work is
// redundant) and nothing crosses warps.  When a class is finished its sums move to the lanes that own the ticks in the output
// (tick 4 lane + i comes from accumulator (i + d) & 3 of the same lane or of lane + 1: d shuffles).
#define ACC_AL_BLOCK 124
#ifndef ACC_AL_MINB
#define ACC_AL_MINB 8
#endif
template <int D>
__device__ __forceinline__ void acc_class_flush(double (&cacc)[4], double (&dacc)[4]) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double v = cacc[(i + D) & 3];
        if (i + D >= 4) v = __shfl_down_sync(0xffffffffu, v, 1);
        dacc[i] += v;
        }
#pragma unroll
    for (int j = 0; j < 4; j++) cacc[j] = 0.0;
}
__device__ __forceinline__ void acc_run_apply(const float4* __restrict__ lut4, int n4m1, int QL, int rec, float& x, float& y, float& z, float& w) {
    const float4 t = __ldg(lut4 + min((int)((unsigned)rec >> RUN_COUNT_BITS) + QL, n4m1));
    // the count as a float without a conversion: 2^23 + c has c in its low mantissa bits
    const float c = __fadd_rn(__int_as_float(0x4B000000 | (rec & ((1 << RUN_COUNT_BITS) - 1))), -8388608.0f);
    x = __fmaf_rn(c, t.x, x); y = __fmaf_rn(c, t.y, y); z = __fmaf_rn(c, t.z, z); w = __fmaf_rn(c, t.w, w);
}
__device__ __forceinline__ void acc_run_apply4(const float4* __restrict__ lut4, int n4m1, int QL, const int4& a, float& x, float& y, float& z, float& w) {
    acc_run_apply(lut4, n4m1, QL, a.x, x, y, z, w); acc_run_apply(lut4, n4m1, QL, a.y, x, y, z, w);
    acc_run_apply(lut4, n4m1, QL, a.z, x, y, z, w); acc_run_apply(lut4, n4m1, QL, a.w, x, y, z, w);
}
__device__ __forceinline__ void acc_class_run(const float4* __restrict__ lut4, int n4m1, const int* __restrict__ recs, int n, int QL,
                                              double (&cacc)[4]) {
    // records one at a time up to the first 16-byte boundary, then eight per trip (two LDG.128 of records fetched one trip ahead,
    // eight LDG.128 of the table, 32 FFMA, one fold of the <= 8-term float32 sums into float64), then the rest.
    // Measured and dropped (B200, tracks_current_mc stage of one ND-LAr unit, 8.30 ms as written): `prefetch.global.L1` of the next
    // trip's table lines 11.1 ms; quads masked at both ends instead of the single-record head and tail 8.50; 80 registers / 6 CTAs
    // per SM 9.35; 48 registers / 10 CTAs per SM (spills) 10.3; warp = class instead of warp = block (the four warps walk the
    // four class lists of the same block together so that they share table lines in L1; sums meet in shared memory) 9.08.
    int r = 0;
    float x = 0.f, y = 0.f, z = 0.f, w = 0.f;
    const int lead = (int)((16u - ((unsigned)reinterpret_cast<uintptr_t>(recs) & 15u)) & 15u) >> 2;
    for (; r < n && r < lead; r++) acc_run_apply(lut4, n4m1, QL, __ldg(recs + r), x, y, z, w);
    if (r > 0) { cacc[0] += (double)x; cacc[1] += (double)y; cacc[2] += (double)z; cacc[3] += (double)w; x = y = z = w = 0.f; }
    if (r + 8 <= n) {
        int4 a = __ldg(reinterpret_cast<const int4*>(recs + r)), b = __ldg(reinterpret_cast<const int4*>(recs + r + 4));
        for (; r + 8 <= n; r += 8) {
            int4 na = a, nb = b;
            if (r + 16 <= n) { na = __ldg(reinterpret_cast<const int4*>(recs + r + 8)); nb = __ldg(reinterpret_cast<const int4*>(recs + r + 12)); }
            acc_run_apply4(lut4, n4m1, QL, a, x, y, z, w);
            acc_run_apply4(lut4, n4m1, QL, b, x, y, z, w);
            cacc[0] += (double)x; cacc[1] += (double)y; cacc[2] += (double)z; cacc[3] += (double)w; x = y = z = w = 0.f;
            a = na; b = nb;
        }
    }
    if (r + 4 <= n) {
        const int4 a = __ldg(reinterpret_cast<const int4*>(recs + r));
        acc_run_apply4(lut4, n4m1, QL, a, x, y, z, w);
        r += 4;
    }
    for (; r < n; r++) acc_run_apply(lut4, n4m1, QL, __ldg(recs + r), x, y, z, w);
    cacc[0] += (double)x; cacc[1] += (double)y; cacc[2] += (double)z; cacc[3] += (double)w;
}

template <typename TL, int STRIDE, int FAST>
__device__ __forceinline__ void mc_accumulate_pair(const McParams& p, long long pr, const PairRec* __restrict__ pairs,
                                                   const SampleRec* __restrict__ samples,
                                                   const int* __restrict__ offs32, const GroupRecT* __restrict__ groups,
                                                   const TL* __restrict__ lut, float* __restrict__ signals,
                                                   const TL* __restrict__ lut_exact) {
    // lut: the table the affine samples index with `rowoff + shift + STRIDE * tick` (the phase-split copy when p.split == 2);
    // lut_exact: the table as the caller passed it, read by the exact per-tick path of the irregular samples
    const PairRec* gp = pairs + pr;
    if (!gp->valid) {
        if (p.ranges && threadIdx.x == 0) p.ranges[(p.seg0 + pr / p.P) * p.P + (pr % p.P)] = make_int2(0, -1);
        return;
    }
    // one staging buffer for the phases that follow each other (each separated by a barrier): sample offsets of the generic
    // interior, 16-byte edge records, full sample records of the irregular path -- less shared memory is more L1 for the table
    __shared__ __align__(16) char s_stage[ACC_CHUNK * 16];
    int* s_off = reinterpret_cast<int*>(s_stage);
    SampleRec* s_rec = reinterpret_cast<SampleRec*>(s_stage);
    static_assert(sizeof(SampleRec) * ACC_TPB <= ACC_CHUNK * 16, "staging buffer too small for the irregular path");
    const int tid = threadIdx.x;
    const int T = p.T;
    const int it_first = gp->it_first, n_live = gp->n_live;
    const double charge = gp->charge, t_start = gp->t_start;
    const long long soff = gp->sample_off;
    long long itrk = p.seg0 + pr / p.P;
    float* out = signals + (itrk * p.P + (pr % p.P)) * (long long)T;
    const int n_irr = gp->n_irregular;
    const int uni_lo = gp->uni_lo, uni_hi = gp->uni_hi;
    int int_lo = gp->int_lo, int_hi = gp->int_hi;
    if (STRIDE == 0 || n_live - n_irr <= 0 || int_lo > int_hi) { int_lo = 0; int_hi = -1; }   // no interior

    // ---- interior ticks, grouped path ------------------------------------------------
    // Tick blocks are dealt round-robin to the warps; warps run independently (no barrier).
    if constexpr (FAST == 1) {
        constexpr int NW = ACC_TPB / 32;
        const int lane = tid & 31, warp = tid >> 5;
        const int ng = gp->n_groups;
#if ACC_GW == 8
        // lane = 8 consecutive ticks, block = 256 ticks
        const int n8m2 = (int)(((long long)p.Rx * p.Ry * p.Rt) >> 3) - 2;
        const GroupRec8* grp = groups + soff;
        for (int tb = int_lo + 256 * warp; tb <= int_hi; tb += 256 * NW) {
            double dacc[8];
#pragma unroll
            for (int j = 0; j < 8; j++) dacc[j] = 0.0;
            acc_gather8(reinterpret_cast<const float*>(lut), n8m2, grp, ng, ((tb - int_lo) >> 3) + lane, dacc);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int it = tb + 8 * lane + j;
                if (it <= int_hi) out[it] = __double2float_rn(charge * dacc[j]);
            }
        }
#else
        const float4* lut4 = reinterpret_cast<const float4*>(lut);
        const int n4m2 = (int)(((long long)p.Rx * p.Ry * p.Rt) >> 2) - 2;
        const GroupRec* grp = groups + soff;
        for (int tb = int_lo; tb <= int_hi; tb += 128 * NW * ACC_FB) {
            double dacc[ACC_FB][4];
            int Qb[ACC_FB];
            int nR = 0;
#pragma unroll
            for (int r = 0; r < ACC_FB; r++) {
                const int t0 = (warp + NW * r) * 128;                 // first tick of the block, relative to tb
                if (tb + t0 <= int_hi) nR = r + 1;
                Qb[r] = ((tb - int_lo + t0) >> 2) + lane;
#pragma unroll
                for (int j = 0; j < 4; j++) dacc[r][j] = 0.0;
            }
            if (nR == 1) acc_gather<1>(lut4, n4m2, grp, ng, Qb, dacc);
            else if (nR == 2) acc_gather<(ACC_FB > 1 ? 2 : 1)>(lut4, n4m2, grp, ng, Qb, dacc);
#pragma unroll
            for (int r = 0; r < ACC_FB; r++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int it = tb + (warp + NW * r) * 128 + 4 * lane + j;
                    if (r < nR && it <= int_hi) out[it] = __double2float_rn(charge * dacc[r][j]);
                }
        }
#endif
    }
    if constexpr (FAST == 4) {
        constexpr int NW = ACC_TPB / 32;
        const int lane = tid & 31, warp = tid >> 5;
        const int n_items = gp->n_groups;
        const float4* lut4 = reinterpret_cast<const float4*>(lut);
        const int n4m1 = (int)(((long long)p.Rx * p.Ry * p.Rt) >> 2) - 1;
        const int* stream = run_stream(groups, soff);
        for (int tb = int_lo + ACC_AL_BLOCK * warp; tb <= int_hi && n_items > 0; tb += ACC_AL_BLOCK * NW) {
            const int QL = ((tb - int_lo) >> 2) + lane;
            double dacc[4] = {0.0, 0.0, 0.0, 0.0}, cacc[4] = {0.0, 0.0, 0.0, 0.0};
            for (int item = 0; item < n_items;) {
                const int2 hdr = __ldg(reinterpret_cast<const int2*>(stream + item));
                const int* recs = stream + item + 2;
                const int c1 = hdr.x & 1023, c2 = (hdr.x >> 10) & 1023, c3 = (hdr.x >> 20) & 1023, c4 = hdr.y;
                if (c1 > 0) { acc_class_run(lut4, n4m1, recs, c1, QL, cacc); acc_class_flush<0>(cacc, dacc); }
                if (c2 > c1) { acc_class_run(lut4, n4m1, recs + c1, c2 - c1, QL, cacc); acc_class_flush<1>(cacc, dacc); }
                if (c3 > c2) { acc_class_run(lut4, n4m1, recs + c2, c3 - c2, QL, cacc); acc_class_flush<2>(cacc, dacc); }
                if (c4 > c3) { acc_class_run(lut4, n4m1, recs + c3, c4 - c3, QL, cacc); acc_class_flush<3>(cacc, dacc); }
                item += 2 + c4 + (c4 & 1);                                  // headers stay 8-byte aligned
            }
            if (lane < 31) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int it = tb + 4 * lane + j;
                    if (it <= int_hi) out[it] = __double2float_rn(charge * dacc[j]);
                }
            }
        }
    }
    if constexpr (FAST == 3) {
        constexpr int NW = ACC_TPB / 32, NS = ACC_TMA_STAGES;
        __shared__ __align__(128) float s_win[NW][NS][ACC_TMA_WORDS];
        __shared__ __align__(8) unsigned long long s_bar[NW][NS];
        const int lane = tid & 31, warp = tid >> 5;
        const int ng = gp->n_groups;
        const long long n_words4 = ((long long)p.Rx * p.Ry * p.Rt) & ~3LL;     // whole 16-byte words of the table
        const float* table = reinterpret_cast<const float*>(lut);
        const int4* recs = reinterpret_cast<const int4*>(reinterpret_cast<const GroupRec*>(groups) + soff);
        if (lane == 0) {
#pragma unroll
            for (int s2 = 0; s2 < NS; s2++) mbar_init(smem_u32(&s_bar[warp][s2]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        unsigned it = 0;                                    // copies consumed so far by this warp (stage = it % NS, parity = (it / NS) & 1)
        unsigned issued = 0;
        for (int tb = int_lo + 128 * warp; tb <= int_hi && ng > 0; tb += 128 * NW) {
            const long long base4 = (long long)(tb - int_lo);              // first table word of the block's window, relative to 4q
            auto issue = [&](int g) {
                if (lane == 0) {
                    const int q = __ldg(&recs[g].x);
                    long long w0 = 4LL * q + base4;
                    long long nw = n_words4 - w0;
                    if (nw > 132) nw = 132;
                    if (nw < 4) { nw = 4; w0 = n_words4 - 4; }              // (never needed by a stored tick)
                    const unsigned st = issued % NS;
                    const uint32_t bar = smem_u32(&s_bar[warp][st]);
                    mbar_expect_tx(bar, (uint32_t)nw * 4u);
                    tma_load_1d(smem_u32(&s_win[warp][st][0]), table + w0, (uint32_t)nw * 4u, bar);
                }
                issued++;
            };
            const int pre = ng < NS - 1 ? ng : NS - 1;
            for (int g = 0; g < pre; g++) issue(g);
            double dacc[4] = {0.0, 0.0, 0.0, 0.0};
            float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
            for (int g = 0; g < ng; g++) {
                const int4 rec = __ldg(recs + g);
                const unsigned st = it % NS;
                mbar_wait(smem_u32(&s_bar[warp][st]), (it / NS) & 1u);
                float4 lo[1], hi[1];
                const float4* w = reinterpret_cast<const float4*>(&s_win[warp][st][0]) + lane;
                lo[0] = w[0]; hi[0] = w[1];
                it++;
                __syncwarp();                                               // every lane has read the stage the next copy overwrites
                if (g + NS - 1 < ng) issue(g + NS - 1);
                acc_apply<1>(rec, lo, hi, acc);
                if (g & 1) {
#pragma unroll
                    for (int j = 0; j < 4; j++) { dacc[j] += (double)acc[0][j]; acc[0][j] = 0.f; }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                dacc[j] += (double)acc[0][j];
                const int it2 = tb + 4 * lane + j;
                if (it2 <= int_hi) out[it2] = __double2float_rn(charge * dacc[j]);
            }
        }
    }
    if constexpr (FAST == 2) {
        // lane = 8 consecutive ticks, warp = 256 ticks, 4-word groups
        constexpr int NW = ACC_TPB / 32;
        const int lane = tid & 31, warp = tid >> 5;
        const int ng = gp->n_groups;
        const float4* lut4 = reinterpret_cast<const float4*>(lut);
        const int n4m3 = (int)(((long long)p.Rx * p.Ry * p.Rt) >> 2) - 3;
        const GroupRec* grp = reinterpret_cast<const GroupRec*>(groups) + soff;
        for (int tb = int_lo + 256 * warp; tb <= int_hi; tb += 256 * NW) {
            double dacc[8];
#pragma unroll
            for (int j = 0; j < 8; j++) dacc[j] = 0.0;
            acc_gather_t8(lut4, n4m3, grp, ng, ((tb - int_lo) >> 2) + 2 * lane, dacc);
            const int it0 = tb + 8 * lane;
            if (it0 + 7 <= int_hi && ((reinterpret_cast<uintptr_t>(out + it0) & 15) == 0)) {
                *reinterpret_cast<float4*>(out + it0) = make_float4(__double2float_rn(charge * dacc[0]), __double2float_rn(charge * dacc[1]),
                                                                    __double2float_rn(charge * dacc[2]), __double2float_rn(charge * dacc[3]));
                *reinterpret_cast<float4*>(out + it0 + 4) = make_float4(__double2float_rn(charge * dacc[4]), __double2float_rn(charge * dacc[5]),
                                                                        __double2float_rn(charge * dacc[6]), __double2float_rn(charge * dacc[7]));
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (it0 + j <= int_hi) out[it0 + j] = __double2float_rn(charge * dacc[j]);
            }
        }
    }
    // ---- interior ticks, generic path: unconditional gather stream ---------------------
    if (STRIDE > 0 && !FAST) {
        for (int base = int_lo; base <= int_hi; base += ACC_TPB * ACC_RMAX) {
            int n_int = int_hi - base + 1;
            if (n_int > ACC_TPB * ACC_RMAX) n_int = ACC_TPB * ACC_RMAX;
            const int nfull = n_int / ACC_TPB, rem = n_int - nfull * ACC_TPB;
            const bool rem_ok = tid < rem;
            double dacc[ACC_RMAX];
#pragma unroll
            for (int r = 0; r < ACC_RMAX; r++) dacc[r] = 0.0;
            for (int c0 = 0; c0 < n_live; c0 += ACC_CHUNK) {
                int ns = n_live - c0 < ACC_CHUNK ? n_live - c0 : ACC_CHUNK;
                __syncthreads();
                for (int q = tid; q < ns; q += ACC_TPB) s_off[q] = offs32[soff + c0 + q];
                __syncthreads();
#define ACC_CASE(N)                                                                                                   \
    case N:                                                                                                           \
        if (n_irr) acc_interior<TL, (STRIDE > 0 ? STRIDE : 1), N, true>(lut, s_off, ns, base + tid, rem_ok, dacc);   \
        else acc_interior<TL, (STRIDE > 0 ? STRIDE : 1), N, false>(lut, s_off, ns, base + tid, rem_ok, dacc);        \
        break;
                switch (nfull) {
                    ACC_CASE(0) ACC_CASE(1) ACC_CASE(2) ACC_CASE(3) ACC_CASE(4) ACC_CASE(5) ACC_CASE(6) ACC_CASE(7)
                    default:
                        if (n_irr) acc_interior<TL, (STRIDE > 0 ? STRIDE : 1), 8, true>(lut, s_off, ns, base + tid, rem_ok, dacc);
                        else acc_interior<TL, (STRIDE > 0 ? STRIDE : 1), 8, false>(lut, s_off, ns, base + tid, rem_ok, dacc);
                }
#undef ACC_CASE
            }
            // irregular samples (off < 0) contribute to interior ticks through the exact path below,
            // so interior results are written with "+=" semantics into a zeroed output: store now,
            // the exact path adds on top.
#pragma unroll
            for (int r = 0; r < ACC_RMAX; r++) {
                int it = base + tid + r * ACC_TPB;
                if (it <= int_hi && (r < nfull || (r == nfull && rem_ok))) out[it] = __double2float_rn(charge * dacc[r]);
            }
        }
    }
    __syncthreads();

    // ---- (a) ticks >= it_first that no sample covers: the reference stores total_current = 0 ----
    if (p.ranges) {
        // sparse rows: only [uni_lo, uni_hi] is stored.  The interior / edge paths write every tick of it; if there are only
        // irregular samples it is zeroed here and (c) adds them on top.
        if (tid == 0) p.ranges[itrk * p.P + (pr % p.P)] = n_live > 0 ? make_int2(uni_lo, uni_hi) : make_int2(0, -1);
        if (n_live > 0 && n_live - n_irr <= 0)
            for (int it = uni_lo + tid; it <= uni_hi; it += ACC_TPB) out[it] = 0.f;
    } else {
        for (int it = it_first + tid; it < T; it += ACC_TPB)
            if (n_live - n_irr <= 0 || it < uni_lo || it > uni_hi) out[it] = 0.f;   // (c) adds the irregular samples on top
    }

    // ---- (b) edge ticks: inside the union of the sample windows but outside the interior -------
    if (STRIDE > 0 && n_live - n_irr > 0) {
        int n_left, right0;
        if (int_lo <= int_hi) { n_left = int_lo - uni_lo; right0 = int_hi + 1; }
        else { n_left = uni_hi - uni_lo + 1; right0 = uni_hi + 1; }
        int n_edge = n_left + (uni_hi - right0 + 1);
        // 32 edge ticks per pass (lane = tick); the ACC_TPB/32 warps split the samples and their float64
        // partial sums are combined in a fixed order.  Samples are staged as 16-byte {offset, lo, hi - lo}
        // records (one LDS.128 and one unsigned range test per sample); four predicated loads are in
        // flight together and their float32 sum is folded into float64.
        constexpr int NG = ACC_TPB / 32;
        __shared__ double s_part[NG][32];
        int4* s_erec = reinterpret_cast<int4*>(s_stage);
        const int lane = tid & 31, grp = tid >> 5;
        for (int e0 = 0; e0 < n_edge; e0 += 32) {
            const int e = e0 + lane;
            const bool active = e < n_edge;
            const int it = e < n_left ? uni_lo + e : right0 + (e - n_left);
            const TL* lutt = lut + STRIDE * it;
            double sum = 0.0;
            for (int c0 = 0; c0 < n_live; c0 += ACC_CHUNK) {
                int ns = n_live - c0 < ACC_CHUNK ? n_live - c0 : ACC_CHUNK;
                __syncthreads();
                for (int q = tid; q < ns; q += ACC_TPB) {
                    const int off = offs32[soff + c0 + q];
                    const unsigned int lh = p.lohi[soff + c0 + q];
                    const int lo_s = (int)(lh & 0xffffu), hi_s = (int)(lh >> 16);
                    // irregular samples never pass the range test
                    s_erec[q] = off == OFF_IRREGULAR ? make_int4(0, 0x3fffffff, 0, 0) : make_int4(off, lo_s, hi_s - lo_s, 0);
                }
                __syncthreads();
                if (!active) continue;
                int sidx = grp;
                for (; sidx + 3 * NG < ns; sidx += 4 * NG) {
                    float v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int4 r = s_erec[sidx + u * NG];
                        v[u] = 0.f;
                        if ((unsigned)(it - r.y) <= (unsigned)r.z) v[u] = (float)__ldg(lutt + r.x);
                    }
                    sum += (double)((v[0] + v[1]) + (v[2] + v[3]));
                }
                for (; sidx < ns; sidx += NG) {
                    const int4 r = s_erec[sidx];
                    if ((unsigned)(it - r.y) <= (unsigned)r.z) sum += (double)(float)__ldg(lutt + r.x);
                }
            }
            s_part[grp][lane] = sum;
            __syncthreads();
            if (grp == 0 && active) {
                double tot = s_part[0][lane];
#pragma unroll
                for (int g = 1; g < NG; g++) tot += s_part[g][lane];
                out[it] = __double2float_rn(charge * tot);
            }
        }
    }

    // ---- (c) irregular samples (tick shift not affine, or non-integral sampling ratio): exact
    //          per-tick evaluation of detsim.py:333 and :211-218, added on top ----------------
    if (n_irr > 0) {
        const double TS = d_c.time_sampling, W = d_c.time_window;
        __syncthreads();
        for (int it0 = it_first; it0 < T; it0 += ACC_TPB) {
            int it = it0 + tid;
            bool active = it < T;
            double sum = 0.0;
            bool any = false;
            for (int c0 = 0; c0 < n_live; c0 += ACC_TPB) {
                int ns = n_live - c0 < ACC_TPB ? n_live - c0 : ACC_TPB;
                __syncthreads();
                if (tid < ns) {
                    if (offs32[soff + c0 + tid] == OFF_IRREGULAR) s_rec[tid] = samples[soff + c0 + tid];
                    else s_rec[tid].shift = 0;                       // regular: skipped below
                }
                __syncthreads();
                if (!active) continue;
                for (int s = 0; s < ns; s++) {
                    const SampleRec& r = s_rec[s];
                    if (r.shift != SHIFT_IRREGULAR) continue;
                    double tick = t_start + (double)it * TS;
                    if (tick < 0) continue;
                    if (!(r.t0 < tick && tick < r.t0 + W)) continue;
                    long long k = __double2ll_rn((tick - r.t0) / d_c.response_sampling);
                    if (0 <= k && k < p.Rt) { sum += (double)lut_exact[(long long)r.rowoff + k]; any = true; }
                }
            }
            if (active && any) out[it] = __double2float_rn((double)out[it] + charge * sum);
        }
    }
}

// one CTA per (segment, pixel) pair.  (A persistent grid fetching pairs from a device counter was measured on B200: 3.6 % slower
// alone -- an atomic and two barriers per pair -- and no better with three batches in flight, profiles/r02_spill_pipeline.md.)
template <typename TL, int STRIDE, int FAST>
__global__ void __launch_bounds__(ACC_TPB, FAST >= 4 ? ACC_AL_MINB : FAST == 1 ? ACC_MINB : 8) k_mc_accumulate(McParams p, const PairRec* __restrict__ pairs,
                                                           const SampleRec* __restrict__ samples,
                                                           const int* __restrict__ offs32, const GroupRecT* __restrict__ groups,
                                                           const TL* __restrict__ lut, float* __restrict__ signals,
                                                           const TL* __restrict__ lut_exact) {
    MC_GUARD(p);
    mc_accumulate_pair<TL, STRIDE, FAST>(p, (long long)blockIdx.x, pairs, samples, offs32, groups, lut, signals, lut_exact);
}

// replay mode: literal reference order, one thread per pair (detsim.py:324-348)
template <typename TL>
__global__ void k_mc_replay(McParams p, const PairRec* __restrict__ pairs, const TL* __restrict__ lut,
                            float* __restrict__ signals, unsigned long long* __restrict__ rng_states) {
    long long pr = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (pr >= p.S * p.P) return;
    PairRec g = pairs[pr];
    if (!g.valid) return;
    long long itrk = p.seg0 + pr / p.P;
    int ipix = (int)(pr % p.P);
    unsigned long long* sp = rng_states + 2 * (itrk + p.rng_stride * ipix);
    Rng rng; rng.s0 = sp[0]; rng.s1 = sp[1];
    float* out = signals + (itrk * p.P + ipix) * (long long)p.T;
    const int M = d_c.mc_sample_multiplier;
    for (int it = g.it_first; it < p.T; it++) {
        double time_tick = g.t_start + (double)it * d_c.time_sampling;
        if (time_tick < 0) continue;
        double total = 0;
        for (long long istep = 0; istep < g.nstep; istep++)
            for (int m = 0; m < M; m++) {
                double f = (double)istep + 0.5;
                double x = g.sub_start[0] + g.step * f * g.dir[0];
                double y = g.sub_start[1] + g.step * f * g.dir[1];
                double z = g.sub_start[2] + g.step * f * g.dir[2];
                z += R32((double)rng_normal_f32(rng) * g.sig_l, g.s32);
                double t0 = fabs(z - g.z_anode) / d_c.v_drift - d_c.time_window;
                if (!(t0 < time_tick && time_tick < t0 + d_c.time_window)) continue;
                x += R32((double)rng_normal_f32(rng) * g.sig_t, g.s32);
                y += R32((double)rng_normal_f32(rng) * g.sig_t, g.s32);
                double x_dist = fabs(g.x_p - x), y_dist = fabs(g.y_p - y);
                if (x_dist > d_c.response_bin_size * p.Rx) continue;
                if (y_dist > d_c.response_bin_size * p.Ry) continue;
                long long i = __double2ll_rn(x_dist / d_c.response_bin_size - 0.5);
                long long j = __double2ll_rn(y_dist / d_c.response_bin_size - 0.5);
                long long k = __double2ll_rn((time_tick - t0) / d_c.response_sampling);
                if (0 <= i && i < p.Rx && 0 <= j && j < p.Ry && 0 <= k && k < p.Rt)
                    total += g.charge * (double)lut[(i * p.Ry + j) * (long long)p.Rt + k];
            }
        out[it] = __double2float_rn(total);
    }
    sp[0] = rng.s0; sp[1] = rng.s1;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct McWs {
    PairRec* pairs; uint32_t* nsamp; long long* offs; long long* block_sums; long long* total;
    int* perm; int* bucket;   // pairs ordered by descending sample count (k_mc_uniforms), 2 x MC_NBUCKET counters
    SampleRec* samples; int* offs32; unsigned int* lohi; SampleU* uu; long long sample_cap;
};
#define MC_BYTES_PER_SAMPLE ((long long)(sizeof(SampleRec) + 4 + 4 + sizeof(SampleU)))
static inline long long align_up(long long x, long long a) { return (x + a - 1) / a * a; }
static inline long long mc_fixed_bytes(long long npair) {
    return align_up(npair * (long long)sizeof(PairRec), 256) + align_up(npair * 4, 256) + align_up(npair * 8, 256) +
           align_up((scan_num_blocks(npair) + 1) * 8, 256) + 256 + align_up(npair * 4, 256) + 1024;
}
LSB_EXPORT int64_t lsb_tracks_current_mc_workspace_bytes(int64_t S, int32_t P, int64_t max_steps_total) {
    return mc_fixed_bytes(S * (long long)P) + align_up(max_steps_total * (long long)sizeof(SampleRec), 256) +
           2 * align_up(max_steps_total * 4, 256) + align_up(max_steps_total * (long long)sizeof(SampleU), 256) + 2048;
}
static inline bool mc_carve(void* ws, long long bytes, long long npair, McWs& w) {
    char* p = (char*)ws;
    long long fixed = mc_fixed_bytes(npair);
    if (bytes < fixed + MC_BYTES_PER_SAMPLE * 64 + 2048) return false;
    w.pairs = (PairRec*)p; p += align_up(npair * (long long)sizeof(PairRec), 256);
    w.nsamp = (uint32_t*)p; p += align_up(npair * 4, 256);
    w.offs = (long long*)p; p += align_up(npair * 8, 256);
    w.block_sums = (long long*)p; p += align_up((scan_num_blocks(npair) + 1) * 8, 256);
    w.total = (long long*)p; p += 256;
    w.perm = (int*)p; p += align_up(npair * 4, 256);
    w.bucket = (int*)p; p += 1024;
    long long rest = bytes - fixed;
    w.sample_cap = (rest - 2048) / MC_BYTES_PER_SAMPLE;
    w.samples = (SampleRec*)p; p += align_up(w.sample_cap * (long long)sizeof(SampleRec), 256);
    w.offs32 = (int*)p; p += align_up(w.sample_cap * 4, 256);
    w.lohi = (unsigned int*)p; p += align_up(w.sample_cap * 4, 256);
    w.uu = (SampleU*)p;
    return w.sample_cap > 0;
}

// Interior-tick strategy of k_mc_accumulate for float tables at unit sampling ratio: 1 = grouped path (k_mc_sort +
// register windows; default), 0 = generic gather stream.  Measured on B200 (profiles/r01_mc_grouped.md): 6x fewer L1
// requests, 4.44 + 0.69 ms against 5.34 ms per 1e4-segment batch.
static int g_mc_grouped = -1;
LSB_EXPORT void lsb_mc_set_grouped(int32_t on) { g_mc_grouped = on ? 1 : 0; }
LSB_EXPORT int32_t lsb_mc_get_grouped(void) {
    if (g_mc_grouped < 0) { const char* e = getenv("LSB_MC_GROUPED"); g_mc_grouped = (e && e[0] == '0') ? 0 : 1; }
    return g_mc_grouped;
}

// phase-aligned interior (k_mc_accumulate<float, 1, 4>): -1 = automatic (on for phase-split tables, where few offsets share an aligned
// 4-word block), 0 = off, 1 = on; LSB_ACC_ALIGNED / lsb_mc_set_aligned.  profiles/r02_acc_aligned.md
static int g_mc_aligned = -2;
LSB_EXPORT void lsb_mc_set_aligned(int32_t mode) { g_mc_aligned = mode < 0 ? -1 : (mode ? 1 : 0); }
LSB_EXPORT int32_t lsb_mc_get_aligned(void) {
    if (g_mc_aligned == -2) { const char* e = getenv("LSB_ACC_ALIGNED"); g_mc_aligned = !e ? -1 : (e[0] == '1' ? 1 : 0); }
    return g_mc_aligned;
}

// ticks per lane of the grouped interior path: 4 (8-word windows; default) or 8 (12-word windows: 1.5 instead of 2 table words per
// tick, but 64 registers / 32 warps per SM -- measured slower on B200: 7.39 against 6.26 ms per module0 batch, 12.9 against 11.0 ms per
// ND-LAr unit, profiles/r02_spill_pipeline.md); LSB_ACC_LANE_TICKS / lsb_mc_set_lane_ticks
static int g_mc_lane_ticks = -1;
LSB_EXPORT void lsb_mc_set_lane_ticks(int32_t n) { g_mc_lane_ticks = n == 8 ? 8 : 4; }
LSB_EXPORT int32_t lsb_mc_get_lane_ticks(void) {
    if (g_mc_lane_ticks < 0) { const char* e = getenv("LSB_ACC_LANE_TICKS"); g_mc_lane_ticks = (e && atoi(e) == 8) ? 8 : 4; }
    return g_mc_lane_ticks;
}

template <typename TL>
static int mc_launch_accumulate(const McParams& p, const McWs& w, const TL* lut, const TL* lut_split, float* signals, cudaStream_t st) {
    unsigned grid = (unsigned)(p.S * p.P);
    if (p.stride == 1 || p.split == 2) {
        if constexpr (sizeof(TL) == 4) {
            const TL* table = p.split == 2 ? lut_split : lut;
            if (lsb_mc_get_grouped() && ((uintptr_t)table & (4 * ACC_GW - 1)) == 0) {
                // grouped path: equal / adjacent offsets must be neighbours
                // group records reuse the uniforms buffer (dead after k_mc_sampler; 24 bytes per sample >= one 16-byte record)
                GroupRecT* groups = reinterpret_cast<GroupRecT*>(w.uu);
                // phase-aligned path (LSB_ACC_ALIGNED=1): one record and one LDG.128 per distinct offset; the class bits of its sort key
                // need offsets + ticks below 2^24 words
                const int mode = lsb_mc_get_aligned();
                const bool aligned = mode < 0 ? p.split == 2 : mode == 1;
                if (aligned && ACC_GW == 4 && (long long)p.Rx * p.Ry * p.Rt + p.T < RUN_MAX_WORDS) {
                    k_mc_sort<true><<<lsb_blocks(p.S * p.P, SORT_WARPS), 32 * SORT_WARPS, 0, st>>>(p, w.pairs, w.offs32, groups);
                    LSB_LAUNCH_CHECK("k_mc_sort");
                    k_mc_accumulate<TL, 1, 4><<<grid, ACC_TPB, 0, st>>>(p, w.pairs, w.samples, w.offs32, groups, table, signals, lut);
                    LSB_LAUNCH_CHECK("k_mc_accumulate");
                    return 0;
                }
                k_mc_sort<false><<<lsb_blocks(p.S * p.P, SORT_WARPS), 32 * SORT_WARPS, 0, st>>>(p, w.pairs, w.offs32, groups);
                LSB_LAUNCH_CHECK("k_mc_sort");
                static int use_tma = -1;
                if (use_tma < 0) { const char* e = getenv("LSB_ACC_TMA"); use_tma = (e && e[0] == '1') ? 1 : 0; }
                if (use_tma) k_mc_accumulate<TL, 1, 3><<<grid, ACC_TPB, 0, st>>>(p, w.pairs, w.samples, w.offs32, groups, table, signals, lut);
                else if (lsb_mc_get_lane_ticks() == 8) k_mc_accumulate<TL, 1, 2><<<grid, ACC_TPB, 0, st>>>(p, w.pairs, w.samples, w.offs32, groups, table, signals, lut);
                else k_mc_accumulate<TL, 1, 1><<<grid, ACC_TPB, 0, st>>>(p, w.pairs, w.samples, w.offs32, groups, table, signals, lut);
                LSB_LAUNCH_CHECK("k_mc_accumulate");
                return 0;
            }
            if (p.split == 2) {
                k_mc_accumulate<TL, 1, 0><<<grid, ACC_TPB, 0, st>>>(p, w.pairs, w.samples, w.offs32, nullptr, table, signals, lut);
                LSB_LAUNCH_CHECK("k_mc_accumulate");
                return 0;
            }
        }
        k_mc_accumulate<TL, 1, 0><<<grid, ACC_TPB, 0, st>>>(p, w.pairs, w.samples, w.offs32, nullptr, lut, signals, lut);
    } else if (p.stride == 2) k_mc_accumulate<TL, 2, 0><<<grid, ACC_TPB, 0, st>>>(p, w.pairs, w.samples, w.offs32, nullptr, lut, signals, lut);
    else k_mc_accumulate<TL, 0, 0><<<grid, ACC_TPB, 0, st>>>(p, w.pairs, w.samples, w.offs32, nullptr, lut, signals, lut);
    LSB_LAUNCH_CHECK("k_mc_accumulate");
    return 0;
}

// phase-split copy of a float table sampled at half the tick length (RESPONSE_SAMPLING = TIME_SAMPLING / 2, e.g. ndlar-module.yaml):
// every row [Rt] becomes [even samples (ceil(Rt/2)) | odd samples (floor(Rt/2))]
__global__ void k_lut_phase_split(const float* __restrict__ src, float* __restrict__ dst, long long rows, int Rt) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= rows * Rt) return;
    const long long row = i / Rt; const int k = (int)(i - row * Rt);
    const int half = (Rt + 1) >> 1;
    dst[row * Rt + (k & 1) * half + (k >> 1)] = src[i];
}
static int mc_build_split(const void* response, int Rx, int Ry, int Rt, float* dst, cudaStream_t st) {
    const long long rows = (long long)Rx * Ry;
    k_lut_phase_split<<<lsb_blocks(rows * Rt, 256), 256, 0, st>>>((const float*)response, dst, rows, Rt);
    LSB_LAUNCH_CHECK("k_lut_phase_split");
    return 0;
}
static inline bool mc_wants_split(const lsb_consts* c, int f64) {
    return !f64 && c->time_sampling / c->response_sampling == 2.0 && lsb_mc_get_grouped();
}

// statistics of the last call (roofline accounting)
static long long g_mc_last_samples = 0;

static int mc_run_range(const Layout& L, const void* tracks, const int32_t* pixels, float* signals, const void* response,
                        const void* response_split, int f64, unsigned long long* rng, int mode, McParams p, void* ws, long long ws_bytes,
                        cudaStream_t st, int depth) {
    McWs w;
    long long npair = p.S * p.P;
    if (!mc_carve(ws, ws_bytes, npair, w)) return lsb_fail_arg("tracks_current_mc: workspace too small");
    k_mc_pairs<<<lsb_blocks(npair, 128), 128, 0, st>>>(L, (const char*)tracks, pixels, p, w.pairs, w.nsamp);
    LSB_LAUNCH_CHECK("k_mc_pairs");
    if (mode == 1) {
        if (f64) k_mc_replay<double><<<lsb_blocks(npair, 64), 64, 0, st>>>(p, w.pairs, (const double*)response, signals, rng);
        else k_mc_replay<float><<<lsb_blocks(npair, 64), 64, 0, st>>>(p, w.pairs, (const float*)response, signals, rng);
        LSB_LAUNCH_CHECK("k_mc_replay");
        return 0;
    }
    int rc = exclusive_scan<uint32_t, long long>(w.nsamp, npair, w.offs, w.block_sums, w.total, st);
    if (rc) return rc;
    if (p.overflow) {
        // sync-free: the caller provisioned the workspace from an upper bound; the kernels check it on the device
        p.total_dev = w.total; p.cap = w.sample_cap;
    } else {
        long long total = 0;
        LSB_CUDA(cudaMemcpyAsync(&total, w.total, 8, cudaMemcpyDeviceToHost, st));
        LSB_CUDA(cudaStreamSynchronize(st));
        if (total > w.sample_cap) {
            if (p.S <= 1) return lsb_fail_arg("tracks_current_mc: workspace too small for a single segment");
            if (depth > 40) return lsb_fail_arg("tracks_current_mc: workspace split too deep");
            McParams a = p, b = p;
            a.S = p.S / 2; b.S = p.S - a.S; b.seg0 = p.seg0 + a.S;
            rc = mc_run_range(L, tracks, pixels, signals, response, response_split, f64, rng, mode, a, ws, ws_bytes, st, depth + 1);
            if (rc) return rc;
            return mc_run_range(L, tracks, pixels, signals, response, response_split, f64, rng, mode, b, ws, ws_bytes, st, depth + 1);
        }
        g_mc_last_samples += total;
        if (total == 0) return 0;
    }
    k_mc_set_offsets<<<lsb_blocks(npair, 256), 256, 0, st>>>(p, w.pairs, w.offs, npair);
    LSB_LAUNCH_CHECK("k_mc_set_offsets");
    LSB_CUDA(cudaMemsetAsync(w.bucket, 0, 2 * MC_NBUCKET * sizeof(int), st));
    k_mc_bucket_count<<<lsb_blocks(npair, 256), 256, 0, st>>>(w.nsamp, npair, w.bucket);
    LSB_LAUNCH_CHECK("k_mc_bucket_count");
    k_mc_bucket_scan<<<1, 32, 0, st>>>(w.bucket);
    LSB_LAUNCH_CHECK("k_mc_bucket_scan");
    k_mc_bucket_scatter<<<lsb_blocks(npair, 256), 256, 0, st>>>(w.nsamp, npair, w.bucket, w.perm);
    LSB_LAUNCH_CHECK("k_mc_bucket_scatter");
    k_mc_uniforms<<<lsb_blocks(npair, UNI_TPB), UNI_TPB, 0, st>>>(p, w.pairs, w.perm, w.uu, rng);
    LSB_LAUNCH_CHECK("k_mc_uniforms");
    p.lohi = w.lohi;
    k_mc_sampler<<<lsb_blocks(npair, SMP_WARPS), 32 * SMP_WARPS, 0, st>>>(p, w.pairs, w.uu, w.samples, w.offs32);
    LSB_LAUNCH_CHECK("k_mc_sampler");
    if (f64) return mc_launch_accumulate<double>(p, w, (const double*)response, nullptr, signals, st);
    return mc_launch_accumulate<float>(p, w, (const float*)response, (const float*)response_split, signals, st);
}

LSB_EXPORT int64_t lsb_tracks_current_mc_last_samples(void) { return g_mc_last_samples; }

static inline int require_current_fields(const lsb_track_layout* L, bool mc) {
    static const int need[] = {LSB_F_X_START, LSB_F_Y_START, LSB_F_Z_START, LSB_F_X_END, LSB_F_Y_END, LSB_F_Z_END,
                               LSB_F_T_START, LSB_F_T0_START, LSB_F_TRAN_DIFF, LSB_F_LONG_DIFF, LSB_F_N_ELECTRONS,
                               LSB_F_PIXEL_PLANE};
    for (int f : need) if (!layout_has(L, f)) return lsb_fail_arg("tracks lacks a field required by tracks_current");
    (void)mc;
    return 0;
}

LSB_EXPORT int lsb_tracks_current_mc(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t S,
                                     const int32_t* pixels, int32_t P, float* signals, int32_t T, const void* response,
                                     int32_t Rx, int32_t Ry, int32_t Rt, int32_t response_f64, uint64_t* rng_states,
                                     int64_t n_rng, int64_t rng_stride, int32_t rng_mode, void* workspace,
                                     int64_t workspace_bytes, void* stream) {
    LSB_REQUIRE(c && L, "tracks_current_mc: null consts/layout");
    if (S == 0 || P == 0 || T == 0) return 0;
    LSB_REQUIRE(tracks && pixels && signals && response && rng_states && workspace, "tracks_current_mc: null pointer");
    LSB_REQUIRE(Rx > 0 && Ry > 0 && Rt > 0 && (long long)Rx * Ry * Rt < 2147483647LL, "tracks_current_mc: bad response shape");
    LSB_REQUIRE(rng_mode == 0 || rng_mode == 1, "tracks_current_mc: rng_mode must be 0 (cloud) or 1 (replay)");
    LSB_REQUIRE(rng_mode == 1 || T < 65536, "tracks_current_mc: more than 65535 ticks per waveform");
    if (rng_stride <= 0) rng_stride = S;
    LSB_REQUIRE(n_rng >= (S - 1) + rng_stride * (long long)(P - 1) + 1, "tracks_current_mc: rng_states too short");
    if (require_current_fields(L, true)) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = lsb_upload_consts(c, st); if (rc) return rc;
    McParams p;
    p.S = S; p.seg0 = 0; p.rng_stride = rng_stride; p.P = P; p.T = T; p.Rx = Rx; p.Ry = Ry; p.Rt = Rt;
    double ratio = c->time_sampling / c->response_sampling;
    p.stride = (ratio == 1.0) ? 1 : (ratio == 2.0 ? 2 : 0);
    g_mc_last_samples = 0;
    p.total_dev = nullptr; p.cap = 0; p.overflow = nullptr; p.ranges = nullptr; p.nfma = nullptr; p.npairs = nullptr; p.diag = nullptr;
    p.split = 0; p.split_len = 0;
    TmpPool pool(st);
    float* split = nullptr;
    if (rng_mode == 0 && mc_wants_split(c, response_f64)) {          // 31.6 MB re-laid out per call (~20 us); the chain keeps its copy
        LSB_CUDA(pool.get(&split, (long long)Rx * Ry * Rt));
        if ((rc = mc_build_split(response, Rx, Ry, Rt, split, st))) return rc;
        p.split = 2; p.split_len = (Rt + 1) >> 1;
    }
    return mc_run_range(make_layout(L), tracks, pixels, signals, response, split, response_f64, (unsigned long long*)rng_states,
                        rng_mode, p, workspace, workspace_bytes, st, 0);
}

// Fused-chain entry without host synchronisation: `workspace` must hold `lsb_tracks_current_mc_workspace_bytes`
// of an UPPER BOUND of the sample count; *total_out (device) receives the actual count, *overflow (device)
// is raised -- and nothing is written -- if the bound was wrong.  Cloud mode only.  `signals` is stored sparsely: row e
// holds data in ticks ranges[e].x .. ranges[e].y only (see McParams::ranges); the caller does not pre-fill it.
static int mc_run_nosync(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t S, const int32_t* pixels,
                         int32_t P, float* signals, int32_t T, const void* response, const void* response_split, int32_t Rx, int32_t Ry, int32_t Rt,
                         int32_t response_f64, uint64_t* rng_states, int64_t rng_stride, void* workspace,
                         int64_t workspace_bytes, long long* total_out, int* overflow, int2* ranges, unsigned long long* nfma,
                         unsigned long long* npairs, unsigned long long* diag, cudaStream_t st) {
    if (S == 0 || P == 0 || T == 0) return 0;
    if (require_current_fields(L, true)) return -1;
    LSB_REQUIRE(T < 65536, "tracks_current_mc: more than 65535 ticks per waveform");
    int rc = lsb_upload_consts(c, st); if (rc) return rc;
    McParams p;
    p.S = S; p.seg0 = 0; p.rng_stride = rng_stride; p.P = P; p.T = T; p.Rx = Rx; p.Ry = Ry; p.Rt = Rt;
    double ratio = c->time_sampling / c->response_sampling;
    p.stride = (ratio == 1.0) ? 1 : (ratio == 2.0 ? 2 : 0);
    p.total_dev = nullptr; p.cap = 0; p.overflow = overflow; p.ranges = ranges; p.nfma = nfma; p.npairs = npairs; p.diag = diag;
    p.split = response_split ? 2 : 0; p.split_len = (Rt + 1) >> 1;
    rc = mc_run_range(make_layout(L), tracks, pixels, signals, response, response_split, response_f64, (unsigned long long*)rng_states, 0, p,
                      workspace, workspace_bytes, st, 0);
    if (rc) return rc;
    McWs w;
    mc_carve(workspace, workspace_bytes, S * (long long)P, w);
    LSB_CUDA(cudaMemcpyAsync(total_out, w.total, 8, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// =======================================================================================
// tracks_current (detsim.py:351-453): deterministic Gaussian-smeared line charge on a
// SAMPLED_POINTS^2 x z_steps grid.  CTA per (segment,pixel); the charge grid of one z slab is
// evaluated once per CTA (rho: detsim.py:120-159) and shared by all ticks through shared memory;
// accumulation order (iz, ix, iy) and FP64 accumulation follow the reference.
// =======================================================================================
#define TC_TPB 256
#define TC_RMAX 8
#define TC_MAXNP 64
#define TC_MAXBIN 40          // response bins per axis the binned path handles (the grid of one pair spans < 2 cm)
#ifndef TC_BINNED
#define TC_BINNED 1
#endif

struct TcPair {
    double start[3], seg[3], dir[3], sig[3];
    double x_p, y_p, q, z_anode, t_start;
    double x_start, y_start, x_step, y_step, z_start_int, z_step;
    long long z_steps;
    int valid, c32, s32;
    // point-independent parts of rho (same expressions as rho_dev, evaluated once per pair)
    double r_ux, r_uy, r_uz, r_d0, r_d1, r_d2, r_e0, r_e1, r_e2, r_sqrt_a_2, r_two_a_Dr, r_four_a, r_factor, r_log_factor;
};

// rho with the pair's invariants taken from TcPair (bit-identical to rho_dev)
__device__ __forceinline__ void rho_prepare(TcPair& g) {
    const bool c32 = g.c32, s32 = g.s32;
    const double* seg = g.seg; const double* sig = g.sig;
    double Dr = seg_length(seg, c32);
    g.r_ux = R32(seg[0] / Dr, c32); g.r_uy = R32(seg[1] / Dr, c32); g.r_uz = R32(seg[2] / Dr, c32);
    double a = R32(g.r_ux * g.r_ux, c32) / (2 * sig[0] * sig[0]) + R32(g.r_uy * g.r_uy, c32) / (2 * sig[1] * sig[1]) +
               R32(g.r_uz * g.r_uz, c32) / (2 * sig[2] * sig[2]);
    double sprod = R32(R32(sig[0] * sig[1], s32) * sig[2], s32);
    g.r_factor = g.q / Dr / (sprod * sqrt(8 * M_PI * M_PI * M_PI));
    g.r_log_factor = log(g.r_factor);
    g.r_sqrt_a_2 = 2 * sqrt(a);
    g.r_d0 = R32(sig[0] * sig[0], s32); g.r_d1 = R32(sig[1] * sig[1], s32); g.r_d2 = R32(sig[2] * sig[2], s32);
    g.r_e0 = 2 * sig[0] * sig[0]; g.r_e1 = 2 * sig[1] * sig[1]; g.r_e2 = 2 * sig[2] * sig[2];
    g.r_two_a_Dr = 2 * a * Dr;
    g.r_four_a = 4 * a;
}
// erf(hi) - erf(lo), lo <= hi (detsim.py:149-151 writes -erf(lo) + erf(hi)): taken between the complementary functions when both
// arguments lie on the same side of zero, where the reference's form cancels (grid points beyond the ends of the segment) and a
// one-ulp difference between two erf implementations would change the waveform at the 1e-5..1 level; identical to rounding wherever
// the reference's form is well-conditioned.  The CPU checker the tests compare with evaluates the difference the same way.
__device__ __forceinline__ double erf_diff(double lo, double hi) {
    if (lo >= 0.0) return erfc(lo) - erfc(hi);
    if (hi <= 0.0) return erfc(-hi) - erfc(-lo);
    return erf(hi) - erf(lo);
}
__device__ __forceinline__ double rho_fast(double x, double y, double z, const TcPair& g) {
    const double dx = x - g.start[0], dy = y - g.start[1], dz = z - g.start[2];
    double b = -(dx / g.r_d0 * g.r_ux + dy / g.r_d1 * g.r_uy + dz / g.r_d2 * g.r_uz);
    double delta = dx * dx / g.r_e0 + dy * dy / g.r_e1 + dz * dz / g.r_e2;
    double integral = sqrt(M_PI) * erf_diff(b / g.r_sqrt_a_2, (b + g.r_two_a_Dr) / g.r_sqrt_a_2) / g.r_sqrt_a_2;
    double expo = 0;
    if (g.r_factor != 0 && integral != 0) expo = exp(b * b / g.r_four_a - delta + g.r_log_factor + log(integral));
    return expo;
}

__device__ double rho_dev(const double* pt, const TcPair& g) {
    const bool c32 = g.c32, s32 = g.s32;
    const double* seg = g.seg; const double* sig = g.sig; const double* start = g.start;
    double Dr = seg_length(seg, c32);
    double ux = R32(seg[0] / Dr, c32), uy = R32(seg[1] / Dr, c32), uz = R32(seg[2] / Dr, c32);
    double a = R32(ux * ux, c32) / (2 * sig[0] * sig[0]) + R32(uy * uy, c32) / (2 * sig[1] * sig[1]) +
               R32(uz * uz, c32) / (2 * sig[2] * sig[2]);
    double sprod = R32(R32(sig[0] * sig[1], s32) * sig[2], s32);
    double factor = g.q / Dr / (sprod * sqrt(8 * M_PI * M_PI * M_PI));
    double sqrt_a_2 = 2 * sqrt(a);
    double x = pt[0], y = pt[1], z = pt[2];
    double b = -((x - start[0]) / R32(sig[0] * sig[0], s32) * ux + (y - start[1]) / R32(sig[1] * sig[1], s32) * uy +
                 (z - start[2]) / R32(sig[2] * sig[2], s32) * uz);
    double delta = (x - start[0]) * (x - start[0]) / (2 * sig[0] * sig[0]) + (y - start[1]) * (y - start[1]) / (2 * sig[1] * sig[1]) +
                   (z - start[2]) * (z - start[2]) / (2 * sig[2] * sig[2]);
    double integral = sqrt(M_PI) * erf_diff(b / sqrt_a_2, (b + 2 * a * Dr) / sqrt_a_2) / sqrt_a_2;
    double expo = 0;
    if (factor != 0 && integral != 0) expo = exp(b * b / (4 * a) - delta + log(factor) + log(integral));
    return expo;
}

// z_interval (detsim.py:42-112)
__device__ void z_interval_dev(const double* sp, const double* ep, double x_p, double y_p, double tol, bool c32,
                               double& z_poca, double& z_lo, double& z_hi) {
    z_poca = z_lo = z_hi = 0;
    const double *start, *end;
    if (sp[0] > ep[0]) { start = ep; end = sp; } else if (sp[0] < ep[0]) { start = sp; end = ep; } else return;
    double xs = start[0], ys = start[1], xe = end[0], ye = end[1];
    double dxe = R32(xe - xs, c32);
    double m = R32(R32(ye - ys, c32) / dxe, c32);
    double q = R32(R32(R32(xe * ys, c32) - R32(xs * ye, c32), c32) / dxe, c32);
    double a = m, b = -1, cc = q;
    double x_poca = (b * (b * x_p - a * y_p) - R32(a * cc, c32)) / (R32(a * a, c32) + b * b);
    double d[3] = {R32(end[0] - start[0], c32), R32(end[1] - start[1], c32), R32(end[2] - start[2], c32)};
    double length = seg_length(d, c32);
    double dir3[3] = {R32(d[0] / length, c32), R32(d[1] / length, c32), R32(d[2] / length, c32)};
    double doca;
    if (x_poca < start[0]) {
        doca = sqrt((x_p - start[0]) * (x_p - start[0]) + (y_p - start[1]) * (y_p - start[1]));
        x_poca = start[0];
    } else if (x_poca > end[0]) {
        doca = sqrt((x_p - end[0]) * (x_p - end[0]) + (y_p - end[1]) * (y_p - end[1]));
        x_poca = end[0];
    } else {
        doca = fabs(a * x_p + b * y_p + cc) / sqrt(R32(a * a, c32) + b * b);
    }
    double zp = start[2] + (x_poca - start[0]) / dir3[0] * dir3[2];
    if (tol > doca) {
        double dx2 = R32(xe - xs, c32), dy2 = R32(ye - ys, c32);
        double length2D = R32(sqrt(R32(R32(dx2 * dx2, c32) + R32(dy2 * dy2, c32), c32)), c32);
        double dir2D0 = R32(d[0] / length2D, c32);
        double deltaL2D = sqrt(tol * tol - doca * doca);
        double x_plus = x_poca + deltaL2D * dir2D0, x_minus = x_poca - deltaL2D * dir2D0;
        double plusL = (x_plus - start[0]) / dir3[0], minusL = (x_minus - start[0]) / dir3[0];
        double plusZ = start[2] + dir3[2] * plusL, minusZ = start[2] + dir3[2] * minusL;
        z_poca = zp; z_lo = fmin(minusZ, plusZ); z_hi = fmax(minusZ, plusZ);
    }
}

template <typename TL>
__global__ void __launch_bounds__(TC_TPB) k_tracks_current(Layout L, const char* __restrict__ tracks,
                                                           const int32_t* __restrict__ pixels, long long S, int P, int T,
                                                           const TL* __restrict__ lut, int Rx, int Ry, int Rt,
                                                           float* __restrict__ signals) {
    __shared__ TcPair g;
    __shared__ double s_charge[TC_MAXNP * TC_MAXNP];
    __shared__ double s_bin[TC_MAXBIN * TC_MAXBIN];
    __shared__ int s_xoff[TC_MAXBIN + 1], s_yoff[TC_MAXBIN + 1], s_xmem[TC_MAXNP], s_ymem[TC_MAXNP];
    __shared__ int s_i[TC_MAXNP], s_j[TC_MAXNP];
    __shared__ int s_xok[TC_MAXNP], s_yok[TC_MAXNP];
    __shared__ int s_anyx;
    const long long pr = blockIdx.x;
    const long long itrk = pr / P;
    const int ipix = (int)(pr % P);
    const int tid = threadIdx.x;
    const int NP = d_c.sampled_points;
    if (tid == 0) {
        g.valid = 0;
        const char* t = tracks + itrk * L.itemsize;
        int pID = pixels[itrk * P + ipix];
        bool c32;
        do {
            if (!pixel_center(pID, g.x_p, g.y_p)) break;
            double end[3];
            load_endpoints(L, t, g.start, end, c32);
            g.c32 = c32;
            g.s32 = fld_f32(L, LSB_F_TRAN_DIFF) && fld_f32(L, LSB_F_LONG_DIFF);
            for (int k = 0; k < 3; k++) g.seg[k] = R32(end[k] - g.start[k], c32);
            double length = seg_length(g.seg, c32);
            for (int k = 0; k < 3; k++) g.dir[k] = R32(g.seg[k] / length, c32);
            g.sig[0] = g.sig[1] = fld_get(L, t, LSB_F_TRAN_DIFF);
            g.sig[2] = fld_get(L, t, LSB_F_LONG_DIFF);
            double s5x = 5 * g.sig[0], s5y = 5 * g.sig[1];
            double imp1 = sqrt(s5x * s5x + s5y * s5y);
            double imp2 = sqrt(d_c.pixel_pitch * d_c.pixel_pitch + d_c.pixel_pitch * d_c.pixel_pitch) / 2;
            double impact = fmax(imp1, imp2) * 2;
            double z_poca, z_start, z_end;
            z_interval_dev(g.start, end, g.x_p, g.y_p, impact, c32, z_poca, z_start, z_end);
            if (z_poca == 0) break;
            g.z_start_int = z_start - 4 * g.sig[2];
            double z_end_int = z_end + 4 * g.sig[2];
            double l0 = (z_start - g.start[2]) / g.dir[2], l1 = (z_end - g.start[2]) / g.dir[2];
            g.x_start = g.start[0] + l0 * g.dir[0]; g.y_start = g.start[1] + l0 * g.dir[1];
            double x_end = g.start[0] + l1 * g.dir[0], y_end = g.start[1] + l1 * g.dir[1];
            g.y_step = (fabs(y_end - g.y_start) + 8 * g.sig[1]) / (NP - 1);
            g.x_step = (fabs(x_end - g.x_start) + 8 * g.sig[0]) / (NP - 1);
            double z_sampling = d_c.time_sampling / 2.;
            long long zc = (long long)ceil(fabs(z_end_int - g.z_start_int) / z_sampling);
            g.z_steps = zc > NP ? zc : NP;
            g.z_step = (z_end_int - g.z_start_int) / (double)(g.z_steps - 1);
            g.t_start = (double)__double2ll_rn((fld_get(L, t, LSB_F_T_START) - fld_get(L, t, LSB_F_T0_START) - d_c.time_padding) /
                                               d_c.time_sampling) * d_c.time_sampling;
            g.q = fld_get(L, t, LSB_F_N_ELECTRONS);
            long long plane = (long long)fld_get(L, t, LSB_F_PIXEL_PLANE);
            if (plane < 0 || plane >= d_c.n_tpc) break;
            g.z_anode = d_c.tpc_borders[plane][2][0];
            rho_prepare(g);
            g.valid = 1;
        } while (0);
        s_anyx = 0;
    }
    __syncthreads();
    if (!g.valid) return;
    const double sgx = g.dir[0] >= 0 ? 1.0 : -1.0, sgy = g.dir[1] >= 0 ? 1.0 : -1.0;
    if (tid < NP) {
        double x = g.x_start + sgx * (tid * g.x_step - 4 * g.sig[0]);
        double x_dist = fabs(g.x_p - x);
        int ok = !(x_dist > d_c.response_bin_size * Rx);
        long long i = __double2ll_rn(x_dist / d_c.response_bin_size - 0.5);
        s_xok[tid] = ok; s_i[tid] = (ok && 0 <= i && i < Rx) ? (int)i : -1;
        if (ok) s_anyx = 1;
        double y = g.y_start + sgy * (tid * g.y_step - 4 * g.sig[1]);
        double y_dist = fabs(g.y_p - y);
        int oky = !(y_dist > d_c.response_bin_size * Ry);
        long long j = __double2ll_rn(y_dist / d_c.response_bin_size - 0.5);
        s_yok[tid] = oky; s_j[tid] = (oky && 0 <= j && j < Ry) ? (int)j : -1;
    }
    __syncthreads();
    const int anyx = s_anyx;
    // response bins touched by the grid: [imin, imin + nbi) x [jmin, jmin + nbj); the binned path needs them to fit s_bin
    int imin = 1 << 30, imax = -1, jmin = 1 << 30, jmax = -1;
    for (int q = 0; q < NP; q++) {
        if (s_i[q] >= 0) { imin = s_i[q] < imin ? s_i[q] : imin; imax = s_i[q] > imax ? s_i[q] : imax; }
        if (s_j[q] >= 0) { jmin = s_j[q] < jmin ? s_j[q] : jmin; jmax = s_j[q] > jmax ? s_j[q] : jmax; }
    }
    const int nbi = imax >= imin ? imax - imin + 1 : 0, nbj = jmax >= jmin ? jmax - jmin + 1 : 0;
    const bool binned = TC_BINNED && nbi <= TC_MAXBIN && nbj <= TC_MAXBIN;
    if (binned && tid < 2) {
        // grid indices grouped by response bin (counting sort, once per pair; bins do not depend on the z slab)
        const int* key = tid == 0 ? s_i : s_j;
        const int kmin = tid == 0 ? imin : jmin, nb = tid == 0 ? nbi : nbj;
        int* off = tid == 0 ? s_xoff : s_yoff;
        int* mem = tid == 0 ? s_xmem : s_ymem;
        int run = 0;
        for (int b = 0; b < nb; b++) {
            off[b] = run;
            for (int q = 0; q < NP; q++) if (key[q] == kmin + b) mem[run++] = q;
        }
        off[nb] = run;
    }
    __syncthreads();
    float* out = signals + (itrk * P + ipix) * (long long)T;
    const double vol = fabs(g.x_step) * fabs(g.y_step) * fabs(g.z_step);
    for (int base = 0; base < T; base += TC_TPB * TC_RMAX) {
        double total[TC_RMAX];
        bool written[TC_RMAX];
#pragma unroll
        for (int r = 0; r < TC_RMAX; r++) { total[r] = 0; written[r] = false; }
        for (long long iz = 0; iz < g.z_steps; iz++) {
            double z = g.z_start_int + (double)iz * g.z_step;
            double t0 = fabs(z - g.z_anode) / d_c.v_drift - d_c.time_window;
            __syncthreads();
            for (int c = tid; c < NP * NP; c += TC_TPB) {
                int ix = c / NP, iy = c % NP;
                double ch = 0;
                if (s_xok[ix] && s_yok[iy]) {
                    double pt[3] = {g.x_start + sgx * (ix * g.x_step - 4 * g.sig[0]),
                                    g.y_start + sgy * (iy * g.y_step - 4 * g.sig[1]), z};
                    ch = rho_fast(pt[0], pt[1], pt[2], g) * fabs(g.x_step) * fabs(g.y_step) * fabs(g.z_step);
                }
                s_charge[c] = ch;
            }
            __syncthreads();
            if (binned) {
                // charge per response bin, summed in the reference's (ix, iy) order within the bin
                for (int b = tid; b < nbi * nbj; b += TC_TPB) {
                    const int bi = b / nbj, bj = b % nbj;
                    double sum = 0.0;
                    for (int a = s_xoff[bi]; a < s_xoff[bi + 1]; a++) {       // members of the bin, ascending ix / iy
                        const int ix = s_xmem[a];
                        for (int c2 = s_yoff[bj]; c2 < s_yoff[bj + 1]; c2++) sum += s_charge[ix * NP + s_ymem[c2]];
                    }
                    s_bin[bi * TC_MAXBIN + bj] = sum;
                }
                __syncthreads();
            }
            // per tick of this thread: inside the slab's window?  table index?
            long long kr[TC_RMAX];
            bool use[TC_RMAX];
            bool any_use = false;
#pragma unroll
            for (int r = 0; r < TC_RMAX; r++) {
                use[r] = false; kr[r] = 0;
                int it = base + tid + r * TC_TPB;
                if (it >= T) continue;
                double tick = g.t_start + (double)it * d_c.time_sampling;
                if (tick < 0.) continue;
                if (!(t0 < tick && tick < t0 + d_c.time_window)) continue;
                if (anyx) written[r] = true;
                long long k = __double2ll_rn((tick - t0) / d_c.response_sampling);
                kr[r] = k;
                use[r] = 0 <= k && k < Rt;
                any_use |= use[r];
            }
            if (binned) {
                // the grid points of the slab fall into a few response bins: one table read per bin and tick, weighted with the
                // bin's summed charge (the reference reads the table once per grid point); the bin loop is outermost so one
                // row pointer serves the thread's TC_RMAX ticks
                if (any_use)
                    for (int bi = 0; bi < nbi; bi++) {
                        const TL* prow = lut + ((long long)(imin + bi) * Ry + jmin) * Rt;
                        for (int bj = 0; bj < nbj; bj++, prow += Rt) {
                            const double B = s_bin[bi * TC_MAXBIN + bj];
#pragma unroll
                            for (int r = 0; r < TC_RMAX; r++)
                                if (use[r]) total[r] += (double)prow[kr[r]] * B;
                        }
                    }
            } else {
#pragma unroll
                for (int r = 0; r < TC_RMAX; r++) {
                    int it = base + tid + r * TC_TPB;
                    if (it >= T) continue;
                    double tick = g.t_start + (double)it * d_c.time_sampling;
                    if (tick < 0.) continue;
                    if (!(t0 < tick && tick < t0 + d_c.time_window)) continue;
                    const long long k = kr[r];
                    const bool kok = use[r];
                    double tot = total[r];
                    for (int ix = 0; ix < NP; ix++) {
                        if (!s_xok[ix]) continue;
                        int i = s_i[ix];
                        for (int iy = 0; iy < NP; iy++) {
                            if (!s_yok[iy]) continue;
                            int j = s_j[iy];
                            double w = 0;
                            if (kok && i >= 0 && j >= 0) w = (double)lut[((long long)i * Ry + j) * Rt + k];
                            tot += w * s_charge[ix * NP + iy];
                        }
                    }
                    total[r] = tot;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < TC_RMAX; r++) {
            int it = base + tid + r * TC_TPB;
            if (it < T && written[r]) out[it] = __double2float_rn(total[r]);
        }
    }
    (void)vol;
}

LSB_EXPORT int lsb_tracks_current(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t S,
                                  const int32_t* pixels, int32_t P, float* signals, int32_t T, const void* response,
                                  int32_t Rx, int32_t Ry, int32_t Rt, int32_t response_f64, void* stream) {
    LSB_REQUIRE(c && L, "tracks_current: null consts/layout");
    if (S == 0 || P == 0 || T == 0) return 0;
    LSB_REQUIRE(tracks && pixels && signals && response, "tracks_current: null pointer");
    LSB_REQUIRE(c->sampled_points >= 2 && c->sampled_points <= TC_MAXNP, "tracks_current: SAMPLED_POINTS must be in [2,64]");
    if (require_current_fields(L, false)) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = lsb_upload_consts(c, st); if (rc) return rc;
    unsigned grid = (unsigned)(S * P);
    if (response_f64) k_tracks_current<double><<<grid, TC_TPB, 0, st>>>(make_layout(L), (const char*)tracks, pixels, S, P, T,
                                                                         (const double*)response, Rx, Ry, Rt, signals);
    else k_tracks_current<float><<<grid, TC_TPB, 0, st>>>(make_layout(L), (const char*)tracks, pixels, S, P, T,
                                                          (const float*)response, Rx, Ry, Rt, signals);
    LSB_LAUNCH_CHECK("k_tracks_current");
    return 0;
}

// zero everything outside the stored tick range of every row (sparse -> dense `signals`)
__global__ void k_signals_dense(float* __restrict__ signals, const int2* __restrict__ ranges, long long n_rows, int T) {
    const long long row = blockIdx.x;
    if (row >= n_rows) return;
    const int2 g = ranges[row];
    float* o = signals + row * (long long)T;
    for (int it = threadIdx.x; it < T; it += blockDim.x)
        if (it < g.x || it > g.y) o[it] = 0.f;
}
// full ranges (dense producer, e.g. replay mode)
__global__ void k_ranges_full(int2* __restrict__ ranges, long long n_rows, int T) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n_rows) ranges[i] = make_int2(0, T - 1);
}
