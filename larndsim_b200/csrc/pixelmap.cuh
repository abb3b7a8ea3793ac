// pixelmap.cuh -- per-pixel reduction of the per-segment induced currents:
//   get_track_pixel_map  (detsim.py:529-562), get_track_pixel_map2 (detsim.py:564-607),
//   sum_pixel_signals    (detsim.py:468-527).
//
// The reference scans all S x P (segment,pixel) entries once per unique pixel (and per distance
// class).  Here every entry is bucketed by its unique-pixel index (count -> scan -> fill), and
// the order the reference produces is recovered by RANKING the entries of a bucket on an unique
// key, so the result does not depend on the order atomics filled the bucket:
//   map   : key = segment index                   (first K distinct segments)
//   map2  : key = (distance class, segment index) (classes 0..max_distance-1 only)
//   sum   : key = flat entry index itrk*P+ipix    (contributions added in ascending segment order:
//           reproducible sums; the reference's float64 atomics leave the order undefined)
// All of it is integer / HBM-bound work.
#pragma once
#include "common.cuh"
#include "glue.cuh"

struct PmEntry { int e; int key_hi; };   // flat entry index, distance class (map2) / slot (sum)

// bucket index of every (segment,pixel) entry: position of the pixel id in the sorted unique list,
// -1 if absent.  shadow_dups: drop an entry when an earlier column of the same row holds the same id
// (map2 stops at the first match of a row, detsim.py:607; map re-inserts the same segment: no-op).
__global__ void k_pm_bucket(const int32_t* __restrict__ pixels, long long n_entries, int P, const int32_t* __restrict__ uniq,
                            long long U, int* __restrict__ bucket, int* __restrict__ counts) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    int32_t p = pixels[e];
    long long lo = 0, hi = U;
    while (lo < hi) { long long mid = (lo + hi) >> 1; if (uniq[mid] < p) lo = mid + 1; else hi = mid; }
    int b = (lo < U && uniq[lo] == p) ? (int)lo : -1;
    if (b >= 0) {
        long long row0 = e - (e % P);
        for (long long q = row0; q < e; q++) if (pixels[q] == p) { b = -1; break; }
    }
    bucket[e] = b;
    if (b >= 0) atomicAdd(counts + b, 1);
}
__global__ void k_pm_fill(const int* __restrict__ bucket, const int32_t* __restrict__ key_hi, long long n_entries,
                          const long long* __restrict__ offs, int* __restrict__ cursor, PmEntry* __restrict__ entries) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    int b = bucket[e];
    if (b < 0) return;
    int slot = atomicAdd(cursor + b, 1);
    PmEntry r; r.e = (int)e; r.key_hi = key_hi ? key_hi[e] : 0;
    entries[offs[b] + slot] = r;
}
// rank inside the bucket and write the segment index at that rank (if < K)
__global__ void k_pm_rank_write(const int* __restrict__ bucket, const int32_t* __restrict__ key_hi, long long n_entries, int P,
                                const long long* __restrict__ offs, const int* __restrict__ counts,
                                const PmEntry* __restrict__ entries, int max_distance, int use_dist,
                                long long* __restrict__ tpm, int K) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    int b = bucket[e];
    if (b < 0) return;
    int d = use_dist ? key_hi[e] : 0;
    if (use_dist && (d < 0 || d >= max_distance)) return;      // never visited by `for target_dist in range(max_distance)`
    long long itrk = e / P;
    const PmEntry* L = entries + offs[b];
    int n = counts[b], rank = 0;
    for (int i = 0; i < n; i++) {
        PmEntry o = L[i];
        if (o.e == (int)e) continue;
        if (use_dist) {
            if (o.key_hi < 0 || o.key_hi >= max_distance) continue;
            long long ot = o.e / P;
            if (o.key_hi < d || (o.key_hi == d && ot < itrk)) rank++;
        } else {
            if (o.e / P < itrk) rank++;
        }
    }
    if (rank < K) tpm[(long long)b * K + rank] = itrk;
}
__global__ void k_check_sorted(const int32_t* __restrict__ uniq, long long U, int* __restrict__ flag) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i + 1 < U && !(uniq[i] < uniq[i + 1])) *flag = 1;
}
// literal fallback (unsorted / repeated unique_pix): thread per unique pixel, reference loops
__global__ void k_tpm_bruteforce(long long* __restrict__ tpm, int K, const int32_t* __restrict__ uniq, long long U,
                                 const int32_t* __restrict__ pixels, const int32_t* __restrict__ dist, long long S, int P,
                                 int max_distance, int v2) {
    long long index = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (index >= U) return;
    int32_t upix = uniq[index];
    long long* row = tpm + index * K;
    if (!v2) {
        for (long long itrk = 0; itrk < S; itrk++)
            for (int ipix = 0; ipix < P; ipix++) {
                if (upix != pixels[itrk * P + ipix]) continue;
                int imap = 0;
                while (imap < K && row[imap] != -1 && row[imap] != itrk) imap++;
                if (imap < K) row[imap] = itrk;
            }
        return;
    }
    for (int target = 0; target < max_distance; target++)
        for (long long itrk = 0; itrk < S; itrk++)
            for (int ipix = 0; ipix < P; ipix++) {
                if (upix != pixels[itrk * P + ipix]) continue;
                if (dist[itrk * P + ipix] == target) {
                    int imap = 0;
                    while (imap < K) {
                        if (row[imap] == itrk) { imap = -1; break; }
                        if (row[imap] == -1) break;
                        imap++;
                    }
                    if (imap >= 0 && imap < K) row[imap] = itrk;
                }
                break;
            }
}

static int tpm_run(long long* tpm, int K, const int32_t* uniq, long long U, const int32_t* pixels, const int32_t* dist,
                   long long S, int P, int max_distance, int v2, cudaStream_t st) {
    if (U == 0 || S == 0 || P == 0 || K == 0) return 0;
    long long n_entries = S * P;
    LSB_REQUIRE(n_entries < 2147483647LL && U < 2147483647LL, "track_pixel_map: S*P and U must be < 2^31");
    TmpPool tp(st);
    int* flag; int* bucket; int* counts; int* cursor; long long* offs; long long* bsums; PmEntry* entries;
    LSB_CUDA(tp.get(&flag, 1));
    LSB_CUDA(cudaMemsetAsync(flag, 0, 4, st));
    k_check_sorted<<<lsb_blocks(U, 256), 256, 0, st>>>(uniq, U, flag);
    LSB_LAUNCH_CHECK("k_check_sorted");
    int h_flag = 0;
    LSB_CUDA(cudaMemcpyAsync(&h_flag, flag, 4, cudaMemcpyDeviceToHost, st));
    LSB_CUDA(cudaStreamSynchronize(st));
    if (h_flag) {
        k_tpm_bruteforce<<<lsb_blocks(U, 32), 32, 0, st>>>(tpm, K, uniq, U, pixels, dist, S, P, max_distance, v2);
        LSB_LAUNCH_CHECK("k_tpm_bruteforce");
        return 0;
    }
    LSB_CUDA(tp.get(&bucket, n_entries));
    LSB_CUDA(tp.get(&counts, U));
    LSB_CUDA(tp.get(&cursor, U));
    LSB_CUDA(tp.get(&offs, U));
    LSB_CUDA(tp.get(&bsums, scan_num_blocks(U) + 1));
    LSB_CUDA(tp.get(&entries, n_entries));
    LSB_CUDA(cudaMemsetAsync(counts, 0, U * 4, st));
    LSB_CUDA(cudaMemsetAsync(cursor, 0, U * 4, st));
    k_pm_bucket<<<lsb_blocks(n_entries, 256), 256, 0, st>>>(pixels, n_entries, P, uniq, U, bucket, counts);
    LSB_LAUNCH_CHECK("k_pm_bucket");
    int rc = exclusive_scan<int, long long>(counts, U, offs, bsums, nullptr, st);
    if (rc) return rc;
    k_pm_fill<<<lsb_blocks(n_entries, 256), 256, 0, st>>>(bucket, v2 ? dist : nullptr, n_entries, offs, cursor, entries);
    LSB_LAUNCH_CHECK("k_pm_fill");
    k_pm_rank_write<<<lsb_blocks(n_entries, 256), 256, 0, st>>>(bucket, v2 ? dist : nullptr, n_entries, P, offs, counts, entries,
                                                               max_distance, v2, tpm, K);
    LSB_LAUNCH_CHECK("k_pm_rank_write");
    return 0;
}

LSB_EXPORT int lsb_get_track_pixel_map(int64_t* track_pixel_map, int32_t K, const int32_t* unique_pix, int64_t U,
                                       const int32_t* pixels, int64_t S, int32_t P, void* stream) {
    LSB_REQUIRE((U == 0 || S == 0 || P == 0 || K == 0) || (track_pixel_map && unique_pix && pixels), "get_track_pixel_map: null pointer");
    return tpm_run((long long*)track_pixel_map, K, unique_pix, U, pixels, nullptr, S, P, 0, 0, (cudaStream_t)stream);
}
LSB_EXPORT int lsb_get_track_pixel_map2(int64_t* track_pixel_map, int32_t K, const int32_t* unique_pix, int64_t U,
                                        const int32_t* pixels, const int32_t* distances, int64_t S, int32_t P,
                                        int32_t max_distance, void* stream) {
    LSB_REQUIRE((U == 0 || S == 0 || P == 0 || K == 0) || (track_pixel_map && unique_pix && pixels && distances),
                "get_track_pixel_map2: null pointer");
    return tpm_run((long long*)track_pixel_map, K, unique_pix, U, pixels, distances, S, P, max_distance, 1, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------
// sum_pixel_signals
// ---------------------------------------------------------------------------------------
// one (segment, pixel) waveform feeding a pixel: flat row index, slot in track_pixel_map, the ticks [lo, hi] of the row that
// hold data (everything outside is zero and need not be stored), tick of the row's first sample in the pixel's time frame
struct SumEntry { int e; int slot; int lo; int hi; long long start_tick; };

// bucket = pixel_index_map value; slot = position of the segment in track_pixel_map[pixel] (first match)
__global__ void k_sum_bucket(const long long* __restrict__ pim, long long n_entries, int P, long long U,
                             const long long* __restrict__ tpm, int K, int* __restrict__ slot_of, int* __restrict__ counts,
                             double* __restrict__ overflow_flag) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    long long p = pim[e];
    int slot = -1;
    if (p >= 0 && p < U) {
        long long itrk = e / P;
        const long long* row = tpm + p * K;
        for (int k = 0; k < K; k++) if (row[k] == itrk) { slot = k; break; }
        if (slot < 0) overflow_flag[p] = 1.0;                 // detsim.py:526-527
        else atomicAdd(counts + p, 1);
    }
    slot_of[e] = slot;
}
__global__ void k_sum_fill(const long long* __restrict__ pim, const int* __restrict__ slot_of, long long n_entries,
                           const long long* __restrict__ offs, int* __restrict__ cursor, int* __restrict__ raw) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    if (slot_of[e] < 0) return;
    long long p = pim[e];
    int s = atomicAdd(cursor + p, 1);
    raw[offs[p] + s] = (int)e;
}
__global__ void k_sum_sort(const long long* __restrict__ pim, const int* __restrict__ slot_of, long long n_entries, int P,
                           const long long* __restrict__ offs, const int* __restrict__ counts, const int* __restrict__ raw,
                           const double* __restrict__ track_starts, const int2* __restrict__ ranges, int T,
                           SumEntry* __restrict__ sorted) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    int slot = slot_of[e];
    if (slot < 0) return;
    long long p = pim[e];
    const int* L = raw + offs[p];
    int n = counts[p], rank = 0;
    for (int i = 0; i < n; i++) if (L[i] < (int)e) rank++;
    SumEntry r; r.e = (int)e; r.slot = slot;
    if (ranges) { const int2 g = ranges[e]; r.lo = g.x; r.hi = g.y; } else { r.lo = 0; r.hi = T - 1; }
    r.start_tick = __double2ll_rn(track_starts[e / P] / d_c.time_sampling);      // detsim.py:504
    sorted[offs[p] + rank] = r;
}

#define SUM_TPB 256
#define SUM_CHUNK 64
#define SUM_R 4             // ticks per thread (t, t + TPB, ...): four independent loads per entry in flight -- with one tick per thread
                            // the kernel was latency-bound at 0.26 of the HBM peak (2048 threads x 4 B in flight per SM)
// FRESH: pixels_signals holds no earlier contributions (fused chain): it is written without being read or pre-zeroed
template <bool WITH_PTS, bool FRESH>
__global__ void __launch_bounds__(SUM_TPB) k_sum_pixel_signals(double* __restrict__ pixels_signals, long long U, int Tt,
                                                               const float* __restrict__ signals, int T,
                                                               const long long* __restrict__ offs, const int* __restrict__ counts,
                                                               const SumEntry* __restrict__ sorted, int K,
                                                               double* __restrict__ pts) {
    __shared__ SumEntry s_e[SUM_CHUNK];
    const long long p = blockIdx.x;
    const int n = counts[p];
    const int t0 = blockIdx.y * (SUM_TPB * SUM_R) + threadIdx.x;
    double* out = pixels_signals + p * Tt;
    if (n == 0) {
        if (FRESH) {
#pragma unroll
            for (int r = 0; r < SUM_R; r++) { const int t = t0 + r * SUM_TPB; if (t < Tt) out[t] = 0.0; }
        }
        return;
    }
    const SumEntry* L = sorted + offs[p];
    double acc[SUM_R];
#pragma unroll
    for (int r = 0; r < SUM_R; r++) { const int t = t0 + r * SUM_TPB; acc[r] = (t < Tt && !FRESH) ? out[t] : 0.0; }
    const int cta_lo = blockIdx.y * (SUM_TPB * SUM_R), cta_hi = cta_lo + SUM_TPB * SUM_R - 1;
    for (int c0 = 0; c0 < n; c0 += SUM_CHUNK) {
        int nc = n - c0 < SUM_CHUNK ? n - c0 : SUM_CHUNK;
        __syncthreads();
        if ((int)threadIdx.x < nc) s_e[threadIdx.x] = L[c0 + threadIdx.x];
        __syncthreads();
        for (int i = 0; i < nc; i++) {
            const SumEntry e = s_e[i];
            // ticks of this entry: start_tick + [lo, hi]; skip the entry for the whole CTA if it misses its tick block
            if (e.start_tick + e.hi < cta_lo || e.start_tick + e.lo > cta_hi) continue;
            const float* row = signals + (long long)e.e * T;
            float v[SUM_R];
#pragma unroll
            for (int r = 0; r < SUM_R; r++) {
                const long long itick = (long long)(t0 + r * SUM_TPB) - e.start_tick;
                v[r] = (itick >= e.lo && itick <= e.hi && t0 + r * SUM_TPB < Tt) ? __ldg(row + itick) : 0.f;
            }
#pragma unroll
            for (int r = 0; r < SUM_R; r++) {
                if (v[r] == 0.f) continue;                         // x + 0 == x: skipping is exact
                acc[r] += (double)v[r];
                if (WITH_PTS) pts[(p * Tt + t0 + r * SUM_TPB) * (long long)K + e.slot] += (double)v[r];
            }
        }
    }
#pragma unroll
    for (int r = 0; r < SUM_R; r++) { const int t = t0 + r * SUM_TPB; if (t < Tt) out[t] = acc[r]; }
}

// tick ranges of sparsely stored rows -> entries (thread per pixel)
__global__ void k_sum_apply_ranges(const long long* __restrict__ offs, const int* __restrict__ counts, long long U,
                                   const int2* __restrict__ ranges, SumEntry* __restrict__ sorted) {
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= U) return;
    SumEntry* L = sorted + offs[p];
    for (int i = 0; i < counts[p]; i++) { const int2 g = ranges[L[i].e]; L[i].lo = g.x; L[i].hi = g.y; }
}

// entries of every pixel, sorted by flat (segment,pixel) index, in caller-provided buffers
struct SumCtx {
    int* slot_of; int* counts; int* cursor; int* raw; long long* offs; long long* bsums; SumEntry* sorted;
};
static int sum_build_entries(const SumCtx& x, long long U, long long S, int P, const double* track_starts,
                             const long long* pim, const long long* tpm, int K, double* overflow_flag, const int2* ranges, int T,
                             cudaStream_t st) {
    long long n_entries = S * P;
    LSB_CUDA(cudaMemsetAsync(x.counts, 0, U * 4, st));
    LSB_CUDA(cudaMemsetAsync(x.cursor, 0, U * 4, st));
    k_sum_bucket<<<lsb_blocks(n_entries, 256), 256, 0, st>>>(pim, n_entries, P, U, tpm, K, x.slot_of, x.counts, overflow_flag);
    LSB_LAUNCH_CHECK("k_sum_bucket");
    int rc = exclusive_scan<int, long long>(x.counts, U, x.offs, x.bsums, nullptr, st);
    if (rc) return rc;
    k_sum_fill<<<lsb_blocks(n_entries, 256), 256, 0, st>>>(pim, x.slot_of, n_entries, x.offs, x.cursor, x.raw);
    LSB_LAUNCH_CHECK("k_sum_fill");
    k_sum_sort<<<lsb_blocks(n_entries, 256), 256, 0, st>>>(pim, x.slot_of, n_entries, P, x.offs, x.counts, x.raw, track_starts, ranges, T, x.sorted);
    LSB_LAUNCH_CHECK("k_sum_sort");
    return 0;
}
static int sum_run(const SumCtx& x, double* pixels_signals, long long U, int Tt, const float* signals, int T, int K, double* pts,
                   cudaStream_t st, bool fresh = false) {
    if (Tt <= 0) return 0;
    if (T <= 0) { if (fresh) LSB_CUDA(cudaMemsetAsync(pixels_signals, 0, (size_t)U * Tt * 8, st)); return 0; }
    dim3 grid((unsigned)U, (unsigned)((Tt + SUM_TPB * SUM_R - 1) / (SUM_TPB * SUM_R)));
    if (pts) k_sum_pixel_signals<true, false><<<grid, SUM_TPB, 0, st>>>(pixels_signals, U, Tt, signals, T, x.offs, x.counts, x.sorted, K, pts);
    else if (fresh) k_sum_pixel_signals<false, true><<<grid, SUM_TPB, 0, st>>>(pixels_signals, U, Tt, signals, T, x.offs, x.counts, x.sorted, K, nullptr);
    else k_sum_pixel_signals<false, false><<<grid, SUM_TPB, 0, st>>>(pixels_signals, U, Tt, signals, T, x.offs, x.counts, x.sorted, K, nullptr);
    LSB_LAUNCH_CHECK("k_sum_pixel_signals");
    return 0;
}

LSB_EXPORT int lsb_sum_pixel_signals(const lsb_consts* c, double* pixels_signals, int64_t U, int32_t Tt, const float* signals,
                                     int64_t S, int32_t P, int32_t T, const double* track_starts,
                                     const int64_t* pixel_index_map, const int64_t* track_pixel_map, int32_t K,
                                     double* pixels_tracks_signals, double* overflow_flag, void* stream) {
    LSB_REQUIRE(c, "sum_pixel_signals: null consts");
    if (U == 0 || S == 0 || P == 0) return 0;
    LSB_REQUIRE(pixels_signals && signals && track_starts && pixel_index_map && track_pixel_map && pixels_tracks_signals && overflow_flag,
                "sum_pixel_signals: null pointer");
    long long n_entries = S * P;
    LSB_REQUIRE(n_entries < 2147483647LL && U < 2147483647LL, "sum_pixel_signals: S*P and U must be < 2^31");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = lsb_upload_consts(c, st); if (rc) return rc;
    TmpPool tp(st);
    SumCtx x;
    LSB_CUDA(tp.get(&x.slot_of, n_entries));
    LSB_CUDA(tp.get(&x.counts, U));
    LSB_CUDA(tp.get(&x.cursor, U));
    LSB_CUDA(tp.get(&x.raw, n_entries));
    LSB_CUDA(tp.get(&x.offs, U));
    LSB_CUDA(tp.get(&x.bsums, scan_num_blocks(U) + 1));
    LSB_CUDA(tp.get(&x.sorted, n_entries));
    rc = sum_build_entries(x, U, S, P, track_starts, (const long long*)pixel_index_map, (const long long*)track_pixel_map, K,
                           overflow_flag, nullptr, T, st);
    if (rc) return rc;
    return sum_run(x, pixels_signals, U, Tt, signals, T, K, pixels_tracks_signals, st);
}
