// batching.cuh -- the callers that cut the segment array into simulation units
//   active_volume.select_active_volume   (larndsim/active_volume.py:4-46)
//   util.batching.TPCBatcher             (larndsim/util/batching.py:17-67)
// The reference walks events x TPC batches on the host and evaluates, for every batch, twelve comparisons over the
// WHOLE segment array (O(E*B*S) NumPy passes) to produce one boolean mask.  Here one pass classifies every segment
// (first TPC that contains its start or end point), a second pass turns (event, TPC batch) into a unit key, and a
// stable LSD radix sort of the keys yields every batch of the run at once: `order` lists the segment indices unit
// by unit, ascending inside a unit -- exactly the rows the reference's mask selects, in the same order.
// Integer / HBM work, bit-exact.
#pragma once
#include "common.cuh"
#include "glue.cuh"

// ---------------------------------------------------------------------------------------
// active volume
// ---------------------------------------------------------------------------------------
// first_tpc[i] = lowest TPC whose OPEN box contains the start or the end point of segment i, -1 if none.
// Comparisons are made in float64 (a float32 field promoted against the float64 border, NumPy >= 2 promotion).
__global__ void k_active_volume(Layout L, const char* __restrict__ tracks, long long n, const double* __restrict__ borders,
                                int tpc_lo, int tpc_hi, int32_t* __restrict__ first_tpc, uint32_t* __restrict__ flag) {
    extern __shared__ double s_b[];                     // [ntpc][3][2] sorted (lo, hi)
    const int nb = (tpc_hi - tpc_lo) * 3;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) {
        const double a = borders[(size_t)(tpc_lo * 3 + k) * 2], b = borders[(size_t)(tpc_lo * 3 + k) * 2 + 1];
        s_b[2 * k] = fmin(a, b);
        s_b[2 * k + 1] = fmax(a, b);
    }
    __syncthreads();
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const char* t = tracks + (size_t)i * L.itemsize;
    const double xs = fld_get(L, t, LSB_F_X_START), ys = fld_get(L, t, LSB_F_Y_START), zs = fld_get(L, t, LSB_F_Z_START);
    const double xe = fld_get(L, t, LSB_F_X_END), ye = fld_get(L, t, LSB_F_Y_END), ze = fld_get(L, t, LSB_F_Z_END);
    int found = -1;
    for (int k = 0; k < tpc_hi - tpc_lo; k++) {
        const double* b = s_b + 6 * k;
        const bool e = xe > b[0] && xe < b[1] && ye > b[2] && ye < b[3] && ze > b[4] && ze < b[5];
        const bool s = xs > b[0] && xs < b[1] && ys > b[2] && ys < b[3] && zs > b[4] && zs < b[5];
        if (e || s) { found = tpc_lo + k; break; }
    }
    first_tpc[i] = found;
    if (flag) flag[i] = found >= 0 ? 1u : 0u;
}

__global__ void k_emit_indices(const uint32_t* __restrict__ flag, const long long* __restrict__ pos, long long n,
                               long long* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n && flag[i]) out[pos[i]] = i;
}

static inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

LSB_EXPORT int64_t lsb_active_volume_ws_bytes(int64_t n) {
    return (int64_t)(align16((size_t)n * 4) + align16((size_t)n * 8) + align16((size_t)(scan_num_blocks(n) + 1) * 8));
}

LSB_EXPORT int lsb_active_volume(const lsb_track_layout* L, const void* tracks, int64_t n, const double* borders, int32_t n_tpc,
                                 int32_t tpc_lo, int32_t tpc_hi, int32_t* first_tpc, int64_t* indices, int64_t* n_selected,
                                 void* ws, int64_t ws_bytes, void* stream) {
    LSB_REQUIRE(L && (tracks || n == 0) && (first_tpc || n == 0), "active_volume: null pointer");
    LSB_REQUIRE(n_tpc >= 0 && n_tpc <= LSB_MAX_TPC && tpc_lo >= 0 && tpc_lo <= tpc_hi && tpc_hi <= n_tpc,
                "active_volume: bad TPC range");
    LSB_REQUIRE(borders || tpc_hi == tpc_lo, "active_volume: null borders");
    static const int need[] = {LSB_F_X_START, LSB_F_Y_START, LSB_F_Z_START, LSB_F_X_END, LSB_F_Y_END, LSB_F_Z_END};
    for (int f : need) LSB_REQUIRE(layout_has(L, f), "active_volume: tracks lacks x/y/z_start or x/y/z_end");
    LSB_REQUIRE(!indices || (n_selected && ws && ws_bytes >= lsb_active_volume_ws_bytes(n)), "active_volume: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        if (n_selected) { cudaMemsetAsync(n_selected, 0, 8, st); }
        return 0;
    }
    uint32_t* flag = nullptr; long long* pos = nullptr; long long* bs = nullptr;
    if (indices) {
        char* p = (char*)ws;
        flag = (uint32_t*)p; p += align16((size_t)n * 4);
        pos = (long long*)p; p += align16((size_t)n * 8);
        bs = (long long*)p;
    }
    const size_t smem = (size_t)(tpc_hi - tpc_lo) * 6 * sizeof(double);
    k_active_volume<<<lsb_blocks(n, 256), 256, smem, st>>>(make_layout(L), (const char*)tracks, n, borders, tpc_lo, tpc_hi, first_tpc, flag);
    LSB_LAUNCH_CHECK("k_active_volume");
    if (indices) {
        int rc = exclusive_scan<uint32_t, long long>(flag, n, pos, bs, (long long*)n_selected, st);
        if (rc) return rc;
        k_emit_indices<<<lsb_blocks(n, 256), 256, 0, st>>>(flag, pos, n, (long long*)indices);
        LSB_LAUNCH_CHECK("k_emit_indices");
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// unit keys:  unit = rank(event) * n_tpc_batches + first_tpc / tpc_batch_size ; n_units for "in no batch"
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ long long load_int_field(const char* p, int dt) {
    switch (dt) {
        case LSB_I32: return *(const int32_t*)p;
        case LSB_U32: return *(const uint32_t*)p;
        case LSB_I64: return *(const long long*)p;
        case LSB_U64: return (long long)*(const unsigned long long*)p;
        case LSB_F32: return (long long)*(const float*)p;
        case LSB_F64: return (long long)*(const double*)p;
    }
    return 0;
}

__global__ void k_unit_keys(const char* __restrict__ tracks, long long n, int itemsize, int ev_off, int ev_dt,
                            const long long* __restrict__ events, long long n_events, const int32_t* __restrict__ first_tpc,
                            int bs, int nB, uint32_t* __restrict__ keys, int32_t* __restrict__ idx) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long ev = load_int_field(tracks + (size_t)i * itemsize + ev_off, ev_dt);
    long long lo = 0, hi = n_events - 1, at = -1;
    while (lo <= hi) {
        const long long mid = (lo + hi) >> 1;
        const long long k = __ldg(events + mid);
        if (k == ev) { at = mid; break; }
        if (k < ev) lo = mid + 1; else hi = mid - 1;
    }
    const int ft = first_tpc[i];
    const long long nU = n_events * nB;
    keys[i] = (at >= 0 && ft >= 0) ? (uint32_t)(at * nB + ft / bs) : (uint32_t)nU;
    idx[i] = (int32_t)i;
}

// ---------------------------------------------------------------------------------------
// stable LSD radix sort, 8 bits per pass.  Tile = RS_TPB threads x RS_ROUNDS consecutive rows of RS_TPB keys.
// hist is digit-major (hist[d * n_tiles + tile]) so that one exclusive scan gives every (digit, tile) its base.
// ---------------------------------------------------------------------------------------
#define RS_TPB 256
#define RS_ROUNDS 4
#define RS_TILE (RS_TPB * RS_ROUNDS)

__global__ void __launch_bounds__(RS_TPB) k_rs_hist(const uint32_t* __restrict__ keys, long long n, int shift, long long n_tiles,
                                                    uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = blockIdx.x * (long long)RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const long long i = base + r * RS_TPB + threadIdx.x;
        if (i < n) atomicAdd(&s_h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = s_h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_TPB) k_rs_scatter(const uint32_t* __restrict__ keys, const int32_t* __restrict__ idx, long long n,
                                                       int shift, long long n_tiles, const long long* __restrict__ base_of,
                                                       uint32_t* __restrict__ keys_out, int32_t* __restrict__ idx_out) {
    __shared__ uint32_t s_cnt[RS_TPB / 32][256];       // keys of each warp per digit, this round
    __shared__ long long s_base[256];                  // next output row of each digit
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    s_base[threadIdx.x] = base_of[(size_t)threadIdx.x * n_tiles + blockIdx.x];
    const long long tile0 = blockIdx.x * (long long)RS_TILE;
    for (int r = 0; r < RS_ROUNDS; r++) {
#pragma unroll
        for (int k = 0; k < RS_TPB / 32; k++) s_cnt[k][threadIdx.x] = 0;
        __syncthreads();
        const long long i = tile0 + r * RS_TPB + threadIdx.x;
        const bool ok = i < n;
        uint32_t key = 0; int32_t id = 0;
        if (ok) { key = keys[i]; id = idx[i]; }
        const uint32_t d = ok ? ((key >> shift) & 255u) : (256u + lane);     // idle lanes match nobody
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (ok && rank == 0) s_cnt[w][d] = __popc(peers);
        __syncthreads();
        if (ok) {
            long long at = s_base[d] + rank;
            for (int k = 0; k < w; k++) at += s_cnt[k][d];
            keys_out[at] = key;
            idx_out[at] = id;
        }
        __syncthreads();
        uint32_t tot = 0;
#pragma unroll
        for (int k = 0; k < RS_TPB / 32; k++) tot += s_cnt[k][threadIdx.x];
        s_base[threadIdx.x] += tot;
        __syncthreads();
    }
}

// unit_offsets[u] = first row of `keys` (sorted) that is >= u, u = 0 .. n_units ; order = sorted indices as int64
__global__ void k_unit_offsets(const uint32_t* __restrict__ keys, long long n, long long n_units, long long* __restrict__ offsets) {
    const long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (u > n_units) return;
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if ((long long)__ldg(keys + mid) < u) lo = mid + 1; else hi = mid;
    }
    offsets[u] = lo;
}
__global__ void k_widen_idx(const int32_t* __restrict__ idx, long long n, long long* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = idx[i];
}

static inline long long rs_tiles(long long n) { return (n + RS_TILE - 1) / RS_TILE; }

// stable sort of (key, index) pairs by the low `bits` bits of the key; the sorted arrays end up in *keyA / *idxA (the pointers are
// swapped per pass); scratch from `pool`
static int rs_sort_pairs(uint32_t** keyA, uint32_t** keyB, int32_t** idxA, int32_t** idxB, long long n, int bits, TmpPool& pool, cudaStream_t st) {
    const long long nt = rs_tiles(n), nh = 256 * nt;
    uint32_t* hist; long long* base; long long* bs; long long* total;
    LSB_CUDA(pool.get(&hist, nh)); LSB_CUDA(pool.get(&base, nh)); LSB_CUDA(pool.get(&bs, scan_num_blocks(nh) + 1)); LSB_CUDA(pool.get(&total, 1));
    for (int shift = 0; shift < bits; shift += 8) {
        k_rs_hist<<<(unsigned)nt, RS_TPB, 0, st>>>(*keyA, n, shift, nt, hist);
        LSB_LAUNCH_CHECK("k_rs_hist");
        int rc = exclusive_scan<uint32_t, long long>(hist, nh, base, bs, total, st);
        if (rc) return rc;
        k_rs_scatter<<<(unsigned)nt, RS_TPB, 0, st>>>(*keyA, *idxA, n, shift, nt, base, *keyB, *idxB);
        LSB_LAUNCH_CHECK("k_rs_scatter");
        uint32_t* tk = *keyA; *keyA = *keyB; *keyB = tk;
        int32_t* ti = *idxA; *idxA = *idxB; *idxB = ti;
    }
    return 0;
}

LSB_EXPORT int64_t lsb_batch_units_ws_bytes(int64_t n) {
    const long long nt = rs_tiles(n), nh = 256 * nt;
    return (int64_t)(4 * align16((size_t)n * 4) + align16((size_t)nh * 4) + align16((size_t)nh * 8) +
                     align16((size_t)(scan_num_blocks(nh) + 1) * 8) + 16);
}

LSB_EXPORT int lsb_batch_units(const void* tracks, int64_t n, int32_t itemsize, int32_t event_offset, int32_t event_dtype,
                               const int64_t* events_sorted, int64_t n_events, const int32_t* first_tpc, int32_t tpc_batch_size,
                               int32_t n_tpc_batches, int64_t* order, int64_t* unit_offsets, void* ws, int64_t ws_bytes,
                               void* stream) {
    LSB_REQUIRE(unit_offsets && (n == 0 || (tracks && first_tpc && order)), "batch_units: null pointer");
    LSB_REQUIRE(n_events >= 0 && (n_events == 0 || events_sorted), "batch_units: null event list");
    LSB_REQUIRE(tpc_batch_size >= 1 && n_tpc_batches >= 0, "batch_units: bad TPC batch size");
    LSB_REQUIRE(event_dtype >= LSB_F32 && event_dtype <= LSB_U64 && event_offset >= 0 && itemsize > 0, "batch_units: bad event field");
    const long long nU = (long long)n_events * n_tpc_batches;
    LSB_REQUIRE(nU < (1ll << 31) && n < (1ll << 31), "batch_units: more than 2^31 units or segments");
    LSB_REQUIRE(n == 0 || (ws && ws_bytes >= lsb_batch_units_ws_bytes(n)), "batch_units: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        cudaMemsetAsync(unit_offsets, 0, (size_t)(nU + 1) * 8, st);
        return 0;
    }
    const long long nt = rs_tiles(n), nh = 256 * nt;
    char* p = (char*)ws;
    uint32_t* keyA = (uint32_t*)p; p += align16((size_t)n * 4);
    uint32_t* keyB = (uint32_t*)p; p += align16((size_t)n * 4);
    int32_t* idxA = (int32_t*)p; p += align16((size_t)n * 4);
    int32_t* idxB = (int32_t*)p; p += align16((size_t)n * 4);
    uint32_t* hist = (uint32_t*)p; p += align16((size_t)nh * 4);
    long long* base = (long long*)p; p += align16((size_t)nh * 8);
    long long* bs = (long long*)p; p += align16((size_t)(scan_num_blocks(nh) + 1) * 8);
    long long* total = (long long*)p;
    k_unit_keys<<<lsb_blocks(n, 256), 256, 0, st>>>((const char*)tracks, n, itemsize, event_offset, event_dtype,
                                                    (const long long*)events_sorted, n_events, first_tpc, tpc_batch_size,
                                                    n_tpc_batches, keyA, idxA);
    LSB_LAUNCH_CHECK("k_unit_keys");
    int bits = 0;
    while ((nU >> bits) != 0) bits++;                  // keys are 0 .. nU
    for (int shift = 0; shift < bits; shift += 8) {
        k_rs_hist<<<(unsigned)nt, RS_TPB, 0, st>>>(keyA, n, shift, nt, hist);
        LSB_LAUNCH_CHECK("k_rs_hist");
        int rc = exclusive_scan<uint32_t, long long>(hist, nh, base, bs, total, st);
        if (rc) return rc;
        k_rs_scatter<<<(unsigned)nt, RS_TPB, 0, st>>>(keyA, idxA, n, shift, nt, base, keyB, idxB);
        LSB_LAUNCH_CHECK("k_rs_scatter");
        uint32_t* tk = keyA; keyA = keyB; keyB = tk;
        int32_t* ti = idxA; idxA = idxB; idxB = ti;
    }
    k_unit_offsets<<<lsb_blocks(nU + 1, 256), 256, 0, st>>>(keyA, n, nU, (long long*)unit_offsets);
    LSB_LAUNCH_CHECK("k_unit_offsets");
    k_widen_idx<<<lsb_blocks(n, 256), 256, 0, st>>>(idxA, n, (long long*)order);
    LSB_LAUNCH_CHECK("k_widen_idx");
    return 0;
}
