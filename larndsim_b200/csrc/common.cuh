// common.cuh -- shared device/host helpers for the sm_100a kernels of larndsim_b200.
// Compiled with -fmad=false: index/gating arithmetic must round exactly like the reference's
// float64 expressions; FMAs are written explicitly where wanted.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/larndsim_b200.h"

#define LSB_EXPORT extern "C" __attribute__((visibility("default")))

// ---------------------------------------------------------------------------------------
// error handling / launch accounting
// ---------------------------------------------------------------------------------------
extern char g_lsb_error[512];
extern long long g_lsb_launches;

static inline int lsb_fail_arg(const char* what) {
    snprintf(g_lsb_error, sizeof(g_lsb_error), "argument error: %s", what);
    return -1;
}
static inline int lsb_fail_cuda(cudaError_t e, const char* where) {
    snprintf(g_lsb_error, sizeof(g_lsb_error), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), where);
    return (int)e;
}
#define LSB_CUDA(call)                                                     \
    do {                                                                   \
        cudaError_t e__ = (call);                                          \
        if (e__ != cudaSuccess) return lsb_fail_cuda(e__, #call);          \
    } while (0)
// per-kernel timing (bench.py): when a profile session is open, an event is recorded after every
// launch / memset; the interval since the previous mark on the stream is attributed to `name`.
void lsb_profile_mark(const char* name, cudaStream_t st);
extern int g_lsb_profiling;
#define LSB_MARK(name, st) do { if (g_lsb_profiling) lsb_profile_mark(name, st); } while (0)
#define LSB_LAUNCH_CHECK_ST(name, st)                                      \
    do {                                                                   \
        g_lsb_launches++;                                                  \
        cudaError_t e__ = cudaGetLastError();                              \
        if (e__ != cudaSuccess) return lsb_fail_cuda(e__, name);           \
        LSB_MARK(name, st);                                                \
    } while (0)
#define LSB_LAUNCH_CHECK(name) LSB_LAUNCH_CHECK_ST(name, st)
#define LSB_REQUIRE(cond, what)                                            \
    do { if (!(cond)) return lsb_fail_arg(what); } while (0)

static inline unsigned int lsb_blocks(long long n, int tpb) { return (unsigned int)((n + tpb - 1) / tpb); }

// ---------------------------------------------------------------------------------------
// constants: one POD snapshot in constant memory, refreshed when the host copy changes
// ---------------------------------------------------------------------------------------
__constant__ lsb_consts d_c;   // single translation unit (lsb.cu)
int lsb_upload_consts(const lsb_consts* c, cudaStream_t st);

// record layout is passed by value to kernels
struct Layout {
    int itemsize;
    int off[LSB_F_COUNT];
    int dt[LSB_F_COUNT];
};
static inline Layout make_layout(const lsb_track_layout* L) {
    Layout o;
    o.itemsize = L->itemsize;
    for (int i = 0; i < LSB_F_COUNT; i++) { o.off[i] = L->offset[i]; o.dt[i] = L->dtype[i]; }
    return o;
}
static inline bool layout_has(const lsb_track_layout* L, int f) { return L->offset[f] >= 0 && L->dtype[f] != LSB_NONE; }

__device__ __forceinline__ double fld_get(const Layout& L, const char* rec, int f) {
    const char* p = rec + L.off[f];
    switch (L.dt[f]) {
        case LSB_F32: return (double)*(const float*)p;
        case LSB_F64: return *(const double*)p;
        case LSB_I32: return (double)*(const int32_t*)p;
        case LSB_U32: return (double)*(const uint32_t*)p;
        case LSB_I64: return (double)*(const long long*)p;
        case LSB_U64: return (double)*(const unsigned long long*)p;
    }
    return 0.0;
}
// store with the conversion Numba applies when assigning a float64 to the field
__device__ __forceinline__ void fld_set(const Layout& L, char* rec, int f, double x) {
    char* p = rec + L.off[f];
    switch (L.dt[f]) {
        case LSB_F32: *(float*)p = __double2float_rn(x); break;
        case LSB_F64: *(double*)p = x; break;
        case LSB_I32: *(int32_t*)p = __double2int_rz(x); break;
        case LSB_U32: *(uint32_t*)p = __double2uint_rz(x); break;
        case LSB_I64: *(long long*)p = __double2ll_rz(x); break;
        case LSB_U64: *(unsigned long long*)p = __double2ull_rz(x); break;
    }
}
__device__ __forceinline__ bool fld_f32(const Layout& L, int f) { return L.dt[f] == LSB_F32; }
// result of an operation Numba types as float32 (both operands float32): computing in double and
// rounding once to float is exact for + - * / sqrt on float operands
__device__ __forceinline__ double R32(double x, bool is32) { return is32 ? (double)__double2float_rn(x) : x; }

// Python float floor division (Numba real_floordiv == CPython float_divmod)
__device__ __forceinline__ double py_floordiv(double a, double b) {
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0.0) {
        if ((b < 0.0) != (mod < 0.0)) { mod += b; div -= 1.0; }
    }
    double fl;
    if (div != 0.0) {
        fl = floor(div);
        if (div - fl > 0.5) fl += 1.0;
    } else {
        fl = copysign(0.0, a / b);
    }
    return fl;
}
__device__ __forceinline__ long long py_mod_ll(long long a, long long b) {
    long long m = a % b;
    if (m != 0 && ((m < 0) != (b < 0))) m += b;
    return m;
}
__device__ __forceinline__ long long py_div_ll(long long a, long long b) {
    long long q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) q -= 1;
    return q;
}
__device__ __forceinline__ bool in_plane(long long x, long long y, long long plane) {
    return 0 <= x && x < d_c.n_pixels[0] && 0 <= y && y < d_c.n_pixels[1] && 0 <= plane && plane < d_c.n_tpc;
}
__device__ __forceinline__ long long pixel2id(long long x, long long y, long long plane) {
    return x + (long long)d_c.n_pixels[0] * (y + (long long)d_c.n_pixels[1] * plane);
}
__device__ __forceinline__ void id2pixel(long long pid, long long& x, long long& y, long long& plane) {
    x = py_mod_ll(pid, d_c.n_pixels[0]);
    y = py_mod_ll(py_div_ll(pid, d_c.n_pixels[0]), d_c.n_pixels[1]);
    plane = py_div_ll(pid, (long long)d_c.n_pixels[0] * d_c.n_pixels[1]);
}

// ---------------------------------------------------------------------------------------
// RNG: numba.cuda.random xoroshiro128+ state layout {s0,s1} and Box-Muller in float32
// (numba/cuda/random.py; on the GPU Numba lowers math.log/cos/sqrt of float32 to libdevice
// __nv_logf/__nv_cosf/__nv_sqrtf, which are what logf/cosf/sqrtf compile to here)
// ---------------------------------------------------------------------------------------
struct Rng { unsigned long long s0, s1; };
__device__ __forceinline__ unsigned long long rotl64(unsigned long long x, int k) { return (x << k) | (x >> (64 - k)); }
__device__ __forceinline__ unsigned long long rng_next(Rng& r) {
    unsigned long long s0 = r.s0, s1 = r.s1;
    unsigned long long result = s0 + s1;
    s1 ^= s0;
    r.s0 = rotl64(s0, 55) ^ s1 ^ (s1 << 14);
    r.s1 = rotl64(s1, 36);
    return result;
}
__device__ __forceinline__ float rng_uniform_f32(Rng& r) {
    // float32((x >> 11) * 2^-53) (numba/cuda/random.py uint64_to_unit_float32): the 53-bit integer is exact in float64 and the
    // scaling is a power of two, so the only rounding is the one to float32 -- taken here straight from the integer (I2F.F32.U64,
    // round to nearest even) followed by the exact scaling; bit-identical, and off the FP64 pipe
    unsigned long long x = rng_next(r);
    return __fmul_rn(__ull2float_rn(x >> 11), 1.1102230246251565e-16f);
}
__device__ __forceinline__ float rng_normal_f32(Rng& r) {
    float u1 = rng_uniform_f32(r);
    float u2 = rng_uniform_f32(r);
    float a = sqrtf(__fmul_rn(-2.0f, logf(u1)));
    float b = cosf(__fmul_rn(6.283185307179586f, u2));
    return __fmul_rn(a, b);
}

// ---------------------------------------------------------------------------------------
// stream-ordered temporaries (cudaMallocAsync from the default pool; the pool keeps freed blocks,
// so steady-state calls do not reach the driver).  Freed when the guard leaves scope.
// ---------------------------------------------------------------------------------------
void lsb_pool_init_once();
// A caller (the fused chain) may lend a pre-allocated device arena: temporaries are then carved from it with
// stack discipline (all users of one arena queue their work on ONE stream, so reuse is stream-ordered) and
// no allocator is involved at all; requests that do not fit fall back to the pool and are remembered so the
// owner can grow the arena for the next batch.
struct TmpArena { char* base; size_t cap, off, high_water, overflow; };
extern TmpArena* g_lsb_arena;
struct TmpPool {
    cudaStream_t st;
    void* ptrs[24];
    int n;
    TmpArena* arena;
    size_t arena_mark;
    explicit TmpPool(cudaStream_t s) : st(s), n(0), arena(g_lsb_arena), arena_mark(g_lsb_arena ? g_lsb_arena->off : 0) { lsb_pool_init_once(); }
    template <typename T>
    cudaError_t get(T** p, long long count) {
        *p = nullptr;
        size_t bytes = (size_t)(count > 0 ? count : 1) * sizeof(T);
        if (arena) {
            size_t need = (bytes + 255) & ~(size_t)255;
            if (arena->off + need <= arena->cap) {
                *p = (T*)(arena->base + arena->off);
                arena->off += need;
                if (arena->off > arena->high_water) arena->high_water = arena->off;
                return cudaSuccess;
            }
            arena->overflow += need;
        }
        if (n >= 24) return cudaErrorMemoryAllocation;
        cudaError_t e = cudaMallocAsync((void**)p, bytes, st);
        if (e == cudaSuccess) ptrs[n++] = (void*)*p;
        return e;
    }
    ~TmpPool() {
        for (int i = n - 1; i >= 0; i--) cudaFreeAsync(ptrs[i], st);
        if (arena) arena->off = arena_mark;
    }
};
