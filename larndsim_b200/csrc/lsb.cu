// lsb.cu -- the one translation unit of liblarndsim_b200.so (sm_100a only).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
//        (see __graft_entry__.build()).  The C ABI is declared in include/larndsim_b200.h.
#include <algorithm>
#include "common.cuh"

char g_lsb_error[512] = "";
long long g_lsb_launches = 0;

static lsb_consts g_consts_host;
static bool g_consts_valid = false;
static int g_consts_device = -1;          // the __constant__ copy lives on ONE device: a cudaSetDevice invalidates the cache

int lsb_upload_consts(const lsb_consts* c, cudaStream_t st) {
    int dev = -1;
    LSB_CUDA(cudaGetDevice(&dev));
    if (dev != g_consts_device) { g_consts_valid = false; g_consts_device = dev; }
    if (g_consts_valid && memcmp(&g_consts_host, c, sizeof(lsb_consts)) == 0) return 0;
    if (c->n_tpc < 0 || c->n_tpc > LSB_MAX_TPC) return lsb_fail_arg("consts: n_tpc out of range");
    // kernels of earlier calls (any stream) may still read the old snapshot
    LSB_CUDA(cudaDeviceSynchronize());
    memcpy(&g_consts_host, c, sizeof(lsb_consts));
    g_consts_valid = false;
    LSB_CUDA(cudaMemcpyToSymbolAsync(d_c, &g_consts_host, sizeof(lsb_consts), 0, cudaMemcpyHostToDevice, st));
    LSB_CUDA(cudaStreamSynchronize(st));
    g_consts_valid = true;
    return 0;
}

TmpArena* g_lsb_arena = nullptr;

// a process-wide time origin for pipeline timelines (recorded on first use)
cudaEvent_t lsb_reference_event() {
    static cudaEvent_t ev = nullptr;
    if (!ev) { cudaEventCreate(&ev); cudaEventRecord(ev, 0); cudaEventSynchronize(ev); }
    return ev;
}

void lsb_pool_init_once() {
    static bool done = false;
    if (done) return;
    done = true;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) return;
    unsigned long long thr = ~0ULL;     // keep freed blocks cached: steady-state calls never reach the driver
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
}

// ---- per-kernel timing --------------------------------------------------------------------
#include <map>
#include <string>
#include <vector>
int g_lsb_profiling = 0;
struct ProfMark { const char* name; cudaEvent_t ev; };
static std::vector<ProfMark> g_prof_marks;
static std::vector<cudaEvent_t> g_prof_pool;
void lsb_profile_mark(const char* name, cudaStream_t st) {
    cudaEvent_t ev;
    if (!g_prof_pool.empty()) { ev = g_prof_pool.back(); g_prof_pool.pop_back(); }
    else if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, st);
    g_prof_marks.push_back({name, ev});
}
LSB_EXPORT int lsb_profile_begin(void* stream) {
    for (auto& m : g_prof_marks) g_prof_pool.push_back(m.ev);
    g_prof_marks.clear();
    g_lsb_profiling = 1;
    lsb_profile_mark("(begin)", (cudaStream_t)stream);
    return 0;
}
// writes "name count total_ms\n" lines, sorted by total time; returns the number of bytes needed
LSB_EXPORT int64_t lsb_profile_end(char* out, int64_t cap) {
    g_lsb_profiling = 0;
    if (!g_prof_marks.empty()) cudaEventSynchronize(g_prof_marks.back().ev);
    std::map<std::string, std::pair<long long, double>> acc;
    for (size_t i = 1; i < g_prof_marks.size(); i++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_prof_marks[i - 1].ev, g_prof_marks[i].ev) != cudaSuccess) continue;
        auto& a = acc[g_prof_marks[i].name];
        a.first++; a.second += ms;
    }
    std::vector<std::pair<double, std::string>> order;
    for (auto& kv : acc) order.push_back({-kv.second.second, kv.first});
    std::sort(order.begin(), order.end());
    std::string txt;
    char line[256];
    for (auto& o : order) {
        auto& a = acc[o.second];
        snprintf(line, sizeof(line), "%s %lld %.6f\n", o.second.c_str(), a.first, a.second);
        txt += line;
    }
    if (out && cap > 0) { size_t n = txt.size() < (size_t)cap - 1 ? txt.size() : (size_t)cap - 1; memcpy(out, txt.data(), n); out[n] = 0; }
    return (int64_t)txt.size() + 1;
}

#include "segments.cuh"
#include "glue.cuh"
#include "current.cuh"
#include "pixelmap.cuh"
#include "fee.cuh"
#include "light.cuh"
#include "packets.cuh"
#include "light_trigger.cuh"
#include "batching.cuh"
#include "rng.cuh"
#include "chain.cuh"
#include "spill.cuh"

LSB_EXPORT int lsb_abi_version(void) { return LSB_ABI_VERSION; }
LSB_EXPORT const char* lsb_last_error(void) { return g_lsb_error; }
LSB_EXPORT int64_t lsb_launch_count(void) { return g_lsb_launches; }

// numba/cuda/random.py: init_xoroshiro128p_states_cpu -- splitmix64(seed) -> s0 = s1, then one
// 2^64 jump per subsequence.  Host side like the reference (create_xoroshiro128p_states runs on the CPU).
static inline uint64_t h_rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline void h_next(uint64_t* s) {
    uint64_t s0 = s[0], s1 = s[1];
    s1 ^= s0;
    s[0] = h_rotl(s0, 55) ^ s1 ^ (s1 << 14);
    s[1] = h_rotl(s1, 36);
}
static void h_jump(uint64_t* s) {
    static const uint64_t J[2] = {0xbeac0467eba5facbULL, 0xd86b048b86aa9922ULL};
    uint64_t a = 0, b = 0;
    for (int i = 0; i < 2; i++)
        for (int bit = 0; bit < 64; bit++) {
            if (J[i] & (1ULL << bit)) { a ^= s[0]; b ^= s[1]; }
            h_next(s);
        }
    s[0] = a; s[1] = b;
}
int lsb_rng_create_states_host_impl(uint64_t* states, int64_t n, uint64_t seed, uint64_t subsequence_start) {
    if (n < 1) return 0;
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    states[0] = z; states[1] = z;
    for (uint64_t i = 0; i < subsequence_start; i++) h_jump(states);
    for (int64_t i = 1; i < n; i++) {
        states[2 * i] = states[2 * i - 2]; states[2 * i + 1] = states[2 * i - 1];
        h_jump(states + 2 * i);
    }
    return 0;
}
LSB_EXPORT int lsb_rng_create_states_host(uint64_t* states_host, int64_t n, uint64_t seed, uint64_t subsequence_start) {
    LSB_REQUIRE(states_host || n == 0, "rng_create_states_host: null pointer");
    return lsb_rng_create_states_host_impl(states_host, n, seed, subsequence_start);
}
