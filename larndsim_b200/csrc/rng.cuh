// rng.cuh -- numba.cuda.random.create_xoroshiro128p_states on the device.
//
// Numba initialises the states on the CPU, sequentially: state 0 = splitmix64(seed) in both words, state i =
// jump(state i-1), where jump advances the generator by 2^64 steps (numba/cuda/random.py init_xoroshiro128p_states_cpu;
// 128 next() calls per jump -- SURVEY.md 8(a) row 15 "CPU-side sequential jump init on every growth").  The xoroshiro128+
// transition is linear over GF(2), so jump is a 128x128 bit matrix J and state i = J^i state 0.  The powers J^(2^k) are built
// once on the host (128 jumps of the basis vectors, then repeated squaring); a thread then reaches state
// `subsequence_start + i` with one matrix-vector product per set bit of that index -- the whole array in one launch, no
// sequential chain, bit-identical to Numba's loop (tests/golden/rng.npz pins the host version, the GPU test compares both).
#pragma once
#include "common.cuh"

#define RNG_JUMP_LEVELS 48           // subsequence indices below 2^48

struct RngJumpTable { uint64_t col[RNG_JUMP_LEVELS][128][2]; };   // J^(2^k): image of basis vector b (bit b of {s0, s1})

static inline void rngj_next(uint64_t* s) {
    uint64_t s0 = s[0], s1 = s[1];
    s1 ^= s0;
    s[0] = ((s0 << 55) | (s0 >> 9)) ^ s1 ^ (s1 << 14);
    s[1] = (s1 << 36) | (s1 >> 28);
}
static inline void rngj_jump(uint64_t* s) {
    static const uint64_t J[2] = {0xbeac0467eba5facbULL, 0xd86b048b86aa9922ULL};
    uint64_t a = 0, b = 0;
    for (int i = 0; i < 2; i++)
        for (int bit = 0; bit < 64; bit++) {
            if (J[i] & (1ULL << bit)) { a ^= s[0]; b ^= s[1]; }
            rngj_next(s);
        }
    s[0] = a; s[1] = b;
}
static inline void rngj_matvec(const uint64_t (*col)[2], const uint64_t* v, uint64_t* out) {
    uint64_t a = 0, b = 0;
    for (int w = 0; w < 2; w++)
        for (int bit = 0; bit < 64; bit++)
            if (v[w] & (1ULL << bit)) { a ^= col[64 * w + bit][0]; b ^= col[64 * w + bit][1]; }
    out[0] = a; out[1] = b;
}
static const RngJumpTable* rng_jump_table_host() {
    static RngJumpTable* T = nullptr;
    if (T) return T;
    T = new RngJumpTable();
    for (int b = 0; b < 128; b++) {
        uint64_t v[2] = {0, 0};
        v[b >> 6] = 1ULL << (b & 63);
        rngj_jump(v);
        T->col[0][b][0] = v[0]; T->col[0][b][1] = v[1];
    }
    for (int k = 1; k < RNG_JUMP_LEVELS; k++)
        for (int b = 0; b < 128; b++) rngj_matvec(T->col[k - 1], T->col[k - 1][b], T->col[k][b]);
    return T;
}
// one device copy per device (the library is used by one process per GPU, but nothing forbids cudaSetDevice)
static const RngJumpTable* rng_jump_table_dev() {
    static const RngJumpTable* dev_tab[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!dev_tab[dev]) {
        RngJumpTable* d = nullptr;
        if (cudaMalloc((void**)&d, sizeof(RngJumpTable)) != cudaSuccess) return nullptr;
        if (cudaMemcpy(d, rng_jump_table_host(), sizeof(RngJumpTable), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
        dev_tab[dev] = d;
    }
    return dev_tab[dev];
}

__global__ void k_rng_create_states(const RngJumpTable* __restrict__ tab, unsigned long long* __restrict__ states, long long n,
                                    unsigned long long seed, unsigned long long start) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL;             // splitmix64 (random.py init_xoroshiro128p_state)
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    unsigned long long s0 = z, s1 = z;
    unsigned long long idx = start + (unsigned long long)i;
    for (int k = 0; idx; k++, idx >>= 1) {
        if (!(idx & 1ULL)) continue;
        if (k >= RNG_JUMP_LEVELS) { s0 = 0; s1 = 0; break; }          // unreachable for supported sizes (checked on the host)
        const ulonglong2* col = reinterpret_cast<const ulonglong2*>(tab->col[k]);
        unsigned long long a = 0, b = 0;
#pragma unroll 4
        for (int bit = 0; bit < 64; bit++) {
            const ulonglong2 c0 = __ldg(col + bit), c1 = __ldg(col + 64 + bit);
            const unsigned long long m0 = 0ULL - ((s0 >> bit) & 1ULL), m1 = 0ULL - ((s1 >> bit) & 1ULL);
            a ^= (c0.x & m0) ^ (c1.x & m1);
            b ^= (c0.y & m0) ^ (c1.y & m1);
        }
        s0 = a; s1 = b;
    }
    states[2 * i] = s0; states[2 * i + 1] = s1;
}

// ---- stepping a stream ahead: T^(RNG_STEP_UNIT * 2^k), T = one next() -------------------------------------------------------------
// A stream that has to be consumed from many positions at once (the per-pixel noise of get_adc_values: ~13 000 draws per pixel,
// strictly sequential in the reference) is cut into chunks of RNG_STEP_UNIT draws; the state at the start of chunk g is
// T^(RNG_STEP_UNIT * g) state 0, one bit-matrix product per set bit of g.
#ifndef RNG_STEP_UNIT
#define RNG_STEP_UNIT 512            // draws per chunk = 256 Box-Muller normals.  The jump to a chunk costs ~3 matrix products (~1300
#endif                               // instructions each): k_fee_rng_chunks per ND-LAr unit 0.69 ms at 128, 0.52 at 256, 0.47 at 512
#define RNG_STEP_LEVELS 20
struct RngStepTable { uint64_t col[RNG_STEP_LEVELS][128][2]; };
static const RngStepTable* rng_step_table_host() {
    static RngStepTable* T = nullptr;
    if (T) return T;
    T = new RngStepTable();
    uint64_t cur[128][2], nxt[128][2];
    for (int b = 0; b < 128; b++) {                                   // T itself
        uint64_t v[2] = {0, 0};
        v[b >> 6] = 1ULL << (b & 63);
        rngj_next(v);
        cur[b][0] = v[0]; cur[b][1] = v[1];
    }
    for (int sq = 0; (1 << sq) < RNG_STEP_UNIT; sq++) {               // T^(2^sq) -> T^RNG_STEP_UNIT
        for (int b = 0; b < 128; b++) rngj_matvec(cur, cur[b], nxt[b]);
        memcpy(cur, nxt, sizeof(cur));
    }
    memcpy(T->col[0], cur, sizeof(cur));
    for (int k = 1; k < RNG_STEP_LEVELS; k++)
        for (int b = 0; b < 128; b++) rngj_matvec(T->col[k - 1], T->col[k - 1][b], T->col[k][b]);
    return T;
}
static const RngStepTable* rng_step_table_dev() {
    static const RngStepTable* dev_tab[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!dev_tab[dev]) {
        RngStepTable* d = nullptr;
        if (cudaMalloc((void**)&d, sizeof(RngStepTable)) != cudaSuccess) return nullptr;
        if (cudaMemcpy(d, rng_step_table_host(), sizeof(RngStepTable), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
        dev_tab[dev] = d;
    }
    return dev_tab[dev];
}
__device__ __forceinline__ void rng_matvec_dev(const ulonglong2* __restrict__ col, unsigned long long& s0, unsigned long long& s1) {
    unsigned long long a = 0, b = 0;
#pragma unroll 4
    for (int bit = 0; bit < 64; bit++) {
        const ulonglong2 c0 = __ldg(col + bit), c1 = __ldg(col + 64 + bit);
        const unsigned long long m0 = 0ULL - ((s0 >> bit) & 1ULL), m1 = 0ULL - ((s1 >> bit) & 1ULL);
        a ^= (c0.x & m0) ^ (c1.x & m1);
        b ^= (c0.y & m0) ^ (c1.y & m1);
    }
    s0 = a; s1 = b;
}

static int rng_create_states_dev(unsigned long long* states_dev, long long n, uint64_t seed, uint64_t start, cudaStream_t st) {
    if (n <= 0) return 0;
    LSB_REQUIRE(start + (uint64_t)n < (1ULL << RNG_JUMP_LEVELS), "rng_create_states: subsequence index beyond 2^48");
    const RngJumpTable* tab = rng_jump_table_dev();
    if (!tab) return lsb_fail_arg("rng_create_states: cannot stage the jump table on the device");
    k_rng_create_states<<<lsb_blocks(n, 128), 128, 0, st>>>(tab, states_dev, n, seed, start);
    LSB_LAUNCH_CHECK("k_rng_create_states");
    return 0;
}

LSB_EXPORT int lsb_rng_create_states(uint64_t* states_dev, int64_t n, uint64_t seed, uint64_t subsequence_start, void* stream) {
    LSB_REQUIRE(states_dev || n == 0, "rng_create_states: null pointer");
    return rng_create_states_dev((unsigned long long*)states_dev, n, seed, subsequence_start, (cudaStream_t)stream);
}
