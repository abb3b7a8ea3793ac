/*
 * larnd_oracle.c -- CPU restatement of the larnd-sim charge/light readout chain.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (`larnd-sim_b200/`) links, imports or
 * executes this file; only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
 * `--impl reference` legs of `bench.py` do, as the checker / the timed CPU baseline.
 *
 * Each function is a literal, loop-for-thread restatement of one reference kernel and cites
 * the reference file:line it follows (paths relative to the reference tree).  Arithmetic
 * follows the *compiled* (Numba-typed) semantics of the reference: expressions that involve
 * a Python-float module constant are float64, expressions whose operands are both float32
 * record fields stay float32, 1-argument round() is llrint (half-to-even), `//` is Python
 * floor division, stores into f4/u4 record fields round / truncate.
 *
 * Parity pin: the functions here are checked in tests/test_oracle_golden.py against
 * fixtures under tests/golden/ that were produced by the reference's own kernels
 * (tools/gen_golden.py: reference source JIT-compiled for the host by numba, and Numba's
 * CUDA simulator), incl. SURVEY.md appendix B vectors.
 *
 * Third-party arithmetic on the path: numba.cuda.random (numba 0.65.0) xoroshiro128+ /
 * Box-Muller -- restated in rng_* below from its published algorithm and pinned against
 * the installed numba's own host functions (golden vectors in tests/golden/rng.npz).
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared; see oracle/Makefile)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/larndsim_b200.h"

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------ */
/* record access                                                                          */
/* ------------------------------------------------------------------------------------ */
static inline double fld_get(const lsb_track_layout* L, const char* rec, int f) {
    const char* p = rec + L->offset[f];
    switch (L->dtype[f]) {
        case LSB_F32: { float v; memcpy(&v, p, 4); return (double)v; }
        case LSB_F64: { double v; memcpy(&v, p, 8); return v; }
        case LSB_I32: { int32_t v; memcpy(&v, p, 4); return (double)v; }
        case LSB_U32: { uint32_t v; memcpy(&v, p, 4); return (double)v; }
        case LSB_I64: { int64_t v; memcpy(&v, p, 8); return (double)v; }
        case LSB_U64: { uint64_t v; memcpy(&v, p, 8); return (double)v; }
    }
    return 0.0;
}
static inline void fld_set(const lsb_track_layout* L, char* rec, int f, double x) {
    char* p = rec + L->offset[f];
    switch (L->dtype[f]) {
        case LSB_F32: { float v = (float)x; memcpy(p, &v, 4); break; }
        case LSB_F64: { memcpy(p, &x, 8); break; }
        case LSB_I32: { int32_t v = (int32_t)x; memcpy(p, &v, 4); break; }
        case LSB_U32: { uint32_t v = (x <= 0.0) ? 0u : (x >= 4294967295.0 ? 4294967295u : (uint32_t)x); memcpy(p, &v, 4); break; }
        case LSB_I64: { int64_t v = (int64_t)x; memcpy(p, &v, 8); break; }
        case LSB_U64: { uint64_t v = (x <= 0.0) ? 0u : (uint64_t)x; memcpy(p, &v, 8); break; }
    }
}
static inline int fld_is_f32(const lsb_track_layout* L, int f) { return L->dtype[f] == LSB_F32; }
/* round an exactly-computed float64 result to float32 when Numba types the op as float32 */
static inline double R32(double x, int is32) { return is32 ? (double)(float)x : x; }

/* Python float floor division (Numba real_floordiv == CPython float_divmod) */
static inline double py_floordiv(double a, double b) {
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

/* ------------------------------------------------------------------------------------ */
/* RNG: numba/cuda/random.py (numba 0.65.0), xoroshiro128+ with Box-Muller                */
/* ------------------------------------------------------------------------------------ */
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(uint64_t* s) {        /* xoroshiro128p_next */
    uint64_t s0 = s[0], s1 = s[1];
    uint64_t result = s0 + s1;
    s1 ^= s0;
    s[0] = rotl64(s0, 55) ^ s1 ^ (s1 << 14);
    s[1] = rotl64(s1, 36);
    return result;
}
static void rng_jump(uint64_t* s) {                   /* xoroshiro128p_jump */
    static const uint64_t J[2] = {0xbeac0467eba5facbULL, 0xd86b048b86aa9922ULL};
    uint64_t s0 = 0, s1 = 0;
    for (int i = 0; i < 2; i++)
        for (int b = 0; b < 64; b++) {
            if (J[i] & (1ULL << b)) { s0 ^= s[0]; s1 ^= s[1]; }
            rng_next(s);
        }
    s[0] = s0; s[1] = s1;
}
static inline float rng_uniform_f32(uint64_t* s) {    /* xoroshiro128p_uniform_float32 */
    uint64_t x = rng_next(s);
    double d = (double)(x >> 11) * (1.0 / 9007199254740992.0);
    return (float)d;
}
static inline float rng_normal_f32(uint64_t* s) {     /* xoroshiro128p_normal_float32 */
    float u1 = rng_uniform_f32(s);
    float u2 = rng_uniform_f32(s);
    float r = sqrtf(-2.0f * logf(u1));
    float c = cosf(6.283185307179586f * u2);
    return r * c;
}
ORC_API void orc_rng_create_states(uint64_t* states, int64_t n, uint64_t seed, uint64_t subsequence_start) {
    if (n < 1) return;                                /* init_xoroshiro128p_states_cpu */
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    states[0] = z; states[1] = z;
    for (uint64_t i = 0; i < subsequence_start; i++) rng_jump(states);
    for (int64_t i = 1; i < n; i++) {
        states[2 * i] = states[2 * i - 2]; states[2 * i + 1] = states[2 * i - 1];
        rng_jump(states + 2 * i);
    }
}
ORC_API void orc_rng_draw(uint64_t* state, int32_t n, int32_t kind, float* out) {
    for (int i = 0; i < n; i++) out[i] = kind ? rng_normal_f32(state) : rng_uniform_f32(state);
}

/* ------------------------------------------------------------------------------------ */
/* quenching.py:11-44                                                                     */
/* ------------------------------------------------------------------------------------ */
ORC_API int orc_quench(const lsb_consts* c, const lsb_track_layout* L, void* tracks, int64_t n, int32_t mode) {
    if (mode != c->mode_box && mode != c->mode_birks) return -1;    /* quenching.py:37-38 */
    for (int64_t i = 0; i < n; i++) {
        char* t = (char*)tracks + i * L->itemsize;
        double dEdx = fld_get(L, t, LSB_F_DEDX);
        double dE = fld_get(L, t, LSB_F_DE);
        double recomb = 0;
        if (mode == c->mode_box) {                                  /* :30-33 */
            double csi = c->box_beta * dEdx / (c->e_field * c->lar_density);
            double r = log(c->box_alpha + csi) / csi;
            recomb = (r > 0) ? r : 0;   /* max(0, r): NaN compares false -> 0 kept like Python max(0,nan) */
            if (isnan(r)) recomb = 0;
        } else {                                                    /* :34-36 */
            recomb = c->birks_ab / (1 + c->birks_kb * dEdx / (c->e_field * c->lar_density));
        }
        fld_set(L, t, LSB_F_N_ELECTRONS, recomb * dE / c->w_ion);   /* :43 (u4 truncation) */
        double ne = fld_get(L, t, LSB_F_N_ELECTRONS);               /* :44 uses the stored value */
        fld_set(L, t, LSB_F_N_PHOTONS, (dE / c->w_ph - ne) * c->scint_prescale);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* drifting.py:11-58                                                                      */
/* ------------------------------------------------------------------------------------ */
ORC_API int orc_drift(const lsb_consts* c, const lsb_track_layout* L, void* tracks, int64_t n) {
    for (int64_t i = 0; i < n; i++) {
        char* t = (char*)tracks + i * L->itemsize;
        double x = fld_get(L, t, LSB_F_X), y = fld_get(L, t, LSB_F_Y), z = fld_get(L, t, LSB_F_Z);
        int32_t plane = c->default_plane_index;
        for (int ip = 0; ip < c->n_tpc; ip++) {                     /* :32-37 */
            const double (*b)[2] = c->tpc_borders[ip];
            double zlo = fmin(b[2][1] - 2e-2, b[2][0] - 2e-2), zhi = fmax(b[2][1] + 2e-2, b[2][0] + 2e-2);
            if (b[0][0] - 2e-2 <= x && x <= b[0][1] + 2e-2 && b[1][0] - 2e-2 <= y && y <= b[1][1] + 2e-2 &&
                zlo <= z && z <= zhi) { plane = ip; break; }
        }
        fld_set(L, t, LSB_F_PIXEL_PLANE, plane);                    /* :39 */
        if (plane != c->default_plane_index) {
            double z_anode = c->tpc_borders[plane][2][0];
            double zs = fld_get(L, t, LSB_F_Z_START), ze = fld_get(L, t, LSB_F_Z_END);
            double drift_distance = fabs(z - z_anode);
            double drift_start = fabs(fmin(zs, ze) - z_anode);
            double drift_end = fabs(fmax(zs, ze) - z_anode);
            double drift_time = drift_distance / c->v_drift;
            double lifetime_red = exp(-drift_time / c->electron_lifetime);
            fld_set(L, t, LSB_F_N_ELECTRONS, fld_get(L, t, LSB_F_N_ELECTRONS) * lifetime_red);   /* :51 */
            fld_set(L, t, LSB_F_LONG_DIFF, sqrt(drift_time * 2 * c->long_diff));
            fld_set(L, t, LSB_F_TRAN_DIFF, sqrt(drift_time * 2 * c->tran_diff));
            double t0 = fld_get(L, t, LSB_F_T0);
            fld_set(L, t, LSB_F_T, fld_get(L, t, LSB_F_T) + (drift_time + t0));                  /* :56-58 */
            fld_set(L, t, LSB_F_T_START, fld_get(L, t, LSB_F_T_START) + (fmin(drift_start, drift_end) / c->v_drift + t0));
            fld_set(L, t, LSB_F_T_END, fld_get(L, t, LSB_F_T_END) + (fmax(drift_start, drift_end) / c->v_drift + t0));
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* pixels_from_track.py                                                                   */
/* ------------------------------------------------------------------------------------ */
static inline int in_plane(const lsb_consts* c, int64_t x, int64_t y, int64_t plane) {
    return 0 <= x && x < c->n_pixels[0] && 0 <= y && y < c->n_pixels[1] && 0 <= plane && plane < c->n_tpc;
}
static inline int64_t pixel2id(const lsb_consts* c, int64_t x, int64_t y, int64_t plane) {   /* :13-26 */
    return x + c->n_pixels[0] * (y + c->n_pixels[1] * plane);
}
/* Python floor semantics for % and // on ints (pid = -1 gives x = Nx-1 ...; :28-41) */
static inline int64_t py_mod(int64_t a, int64_t b) { int64_t m = a % b; if (m != 0 && ((m < 0) != (b < 0))) m += b; return m; }
static inline int64_t py_div(int64_t a, int64_t b) { int64_t q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) q -= 1; return q; }
static inline void id2pixel(const lsb_consts* c, int64_t pid, int64_t* x, int64_t* y, int64_t* plane) {
    *x = py_mod(pid, c->n_pixels[0]);
    *y = py_mod(py_div(pid, c->n_pixels[0]), c->n_pixels[1]);
    *plane = py_div(pid, (int64_t)c->n_pixels[0] * c->n_pixels[1]);
}
static int segment_pixels(const lsb_consts* c, const lsb_track_layout* L, const char* t,
                          int64_t* x0, int64_t* y0, int64_t* x1, int64_t* y1, int64_t* plane) {
    *plane = (int64_t)fld_get(L, t, LSB_F_PIXEL_PLANE);
    if (*plane < 0 || *plane >= c->n_tpc) return 0;   /* reference reads TPC_BORDERS out of bounds here */
    const double (*b)[2] = c->tpc_borders[*plane];
    *x0 = (int64_t)py_floordiv(fld_get(L, t, LSB_F_X_START) - b[0][0], c->pixel_pitch);
    *y0 = (int64_t)py_floordiv(fld_get(L, t, LSB_F_Y_START) - b[1][0], c->pixel_pitch);
    *x1 = (int64_t)py_floordiv(fld_get(L, t, LSB_F_X_END) - b[0][0], c->pixel_pitch);
    *y1 = (int64_t)py_floordiv(fld_get(L, t, LSB_F_Y_END) - b[1][0], c->pixel_pitch);
    return 1;
}
/* get_num_active_pixels :111-155 */
static int64_t num_active_pixels(const lsb_consts* c, int64_t x0, int64_t y0, int64_t x1, int64_t y1, int64_t plane) {
    int64_t dx = llabs(x1 - x0), sx = x0 < x1 ? 1 : -1, dy = -llabs(y1 - y0), sy = y0 < y1 ? 1 : -1;
    int64_t err = dx + dy, n = 0;
    if (in_plane(c, x0, y0, plane)) n++;
    while (x0 != x1 || y0 != y1) {
        int64_t e2 = 2 * err;
        if (e2 - dy > dx - e2) { err += dy; x0 += sx; } else { err += dx; y0 += sy; }
        if (in_plane(c, x0, y0, plane)) n++;
    }
    return n;
}
ORC_API int orc_max_pixels(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t n, int64_t* n_max) {
    for (int64_t i = 0; i < n; i++) {                                /* :43-65 */
        const char* t = (const char*)tracks + i * L->itemsize;
        int64_t x0, y0, x1, y1, plane;
        if (!segment_pixels(c, L, t, &x0, &y0, &x1, &y1, &plane)) continue;
        int64_t v = num_active_pixels(c, x0, y0, x1, y1, plane);
        if (v > *n_max) *n_max = v;
    }
    return 0;
}
ORC_API int orc_get_pixels(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t n,
                           int32_t* active, int32_t max_active, int32_t* neigh, int32_t* nrad, int32_t P,
                           double* n_pixels_list, int32_t radius) {
    for (int64_t it = 0; it < n; it++) {                             /* :67-109 */
        const char* t = (const char*)tracks + it * L->itemsize;
        int32_t* act = active + it * max_active;
        int32_t* nb = neigh + it * P;
        int32_t* nr = nrad + it * P;
        int64_t x0, y0, x1, y1, plane;
        if (segment_pixels(c, L, t, &x0, &y0, &x1, &y1, &plane)) {
            /* get_active_pixels :157-199 (index = step count; guarded against the row length) */
            int64_t dx = llabs(x1 - x0), sx = x0 < x1 ? 1 : -1, dy = -llabs(y1 - y0), sy = y0 < y1 ? 1 : -1;
            int64_t err = dx + dy, i = 0;
            if (in_plane(c, x0, y0, plane) && i < max_active) act[i] = (int32_t)pixel2id(c, x0, y0, plane);
            while (x0 != x1 || y0 != y1) {
                i++;
                int64_t e2 = 2 * err;
                if (e2 - dy > dx - e2) { err += dy; x0 += sx; } else { err += dx; y0 += sy; }
                if (in_plane(c, x0, y0, plane) && i < max_active) act[i] = (int32_t)pixel2id(c, x0, y0, plane);
            }
        }
        /* get_neighboring_pixels :201-272 */
        int64_t count = 0;
        for (int pix = 0; pix < max_active; pix++) {
            if (act[pix] == -1) continue;
            for (int xr = -radius; xr <= radius; xr++)
                for (int yr = -radius; yr <= radius; yr++) {
                    int64_t ax, ay, pl;
                    id2pixel(c, act[pix], &ax, &ay, &pl);
                    int64_t nx = ax + xr, ny = ay + yr;
                    if (!in_plane(c, nx, ny, pl)) continue;
                    int64_t np_ = pixel2id(c, nx, ny, pl);
                    int unique = 1;
                    for (int k = 0; k < P; k++) if (nb[k] == np_) { unique = 0; break; }
                    if (!unique) continue;
                    int adx = abs(xr), ady = abs(yr);
                    int dmax = adx > ady ? adx : ady, dmin = adx > ady ? ady : adx, dsum = dmax + dmin;
                    int dist = -1;
                    if (dsum > c->max_neighbor_backtrack_distance) dist = -1;
                    else if (dsum <= 1) dist = dsum;
                    else if (dsum == 2) dist = (dmax == 1) ? 2 : 3;
                    else if (dsum == 3) dist = (dmax == 2) ? 4 : 5;
                    else if (dsum == 4) dist = (dmax == 2) ? 6 : (dmax == 3 ? 7 : 8);
                    else dist = -1;
                    if (count < P) { nb[count] = (int32_t)np_; nr[count] = dist; }
                    count++;
                }
        }
        n_pixels_list[it] = (double)count;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* detsim.py:18-40 time_intervals                                                         */
/* ------------------------------------------------------------------------------------ */
ORC_API int orc_time_intervals(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t n,
                               double* track_starts, int64_t* time_max) {
    for (int64_t i = 0; i < n; i++) {
        const char* t = (const char*)tracks + i * L->itemsize;
        double t_end = (double)llrint((fld_get(L, t, LSB_F_T_END) + 1) / c->time_sampling) * c->time_sampling;
        double t_start = (double)llrint((fld_get(L, t, LSB_F_T_START) - c->time_padding) / c->time_sampling) * c->time_sampling;
        double t_length = t_end - t_start;
        track_starts[i] = t_start;
        int64_t v = (int64_t)ceil(t_length / c->time_sampling);
        if (v > *time_max) *time_max = v;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* glue: cli/simulate_pixels.py:953-956, 1021-1025                                        */
/* ------------------------------------------------------------------------------------ */
static int cmp_i32(const void* a, const void* b) { int32_t x = *(const int32_t*)a, y = *(const int32_t*)b; return (x > y) - (x < y); }
ORC_API int64_t orc_unique_pixels(const int32_t* pixels, int64_t n, int32_t* unique_out) {
    int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    memcpy(tmp, pixels, sizeof(int32_t) * (size_t)n);
    qsort(tmp, (size_t)n, sizeof(int32_t), cmp_i32);
    int64_t u = 0;
    for (int64_t i = 0; i < n; i++) {
        if (tmp[i] == -1) continue;
        if (u == 0 || unique_out[u - 1] != tmp[i]) unique_out[u++] = tmp[i];
    }
    free(tmp);
    return u;
}
ORC_API int orc_pixel_index_map(const int32_t* pixels, int64_t n, const int32_t* unique_pix, int64_t U, int64_t* map) {
    for (int64_t i = 0; i < n; i++) {
        int32_t p = pixels[i];
        int64_t lo = 0, hi = U;
        while (lo < hi) { int64_t mid = (lo + hi) / 2; if (unique_pix[mid] < p) lo = mid + 1; else hi = mid; }
        map[i] = (lo < U && unique_pix[lo] == p) ? lo : -1;   /* -1 ids are not in unique_pix -> stay -1 */
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* detsim.py helpers                                                                      */
/* ------------------------------------------------------------------------------------ */
typedef struct { double v[3]; } vec3;
static inline double resp_at(const void* resp, int f64, int64_t idx) {
    return f64 ? ((const double*)resp)[idx] : (double)((const float*)resp)[idx];
}
/* get_closest_waveform :193-218 */
static inline double closest_waveform(const lsb_consts* c, double x, double y, double t, const void* resp,
                                      int Rx, int Ry, int Rt, int f64) {
    long long i = llrint((x / c->response_bin_size) - 0.5);
    long long j = llrint((y / c->response_bin_size) - 0.5);
    long long k = llrint(t / c->response_sampling);
    if (0 <= i && i < Rx && 0 <= j && j < Ry && 0 <= k && k < Rt) return resp_at(resp, f64, (i * Ry + j) * (int64_t)Rt + k);
    return 0;
}
/* overlapping_segment :220-256; c32 = coordinate fields are float32 (Numba keeps end-start float32) */
static void overlapping_segment(double x, double y, const double* start, const double* end, double radius, int c32,
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

typedef struct {
    int valid;            /* pixel id valid and sub-segment non-empty */
    double x_p, y_p, t_start, z_anode;
    double sub_start[3], dir[3];
    double step, charge, sig_t, sig_l;
    int64_t nstep;
    int s32;              /* sigma fields float32 */
} pair_geom;

/* detsim.py:275-322: everything of tracks_current_mc that does not depend on the tick */
static void mc_pair_geometry(const lsb_consts* c, const lsb_track_layout* L, const char* t, int32_t pID,
                             int Rx, int Ry, pair_geom* g) {
    g->valid = 0;
    int64_t px, py, pl;
    id2pixel(c, pID, &px, &py, &pl);
    if (!(px >= 0 && py >= 0)) return;      /* always true with Python modulo; kept literally (:280) */
    /* pID == -1 (row padding): Python modulo gives pixel (Nx-1, Ny-1) and plane -1, and the negative
     * index wraps to the LAST TPC (detsim.py:185) -- the reference computes a current for it. */
    if (pl < 0) pl += c->n_tpc;
    if (pl < 0 || pl >= c->n_tpc) return;
    const double (*b)[2] = c->tpc_borders[pl];
    g->x_p = px * c->pixel_pitch + b[0][0] + c->pixel_pitch / 2;
    g->y_p = py * c->pixel_pitch + b[1][0] + c->pixel_pitch / 2;
    int c32 = fld_is_f32(L, LSB_F_X_START) && fld_is_f32(L, LSB_F_Y_START) && fld_is_f32(L, LSB_F_Z_START) &&
              fld_is_f32(L, LSB_F_X_END) && fld_is_f32(L, LSB_F_Y_END) && fld_is_f32(L, LSB_F_Z_END);
    g->s32 = fld_is_f32(L, LSB_F_TRAN_DIFF) && fld_is_f32(L, LSB_F_LONG_DIFF);
    double start[3], end[3];
    double zs = fld_get(L, t, LSB_F_Z_START), ze = fld_get(L, t, LSB_F_Z_END);
    if (zs < ze) {
        start[0] = fld_get(L, t, LSB_F_X_START); start[1] = fld_get(L, t, LSB_F_Y_START); start[2] = zs;
        end[0] = fld_get(L, t, LSB_F_X_END); end[1] = fld_get(L, t, LSB_F_Y_END); end[2] = ze;
    } else {
        end[0] = fld_get(L, t, LSB_F_X_START); end[1] = fld_get(L, t, LSB_F_Y_START); end[2] = zs;
        start[0] = fld_get(L, t, LSB_F_X_END); start[1] = fld_get(L, t, LSB_F_Y_END); start[2] = ze;
    }
    g->t_start = (double)llrint((fld_get(L, t, LSB_F_T_START) - fld_get(L, t, LSB_F_T0_START) - c->time_padding) /
                                c->time_sampling) * c->time_sampling;
    double seg[3];
    for (int k = 0; k < 3; k++) seg[k] = R32(end[k] - start[k], c32);
    double length = R32(sqrt(R32(R32(R32(seg[0] * seg[0], c32) + R32(seg[1] * seg[1], c32), c32) + R32(seg[2] * seg[2], c32), c32)), c32);
    for (int k = 0; k < 3; k++) g->dir[k] = R32(seg[k] / length, c32);
    g->sig_t = fld_get(L, t, LSB_F_TRAN_DIFF);
    g->sig_l = fld_get(L, t, LSB_F_LONG_DIFF);
    double impact = sqrt((double)((int64_t)Rx * Rx + (int64_t)Ry * Ry)) * c->response_bin_size;
    double ss[3], se[3];
    overlapping_segment(g->x_p, g->y_p, start, end, impact, c32, ss, se);
    double sub[3] = {se[0] - ss[0], se[1] - ss[1], se[2] - ss[2]};
    double sub_len = sqrt(sub[0] * sub[0] + sub[1] * sub[1] + sub[2] * sub[2]);
    if (sub_len == 0) return;
    long long ns = llrint(sub_len / c->min_step_size);
    g->nstep = ns > 1 ? ns : 1;
    g->step = sub_len / (double)g->nstep;
    g->charge = fld_get(L, t, LSB_F_N_ELECTRONS) * (sub_len / length) / (double)(g->nstep * c->mc_sample_multiplier);
    for (int k = 0; k < 3; k++) g->sub_start[k] = ss[k];
    int64_t plane = (int64_t)fld_get(L, t, LSB_F_PIXEL_PLANE);
    if (plane < 0 || plane >= c->n_tpc) return;
    g->z_anode = c->tpc_borders[plane][2][0];
    g->valid = 1;
}

/* detsim.py:258-348 tracks_current_mc, replay order: one thread at a time, ticks of a
 * (segment,pixel) consumed in ascending order from the shared state rng_states[itrk+S*ipix]. */
ORC_API int orc_tracks_current_mc(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t S,
                                  const int32_t* pixels, int32_t P, float* signals, int32_t T,
                                  const void* resp, int32_t Rx, int32_t Ry, int32_t Rt, int32_t f64,
                                  uint64_t* rng_states, int32_t mode) {
    int64_t npair = S * (int64_t)P;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t pr = 0; pr < npair; pr++) {
        int64_t itrk = pr / P; int32_t ipix = (int32_t)(pr % P);
        const char* t = (const char*)tracks + itrk * L->itemsize;
        int32_t pID = pixels[itrk * P + ipix];
        pair_geom g;
        mc_pair_geometry(c, L, t, pID, Rx, Ry, &g);
        if (!g.valid) continue;
        uint64_t* st = rng_states + 2 * (itrk + S * ipix);
        float* out = signals + (itrk * P + ipix) * (int64_t)T;
        if (mode == 1) {
            /* reference order (:324-348) */
            for (int32_t it = 0; it < T; it++) {
                double time_tick = g.t_start + it * c->time_sampling;
                if (time_tick < 0) continue;                                   /* :299-300 */
                double total = 0;
                for (int64_t istep = 0; istep < g.nstep; istep++)
                    for (int m = 0; m < c->mc_sample_multiplier; m++) {
                        double x = g.sub_start[0] + g.step * (istep + 0.5) * g.dir[0];
                        double y = g.sub_start[1] + g.step * (istep + 0.5) * g.dir[1];
                        double z = g.sub_start[2] + g.step * (istep + 0.5) * g.dir[2];
                        z += R32((double)rng_normal_f32(st) * g.sig_l, g.s32);   /* :331 */
                        double t0 = fabs(z - g.z_anode) / c->v_drift - c->time_window;
                        if (!(t0 < time_tick && time_tick < t0 + c->time_window)) continue;
                        x += R32((double)rng_normal_f32(st) * g.sig_t, g.s32);
                        y += R32((double)rng_normal_f32(st) * g.sig_t, g.s32);
                        double x_dist = fabs(g.x_p - x), y_dist = fabs(g.y_p - y);
                        if (x_dist > c->response_bin_size * Rx) continue;
                        if (y_dist > c->response_bin_size * Ry) continue;
                        total += g.charge * closest_waveform(c, x_dist, y_dist, time_tick - t0, resp, Rx, Ry, Rt, f64);
                    }
                out[it] = (float)total;
            }
        } else {
            /* "cloud" order (the product's production mode, see include/larndsim_b200.h):
             * z, x, y normals drawn once per step, cloud applied to all ticks */
            double* tot = (double*)calloc((size_t)T, sizeof(double));
            for (int64_t istep = 0; istep < g.nstep; istep++)
                for (int m = 0; m < c->mc_sample_multiplier; m++) {
                    double x = g.sub_start[0] + g.step * (istep + 0.5) * g.dir[0];
                    double y = g.sub_start[1] + g.step * (istep + 0.5) * g.dir[1];
                    double z = g.sub_start[2] + g.step * (istep + 0.5) * g.dir[2];
                    z += R32((double)rng_normal_f32(st) * g.sig_l, g.s32);
                    double t0 = fabs(z - g.z_anode) / c->v_drift - c->time_window;
                    x += R32((double)rng_normal_f32(st) * g.sig_t, g.s32);
                    y += R32((double)rng_normal_f32(st) * g.sig_t, g.s32);
                    double x_dist = fabs(g.x_p - x), y_dist = fabs(g.y_p - y);
                    if (x_dist > c->response_bin_size * Rx) continue;
                    if (y_dist > c->response_bin_size * Ry) continue;
                    for (int32_t it = 0; it < T; it++) {
                        double time_tick = g.t_start + it * c->time_sampling;
                        if (time_tick < 0) continue;
                        if (!(t0 < time_tick && time_tick < t0 + c->time_window)) continue;
                        tot[it] += g.charge * closest_waveform(c, x_dist, y_dist, time_tick - t0, resp, Rx, Ry, Rt, f64);
                    }
                }
            for (int32_t it = 0; it < T; it++) {
                double time_tick = g.t_start + it * c->time_sampling;
                if (time_tick < 0) continue;
                out[it] = (float)tot[it];
            }
            free(tot);
        }
    }
    return 0;
}

/* Roofline accounting: N_sp = sample points, N_fma = (sample,tick) pairs that pass
 * every gate, for the cloud formulation with sigma as stored (uses the same RNG draws). */
ORC_API int orc_count_mc_work(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t S,
                              const int32_t* pixels, int32_t P, int64_t* n_pairs, int64_t* n_samples) {
    int64_t np_ = 0, ns = 0;
    for (int64_t pr = 0; pr < S * (int64_t)P; pr++) {
        int64_t itrk = pr / P; int32_t ipix = (int32_t)(pr % P);
        pair_geom g;
        mc_pair_geometry(c, L, (const char*)tracks + itrk * L->itemsize, pixels[itrk * P + ipix], 45, 45, &g);
        if (!g.valid) continue;
        np_++; ns += g.nstep * c->mc_sample_multiplier;
    }
    *n_pairs = np_; *n_samples = ns;
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* detsim.py:351-453 tracks_current (deterministic)                                       */
/* ------------------------------------------------------------------------------------ */
static inline double sgn(double x) { return x >= 0 ? 1.0 : -1.0; }     /* :455-466 */
/* rho :120-159 with _b :114-118.  sigmas float32-typed when s32 (products of two float32 stay float32). */
/* erf(hi) - erf(lo), lo <= hi, of rho (detsim.py:149-151).  The reference writes -erf(lo) + erf(hi): when both arguments
 * lie on the same side of zero beyond ~4 (grid points past the ends of the segment) both terms are +-1 to within 1e-8 and
 * the difference keeps few or no digits -- a one-ulp difference between two libm's erf then changes the result at the
 * 1e-5..1 level (and the reference returns 0 when it cancels completely).  The same difference is taken here between the
 * complementary functions on that side, which keeps full relative precision; wherever the reference's form is
 * well-conditioned the two agree to rounding.  The CUDA kernel (csrc/current.cuh rho_fast / rho_dev) does the same. */
static inline double erf_diff(double lo, double hi) {
    if (lo >= 0.0) return erfc(lo) - erfc(hi);
    if (hi <= 0.0) return erfc(-hi) - erfc(-lo);
    return erf(hi) - erf(lo);
}

static double rho(const double* pt, double q, const double* start, const double* sig, const double* seg, int s32, int c32) {
    /* Numba typing: segment/start float32 (c32), sigmas float32 (s32), point float64; an int literal
     * times a float32 is float64 (2*sigma), float32*float32 stays float32, x**2 keeps the base type. */
    double Dx = seg[0], Dy = seg[1], Dz = seg[2];
    double Dr = R32(sqrt(R32(R32(R32(Dx * Dx, c32) + R32(Dy * Dy, c32), c32) + R32(Dz * Dz, c32), c32)), c32);
    double ux = R32(Dx / Dr, c32), uy = R32(Dy / Dr, c32), uz = R32(Dz / Dr, c32);
    double a = R32(ux * ux, c32) / (2 * sig[0] * sig[0]) + R32(uy * uy, c32) / (2 * sig[1] * sig[1]) +
               R32(uz * uz, c32) / (2 * sig[2] * sig[2]);
    double sprod = R32(R32(sig[0] * sig[1], s32) * sig[2], s32);
    double factor = q / Dr / (sprod * sqrt(8 * M_PI * M_PI * M_PI));
    double sqrt_a_2 = 2 * sqrt(a);
    double x = pt[0], y = pt[1], z = pt[2];
    double b = -((x - start[0]) / R32(sig[0] * sig[0], s32) * ux +
                 (y - start[1]) / R32(sig[1] * sig[1], s32) * uy +
                 (z - start[2]) / R32(sig[2] * sig[2], s32) * uz);
    double delta = (x - start[0]) * (x - start[0]) / (2 * sig[0] * sig[0]) +
                   (y - start[1]) * (y - start[1]) / (2 * sig[1] * sig[1]) +
                   (z - start[2]) * (z - start[2]) / (2 * sig[2] * sig[2]);
    double integral = sqrt(M_PI) * erf_diff(b / sqrt_a_2, (b + 2 * a * Dr) / sqrt_a_2) / sqrt_a_2;
    double expo = 0;
    if (factor != 0 && integral != 0)
        expo = exp(b * b / (4 * a) - delta + log(factor) + log(integral));
    return expo;
}
/* z_interval :42-112 */
static void z_interval(const double* sp, const double* ep, double x_p, double y_p, double tol, int c32,
                       double* z_poca, double* z_lo, double* z_hi) {
    *z_poca = *z_lo = *z_hi = 0;
    const double *start, *end;
    if (sp[0] > ep[0]) { start = ep; end = sp; } else if (sp[0] < ep[0]) { start = sp; end = ep; } else return;
    double xs = start[0], ys = start[1], xe = end[0], ye = end[1];
    double dxe = R32(xe - xs, c32);
    double m = R32(R32(ye - ys, c32) / dxe, c32);
    double q = R32(R32(R32(xe * ys, c32) - R32(xs * ye, c32), c32) / dxe, c32);
    double a = m, b = -1, cc = q;
    /* x_poca = (b*(b*x_p - a*y_p) - a*c)/(a*a+b*b): x_p float64 -> float64 */
    double x_poca = (b * (b * x_p - a * y_p) - a * cc) / (R32(a * a, c32) + b * b);
    double d0 = R32(end[0] - start[0], c32), d1 = R32(end[1] - start[1], c32), d2 = R32(end[2] - start[2], c32);
    double length = R32(sqrt(R32(R32(R32(d0 * d0, c32) + R32(d1 * d1, c32), c32) + R32(d2 * d2, c32), c32)), c32);
    double dir3[3] = {R32(d0 / length, c32), R32(d1 / length, c32), R32(d2 / length, c32)};
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
    *z_poca = start[2] + (x_poca - start[0]) / dir3[0] * dir3[2];
    if (tol > doca) {
        double dx2 = R32(xe - xs, c32), dy2 = R32(ye - ys, c32);
        double length2D = R32(sqrt(R32(R32(dx2 * dx2, c32) + R32(dy2 * dy2, c32), c32)), c32);
        double dir2D0 = R32(d0 / length2D, c32);
        double deltaL2D = sqrt(tol * tol - doca * doca);
        double x_plus = x_poca + deltaL2D * dir2D0, x_minus = x_poca - deltaL2D * dir2D0;
        double plusL = (x_plus - start[0]) / dir3[0], minusL = (x_minus - start[0]) / dir3[0];
        double plusZ = start[2] + dir3[2] * plusL, minusZ = start[2] + dir3[2] * minusL;
        *z_lo = fmin(minusZ, plusZ); *z_hi = fmax(minusZ, plusZ);
        return;
    }
    *z_poca = *z_lo = *z_hi = 0;
}
ORC_API int orc_tracks_current(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t S,
                               const int32_t* pixels, int32_t P, float* signals, int32_t T,
                               const void* resp, int32_t Rx, int32_t Ry, int32_t Rt, int32_t f64) {
    int64_t npair = S * (int64_t)P;
    int NP = c->sampled_points;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t pr = 0; pr < npair; pr++) {
        int64_t itrk = pr / P; int32_t ipix = (int32_t)(pr % P);
        const char* t = (const char*)tracks + itrk * L->itemsize;
        int32_t pID = pixels[itrk * P + ipix];
        int64_t px, py, pl;
        id2pixel(c, pID, &px, &py, &pl);
        if (pl < 0) pl += c->n_tpc;
        if (pl < 0 || pl >= c->n_tpc) continue;
        const double (*bd)[2] = c->tpc_borders[pl];
        double x_p = px * c->pixel_pitch + bd[0][0] + c->pixel_pitch / 2;
        double y_p = py * c->pixel_pitch + bd[1][0] + c->pixel_pitch / 2;
        int c32 = fld_is_f32(L, LSB_F_X_START) && fld_is_f32(L, LSB_F_Y_START) && fld_is_f32(L, LSB_F_Z_START) &&
                  fld_is_f32(L, LSB_F_X_END) && fld_is_f32(L, LSB_F_Y_END) && fld_is_f32(L, LSB_F_Z_END);
        int s32 = fld_is_f32(L, LSB_F_TRAN_DIFF) && fld_is_f32(L, LSB_F_LONG_DIFF);
        double start[3], end[3];
        double zs = fld_get(L, t, LSB_F_Z_START), ze = fld_get(L, t, LSB_F_Z_END);
        if (zs < ze) {
            start[0] = fld_get(L, t, LSB_F_X_START); start[1] = fld_get(L, t, LSB_F_Y_START); start[2] = zs;
            end[0] = fld_get(L, t, LSB_F_X_END); end[1] = fld_get(L, t, LSB_F_Y_END); end[2] = ze;
        } else {
            end[0] = fld_get(L, t, LSB_F_X_START); end[1] = fld_get(L, t, LSB_F_Y_START); end[2] = zs;
            start[0] = fld_get(L, t, LSB_F_X_END); start[1] = fld_get(L, t, LSB_F_Y_END); start[2] = ze;
        }
        double seg[3];
        for (int k = 0; k < 3; k++) seg[k] = R32(end[k] - start[k], c32);
        double length = R32(sqrt(R32(R32(R32(seg[0] * seg[0], c32) + R32(seg[1] * seg[1], c32), c32) + R32(seg[2] * seg[2], c32), c32)), c32);
        double dir[3];
        for (int k = 0; k < 3; k++) dir[k] = R32(seg[k] / length, c32);
        double sig[3] = {fld_get(L, t, LSB_F_TRAN_DIFF), fld_get(L, t, LSB_F_TRAN_DIFF), fld_get(L, t, LSB_F_LONG_DIFF)};
        /* impact_factor :388-389 */
        double s5x = 5 * sig[0], s5y = 5 * sig[1];              /* int * float32 -> float64 */
        double imp1 = sqrt(s5x * s5x + s5y * s5y);
        double imp2 = sqrt(c->pixel_pitch * c->pixel_pitch + c->pixel_pitch * c->pixel_pitch) / 2;
        double impact = fmax(imp1, imp2) * 2;
        double z_poca, z_start, z_end;
        z_interval(start, end, x_p, y_p, impact, c32, &z_poca, &z_start, &z_end);
        if (z_poca == 0) continue;
        double z_start_int = z_start - 4 * sig[2];
        double z_end_int = z_end + 4 * sig[2];
        /* track_point :161-178 */
        double l0 = (z_start - start[2]) / dir[2], l1 = (z_end - start[2]) / dir[2];
        double x_start = start[0] + l0 * dir[0], y_start = start[1] + l0 * dir[1];
        double x_end = start[0] + l1 * dir[0], y_end = start[1] + l1 * dir[1];
        double y_step = (fabs(y_end - y_start) + 8 * sig[1]) / (NP - 1);
        double x_step = (fabs(x_end - x_start) + 8 * sig[0]) / (NP - 1);
        double z_sampling = c->time_sampling / 2.;
        int64_t zc = (int64_t)ceil(fabs(z_end_int - z_start_int) / z_sampling);
        int64_t z_steps = zc > NP ? zc : NP;
        double z_step = (z_end_int - z_start_int) / (z_steps - 1);
        double t_start = (double)llrint((fld_get(L, t, LSB_F_T_START) - fld_get(L, t, LSB_F_T0_START) - c->time_padding) /
                                        c->time_sampling) * c->time_sampling;
        double q = fld_get(L, t, LSB_F_N_ELECTRONS);
        int64_t plane = (int64_t)fld_get(L, t, LSB_F_PIXEL_PLANE);
        if (plane < 0 || plane >= c->n_tpc) continue;
        double z_anode = c->tpc_borders[plane][2][0];
        float* out = signals + (itrk * P + ipix) * (int64_t)T;
        for (int32_t it = 0; it < T; it++) {
            double time_tick = t_start + it * c->time_sampling;
            if (time_tick < 0.) continue;
            double total = 0;
            int written = 0;
            for (int64_t iz = 0; iz < z_steps; iz++) {
                double z = z_start_int + iz * z_step;
                double t0 = fabs(z - z_anode) / c->v_drift - c->time_window;
                if (!(t0 < time_tick && time_tick < t0 + c->time_window)) continue;
                for (int ix = 0; ix < NP; ix++) {
                    double x = x_start + sgn(dir[0]) * (ix * x_step - 4 * sig[0]);
                    double x_dist = fabs(x_p - x);
                    if (x_dist > c->response_bin_size * Rx) continue;
                    for (int iy = 0; iy < NP; iy++) {
                        double y = y_start + sgn(dir[1]) * (iy * y_step - 4 * sig[1]);
                        double y_dist = fabs(y_p - y);
                        if (y_dist > c->response_bin_size * Ry) continue;
                        double pt[3] = {x, y, z};
                        double charge = rho(pt, q, start, sig, seg, s32, c32) * fabs(x_step) * fabs(y_step) * fabs(z_step);
                        total += closest_waveform(c, x_dist, y_dist, time_tick - t0, resp, Rx, Ry, Rt, f64) * charge;
                    }
                    written = 1;                               /* store inside the ix loop (:453) */
                }
            }
            if (written) out[it] = (float)total;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* detsim.py:529-607 track/pixel maps                                                     */
/* ------------------------------------------------------------------------------------ */
ORC_API int orc_get_track_pixel_map(int64_t* tpm, int32_t K, const int32_t* unique_pix, int64_t U,
                                    const int32_t* pixels, int64_t S, int32_t P) {
    for (int64_t index = 0; index < U; index++) {
        int32_t upix = unique_pix[index];
        for (int64_t itrk = 0; itrk < S; itrk++)
            for (int ipix = 0; ipix < P; ipix++) {
                if (upix != pixels[itrk * P + ipix]) continue;
                int imap = 0;
                while (imap < K && tpm[index * K + imap] != -1 && tpm[index * K + imap] != itrk) imap++;
                if (imap < K) tpm[index * K + imap] = itrk;
            }
    }
    return 0;
}
ORC_API int orc_get_track_pixel_map2(int64_t* tpm, int32_t K, const int32_t* unique_pix, int64_t U,
                                     const int32_t* pixels, const int32_t* distances, int64_t S, int32_t P,
                                     int32_t max_distance) {
#pragma omp parallel for schedule(static)
    for (int64_t index = 0; index < U; index++) {
        int32_t upix = unique_pix[index];
        for (int target = 0; target < max_distance; target++)
            for (int64_t itrk = 0; itrk < S; itrk++)
                for (int ipix = 0; ipix < P; ipix++) {
                    if (upix != pixels[itrk * P + ipix]) continue;
                    if (distances[itrk * P + ipix] == target) {
                        int imap = 0;
                        while (imap < K) {
                            if (tpm[index * K + imap] == itrk) { imap = -1; break; }
                            if (tpm[index * K + imap] == -1) break;
                            imap++;
                        }
                        if (imap >= 0 && imap < K) tpm[index * K + imap] = itrk;
                    }
                    break;                                    /* :607: first match in the row ends the row */
                }
    }
    return 0;
}

/* detsim.py:468-527 sum_pixel_signals, threads in grid order (segment-major): the sum order of
 * the CUDA simulator; the reference's atomics leave the order undefined. */
ORC_API int orc_sum_pixel_signals(const lsb_consts* c, double* pixels_signals, int64_t U, int32_t Tt,
                                  const float* signals, int64_t S, int32_t P, int32_t T,
                                  const double* track_starts, const int64_t* pim, const int64_t* tpm, int32_t K,
                                  double* pts, double* overflow_flag) {
    (void)U;
    for (int64_t itrk = 0; itrk < S; itrk++)
        for (int ipix = 0; ipix < P; ipix++) {
            int64_t pixel_index = pim[itrk * P + ipix];
            long long start_tick = llrint(track_starts[itrk] / c->time_sampling);
            if (pixel_index < 0) continue;
            int counter = -99;
            for (int k = 0; k < K; k++) {
                if (itrk == tpm[pixel_index * K + k]) { counter = k; break; }
            }
            if (counter >= 0) {
                for (int itick = 0; itick < T; itick++) {
                    long long itime = start_tick + itick;
                    if (itime < Tt && itime > -1) {
                        double s = (double)signals[(itrk * P + ipix) * (int64_t)T + itick];
                        pixels_signals[pixel_index * Tt + itime] += s;
                        pts[(pixel_index * Tt + itime) * K + counter] += s;
                    }
                }
            } else {
                overflow_flag[pixel_index] = 1;
            }
        }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* fee.py:499-655                                                                         */
/* ------------------------------------------------------------------------------------ */
ORC_API int orc_digitize(const lsb_consts* c, const double* integral, const double* gain_list, int64_t n, double* adcs) {
    double gain_s = c->gain * c->unit_mV / c->unit_e;               /* default argument fee.py:499 */
    for (int64_t i = 0; i < n; i++) {
        double g = gain_list ? gain_list[i] : gain_s;
        double v = integral[i] * g + c->v_pedestal * c->unit_mV - c->v_cm * c->unit_mV;
        v = v > 0 ? v : 0;
        double a = nearbyint(v * c->adc_counts / (c->v_ref * c->unit_mV - c->v_cm * c->unit_mV));   /* np.around: half-even */
        adcs[i] = a < c->adc_counts - 1 ? a : c->adc_counts - 1;
    }
    return 0;
}

ORC_API int orc_get_adc_values(const lsb_consts* c, const double* pixels_signals, const double* pst,
                               int64_t U, int32_t Tt, int32_t K, const double* time_ticks, int32_t n_tt,
                               double* adc_list, double* adc_ticks_list, int32_t A, double time_padding,
                               uint64_t* rng_states, double* cf, const double* thresholds) {
    const double TS = c->time_sampling, BR = c->buffer_risetime, e = c->unit_e;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t ip = 0; ip < U; ip++) {
        const double* curre = pixels_signals + ip * Tt;
        uint64_t* st = rng_states + 2 * ip;
        int64_t ic = 0, iadc = 0, adc_busy = 0, last_reset = 0;
        double true_q = 0;
        double q_sum = (double)rng_normal_f32(st) * c->reset_noise_charge * e;       /* :557 */
        while (ic < Tt || adc_busy > 0) {
            if (iadc >= c->max_adc_values) break;                                    /* :561-563 */
            if (iadc >= A) break;                                                    /* guard: rows of the outputs */
            double q = 0;
            double* cfr = cf + (ip * A + iadc) * K;
            if (BR > 0) {                                                            /* :566-573 */
                int64_t conv_start = (int64_t)floor(ic - 10 * BR / TS);
                if (last_reset > conv_start) conv_start = last_reset;
                int64_t jend = ic + 1 < Tt ? ic + 1 : Tt;
                for (int64_t jc = conv_start; jc < jend; jc++) {
                    double w = exp((jc - ic) * TS / BR) * (1 - exp(-TS / BR));
                    q += curre[jc] * TS * w;
                    for (int k = 0; k < K; k++) cfr[k] += pst[(ip * Tt + jc) * K + k] * TS * w;
                }
            } else if (ic < Tt) {                                                    /* :575-578 */
                q += curre[ic] * TS;
                for (int k = 0; k < K; k++) cfr[k] += pst[(ip * Tt + ic) * K + k] * TS;
            }
            q_sum += q; true_q += q;
            double q_noise = (double)rng_normal_f32(st) * c->uncorrelated_noise_charge * e;
            double disc_noise = (double)rng_normal_f32(st) * c->discriminator_noise * e;
            if (adc_busy > 0) adc_busy--;
            if (q_sum + q_noise >= thresholds[ip] + disc_noise && adc_busy == 0) {   /* :589 */
                int64_t interval = llrint((3 * c->clock_cycle + c->adc_hold_delay * c->clock_cycle) / TS);
                int64_t integrate_end = ic + interval;
                ic++;
                while (ic <= integrate_end) {                                        /* :595-613 */
                    q = 0;
                    if (BR > 0) {
                        int64_t conv_start = (int64_t)floor(ic - 10 * BR / TS);
                        if (last_reset > conv_start) conv_start = last_reset;
                        int64_t jend = ic + 1 < Tt ? ic + 1 : Tt;
                        for (int64_t jc = conv_start; jc < jend; jc++) {
                            double w = exp((jc - ic) * TS / BR) * (1 - exp(-TS / BR));
                            q += curre[jc] * TS * w;
                            for (int k = 0; k < K; k++) cfr[k] += pst[(ip * Tt + jc) * K + k] * TS * w;
                        }
                    } else if (ic < Tt) {
                        q += curre[ic] * TS;
                        for (int k = 0; k < K; k++) cfr[k] += pst[(ip * Tt + ic) * K + k] * TS;
                    }
                    q_sum += q; true_q += q; ic++;
                }
                double adc = q_sum + (double)rng_normal_f32(st) * c->uncorrelated_noise_charge * e;   /* :616 */
                disc_noise = (double)rng_normal_f32(st) * c->discriminator_noise * e;
                if (adc < thresholds[ip] + disc_noise) {                             /* :619-627 */
                    ic += llrint(c->reset_cycles * c->clock_cycle / TS);
                    q_sum = (double)rng_normal_f32(st) * c->reset_noise_charge * e;
                    true_q = 0;
                    for (int k = 0; k < K; k++) cfr[k] = 0;
                    last_reset = ic;
                    continue;
                }
                if (true_q > 0) for (int k = 0; k < K; k++) cfr[k] /= true_q;        /* :633-635 */
                adc_list[ip * A + iadc] = adc;
                int64_t crossing = ic < n_tt - 1 ? ic : n_tt - 1;                    /* :639 */
                int64_t post = ic - crossing > 0 ? ic - crossing : 0;
                adc_ticks_list[ip * A + iadc] = time_ticks[crossing] + time_padding - 2 + post;
                ic += llrint(c->reset_cycles * c->clock_cycle / TS);
                last_reset = ic;
                adc_busy = llrint(c->adc_busy_delay * c->clock_cycle / TS);
                q_sum = (double)rng_normal_f32(st) * c->reset_noise_charge * e;
                true_q = 0;
                iadc++;
                continue;
            }
            ic++;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* lightLUT.py:15-136                                                                     */
/* ------------------------------------------------------------------------------------ */
ORC_API int orc_calculate_light_incidence(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t S,
                                          const void* lut, const lsb_lut_layout* LL, void* linc, const lsb_linc_layout* LI,
                                          int32_t ndet, int32_t* voxel, const double* eff, const int64_t* ch2tpc) {
    for (int64_t itrk = 0; itrk < S; itrk++) {
        const char* t = (const char*)tracks + itrk * L->itemsize;
        double pos[3] = {fld_get(L, t, LSB_F_X), fld_get(L, t, LSB_F_Y), fld_get(L, t, LSB_F_Z)};
        double n_photons = fld_get(L, t, LSB_F_N_PHOTONS);
        int64_t itpc = (int64_t)fld_get(L, t, LSB_F_PIXEL_PLANE);
        int64_t imod = py_div(itpc, 2);
        if (itpc == c->default_plane_index) continue;
        if (itpc < 0 || itpc >= c->n_tpc) continue;
        const double (*b)[2] = c->tpc_borders[itpc];
        int is_even = b[2][1] > b[2][0];                                             /* get_voxel :28-63 */
        double x_min = b[0][0] - 2e-2, x_max = b[0][1] + 2e-2, y_min = b[1][0] - 2e-2, y_max = b[1][1] + 2e-2;
        double z_min = b[2][0] - 2e-2, z_max = b[2][1] + 2e-2;
        int64_t i, j, k;
        if (is_even) i = (int64_t)((pos[0] - x_min) / (x_max - x_min) * LL->shape[0]);
        else i = (int64_t)((x_max - pos[0]) / (x_max - x_min) * LL->shape[0]);
        j = (int64_t)((y_max - pos[1]) / (y_max - y_min) * LL->shape[1]);
        k = (int64_t)((pos[2] - z_min) / (z_max - z_min) * LL->shape[2]);
        i = i < 0 ? 0 : i; i = i > LL->shape[0] - 1 ? LL->shape[0] - 1 : i;
        j = j < 0 ? 0 : j; j = j > LL->shape[1] - 1 ? LL->shape[1] - 1 : j;
        k = k < 0 ? 0 : k; k = k > LL->shape[2] - 1 ? LL->shape[2] - 1 : k;
        voxel[itrk * 3 + 0] = (int32_t)i; voxel[itrk * 3 + 1] = (int32_t)j; voxel[itrk * 3 + 2] = (int32_t)k;
        const char* vox = (const char*)lut + (((i * LL->shape[1] + j) * LL->shape[2] + k) * (int64_t)LL->shape[3]) * LL->itemsize;
        int64_t channel_offset = (ndet < c->n_op_channel) ? ndet * imod : 0;          /* :121-124 */
        for (int32_t o = 0; o < ndet; o++) {
            int64_t ch = o + channel_offset;
            int64_t li = o % LL->shape[3];
            float vis_f; memcpy(&vis_f, vox + li * LL->itemsize + LL->off_vis, 4);
            /* vis (float32) * bool -> float32; eff (float64) * vis * n_photons -> float64, stored f4 */
            double vis = (double)vis_f * (ch2tpc[ch] == itpc ? 1.0 : 0.0);
            float out = (float)(eff[ch] * vis * n_photons);
            char* rec = (char*)linc + (itrk * ndet + o) * (int64_t)LI->itemsize;
            memcpy(rec + LI->off_n_photons_det, &out, 4);
            if (c->light_trig_mode == 0) {
                float t1f; memcpy(&t1f, vox + li * LL->itemsize + LL->off_t0, 4);
                double t1 = ((double)t1f * c->unit_ns + fld_get(L, t, LSB_F_T0) * c->unit_mus) / c->unit_mus;
                float o2 = (float)t1;
                memcpy(rec + LI->off_t0_det, &o2, 4);
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* light_sim.py:58-129 sum_light_signals                                                  */
/* ------------------------------------------------------------------------------------ */
ORC_API int orc_sum_light_signals(const lsb_consts* c, const lsb_track_layout* L, const void* segments, int64_t S,
                                  const int32_t* seg_voxel, const int64_t* seg_track_id,
                                  const void* linc, const lsb_linc_layout* LI, int32_t ndet_inc,
                                  const int32_t* op_channel, const void* lut, const lsb_lut_layout* LL,
                                  double start_time, float* lsi, int32_t ndet, int32_t nticks,
                                  int64_t* true_id, double* true_ph, int32_t n_true,
                                  const int64_t* sorted_indices, int64_t n_sorted, double t0_profile_length) {
    (void)S;
#pragma omp parallel for collapse(2) schedule(dynamic, 256)
    for (int32_t idet = 0; idet < ndet; idet++)
        for (int32_t itick = 0; itick < nticks; itick++) {
            double start_tick_time = itick * c->light_tick_size + start_time;
            double end_tick_time = start_tick_time + c->light_tick_size;
            int32_t ch = op_channel[idet];
            int64_t idet_lut = py_mod(ch, LL->shape[3]);
            for (int64_t s = 0; s < n_sorted; s++) {
                int64_t itrk = sorted_indices[idet * n_sorted + s];
                const char* rec = (const char*)linc + (itrk * ndet_inc + ch) * (int64_t)LI->itemsize;
                float nph; memcpy(&nph, rec + LI->off_n_photons_det, 4);
                if (!(nph > 0)) continue;
                const int32_t* vox = seg_voxel + itrk * 3;
                double track_time = fld_get(L, (const char*)segments + itrk * L->itemsize, LSB_F_T0);
                double track_end_time = track_time + t0_profile_length * c->unit_ns / c->unit_mus;
                if (track_end_time < start_tick_time || track_time > end_tick_time) continue;
                const char* lrec = (const char*)lut + ((((int64_t)vox[0] * LL->shape[1] + vox[1]) * LL->shape[2] + vox[2]) *
                                                        (int64_t)LL->shape[3] + idet_lut) * LL->itemsize;
                if (c->enable_lut_smearing) {
                    for (int ip = 0; ip < LL->n_time_dist; ip++) {
                        double profile_time = track_time + ip * c->unit_ns / c->unit_mus;
                        if (profile_time < end_tick_time && profile_time > start_tick_time) {
                            float tp; memcpy(&tp, lrec + LL->off_time_dist + 4 * ip, 4);
                            /* float32*float32 -> float32, / float64 -> float64 */
                            double photons = (double)(float)(nph * tp) / c->light_tick_size;
                            float* o = &lsi[idet * (int64_t)nticks + itick];
                            *o = (float)((double)*o + photons);
                            if (photons > c->mc_truth_threshold) {
                                for (int q = 0; q < n_true; q++) {
                                    int64_t* tid = &true_id[((int64_t)idet * nticks + itick) * n_true + q];
                                    if (*tid == -1 || *tid == seg_track_id[itrk]) {
                                        *tid = seg_track_id[itrk];
                                        true_ph[((int64_t)idet * nticks + itick) * n_true + q] += photons;
                                        break;
                                    }
                                }
                            }
                        }
                    }
                } else {
                    float ta; memcpy(&ta, lrec + LL->off_t0_avg, 4);
                    double t0_avg = (double)ta * c->unit_ns / c->unit_mus;
                    double profile_time = track_time + t0_avg;
                    if (profile_time < end_tick_time && profile_time > start_tick_time) {
                        double photons = (double)nph / c->light_tick_size;
                        float* o = &lsi[idet * (int64_t)nticks + itick];
                        *o = (float)((double)*o + photons);
                        if (photons > c->mc_truth_threshold) {
                            for (int q = 0; q < n_true; q++) {
                                int64_t* tid = &true_id[((int64_t)idet * nticks + itick) * n_true + q];
                                if (*tid == -1 || *tid == seg_track_id[itrk]) {
                                    *tid = seg_track_id[itrk];
                                    true_ph[((int64_t)idet * nticks + itick) * n_true + q] += photons;
                                    break;
                                }
                            }
                        }
                    }
                }
            }
        }
    return 0;
}

/* light_sim.py:131-145 scintillation_model */
static inline double scint_model(const lsb_consts* c, int64_t tt) {
    double p1 = c->singlet_fraction * exp(-tt * c->light_tick_size / c->tau_s) * (1 - exp(-c->light_tick_size / c->tau_s));
    double p3 = (1 - c->singlet_fraction) * exp(-tt * c->light_tick_size / c->tau_t) * (1 - exp(-c->light_tick_size / c->tau_t));
    return (p1 + p3) * (tt >= 0 ? 1.0 : 0.0);
}
/* light_sim.py:148-183 calc_scintillation_effect */
ORC_API int orc_calc_scintillation_effect(const lsb_consts* c, const float* inc, const int64_t* in_id, const double* in_ph,
                                          float* out, int64_t* out_id, double* out_ph, int32_t ndet, int32_t nticks,
                                          int32_t n_in, int32_t n_out) {
    int64_t conv_ticks = (int64_t)ceil((c->light_window[1] - c->light_window[0]) / c->light_tick_size);
#pragma omp parallel for collapse(2) schedule(static)
    for (int32_t idet = 0; idet < ndet; idet++)
        for (int32_t itick = 0; itick < nticks; itick++) {
            int64_t j0 = itick - conv_ticks > 0 ? itick - conv_ticks : 0;
            for (int64_t j = j0; j <= itick; j++) {
                float v = inc[idet * (int64_t)nticks + j];
                if (v == 0) continue;
                double w = scint_model(c, itick - j);
                float* o = &out[idet * (int64_t)nticks + itick];
                *o = (float)((double)*o + w * (double)v);
                for (int it = 0; it < n_in; it++) {
                    int64_t id = in_id[((int64_t)idet * nticks + j) * n_in + it];
                    if (id == -1) break;
                    double ph = in_ph[((int64_t)idet * nticks + j) * n_in + it];
                    if (w * ph < c->mc_truth_threshold) continue;
                    for (int jt = 0; jt < n_out; jt++) {
                        int64_t* oid = &out_id[((int64_t)idet * nticks + itick) * n_out + jt];
                        if (*oid == id || *oid == -1) {
                            *oid = id;
                            out_ph[((int64_t)idet * nticks + itick) * n_out + jt] += w * ph;
                            break;
                        }
                    }
                }
            }
        }
    return 0;
}

/* light_sim.py:186-238 */
static int32_t poisson_i32(double mean, uint64_t* st) {
    if (mean <= 0) return 0;
    if (mean < 30) {
        float u = rng_uniform_f32(st);
        int32_t x = 0;
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
    double v = (double)rng_normal_f32(st) * sqrt(mean) + mean;
    int64_t iv = (int64_t)v;
    return (int32_t)(iv > 0 ? iv : 0);
}
ORC_API int orc_calc_stat_fluctuations(const lsb_consts* c, const float* inc, float* out, int32_t ndet, int32_t nticks,
                                       uint64_t* rng_states) {
    for (int32_t idet = 0; idet < ndet; idet++)
        for (int32_t itick = 0; itick < nticks; itick++) {
            int64_t idx = idet * (int64_t)nticks + itick;
            if (inc[idx] > 0) {
                /* float32 * float64 -> float64 mean */
                double mean = (double)inc[idx] * c->light_tick_size;
                out[idx] = (float)(1. / c->light_tick_size * poisson_i32(mean, rng_states + 2 * idx));
            } else out[idx] = 0.f;
        }
    return 0;
}

/* light_sim.py:241-300 interp, sipm_response_model */
static double interp_arr(double idx, const double* arr, int n, double low, double high) {
    int64_t i0 = (int64_t)floor(idx);
    if (i0 < 0) return low;
    if (i0 > n - 1) return high;
    if ((double)i0 == idx) return arr[i0];
    if (i0 > n - 2) return high;
    double v0 = arr[i0], v1 = arr[i0 + 1];
    return v0 + (v1 - v0) * (idx - i0);
}
static double sipm_model(const lsb_consts* c, int64_t tt, const double* impulse, int n_imp) {
    if (c->sipm_response_model == 0) {
        double t = tt * c->light_tick_size;
        double imp = (t >= 0 ? 1.0 : 0.0) * exp(-t / c->light_response_time) * sin(t / c->light_oscillation_period);
        imp /= c->light_oscillation_period * (c->light_response_time * c->light_response_time);
        imp *= c->light_oscillation_period * c->light_oscillation_period + c->light_response_time * c->light_response_time;
        return imp * c->light_tick_size;
    }
    double imp = interp_arr(tt * c->light_tick_size / c->impulse_tick_size, impulse, n_imp, 0, 0);
    imp /= c->impulse_tick_size / c->light_tick_size;
    return imp;
}
/* light_sim.py:303-336 calc_light_detector_response (truth block indexes itick where jtick is
 * meant, :333-335 -- replicated) */
ORC_API int orc_calc_light_detector_response(const lsb_consts* c, const float* inc, const int64_t* in_id, const double* in_ph,
                                             float* out, int64_t* out_id, double* out_ph, int32_t ndet, int32_t nticks,
                                             int32_t n_in, int32_t n_out, const double* gain, const double* impulse, int32_t n_imp) {
    int64_t conv_ticks = (int64_t)ceil((c->light_window[1] - c->light_window[0]) / c->light_tick_size);
#pragma omp parallel for collapse(2) schedule(static)
    for (int32_t idet = 0; idet < ndet; idet++)
        for (int32_t itick = 0; itick < nticks; itick++) {
            int64_t j0 = itick - conv_ticks > 0 ? itick - conv_ticks : 0;
            for (int64_t j = j0; j <= itick; j++) {
                double w = sipm_model(c, itick - j, impulse, n_imp);
                float* o = &out[idet * (int64_t)nticks + itick];
                *o = (float)((double)*o + gain[idet] * w * (double)inc[idet * (int64_t)nticks + j]);
                for (int it = 0; it < n_in; it++) {
                    int64_t base_j = ((int64_t)idet * nticks + j) * n_in, base_i = ((int64_t)idet * nticks + itick) * n_in;
                    if (in_id[base_j + it] == -1) break;
                    if (fabs(w * in_ph[base_j + it]) < c->mc_truth_threshold) continue;
                    for (int jt = 0; jt < n_out; jt++) {
                        if (in_id[base_i + jt] == in_id[base_i + it] || in_id[base_i + jt] == -1) {
                            out_id[((int64_t)idet * nticks + itick) * n_out + jt] = in_id[base_i + it];
                            out_ph[((int64_t)idet * nticks + itick) * n_out + jt] += w * in_ph[base_j + it];
                            break;
                        }
                    }
                }
            }
        }
    return 0;
}

ORC_API int orc_abi_version(void) { return LSB_ABI_VERSION; }
