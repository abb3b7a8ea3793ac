// segments.cuh -- per-segment kernels: quench, drift, max_pixels, get_pixels, time_intervals.
// One thread per segment; these are HBM-bound record updates (2 x itemsize x S bytes).
#pragma once
#include "common.cuh"

// quenching.py:11-44
__global__ void k_quench(Layout L, char* __restrict__ tracks, long long n, int mode) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    char* t = tracks + i * L.itemsize;
    double dEdx = fld_get(L, t, LSB_F_DEDX);
    double dE = fld_get(L, t, LSB_F_DE);
    double recomb = 0.0;
    if (mode == d_c.mode_box) {
        double csi = d_c.box_beta * dEdx / (d_c.e_field * d_c.lar_density);
        double r = log(d_c.box_alpha + csi) / csi;
        recomb = (r > 0.0) ? r : 0.0;          // max(0, r); NaN -> 0 like Python's max(0, nan)
    } else {
        recomb = d_c.birks_ab / (1.0 + d_c.birks_kb * dEdx / (d_c.e_field * d_c.lar_density));
    }
    fld_set(L, t, LSB_F_N_ELECTRONS, recomb * dE / d_c.w_ion);   // u4 field: truncation
    double ne = fld_get(L, t, LSB_F_N_ELECTRONS);                // n_photons uses the STORED value
    fld_set(L, t, LSB_F_N_PHOTONS, (dE / d_c.w_ph - ne) * d_c.scint_prescale);
}

// drifting.py:11-58
__global__ void k_drift(Layout L, char* __restrict__ tracks, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    char* t = tracks + i * L.itemsize;
    double x = fld_get(L, t, LSB_F_X), y = fld_get(L, t, LSB_F_Y), z = fld_get(L, t, LSB_F_Z);
    int plane = d_c.default_plane_index;
    for (int ip = 0; ip < d_c.n_tpc; ip++) {
        const double (*b)[2] = d_c.tpc_borders[ip];
        double zlo = fmin(b[2][1] - 2e-2, b[2][0] - 2e-2), zhi = fmax(b[2][1] + 2e-2, b[2][0] + 2e-2);
        if (b[0][0] - 2e-2 <= x && x <= b[0][1] + 2e-2 && b[1][0] - 2e-2 <= y && y <= b[1][1] + 2e-2 &&
            zlo <= z && z <= zhi) { plane = ip; break; }
    }
    fld_set(L, t, LSB_F_PIXEL_PLANE, (double)plane);
    if (plane != d_c.default_plane_index) {
        double z_anode = d_c.tpc_borders[plane][2][0];
        double zs = fld_get(L, t, LSB_F_Z_START), ze = fld_get(L, t, LSB_F_Z_END);
        double drift_distance = fabs(z - z_anode);
        double drift_start = fabs(fmin(zs, ze) - z_anode);
        double drift_end = fabs(fmax(zs, ze) - z_anode);
        double drift_time = drift_distance / d_c.v_drift;
        double lifetime_red = exp(-drift_time / d_c.electron_lifetime);
        fld_set(L, t, LSB_F_N_ELECTRONS, fld_get(L, t, LSB_F_N_ELECTRONS) * lifetime_red);
        fld_set(L, t, LSB_F_LONG_DIFF, sqrt(drift_time * 2 * d_c.long_diff));
        fld_set(L, t, LSB_F_TRAN_DIFF, sqrt(drift_time * 2 * d_c.tran_diff));
        double t0 = fld_get(L, t, LSB_F_T0);
        fld_set(L, t, LSB_F_T, fld_get(L, t, LSB_F_T) + (drift_time + t0));
        fld_set(L, t, LSB_F_T_START, fld_get(L, t, LSB_F_T_START) + (fmin(drift_start, drift_end) / d_c.v_drift + t0));
        fld_set(L, t, LSB_F_T_END, fld_get(L, t, LSB_F_T_END) + (fmax(drift_start, drift_end) / d_c.v_drift + t0));
    }
}

// start/end pixel of a segment (pixels_from_track.py:54-60, 90-102)
__device__ __forceinline__ bool segment_pixels(const Layout& L, const char* t, long long& x0, long long& y0,
                                               long long& x1, long long& y1, long long& plane) {
    plane = (long long)fld_get(L, t, LSB_F_PIXEL_PLANE);
    if (plane < 0 || plane >= d_c.n_tpc) return false;   // the reference indexes TPC_BORDERS out of bounds here
    const double (*b)[2] = d_c.tpc_borders[plane];
    x0 = (long long)py_floordiv(fld_get(L, t, LSB_F_X_START) - b[0][0], d_c.pixel_pitch);
    y0 = (long long)py_floordiv(fld_get(L, t, LSB_F_Y_START) - b[1][0], d_c.pixel_pitch);
    x1 = (long long)py_floordiv(fld_get(L, t, LSB_F_X_END) - b[0][0], d_c.pixel_pitch);
    y1 = (long long)py_floordiv(fld_get(L, t, LSB_F_Y_END) - b[1][0], d_c.pixel_pitch);
    return true;
}

// pixels_from_track.py:43-65 + get_num_active_pixels :111-155
__global__ void k_max_pixels(Layout L, const char* __restrict__ tracks, long long n, long long* n_max) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long cnt = 0;
    if (i < n) {
        long long x0, y0, x1, y1, plane;
        if (segment_pixels(L, tracks + i * L.itemsize, x0, y0, x1, y1, plane)) {
            long long dx = llabs(x1 - x0), sx = x0 < x1 ? 1 : -1, dy = -llabs(y1 - y0), sy = y0 < y1 ? 1 : -1;
            long long err = dx + dy;
            if (in_plane(x0, y0, plane)) cnt++;
            while (x0 != x1 || y0 != y1) {
                long long e2 = 2 * err;
                if (e2 - dy > dx - e2) { err += dy; x0 += sx; } else { err += dx; y0 += sy; }
                if (in_plane(x0, y0, plane)) cnt++;
            }
        }
    }
    // warp-level max, one atomic per warp
    for (int o = 16; o > 0; o >>= 1) { long long v = __shfl_xor_sync(0xffffffffu, cnt, o); cnt = v > cnt ? v : cnt; }
    if ((threadIdx.x & 31) == 0 && cnt > 0) atomicMax(n_max, cnt);
}

// pixels_from_track.py:67-109: get_active_pixels :157-199 then get_neighboring_pixels :201-272
__global__ void k_get_pixels(Layout L, const char* __restrict__ tracks, long long n, int32_t* __restrict__ active,
                             int max_active, int32_t* __restrict__ neigh, int32_t* __restrict__ nrad, int P,
                             double* __restrict__ n_pixels_list, int radius) {
    long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (it >= n) return;
    int32_t* act = active + it * max_active;
    int32_t* nb = neigh + it * P;
    int32_t* nr = nrad + it * P;
    long long x0, y0, x1, y1, plane;
    if (segment_pixels(L, tracks + it * L.itemsize, x0, y0, x1, y1, plane)) {
        long long dx = llabs(x1 - x0), sx = x0 < x1 ? 1 : -1, dy = -llabs(y1 - y0), sy = y0 < y1 ? 1 : -1;
        long long err = dx + dy, i = 0;
        if (in_plane(x0, y0, plane) && i < max_active) act[i] = (int32_t)pixel2id(x0, y0, plane);
        while (x0 != x1 || y0 != y1) {
            i++;
            long long e2 = 2 * err;
            if (e2 - dy > dx - e2) { err += dy; x0 += sx; } else { err += dx; y0 += sy; }
            // index = step count (reference); guarded against the row length
            if (in_plane(x0, y0, plane) && i < max_active) act[i] = (int32_t)pixel2id(x0, y0, plane);
        }
    }
    long long count = 0;
    for (int pix = 0; pix < max_active; pix++) {
        int32_t a = act[pix];
        if (a == -1) continue;
        long long ax, ay, pl;
        id2pixel(a, ax, ay, pl);
        for (int xr = -radius; xr <= radius; xr++)
            for (int yr = -radius; yr <= radius; yr++) {
                long long nx = ax + xr, ny = ay + yr;
                if (!in_plane(nx, ny, pl)) continue;
                int32_t np_ = (int32_t)pixel2id(nx, ny, pl);
                bool unique = true;
                for (int k = 0; k < P; k++) if (nb[k] == np_) { unique = false; break; }   // scans the whole row
                if (!unique) continue;
                int adx = abs(xr), ady = abs(yr);
                int dmax = max(adx, ady), dmin = min(adx, ady), dsum = dmax + dmin;
                int dist = -1;
                if (dsum > d_c.max_neighbor_backtrack_distance) dist = -1;
                else if (dsum <= 1) dist = dsum;
                else if (dsum == 2) dist = (dmax == 1) ? 2 : 3;
                else if (dsum == 3) dist = (dmax == 2) ? 4 : 5;
                else if (dsum == 4) dist = (dmax == 2) ? 6 : (dmax == 3 ? 7 : 8);
                if (count < P) { nb[count] = np_; nr[count] = dist; }
                count++;
            }
    }
    n_pixels_list[it] = (double)count;
}

// detsim.py:18-40
__global__ void k_time_intervals(Layout L, const char* __restrict__ tracks, long long n, double* __restrict__ track_starts,
                                 long long* time_max) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long v = 0;
    if (i < n) {
        const char* t = tracks + i * L.itemsize;
        double t_end = (double)__double2ll_rn((fld_get(L, t, LSB_F_T_END) + 1) / d_c.time_sampling) * d_c.time_sampling;
        double t_start = (double)__double2ll_rn((fld_get(L, t, LSB_F_T_START) - d_c.time_padding) / d_c.time_sampling) * d_c.time_sampling;
        double t_length = t_end - t_start;
        track_starts[i] = t_start;
        v = (long long)ceil(t_length / d_c.time_sampling);
    }
    for (int o = 16; o > 0; o >>= 1) { long long w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(time_max, v);
}

// ---------------------------------------------------------------------------------------
LSB_EXPORT int lsb_quench(const lsb_consts* c, const lsb_track_layout* L, void* tracks, int64_t n, int32_t mode, void* stream) {
    LSB_REQUIRE(c && L && (tracks || n == 0), "quench: null pointer");
    LSB_REQUIRE(mode == c->mode_box || mode == c->mode_birks,
                "Invalid recombination mode: must be 'physics.BOX' or 'physics.BIRKS'");   // quenching.py:38
    LSB_REQUIRE(layout_has(L, LSB_F_DEDX) && layout_has(L, LSB_F_DE) && layout_has(L, LSB_F_N_ELECTRONS) &&
                layout_has(L, LSB_F_N_PHOTONS), "quench: tracks lacks dEdx/dE/n_electrons/n_photons");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = lsb_upload_consts(c, st); if (rc) return rc;
    k_quench<<<lsb_blocks(n, 256), 256, 0, st>>>(make_layout(L), (char*)tracks, n, mode);
    LSB_LAUNCH_CHECK("k_quench");
    return 0;
}

LSB_EXPORT int lsb_drift(const lsb_consts* c, const lsb_track_layout* L, void* tracks, int64_t n, void* stream) {
    LSB_REQUIRE(c && L && (tracks || n == 0), "drift: null pointer");
    static const int need[] = {LSB_F_X, LSB_F_Y, LSB_F_Z, LSB_F_Z_START, LSB_F_Z_END, LSB_F_PIXEL_PLANE, LSB_F_N_ELECTRONS,
                               LSB_F_LONG_DIFF, LSB_F_TRAN_DIFF, LSB_F_T, LSB_F_T_START, LSB_F_T_END, LSB_F_T0};
    for (int f : need) LSB_REQUIRE(layout_has(L, f), "drift: tracks lacks a required field");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = lsb_upload_consts(c, st); if (rc) return rc;
    k_drift<<<lsb_blocks(n, 256), 256, 0, st>>>(make_layout(L), (char*)tracks, n);
    LSB_LAUNCH_CHECK("k_drift");
    return 0;
}

static inline int require_pixel_fields(const lsb_track_layout* L) {
    static const int need[] = {LSB_F_PIXEL_PLANE, LSB_F_X_START, LSB_F_Y_START, LSB_F_X_END, LSB_F_Y_END};
    for (int f : need) if (!layout_has(L, f)) return lsb_fail_arg("tracks lacks pixel_plane/x_start/y_start/x_end/y_end");
    return 0;
}

LSB_EXPORT int lsb_max_pixels(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t n,
                              int64_t* n_max_pixels, void* stream) {
    LSB_REQUIRE(c && L && n_max_pixels && (tracks || n == 0), "max_pixels: null pointer");
    if (require_pixel_fields(L)) return -1;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = lsb_upload_consts(c, st); if (rc) return rc;
    k_max_pixels<<<lsb_blocks(n, 128), 128, 0, st>>>(make_layout(L), (const char*)tracks, n, (long long*)n_max_pixels);
    LSB_LAUNCH_CHECK("k_max_pixels");
    return 0;
}

LSB_EXPORT int lsb_get_pixels(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t n,
                              int32_t* active_pixels, int32_t max_active, int32_t* neighboring_pixels,
                              int32_t* neighboring_radius, int32_t max_neighbors, double* n_pixels_list,
                              int32_t radius, void* stream) {
    LSB_REQUIRE(c && L && (n == 0 || (tracks && active_pixels && neighboring_pixels && neighboring_radius && n_pixels_list)),
                "get_pixels: null pointer");
    LSB_REQUIRE(max_active >= 0 && max_neighbors >= 0 && radius >= 0, "get_pixels: negative size");
    if (require_pixel_fields(L)) return -1;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = lsb_upload_consts(c, st); if (rc) return rc;
    k_get_pixels<<<lsb_blocks(n, 128), 128, 0, st>>>(make_layout(L), (const char*)tracks, n, active_pixels, max_active,
                                                    neighboring_pixels, neighboring_radius, max_neighbors,
                                                    n_pixels_list, radius);
    LSB_LAUNCH_CHECK("k_get_pixels");
    return 0;
}

LSB_EXPORT int lsb_time_intervals(const lsb_consts* c, const lsb_track_layout* L, const void* tracks, int64_t n,
                                  double* track_starts, int64_t* time_max, void* stream) {
    LSB_REQUIRE(c && L && time_max && (n == 0 || (tracks && track_starts)), "time_intervals: null pointer");
    LSB_REQUIRE(layout_has(L, LSB_F_T_START) && layout_has(L, LSB_F_T_END), "time_intervals: tracks lacks t_start/t_end");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = lsb_upload_consts(c, st); if (rc) return rc;
    k_time_intervals<<<lsb_blocks(n, 128), 128, 0, st>>>(make_layout(L), (const char*)tracks, n, track_starts, (long long*)time_max);
    LSB_LAUNCH_CHECK("k_time_intervals");
    return 0;
}
