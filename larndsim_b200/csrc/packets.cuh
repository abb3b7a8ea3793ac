// packets.cuh -- hit compaction and LArPix packet building: fee.export_to_hdf5 (fee.py:84-359) without the
// file I/O.  The reference walks the hits sequentially (pixel-major, then ADC slot) and carries three pieces of
// state: the clock-rollover count (event_start_time_list[itick:] -= CLOCK_RESET_PERIOD, :176-189), the event of
// the previous hit (:197-238) and the timestamp of the previous data packet (:272-276).  All three are prefix
// quantities over the slot index s = pixel * A + iadc:
//   rollovers(s)      = max over slots <= s of need(slot)  (the while-loop conditions are monotone in the count)
//   previous event    = event of the last valid slot < s
//   previous timestamp = time tick of the last slot < s that produced a data packet
// so the builder is: k_pkt_slots (per pixel: decode, readout-table lookups, per-slot validity and need) ->
// inclusive max-scan of three int32 channels -> k_pkt_count (packets per slot) -> exclusive sum-scan ->
// k_pkt_write (packets in the reference's order) -> k_pkt_assn (warp per packet: fraction-sorted truth rows,
// fee.py:287-342).  Integer / byte work, HBM-bound (the [U, A, K] fraction table is read once).
#pragma once
#include "common.cuh"
#include "glue.cuh"

struct PktTables {
    double clock_cycle, adc_pedestal, mus, s;
    long long reset_period;
    int light_trig_mode;
    int npx, npy, nptx, npty, ntx, nty;       // pixels per plane, per tile; tiles per anode
    int n_tiles, n_modules, max_groups, n_iog, n_bad;
    const int* tile_map;        // [2][ntx][nty]
    const int* tile_orient;     // [n_tiles][2]  sign of the x / y axis
    const int* pix_conn;        // [nptx][npty]  chip * 1000 + channel, -1: not connected
    const int* tile_chip_io;    // [n_tiles][256]  io_group * 1000 + io_channel, -1: absent
    const int* module_ng;       // [n_modules]  0: module not in MODULE_TO_IO_GROUPS
    const int* module_io;       // [n_modules][max_groups]
    const int* io_groups;       // [n_iog]  groups that get the per-event timestamp / sync packets
    const long long* bad;       // [n_bad] sorted keys ((io_group * 1000 + io_channel) * 1000 + chip) * 64 + channel
};
struct PixInfo { int ok_module, ok_conn, chip, channel, io_group, io_channel; };
struct Scan3 { int r, pv, pe; };                 // rollover count, last valid slot, last emitting slot

__device__ __forceinline__ Scan3 scan3_max(const Scan3& a, const Scan3& b) {
    Scan3 o; o.r = a.r > b.r ? a.r : b.r; o.pv = a.pv > b.pv ? a.pv : b.pv; o.pe = a.pe > b.pe ? a.pe : b.pe; return o;
}
__device__ __forceinline__ Scan3 scan3_id() { Scan3 o; o.r = 0; o.pv = -1; o.pe = -1; return o; }

// time tick of a hit after `r` rollovers (fee.py:180-181) -- the expression is evaluated exactly as written
__device__ __forceinline__ long long pkt_time_tick(double t, long long e0, long long r, const PktTables& T) {
    return (long long)floor(t / T.clock_cycle + (double)(e0 - r * T.reset_period));
}

__global__ void k_pkt_slots(PktTables T, long long U, int A, const int32_t* __restrict__ unique_pix, const double* __restrict__ adc,
                            const double* __restrict__ ticks, const long long* __restrict__ pix_t0, PixInfo* __restrict__ info,
                            Scan3* __restrict__ sc) {
    const long long ip = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (ip >= U) return;
    PixInfo pi; pi.ok_module = 0; pi.ok_conn = 0; pi.chip = 0; pi.channel = 0; pi.io_group = 0; pi.io_channel = 0;
    const long long pid = unique_pix[ip];
    // id2pixel (pixels_from_track.py:28-41), Python floor semantics
    const long long npl = (long long)T.npx * T.npy;
    long long plane = pid / npl; if (pid % npl != 0 && pid < 0) plane -= 1;
    const long long px = py_mod_ll(pid, T.npx);
    long long q = pid / T.npx; if (pid % T.npx != 0 && pid < 0) q -= 1;
    const long long py = py_mod_ll(q, T.npy);
    long long module = plane / 2; if (plane % 2 != 0 && plane < 0) module -= 1;
    module += 1;
    if (module >= 0 && module < T.n_modules && T.module_ng[module] > 0) {
        pi.ok_module = 1;
        const int tile_x = (int)(px / T.nptx), tile_y = (int)(py / T.npty);
        const int anode = (py_mod_ll(plane, 2) == 0) ? 0 : 1;
        const int tile = (tile_x < T.ntx && tile_y < T.nty) ? T.tile_map[(anode * T.ntx + tile_x) * T.nty + tile_y] : -1;
        int rx = (int)(px % T.nptx), ry = (int)(py % T.npty);
        if (tile >= 0 && tile < T.n_tiles) {
            if (T.tile_orient[2 * tile] < 0) rx = T.nptx - rx - 1;            // rotate_tile (fee.py:40-64)
            if (T.tile_orient[2 * tile + 1] < 0) ry = T.npty - ry - 1;
            const int cc = T.pix_conn[rx * T.npty + ry];
            if (cc >= 0) {
                pi.chip = cc / 1000; pi.channel = cc % 1000;
                const int gio = pi.chip < 256 ? T.tile_chip_io[tile * 256 + pi.chip] : -1;
                if (gio >= 0) {
                    const int g = gio / 1000;
                    pi.io_channel = gio % 1000;
                    if (g >= 1 && g <= T.module_ng[module]) {
                        pi.io_group = T.module_io[module * T.max_groups + g - 1];
                        pi.ok_conn = 1;
                        const long long key = (((long long)pi.io_group * 1000 + pi.io_channel) * 1000 + pi.chip) * 64 + pi.channel;
                        int lo = 0, hi = T.n_bad - 1;                            // bad channels (fee.py:256-260)
                        while (lo <= hi) {
                            const int mid = (lo + hi) >> 1;
                            const long long v = T.bad[mid];
                            if (v == key) { pi.ok_conn = 0; break; }
                            if (v < key) lo = mid + 1; else hi = mid - 1;
                        }
                    }
                }
            }
        }
    }
    info[ip] = pi;
    const long long e0 = pix_t0[ip];
    bool open = pi.ok_module != 0;
    for (int ia = 0; ia < A; ia++) {
        const long long s = ip * A + ia;
        Scan3 v = scan3_id();
        if (open && adc[s] > T.adc_pedestal) {                                   // fee.py:171 (else: break)
            long long r = 0;
            if (e0 > T.reset_period - 1) r = (e0 - (T.reset_period - 1) + T.reset_period - 1) / T.reset_period;
            while (pkt_time_tick(ticks[s], e0, r, T) > T.reset_period - 1) r++;   // :176-189
            v.r = (int)r; v.pv = (int)s; v.pe = pi.ok_conn ? (int)s : -1;
        } else {
            open = false;
        }
        sc[s] = v;
    }
}

// ---- inclusive max-scan of Scan3 (same three-phase structure as glue.cuh's sum scan) -----------------
__device__ __forceinline__ Scan3 block_inclusive_scan3(Scan3 v, Scan3* total) {
    __shared__ Scan3 warp_tot[SCAN_TPB / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        Scan3 y;
        y.r = __shfl_up_sync(0xffffffffu, v.r, o); y.pv = __shfl_up_sync(0xffffffffu, v.pv, o); y.pe = __shfl_up_sync(0xffffffffu, v.pe, o);
        if (lane >= o) v = scan3_max(v, y);
    }
    if (lane == 31) warp_tot[wid] = v;
    __syncthreads();
    Scan3 pre = scan3_id();
    for (int w = 0; w < wid; w++) pre = scan3_max(pre, warp_tot[w]);
    if (total && threadIdx.x == SCAN_TPB - 1) *total = scan3_max(pre, v);
    __syncthreads();
    return scan3_max(pre, v);
}
__global__ void k_scan3_reduce(const Scan3* __restrict__ in, long long n, Scan3* __restrict__ tile_tot) {
    const long long base = (long long)blockIdx.x * SCAN_TILE;
    Scan3 s = scan3_id();
    for (int k = 0; k < SCAN_IPT; k++) {
        const long long i = base + k * SCAN_TPB + threadIdx.x;
        if (i < n) s = scan3_max(s, in[i]);
    }
    __shared__ Scan3 tot;
    block_inclusive_scan3(s, &tot);
    __syncthreads();
    if (threadIdx.x == 0) tile_tot[blockIdx.x] = tot;
}
__global__ void k_scan3_tiles(Scan3* __restrict__ tile_tot, long long nb) {
    // single block: tile totals -> exclusive prefixes
    __shared__ Scan3 s_inc[SCAN_TPB];
    __shared__ Scan3 carry_s;
    if (threadIdx.x == 0) carry_s = scan3_id();
    __syncthreads();
    for (long long base = 0; base < nb; base += SCAN_TPB) {
        const long long i = base + threadIdx.x;
        const Scan3 v = i < nb ? tile_tot[i] : scan3_id();
        s_inc[threadIdx.x] = block_inclusive_scan3(v, nullptr);
        __syncthreads();
        const Scan3 carry = carry_s;
        if (i < nb) tile_tot[i] = scan3_max(carry, threadIdx.x > 0 ? s_inc[threadIdx.x - 1] : scan3_id());
        __syncthreads();
        if (threadIdx.x == 0) carry_s = scan3_max(carry, s_inc[SCAN_TPB - 1]);
        __syncthreads();
    }
}
__global__ void k_scan3_apply(Scan3* __restrict__ io, long long n, const Scan3* __restrict__ tile_pre) {
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_IPT;
    Scan3 v[SCAN_IPT];
    Scan3 s = scan3_id();
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) {
        const long long i = base + k;
        v[k] = i < n ? io[i] : scan3_id();
        s = scan3_max(s, v[k]);
    }
    const Scan3 inc = block_inclusive_scan3(s, nullptr);
    // exclusive prefix of this thread = inclusive of the previous thread
    __shared__ Scan3 s_inc[SCAN_TPB];
    s_inc[threadIdx.x] = inc;
    __syncthreads();
    Scan3 run = scan3_max(tile_pre[blockIdx.x], threadIdx.x > 0 ? s_inc[threadIdx.x - 1] : scan3_id());
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) {
        const long long i = base + k;
        run = scan3_max(run, v[k]);
        if (i < n) io[i] = run;
    }
}

struct PktTrig { const double* t; const long long* ev; const int* module; int n; };

// packets emitted by slot s: [event-change block][timestamp][data]   (fee.py:197-285)
__device__ __forceinline__ int pkt_slot_plan(const PktTables& T, const PktTrig& G, long long s, int A, const Scan3* sc,
                                             const double* ticks, const long long* pix_t0, const long long* event_id,
                                             bool& evchange, bool& ts, bool& data, long long& time_tick, long long& r) {
    const Scan3 me = sc[s];
    evchange = ts = data = false;
    if (me.pv != (int)s) return 0;                        // not a valid hit
    const Scan3 prev = s > 0 ? sc[s - 1] : scan3_id();
    r = me.r;
    const long long ip = s / A;
    time_tick = py_mod_ll(pkt_time_tick(ticks[s], pix_t0[ip], r, T), T.reset_period);
    int n = 0;
    if (T.light_trig_mode != 1) {
        const long long last_event = prev.pv >= 0 ? event_id[prev.pv] : -1;
        const long long ev = event_id[s];
        if (ev != last_event) {
            evchange = true;
            n += 2 * T.n_iog;
            for (int k = 0; k < G.n; k++)
                if (G.ev[k] == ev) { const int m = G.module[k]; n += (m >= 0 && m < T.n_modules) ? T.module_ng[m] : 0; }
        }
    }
    if (me.pe == (int)s) {
        data = true;
        long long last_tt = -1;
        if (prev.pe >= 0) last_tt = py_mod_ll(pkt_time_tick(ticks[prev.pe], pix_t0[prev.pe / A], sc[prev.pe].r, T), T.reset_period);
        ts = time_tick != last_tt;
        n += ts ? 2 : 1;
    }
    return n;
}
__global__ void k_pkt_count(PktTables T, PktTrig G, long long N, int A, const Scan3* __restrict__ sc, const double* __restrict__ ticks,
                            const long long* __restrict__ pix_t0, const long long* __restrict__ event_id, uint32_t* __restrict__ count) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= N) return;
    bool a, b, c; long long tt, r;
    count[s] = (uint32_t)pkt_slot_plan(T, G, s, A, sc, ticks, pix_t0, event_id, a, b, c, tt, r);
}

__device__ __forceinline__ int pkt_parity(int chip, int channel, long long timestamp, int first_packet, int dataword) {
    // odd parity over bits 0..62 of the Packet_v2 word (type 0, flags 0): chip[2:10] channel[10:16] timestamp[16:47]
    // first_packet[47] dataword[48:56]
    unsigned long long w = ((unsigned long long)(chip & 0xFF) << 2) | ((unsigned long long)(channel & 0x3F) << 10) |
                           ((unsigned long long)(timestamp & 0x7FFFFFFF) << 16) | ((unsigned long long)(first_packet & 1) << 47) |
                           ((unsigned long long)(dataword & 0xFF) << 48);
    return 1 - (__popcll(w) & 1);
}
__device__ __forceinline__ lsb_packet pkt_blank(int type) {
    lsb_packet p;
    memset(&p, 0, sizeof(p));
    p.packet_type = (uint8_t)type;
    return p;
}
__global__ void k_pkt_write(PktTables T, PktTrig G, long long N, int A, const Scan3* __restrict__ sc, const double* __restrict__ ticks,
                            const double* __restrict__ adc, const long long* __restrict__ pix_t0, const double* __restrict__ pix_t0_us,
                            const long long* __restrict__ event_id, const PixInfo* __restrict__ info, const long long* __restrict__ offs,
                            long long cap, lsb_packet* __restrict__ out, long long* __restrict__ src_slot,
                            const long long* __restrict__ base_dev) {
    // base_dev (optional): device-resident position of this call's first packet in `out` (append mode of the spill runner)
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= N) return;
    bool evchange, ts, data; long long tt, r;
    if (pkt_slot_plan(T, G, s, A, sc, ticks, pix_t0, event_id, evchange, ts, data, tt, r) == 0) return;
    long long o = offs[s] + (base_dev ? *base_dev : 0);
    const long long ip = s / A;
    auto put = [&](const lsb_packet& p, long long src) { if (o < cap) { out[o] = p; src_slot[o] = src; } o++; };
    if (evchange) {
        const long long e0m = py_mod_ll(pix_t0[ip] - r * T.reset_period, T.reset_period);
        for (int k = 0; k < T.n_iog; k++) {
            lsb_packet p = pkt_blank(4);                                         // TimestampPacket, Key(io_group, 0, 0)
            p.io_group = (uint8_t)T.io_groups[k];
            p.timestamp_s = pix_t0_us[ip] * T.mus / T.s;
            put(p, -1);
            lsb_packet q = pkt_blank(6);                                         // SyncPacket 'S'
            q.io_group = (uint8_t)T.io_groups[k]; q.sub_type = 'S'; q.timestamp = (uint64_t)tt;
            put(q, -1);
        }
        const long long ev = event_id[s];
        for (int k = 0; k < G.n; k++) {
            if (G.ev[k] != ev) continue;
            const long long t_trig = py_mod_ll((long long)floor(G.t[k] / T.clock_cycle + (double)e0m), T.reset_period);
            const int m = G.module[k];
            const int ng = (m >= 0 && m < T.n_modules) ? T.module_ng[m] : 0;
            for (int j = 0; j < ng; j++) {
                lsb_packet p = pkt_blank(7);                                     // TriggerPacket type 0x02
                p.io_group = (uint8_t)T.module_io[m * T.max_groups + j]; p.sub_type = 2; p.timestamp = (uint64_t)t_trig;
                put(p, -1);
            }
        }
    }
    if (data) {
        const PixInfo pi = info[ip];
        if (ts) {
            // event_start_time_list[0] has only seen the rollovers that happened while pixel 0 was processed
            const long long r0 = ip == 0 ? r : sc[A - 1].r;
            lsb_packet p = pkt_blank(4);
            p.io_group = (uint8_t)pi.io_group;
            p.timestamp_s = floor((double)(pix_t0[0] - r0 * T.reset_period) * T.clock_cycle * T.mus / T.s);
            put(p, -1);
        }
        lsb_packet p = pkt_blank(0);
        const int dw = (int)adc[s];
        p.io_group = (uint8_t)pi.io_group; p.io_channel = (uint8_t)pi.io_channel; p.chip_id = (uint8_t)pi.chip;
        p.channel_id = (uint8_t)pi.channel; p.dataword = (uint8_t)dw; p.first_packet = 1;
        p.timestamp = (uint64_t)tt; p.receipt_timestamp = (uint32_t)tt;
        p.parity = (uint8_t)pkt_parity(pi.chip, pi.channel, tt, 1, dw);
        put(p, s);
    }
}

// ---- mc_packets_assn (fee.py:287-342): warp per packet --------------------------------------------------
// np_sum_schedule: numpy's float64 add.reduce on a contiguous run (checked against numpy 2.3 on random data): a plain
// loop from 0 below 8 elements, else 8 interleaved accumulators combined as a tree, then the tail (n <= 128)
#define ASSN_MAXK 128
#define ASSN_WARPS 4
__global__ void __launch_bounds__(32 * ASSN_WARPS) k_pkt_assn(long long n_packets, const long long* __restrict__ src_slot, int A, int K, int NA,
                                                              const long long* __restrict__ event_id, const double* __restrict__ cf,
                                                              const long long* __restrict__ track_ids, const long long* __restrict__ traj_ids,
                                                              char* __restrict__ rows, const long long* __restrict__ n_dev,
                                                              const long long* __restrict__ base_dev, long long cap,
                                                              const long long* __restrict__ seg_map, const long long* __restrict__ traj_map) {
    // one output row = the mc_packets_assn record: event_ids i8[1] | segment_ids i8[NA] | fraction f8[NA] | file_traj_ids i8[NA] |
    // fraction_traj f8[NA]  (8 + 32 NA bytes, every field 8-byte aligned).
    // Append mode (spill runner): the packet count and the position of the first packet are read from the device (n_dev,
    // base_dev; the grid is sized without knowing them and strides), and track_ids / traj_ids hold segment indices of the batch
    // that are translated through seg_map / traj_map on the fly (cli/simulate_pixels.py:1114-1115).
    __shared__ double s_f[ASSN_WARPS][ASSN_MAXK];        // fractions in sorted order
    __shared__ long long s_t[ASSN_WARPS][ASSN_MAXK];     // trajectory ids in sorted order
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long n_all = n_dev ? *n_dev : n_packets;
    const long long base = base_dev ? *base_dev : 0;
  for (long long ipk0 = blockIdx.x * (long long)ASSN_WARPS + w; ipk0 < n_all; ipk0 += (long long)gridDim.x * ASSN_WARPS) {
    const long long ipk = base + ipk0;
    if (ipk >= cap) break;
    __syncwarp();
    const long long s = src_slot[ipk];
    char* row = rows + ipk * (8LL + 32LL * NA);
    long long* o_event = reinterpret_cast<long long*>(row);
    long long* seg = reinterpret_cast<long long*>(row + 8); double* fr = reinterpret_cast<double*>(row + 8 + 8LL * NA);
    long long* tj = reinterpret_cast<long long*>(row + 8 + 16LL * NA); double* ftj = reinterpret_cast<double*>(row + 8 + 24LL * NA);
    for (int j = lane; j < NA; j += 32) { seg[j] = -1; fr[j] = 0.0; tj[j] = -1; ftj[j] = 0.0; }
    if (lane == 0) o_event[0] = s >= 0 ? event_id[s] : -1;
    if (s < 0) continue;
    __syncwarp();
    const long long ip = s / A;
    const double* f = cf + s * (long long)K;
    const long long* trk = track_ids + ip * K;
    const long long* trj = traj_ids + ip * K;
    // descending by fraction; equal fractions in descending slot order (= np.flip of a stable ascending argsort)
    for (int i = lane; i < K; i += 32) {
        const double fi = f[i];
        int rank = 0;
        for (int j = 0; j < K; j++) { const double fj = f[j]; rank += (fj > fi) || (fj == fi && j > i); }
        long long tr = trj[i], sg = trk[i];
        if (traj_map && tr >= 0) tr = traj_map[tr];
        if (seg_map && sg >= 0) sg = seg_map[sg];
        s_f[w][rank] = fi; s_t[w][rank] = tr;
        if (rank < NA) { seg[rank] = sg; fr[rank] = fi; }
    }
    __syncwarp();
    // trajectories: ascending unique ids, fractions summed in sorted order, stored through float32
    for (int i = lane; i < K; i += 32) {
        const long long id = s_t[w][i];
        if (id < 0) continue;
        bool first = true;
        for (int j = 0; j < i; j++) if (s_t[w][j] == id) { first = false; break; }
        if (!first) continue;
        int tidx = 0;
        for (int j = 0; j < K; j++) {
            const long long o = s_t[w][j];
            if (o < 0 || o >= id) continue;
            bool fo = true;
            for (int l = 0; l < j; l++) if (s_t[w][l] == o) { fo = false; break; }
            tidx += fo;
        }
        if (tidx >= NA) continue;
        // np.sum over the matching fractions in sorted order (see np_sum_schedule): streamed, no scratch
        int n = 0;
        for (int j = 0; j < K; j++) n += s_t[w][j] == id;
        const int lim = n - (n % 8);
        double res = 0.0, r[8];
        int c = 0;
        for (int j = 0; j < K; j++) {
            if (s_t[w][j] != id) continue;
            const double v = s_f[w][j];
            const int k = c++;
            if (n < 8) res += v;
            else if (k < 8) r[k] = v;
            else if (k < lim) r[k & 7] += v;
        }
        if (n >= 8) {
            res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
            c = 0;
            for (int j = 0; j < K; j++) {
                if (s_t[w][j] != id) continue;
                if (c++ >= lim) res += s_f[w][j];
            }
        }
        tj[tidx] = id;
        ftj[tidx] = (double)__double2float_rn(res);
    }
  }
}

static int scan3_inclusive(Scan3* io, long long n, Scan3* tile_tot, cudaStream_t st) {
    if (n <= 0) return 0;
    const long long nb = scan_num_blocks(n);
    k_scan3_reduce<<<(unsigned)nb, SCAN_TPB, 0, st>>>(io, n, tile_tot);
    LSB_LAUNCH_CHECK("k_scan3_reduce");
    k_scan3_tiles<<<1, SCAN_TPB, 0, st>>>(tile_tot, nb);
    LSB_LAUNCH_CHECK("k_scan3_tiles");
    k_scan3_apply<<<(unsigned)nb, SCAN_TPB, 0, st>>>(io, n, tile_tot);
    LSB_LAUNCH_CHECK("k_scan3_apply");
    return 0;
}

template <typename T>
static int pkt_upload(TmpPool& pool, const T* host, long long n, const T** dev, cudaStream_t st) {
    T* d = nullptr;
    LSB_CUDA(pool.get(&d, n > 0 ? n : 1));
    if (n > 0) LSB_CUDA(cudaMemcpyAsync(d, host, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, st));
    *dev = d;
    return 0;
}

LSB_EXPORT int lsb_export_packets(const lsb_readout_tables* rt, int64_t U, int32_t A, int32_t K, const int64_t* event_id,
                                  const double* adc, const double* adc_ticks, const int32_t* unique_pix, const double* current_fractions,
                                  const int64_t* track_ids, const int64_t* traj_ids, const int64_t* pix_t0_ticks, const double* pix_t0_us,
                                  int32_t n_trig, const double* trig_times, const int64_t* trig_event, const int32_t* trig_module,
                                  int64_t cap_packets, lsb_packet* packets, void* assn_rows, int32_t n_assn, int64_t* n_packets, void* stream) {
    LSB_REQUIRE(rt && n_packets, "export_packets: null tables / n_packets");
    LSB_REQUIRE(K >= 0 && K <= ASSN_MAXK, "export_packets: MAX_TRACKS_PER_PIXEL above the supported 128");
    LSB_REQUIRE(U * (long long)A < 2147483647LL, "export_packets: more than 2^31 hit slots in one call");
    *n_packets = 0;
    if (U == 0 || A == 0) return 0;
    LSB_REQUIRE(event_id && adc && adc_ticks && unique_pix && pix_t0_ticks && pix_t0_us, "export_packets: null input");
    cudaStream_t st = (cudaStream_t)stream;
    TmpPool pool(st);
    PktTables T;
    T.clock_cycle = rt->clock_cycle; T.adc_pedestal = rt->adc_pedestal; T.mus = rt->mus; T.s = rt->s;
    T.reset_period = rt->clock_reset_period; T.light_trig_mode = rt->light_trig_mode;
    T.npx = rt->n_pixels[0]; T.npy = rt->n_pixels[1]; T.nptx = rt->n_pixels_per_tile[0]; T.npty = rt->n_pixels_per_tile[1];
    T.ntx = rt->n_tiles_xy[0]; T.nty = rt->n_tiles_xy[1];
    T.n_tiles = rt->n_tiles; T.n_modules = rt->n_modules; T.max_groups = rt->max_groups; T.n_iog = rt->n_io_groups; T.n_bad = rt->n_bad;
    int rc;
    if ((rc = pkt_upload(pool, rt->tile_map, 2LL * T.ntx * T.nty, &T.tile_map, st))) return rc;
    if ((rc = pkt_upload(pool, rt->tile_orientation, 2LL * T.n_tiles, &T.tile_orient, st))) return rc;
    if ((rc = pkt_upload(pool, rt->pixel_connection, (long long)T.nptx * T.npty, &T.pix_conn, st))) return rc;
    if ((rc = pkt_upload(pool, rt->tile_chip_to_io, 256LL * T.n_tiles, &T.tile_chip_io, st))) return rc;
    if ((rc = pkt_upload(pool, rt->module_n_groups, (long long)T.n_modules, &T.module_ng, st))) return rc;
    if ((rc = pkt_upload(pool, rt->module_io_groups, (long long)T.n_modules * T.max_groups, &T.module_io, st))) return rc;
    if ((rc = pkt_upload(pool, rt->io_groups, (long long)T.n_iog, &T.io_groups, st))) return rc;
    const long long* bad_dev = nullptr;
    if ((rc = pkt_upload(pool, (const long long*)rt->bad_channels, (long long)T.n_bad, &bad_dev, st))) return rc;
    T.bad = bad_dev;
    PktTrig G; G.n = n_trig; G.t = trig_times; G.ev = (const long long*)trig_event; G.module = trig_module;
    const long long N = U * (long long)A;
    PixInfo* info; Scan3* sc; Scan3* tiles; uint32_t* count; long long* offs; long long* bsum; long long* total; long long* src;
    LSB_CUDA(pool.get(&info, U)); LSB_CUDA(pool.get(&sc, N)); LSB_CUDA(pool.get(&tiles, scan_num_blocks(N) + 1));
    LSB_CUDA(pool.get(&count, N)); LSB_CUDA(pool.get(&offs, N)); LSB_CUDA(pool.get(&bsum, scan_num_blocks(N) + 1));
    LSB_CUDA(pool.get(&total, 1)); LSB_CUDA(pool.get(&src, cap_packets > 0 ? cap_packets : 1));
    k_pkt_slots<<<lsb_blocks(U, 128), 128, 0, st>>>(T, U, A, unique_pix, adc, adc_ticks, (const long long*)pix_t0_ticks, info, sc);
    LSB_LAUNCH_CHECK("k_pkt_slots");
    if ((rc = scan3_inclusive(sc, N, tiles, st))) return rc;
    k_pkt_count<<<lsb_blocks(N, 256), 256, 0, st>>>(T, G, N, A, sc, adc_ticks, (const long long*)pix_t0_ticks, (const long long*)event_id, count);
    LSB_LAUNCH_CHECK("k_pkt_count");
    if ((rc = exclusive_scan<uint32_t, long long>(count, N, offs, bsum, total, st))) return rc;
    long long n_out = 0;
    LSB_CUDA(cudaMemcpyAsync(&n_out, total, 8, cudaMemcpyDeviceToHost, st));
    LSB_CUDA(cudaStreamSynchronize(st));
    *n_packets = n_out;
    if (n_out > cap_packets) return lsb_fail_arg("export_packets: output capacity too small (n_packets holds the required count)");
    if (n_out == 0) return 0;
    LSB_REQUIRE(packets && assn_rows && current_fractions && track_ids && traj_ids, "export_packets: null output / truth pointer");
    k_pkt_write<<<lsb_blocks(N, 256), 256, 0, st>>>(T, G, N, A, sc, adc_ticks, adc, (const long long*)pix_t0_ticks, pix_t0_us,
                                                    (const long long*)event_id, info, offs, cap_packets, packets, src, nullptr);
    LSB_LAUNCH_CHECK("k_pkt_write");
    k_pkt_assn<<<lsb_blocks(n_out, ASSN_WARPS), 32 * ASSN_WARPS, 0, st>>>(n_out, src, A, K, n_assn, (const long long*)event_id, current_fractions,
                                                                         (const long long*)track_ids, (const long long*)traj_ids, (char*)assn_rows, nullptr, nullptr, n_out,
                                                                         nullptr, nullptr);
    LSB_LAUNCH_CHECK("k_pkt_assn");
    LSB_CUDA(cudaStreamSynchronize(st));      // the uploaded tables are temporaries of this call
    return 0;
}
