// spill.cuh -- native driver of the reference's batch loop (cli/simulate_pixels.py:864-1117 + save_results :1370-1390,
// fee.export_to_hdf5 fee.py:84-359) for one rank's share of a spill / file.
//
// The reference walks the (event, TPC group) batches one after the other: boolean mask -> tracks[mask] -> 8 kernel launches
// with host round trips -> concatenate -> per-hit Python loop that builds packets.  Here the units handed to this rank are
// gathered once into one contiguous device buffer, run through `depth` chain handles in flight at a time (the
// latency-bound front-end stage of one unit executes under the L1-bound current stage of the next, chain.cuh), and every
// unit's hits are turned into LArPix packets + mc_packets_assn rows ON THE DEVICE, appended to one output buffer per rank in
// unit order.  The host sees two scalars per unit while it runs (buffer sizes, as in the reference) and ONE table of packet
// counts at the end; nothing else crosses PCIe.  Units are independent: with lsb_chain_set_rng_fresh the
// result of a unit is a function of (records, seed) only, so any assignment of units to ranks gives the same packets.
#pragma once
#include "common.cuh"
#include "packets.cuh"
#include "rng.cuh"
#include "chain.cuh"

#define SPILL_MAX_DEPTH 4

struct SpillScratch {                 // per-handle temporaries of the packet stage (alive until the handle is reused)
    DevBuf ev, t0t, t0u, info, sc, tiles, count, offs, bsum, total, trig;
};

struct lsb_spill {
    lsb_consts c; lsb_track_layout L;
    int depth, n_assn, n_sm;
    int serial;                       // 1: every stage of a unit on ONE stream (per-kernel timing)
    lsb_chain* ch[SPILL_MAX_DEPTH];
    SpillScratch sx[SPILL_MAX_DEPTH];
    cudaEvent_t ev_export[SPILL_MAX_DEPTH];
    cudaEvent_t ev_gather;
    PktTables T; DevBuf tab[8];
    DevBuf gathered, seg_ids, traj_ids;
    DevBuf out_packets, out_rows, out_src, unit_table, cursor;
    long long cap_packets;
    long long* host_table;            // pinned: [2 * n_units + 2]
    long long host_table_cap;
};

// records of one unit: order[begin .. begin + n) -> contiguous; the truth ids of the file travel with them
__global__ void k_spill_gather(const char* __restrict__ src, const long long* __restrict__ order, long long n, int itemsize,
                               char* __restrict__ dst, int seg_off, int seg_dt, int traj_off, int traj_dt,
                               long long* __restrict__ seg_ids, long long* __restrict__ traj_ids) {
    const int words = itemsize >> 2;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n * words) return;
    const long long r = i / words; const int w = (int)(i - r * words);
    const long long row = order ? order[r] : r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src + row * itemsize);
    reinterpret_cast<uint32_t*>(dst + r * itemsize)[w] = s[w];
    if (w == 0) {
        const char* rec = src + row * itemsize;
        long long a = -1, b = -1;
        if (seg_off >= 0) {
            const char* p = rec + seg_off;
            a = seg_dt == LSB_U32 ? (long long)*(const uint32_t*)p : seg_dt == LSB_I32 ? (long long)*(const int32_t*)p :
                seg_dt == LSB_I64 ? *(const long long*)p : seg_dt == LSB_U64 ? (long long)*(const unsigned long long*)p :
                seg_dt == LSB_F64 ? (long long)*(const double*)p : (long long)*(const float*)p;
        }
        if (traj_off >= 0) {
            const char* p = rec + traj_off;
            b = traj_dt == LSB_U32 ? (long long)*(const uint32_t*)p : traj_dt == LSB_I32 ? (long long)*(const int32_t*)p :
                traj_dt == LSB_I64 ? *(const long long*)p : traj_dt == LSB_U64 ? (long long)*(const unsigned long long*)p :
                traj_dt == LSB_F64 ? (long long)*(const double*)p : (long long)*(const float*)p;
        }
        seg_ids[r] = a; traj_ids[r] = b;
    }
}
struct SpillTrig { double t; long long ev; int module; int pad; };
__global__ void k_spill_unit_consts(long long* __restrict__ ev, long long n_ev, long long event, long long* __restrict__ t0t,
                                    double* __restrict__ t0u, long long U, long long t0_ticks, double t0_us, SpillTrig* __restrict__ trig) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n_ev) ev[i] = event;
    if (i < U) { t0t[i] = t0_ticks; t0u[i] = t0_us; }
    // charge-only runs: "each event triggers once at perfect t0" (cli/simulate_pixels.py:222-226): time 0, module 1
    if (i == 0) { trig->t = 0.0; trig->ev = event; trig->module = 1; trig->pad = 0; }
}
// after a unit's packets are written: record (first packet, count) and move the cursor; cursor = {next free, overflow}
__global__ void k_spill_advance(long long* __restrict__ cursor, const long long* __restrict__ n_unit, long long cap,
                                long long* __restrict__ table_entry) {
    const long long base = cursor[0], n = *n_unit;
    table_entry[0] = base; table_entry[1] = n;
    if (base + n > cap) cursor[1] = 1;
    cursor[0] = base + n;
}

template <typename T>
static int spill_upload(DevBuf& b, const T* host, long long n, const T** dev) {
    int rc = b.need((size_t)(n > 0 ? n : 1) * sizeof(T));
    if (rc) return rc;
    if (n > 0) LSB_CUDA(cudaMemcpy(b.p, host, (size_t)n * sizeof(T), cudaMemcpyHostToDevice));
    *dev = (const T*)b.p;
    return 0;
}

LSB_EXPORT void lsb_spill_destroy(lsb_spill* sp);
LSB_EXPORT int lsb_spill_set_serial(lsb_spill* sp, int32_t serial) { LSB_REQUIRE(sp, "spill_set_serial: null handle"); sp->serial = serial ? 1 : 0; return 0; }
LSB_EXPORT lsb_spill* lsb_spill_create(const lsb_consts* c, const lsb_track_layout* L, const void* response, int32_t Rx, int32_t Ry,
                                       int32_t Rt, int32_t response_f64, const lsb_readout_tables* rt, int32_t n_assn, int32_t depth) {
    if (!c || !L || !rt || depth < 1 || depth > SPILL_MAX_DEPTH || n_assn < 0) { lsb_fail_arg("spill_create: bad arguments"); return nullptr; }
    lsb_spill* sp = new lsb_spill();
    sp->c = *c; sp->L = *L; sp->depth = depth; sp->n_assn = n_assn; sp->cap_packets = 0; sp->serial = 0; sp->n_sm = 148;
    { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sp->n_sm, cudaDevAttrMultiProcessorCount, dev); if (sp->n_sm < 1) sp->n_sm = 148; } sp->host_table = nullptr; sp->host_table_cap = 0;
    for (int k = 0; k < SPILL_MAX_DEPTH; k++) { sp->ch[k] = nullptr; sp->ev_export[k] = nullptr; }
    cudaEventCreateWithFlags(&sp->ev_gather, cudaEventDisableTiming);
    for (int k = 0; k < depth; k++) {
        sp->ch[k] = lsb_chain_create(c, L, response, Rx, Ry, Rt, response_f64, 0, 0);
        if (!sp->ch[k]) { lsb_spill_destroy(sp); return nullptr; }
        sp->ch[k]->rng_fresh = 1;
        cudaEventCreateWithFlags(&sp->ev_export[k], cudaEventDisableTiming);
    }
    PktTables& T = sp->T;
    T.clock_cycle = rt->clock_cycle; T.adc_pedestal = rt->adc_pedestal; T.mus = rt->mus; T.s = rt->s;
    T.reset_period = rt->clock_reset_period; T.light_trig_mode = rt->light_trig_mode;
    T.npx = rt->n_pixels[0]; T.npy = rt->n_pixels[1]; T.nptx = rt->n_pixels_per_tile[0]; T.npty = rt->n_pixels_per_tile[1];
    T.ntx = rt->n_tiles_xy[0]; T.nty = rt->n_tiles_xy[1];
    T.n_tiles = rt->n_tiles; T.n_modules = rt->n_modules; T.max_groups = rt->max_groups; T.n_iog = rt->n_io_groups; T.n_bad = rt->n_bad;
    const long long* bad = nullptr;
    if (spill_upload(sp->tab[0], rt->tile_map, 2LL * T.ntx * T.nty, &T.tile_map) ||
        spill_upload(sp->tab[1], rt->tile_orientation, 2LL * T.n_tiles, &T.tile_orient) ||
        spill_upload(sp->tab[2], rt->pixel_connection, (long long)T.nptx * T.npty, &T.pix_conn) ||
        spill_upload(sp->tab[3], rt->tile_chip_to_io, 256LL * T.n_tiles, &T.tile_chip_io) ||
        spill_upload(sp->tab[4], rt->module_n_groups, (long long)T.n_modules, &T.module_ng) ||
        spill_upload(sp->tab[5], rt->module_io_groups, (long long)T.n_modules * T.max_groups, &T.module_io) ||
        spill_upload(sp->tab[6], rt->io_groups, (long long)T.n_iog, &T.io_groups) ||
        spill_upload(sp->tab[7], (const long long*)rt->bad_channels, (long long)T.n_bad, &bad)) { lsb_spill_destroy(sp); return nullptr; }
    T.bad = bad;
    return sp;
}
LSB_EXPORT void lsb_spill_destroy(lsb_spill* sp) {
    if (!sp) return;
    cudaDeviceSynchronize();
    for (int k = 0; k < SPILL_MAX_DEPTH; k++) {
        if (sp->ch[k]) lsb_chain_destroy(sp->ch[k]);
        if (sp->ev_export[k]) cudaEventDestroy(sp->ev_export[k]);
        SpillScratch& x = sp->sx[k];
        DevBuf* all[] = {&x.ev, &x.t0t, &x.t0u, &x.info, &x.sc, &x.tiles, &x.count, &x.offs, &x.bsum, &x.total, &x.trig};
        for (DevBuf* b : all) b->release();
    }
    cudaEventDestroy(sp->ev_gather);
    for (DevBuf& b : sp->tab) b.release();
    DevBuf* all[] = {&sp->gathered, &sp->seg_ids, &sp->traj_ids, &sp->out_packets, &sp->out_rows, &sp->out_src, &sp->unit_table, &sp->cursor};
    for (DevBuf* b : all) b->release();
    if (sp->host_table) cudaFreeHost(sp->host_table);
    delete sp;
}

// hits of the unit a handle has just been given -> packets + truth rows appended to the rank's output (queued on the handle's
// stream; waits for the previous unit's append, the cursor is shared)
static int spill_export_unit(lsb_spill* sp, int k, const lsb_chain_result* r, long long unit_slot, long long gathered_off,
                             long long event, double t0_us, cudaEvent_t prev_export) {
    lsb_chain* h = sp->ch[k];
    SpillScratch& x = sp->sx[k];
    cudaStream_t st = h->hp;
    const long long U = r->n_unique_pixels;
    const int A = sp->c.max_adc_values, K = sp->c.max_tracks_per_pixel;
    LSB_REQUIRE(K <= ASSN_MAXK, "spill: MAX_TRACKS_PER_PIXEL above the supported 128");
    const long long N = U * (long long)A;
    LSB_REQUIRE(N < 2147483647LL, "spill: more than 2^31 hit slots in one unit");
    int rc;
    if ((rc = x.ev.need((size_t)N * 8)) || (rc = x.t0t.need((size_t)U * 8)) || (rc = x.t0u.need((size_t)U * 8)) ||
        (rc = x.info.need((size_t)U * sizeof(PixInfo))) || (rc = x.sc.need((size_t)N * sizeof(Scan3))) ||
        (rc = x.tiles.need((size_t)(scan_num_blocks(N) + 1) * sizeof(Scan3))) || (rc = x.count.need((size_t)N * 4)) ||
        (rc = x.offs.need((size_t)N * 8)) || (rc = x.bsum.need((size_t)(scan_num_blocks(N) + 1) * 8)) || (rc = x.total.need(8)) ||
        (rc = x.trig.need(32))) return rc;
    const long long t0_ticks = (long long)(t0_us / sp->T.clock_cycle);        // int(event_start_time / CLOCK_CYCLE), fee.py:137
    k_spill_unit_consts<<<lsb_blocks(N, 256), 256, 0, st>>>((long long*)x.ev.p, N, event, (long long*)x.t0t.p, (double*)x.t0u.p, U, t0_ticks, t0_us, (SpillTrig*)x.trig.p);
    LSB_LAUNCH_CHECK("k_spill_unit_consts");
    PktTrig G; G.n = 1; G.t = (const double*)x.trig.p; G.ev = (const long long*)((char*)x.trig.p + 8); G.module = (const int*)((char*)x.trig.p + 16);
    const PktTables& T = sp->T;
    k_pkt_slots<<<lsb_blocks(U, 128), 128, 0, st>>>(T, U, A, r->unique_pix, r->adc_digit, r->adc_ticks_list, (const long long*)x.t0t.p,
                                                    (PixInfo*)x.info.p, (Scan3*)x.sc.p);
    LSB_LAUNCH_CHECK("k_pkt_slots");
    if ((rc = scan3_inclusive((Scan3*)x.sc.p, N, (Scan3*)x.tiles.p, st))) return rc;
    k_pkt_count<<<lsb_blocks(N, 256), 256, 0, st>>>(T, G, N, A, (const Scan3*)x.sc.p, r->adc_ticks_list, (const long long*)x.t0t.p,
                                                    (const long long*)x.ev.p, (uint32_t*)x.count.p);
    LSB_LAUNCH_CHECK("k_pkt_count");
    if ((rc = exclusive_scan<uint32_t, long long>((const uint32_t*)x.count.p, N, (long long*)x.offs.p, (long long*)x.bsum.p, (long long*)x.total.p, st))) return rc;
    if (prev_export) LSB_CUDA(cudaStreamWaitEvent(st, prev_export, 0));
    long long* cursor = (long long*)sp->cursor.p;
    k_pkt_write<<<lsb_blocks(N, 256), 256, 0, st>>>(T, G, N, A, (const Scan3*)x.sc.p, r->adc_ticks_list, r->adc_digit, (const long long*)x.t0t.p,
                                                    (const double*)x.t0u.p, (const long long*)x.ev.p, (const PixInfo*)x.info.p,
                                                    (const long long*)x.offs.p, sp->cap_packets, (lsb_packet*)sp->out_packets.p,
                                                    (long long*)sp->out_src.p, cursor);
    LSB_LAUNCH_CHECK("k_pkt_write");
    k_pkt_assn<<<sp->n_sm * 8, 32 * ASSN_WARPS, 0, st>>>(0, (const long long*)sp->out_src.p, A, K, sp->n_assn, (const long long*)x.ev.p,
                                                    r->current_fractions, (const long long*)r->track_pixel_map, (const long long*)r->track_pixel_map,
                                                    (char*)sp->out_rows.p, (const long long*)x.total.p, cursor, sp->cap_packets,
                                                    (const long long*)sp->seg_ids.p + gathered_off, (const long long*)sp->traj_ids.p + gathered_off);
    LSB_LAUNCH_CHECK("k_pkt_assn");
    k_spill_advance<<<1, 1, 0, st>>>(cursor, (const long long*)x.total.p, sp->cap_packets, (long long*)sp->unit_table.p + 2 * unit_slot);
    LSB_LAUNCH_CHECK("k_spill_advance");
    LSB_CUDA(cudaEventRecord(sp->ev_export[k], st));
    return 0;
}

// tracks_dev: the whole (selected, quenched, drifted) record array on the device.  order_dev: the row permutation of
// lsb_batch_units (or NULL: identity).  Unit i of this call = rows order[unit_begin[i] .. + unit_count[i]), event id
// unit_event[i] starting at unit_t0_us[i], RNG seed unit_seed[i].  Output: packets / mc_packets_assn rows of all units, unit
// after unit, in device buffers owned by the runner (valid until the next run); unit_packets_host[i] = packets of unit i.
// Returns LSB_SPILL_OVERFLOW (= -2) if cap_packets was too small: out->n_packets then holds the required capacity.
#define LSB_SPILL_OVERFLOW (-2)
LSB_EXPORT int lsb_spill_run(lsb_spill* sp, const void* tracks_dev, const int64_t* order_dev, int64_t n_units,
                             const int64_t* unit_begin, const int64_t* unit_count, const int64_t* unit_event,
                             const double* unit_t0_us, const uint64_t* unit_seed, int32_t seg_id_offset, int32_t seg_id_dtype,
                             int32_t traj_id_offset, int32_t traj_id_dtype, int64_t cap_packets, int64_t* unit_packets_host,
                             lsb_spill_result* out, void* stream) {
    LSB_REQUIRE(sp && out && (n_units == 0 || (unit_begin && unit_count && unit_event && unit_t0_us && unit_seed)), "spill_run: null pointer");
    memset(out, 0, sizeof(*out));
    cudaStream_t st = (cudaStream_t)stream;
    const int itemsize = sp->L.itemsize;
    LSB_REQUIRE(itemsize % 4 == 0, "spill_run: record size must be a multiple of 4 bytes");
    long long total = 0;
    for (long long i = 0; i < n_units; i++) { LSB_REQUIRE(unit_count[i] >= 0, "spill_run: negative unit size"); total += unit_count[i]; }
    out->n_units = n_units; out->n_segments = total;
    if (n_units == 0) return 0;
    LSB_REQUIRE(tracks_dev || total == 0, "spill_run: null records");
    int rc;
    if (cap_packets < 1024) cap_packets = 1024;
    const long long rowbytes = 8LL + 32LL * sp->n_assn;
    if ((rc = sp->gathered.need((size_t)(total ? total : 1) * itemsize)) || (rc = sp->seg_ids.need((size_t)(total ? total : 1) * 8)) ||
        (rc = sp->traj_ids.need((size_t)(total ? total : 1) * 8)) || (rc = sp->out_packets.need((size_t)cap_packets * sizeof(lsb_packet))) ||
        (rc = sp->out_rows.need((size_t)cap_packets * rowbytes)) || (rc = sp->out_src.need((size_t)cap_packets * 8)) ||
        (rc = sp->unit_table.need((size_t)n_units * 16)) || (rc = sp->cursor.need(16))) return rc;
    sp->cap_packets = cap_packets;
    if (sp->host_table_cap < 2 * n_units + 2) {
        if (sp->host_table) cudaFreeHost(sp->host_table);
        sp->host_table = nullptr;
        LSB_CUDA(cudaMallocHost((void**)&sp->host_table, (size_t)(2 * n_units + 2) * 8));
        sp->host_table_cap = 2 * n_units + 2;
    }
    LSB_CUDA(cudaMemsetAsync(sp->unit_table.p, 0, (size_t)n_units * 16, st));
    LSB_CUDA(cudaMemsetAsync(sp->cursor.p, 0, 16, st));
    // ---- gather the units' records (and their file ids) ----------------------------------------------------------------
    {
        long long off = 0;
        for (long long i = 0; i < n_units; i++) {
            const long long n = unit_count[i];
            if (n == 0) continue;
            k_spill_gather<<<lsb_blocks(n * (itemsize >> 2), 256), 256, 0, st>>>((const char*)tracks_dev, order_dev ? (const long long*)order_dev + unit_begin[i] : nullptr,
                                                                                n, itemsize, (char*)sp->gathered.p + off * itemsize, seg_id_offset, seg_id_dtype,
                                                                                traj_id_offset, traj_id_dtype, (long long*)sp->seg_ids.p + off,
                                                                                (long long*)sp->traj_ids.p + off);
            LSB_LAUNCH_CHECK("k_spill_gather");
            off += n;
        }
        LSB_CUDA(cudaEventRecord(sp->ev_gather, st));
    }
    // ---- the units, `depth` in flight ---------------------------------------------------------------------------------
    cudaEvent_t prev_export = nullptr;
    long long off = 0;
    lsb_chain_result done;
    FILE* timeline = getenv("LSB_SPILL_TIMELINE") ? fopen(getenv("LSB_SPILL_TIMELINE"), "a") : nullptr;     // diagnostics (tools/timeline.py)
    if (timeline) fprintf(timeline, "# run\n");
    auto collect = [&](int k) -> int {
        lsb_chain* h = sp->ch[k];
        if (!h->pending_valid) return 0;
        int rc2 = lsb_chain_wait(h, &done);
        if (rc2) return rc2;
        out->n_hits += done.n_hits; out->n_unique_pixels += done.n_unique_pixels;
        if (done.n_samples > 0) out->n_samples += done.n_samples;
        out->n_fma += done.n_fma;
        if (timeline && done.unique_pix)        // ms since the library's reference event: front begin/end, MC begin/end, FEE begin, done
            fprintf(timeline, "%lld %lld %.4f %.4f %.4f %.4f %.4f %.4f\n", (long long)done.n_segments, (long long)done.n_unique_pixels, done.stage_ms[0],
                    done.stage_ms[1], done.stage_ms[2], done.stage_ms[3], done.stage_ms[4], done.stage_ms[5]);
        out->pair_ticks += done.n_pairs * done.n_ticks; out->pixel_ticks += done.n_unique_pixels * (long long)sp->c.n_time_ticks;
        return 0;
    };
    for (int k = 0; k < sp->depth; k++) LSB_CUDA(cudaStreamWaitEvent(sp->ch[k]->hp, sp->ev_gather, 0));
    for (long long i = 0; i < n_units; i++) {
        const long long n = unit_count[i];
        if (n == 0) continue;
        const int k = (int)(i % sp->depth);
        lsb_chain* h = sp->ch[k];
        if ((rc = collect(k))) return rc;
        if ((rc = chain_enqueue(h, (char*)sp->gathered.p + off * itemsize, n, -1, unit_seed[i], 1, &h->pending, h->hp, sp->serial ? h->hp : h->lp))) return rc;
        if (h->pending.unique_pix && h->pending.n_unique_pixels > 0) {
            if ((rc = spill_export_unit(sp, k, &h->pending, i, off, unit_event[i], unit_t0_us[i], prev_export))) return rc;
            prev_export = sp->ev_export[k];
        }
        LSB_CUDA(cudaEventRecord(h->ev_done, h->hp));
        h->pending_valid = 1;
        off += n;
    }
    for (int k = 0; k < sp->depth; k++) if ((rc = collect(k))) { if (timeline) fclose(timeline); return rc; }
    if (timeline) fclose(timeline);
    // every handle has drained (lsb_chain_wait synchronises on its last event): the table is final
    LSB_CUDA(cudaMemcpyAsync(sp->host_table, sp->unit_table.p, (size_t)n_units * 16, cudaMemcpyDeviceToHost, st));
    LSB_CUDA(cudaMemcpyAsync(sp->host_table + 2 * n_units, sp->cursor.p, 16, cudaMemcpyDeviceToHost, st));
    LSB_CUDA(cudaStreamSynchronize(st));
    long long n_packets = 0;
    for (long long i = 0; i < n_units; i++) { if (unit_packets_host) unit_packets_host[i] = sp->host_table[2 * i + 1]; n_packets += sp->host_table[2 * i + 1]; }
    out->n_packets = n_packets;
    out->packets = (const lsb_packet*)sp->out_packets.p; out->assn_rows = sp->out_rows.p;
    out->records = sp->gathered.p; out->assn_row_bytes = rowbytes;
    if (sp->host_table[2 * n_units + 1] || n_packets > cap_packets) { out->overflow = 1; return LSB_SPILL_OVERFLOW; }
    return 0;
}

// tracks[mask] of the reference (cli/simulate_pixels.py:670, batching): dst[r] = src[order[r]]
LSB_EXPORT int lsb_gather_records(const void* src_dev, const int64_t* order_dev, int64_t n, int32_t itemsize, void* dst_dev, void* stream) {
    if (n <= 0) return 0;
    LSB_REQUIRE(src_dev && dst_dev && itemsize > 0 && itemsize % 4 == 0, "gather_records: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    TmpPool pool(st);
    long long* ids = nullptr;                 // the id outputs of the kernel are not wanted here
    LSB_CUDA(pool.get(&ids, 2 * n));
    k_spill_gather<<<lsb_blocks(n * (itemsize >> 2), 256), 256, 0, st>>>((const char*)src_dev, (const long long*)order_dev, n, itemsize, (char*)dst_dev,
                                                                        -1, 0, -1, 0, ids, ids + n);
    LSB_LAUNCH_CHECK("k_spill_gather");
    return 0;
}

// Variable-length blocks -> one contiguous buffer: block b = src[b] .. + bytes[b] goes to dst + dst_off[b] (16-byte
// granularity when everything is aligned, else bytes).  Used to put the units gathered from all ranks into file order.
struct BlockCopy { const char* src; long long dst_off; long long bytes; };
__global__ void k_copy_blocks(const BlockCopy* __restrict__ blocks, char* __restrict__ dst) {
    const BlockCopy b = blocks[blockIdx.y];
    char* d = dst + b.dst_off;
    const uintptr_t align = (uintptr_t)b.src | (uintptr_t)d | (uintptr_t)b.bytes;
    if ((align & 15) == 0) {
        const long long n = b.bytes >> 4;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
            reinterpret_cast<int4*>(d)[i] = __ldg(reinterpret_cast<const int4*>(b.src) + i);
    } else if ((align & 7) == 0) {                       // mc_packets_assn rows: 8 + 32 n bytes, 8-byte aligned
        const long long n = b.bytes >> 3;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
            reinterpret_cast<long long*>(d)[i] = __ldg(reinterpret_cast<const long long*>(b.src) + i);
    } else {
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < b.bytes; i += (long long)gridDim.x * blockDim.x) d[i] = b.src[i];
    }
}
LSB_EXPORT int lsb_copy_blocks(int64_t n_blocks, const void* const* src_dev, const int64_t* dst_off, const int64_t* bytes, void* dst_dev,
                               void* stream) {
    if (n_blocks <= 0) return 0;
    LSB_REQUIRE(src_dev && dst_off && bytes && dst_dev, "copy_blocks: null pointer");
    LSB_REQUIRE(n_blocks <= 65535, "copy_blocks: more than 65535 blocks in one call");
    cudaStream_t st = (cudaStream_t)stream;
    TmpPool pool(st);
    BlockCopy* d = nullptr;
    LSB_CUDA(pool.get(&d, n_blocks));
    BlockCopy* h = (BlockCopy*)malloc(sizeof(BlockCopy) * (size_t)n_blocks);
    if (!h) return lsb_fail_arg("copy_blocks: out of host memory");
    long long mx = 0;
    for (long long i = 0; i < n_blocks; i++) { h[i].src = (const char*)src_dev[i]; h[i].dst_off = dst_off[i]; h[i].bytes = bytes[i]; if (bytes[i] > mx) mx = bytes[i]; }
    cudaError_t e = cudaMemcpyAsync(d, h, sizeof(BlockCopy) * (size_t)n_blocks, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    free(h);
    if (e != cudaSuccess) return lsb_fail_cuda(e, "copy_blocks upload");
    long long per = (mx / 16 + 255) / 256;
    unsigned gx = (unsigned)(per < 1 ? 1 : (per > 64 ? 64 : per));
    k_copy_blocks<<<dim3(gx, (unsigned)n_blocks), 256, 0, st>>>(d, (char*)dst_dev);
    LSB_LAUNCH_CHECK("k_copy_blocks");
    return 0;
}

// ---- per-rank device -> host copies into a host buffer shared by the ranks of a node ------------------------------------------------
// With N ranks on one node the packets and truth rows of a spill end in ONE host table (what fee.export_to_hdf5 receives).  Rank 0
// maps a shared-memory file, every rank registers the same pages with its CUDA context and copies its own units' blocks straight to
// their file-order positions: N PCIe links instead of one, no gather through rank 0's HBM.
LSB_EXPORT int lsb_host_register(void* host_ptr, int64_t bytes) {
    LSB_REQUIRE(host_ptr && bytes > 0, "host_register: null pointer");
    LSB_CUDA(cudaHostRegister(host_ptr, (size_t)bytes, cudaHostRegisterPortable));
    return 0;
}
LSB_EXPORT int lsb_host_unregister(void* host_ptr) {
    LSB_REQUIRE(host_ptr, "host_unregister: null pointer");
    LSB_CUDA(cudaHostUnregister(host_ptr));
    return 0;
}
LSB_EXPORT int lsb_d2h_blocks(int64_t n_blocks, const void* const* src_dev, const int64_t* dst_off, const int64_t* bytes, void* dst_host,
                              void* stream) {
    if (n_blocks <= 0) return 0;
    LSB_REQUIRE(src_dev && dst_off && bytes && dst_host, "d2h_blocks: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    for (long long i = 0; i < n_blocks; i++) {
        if (bytes[i] <= 0) continue;
        LSB_CUDA(cudaMemcpyAsync((char*)dst_host + dst_off[i], src_dev[i], (size_t)bytes[i], cudaMemcpyDeviceToHost, st));
    }
    return 0;
}
