"""One spill / file through the whole charge path, partitioned over the ranks of a ``torch.distributed`` job.

This is the reference's ``run_simulation`` from the active-volume cut to the packet list
(cli/simulate_pixels.py:667-671, 727-742, 864-1117, save_results :179-236 -> fee.export_to_hdf5 fee.py:84-359) with
the loop itself native (``lsb_spill_run``, csrc/spill.cuh):

    select_active_volume -> quench, drift (whole file) -> TPCBatcher plan (one device pass) -> units assigned to the
    ranks longest-first -> every rank: its units through ``depth`` chains in flight, packets + ``mc_packets_assn`` rows
    appended on the device -> one NCCL exchange to rank 0 -> blocks put into file order on the device -> (optional) one
    D2H copy into pinned host arrays.  With host output and several ranks on one node the exchange is skipped: the ranks map
    one shared host table and each copies its own units' blocks straight to their file-order positions (N PCIe links).

Units (event x TPC group) are independent.  Every unit draws from
``create_xoroshiro128p_states(n, seed = rand_seed + unit number)``, so the packets of a unit do not depend on which rank or
chain processed it: the output of N ranks is bit-identical to the output of one.
"""
import ctypes as C
import os
import socket

import numpy as np
import torch
import torch.distributed as dist

from . import _abi, _launch as _l, consts as _consts
from . import active_volume as _av, packets as _p, quenching as _q, drifting as _d, fee as _fee
from .util import batching as _bt


class _SpillResult(C.Structure):
    _fields_ = [("n_units", C.c_int64), ("n_segments", C.c_int64), ("n_packets", C.c_int64), ("n_hits", C.c_int64),
                ("n_unique_pixels", C.c_int64), ("n_samples", C.c_int64), ("n_fma", C.c_int64), ("pair_ticks", C.c_int64),
                ("pixel_ticks", C.c_int64), ("packets", C.c_void_p),
                ("assn_rows", C.c_void_p), ("records", C.c_void_p), ("assn_row_bytes", C.c_int64), ("overflow", C.c_int32),
                ("pad", C.c_int32)]


def unit_costs(sizes):
    """Work estimate per unit for the longest-first assignment: the chain's time is linear in the number of segments plus
    a per-call floor (front-end state machine over every tick, ~50 launches): measured ~0.9 us per segment + ~0.7 ms."""
    s = np.asarray(sizes, dtype=np.float64)
    return np.where(s > 0, s + 800.0, 0.0)


def assign_units(sizes, world):
    """Longest-processing-time-first (deterministic; every rank computes the same plan) -> list of ascending unit lists."""
    cost = unit_costs(sizes)
    order = sorted(range(len(cost)), key=lambda i: (-cost[i], i))
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for i in order:
        if sizes[i] == 0:
            continue
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += cost[i]
    for lst in out:
        lst.sort()
    return out


def file_order_blocks(unit_packets, plan, n_event_packets, n_tpc_batches):
    """Where every block of packets goes in the output file (the order in which the reference's sequential loop appends them:
    per event its between-batch packets, then its batches in TPC-group order).  ``unit_packets[u]``: packets of unit u;
    ``plan[r]``: ascending units of rank r (its output buffer holds them back to back); ``n_event_packets[e]``: between-batch
    packets of event e.  Returns ``(blocks, n_total)`` with blocks = list of ``(source, source_offset, dest_offset, count)``,
    source = rank number or -1 for the event-level packets (offsets count packets).  Pure host logic: every rank can compute it."""
    n_units = len(unit_packets)
    owner = {}
    for r, lst in enumerate(plan):
        for u in lst:
            owner[int(u)] = r
    run = [0] * len(plan)
    blocks, pos, ev_off = [], 0, 0
    n_events = n_units // n_tpc_batches if n_tpc_batches else 0
    for e in range(n_events):
        ne = int(n_event_packets[e]) if e < len(n_event_packets) else 0
        if ne:
            blocks.append((-1, ev_off, pos, ne))
            pos += ne
            ev_off += ne
        for b in range(n_tpc_batches):
            u = e * n_tpc_batches + b
            n = int(unit_packets[u])
            if n == 0:
                continue
            r = owner[u]
            blocks.append((r, run[r], pos, n))
            run[r] += n
            pos += n
    return blocks, pos


class SharedHostUnavailable(RuntimeError):
    """raised on EVERY rank of the group when the shared host table cannot be set up on one of them"""


class SharedHostTable:
    """A byte table in shared memory mapped by every rank of a process group on ONE node, page-locked in every rank's CUDA
    context (``register(ptr, nbytes)`` / ``unregister(ptr)``; None: plain shared memory, e.g. CPU tests).  ``ensure`` is a
    collective: every rank calls it with the same size."""
    _serial = 0

    def __init__(self, group, rank, register=None, unregister=None, directory="/dev/shm"):
        self.group, self.rank = group, rank
        self._reg, self._unreg = register, unregister
        self.dir = directory
        self.mm = None
        self.nbytes = 0

    def ensure(self, nbytes):
        if self.mm is not None and self.nbytes >= nbytes:
            return self.mm
        self.close()
        size = int(nbytes * 1.25) + 4096
        obj = [None]
        if self.rank == 0:
            SharedHostTable._serial += 1
            path = os.path.join(self.dir, "lsb_%d_%d" % (os.getpid(), SharedHostTable._serial))
            try:
                with open(path, "wb") as f:
                    f.truncate(size)
                obj = [path]
            except OSError:
                obj = [None]
        dist.broadcast_object_list(obj, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        if obj[0] is None:                             # (every rank sees the same None: the failure is collective)
            raise SharedHostUnavailable("cannot create a shared-memory file under %s" % self.dir)
        mm, err = None, None
        try:
            mm = np.memmap(obj[0], dtype=np.uint8, mode="r+", shape=(size,))
        except (OSError, ValueError) as e:
            err = "mmap: %s" % e
        dist.barrier(group=self.group)                 # everybody has tried to map it: the name can go, the pages stay
        if self.rank == 0:
            os.unlink(obj[0])
        registered = False
        if err is None and self._reg is not None:
            try:
                self._reg(mm.ctypes.data, size)
                registered = True
            except Exception as e:                     # e.g. a locked-memory limit: must not leave the other ranks waiting
                err = "register: %s" % e
        world = dist.get_world_size(self.group)
        errs = [None] * world
        dist.all_gather_object(errs, err, group=self.group)
        if any(e is not None for e in errs):
            if registered and self._unreg is not None:
                self._unreg(mm.ctypes.data)
            raise SharedHostUnavailable("; ".join("rank %d: %s" % (r, e) for r, e in enumerate(errs) if e is not None))
        self.mm, self.nbytes = mm, size
        return mm

    def close(self):
        if self.mm is not None:
            if self._unreg is not None:
                try:
                    self._unreg(self.mm.ctypes.data)
                except Exception:
                    pass
            self.mm = None
            self.nbytes = 0


class SpillOutput:
    """What rank 0 holds after a spill (other ranks: ``packets is None``)."""

    def __init__(self):
        self.packets = self.packets_mc_ds = None          # NumPy structured arrays (host=True) or torch uint8 CUDA tensors
        self.tracks = None                                # selected, quenched, drifted records
        self.unit_sizes = self.unit_packets = None
        self.n_segments = self.n_packets = 0
        self.stats = {}


class SpillRunner:
    def __init__(self, track_dtype, response, depth=4, tpc_batch_size=None, event_separator=None, group=None, provider=None,
                 single_rank=False):
        """``single_rank=True``: ignore the process group and simulate every unit here (e.g. to check a distributed run)."""
        self._prov = provider or _consts.provider()
        p = self._prov
        self.dtype = np.dtype(track_dtype)
        self.sep = event_separator or getattr(p.sim, "EVENT_SEPARATOR", "event_id")
        self.tpc_batch_size = int(tpc_batch_size or getattr(p.sim, "EVENT_BATCH_SIZE", 2))
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() and not single_rank else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() and not single_rank else 0
        self._c = _consts.snapshot(p)
        self._L = _abi.track_layout(self.dtype)
        self._resp = _l.dev(response, name="response")
        self.tables = _p.ReadoutTables.from_consts(p)
        self.n_assn = int(self.tables.association_count)
        self.assn_dtype = _p.assn_dtype(self.n_assn)
        r = self._resp
        lib = _l.lib()
        lib.lsb_spill_create.restype = C.c_void_p
        self._h = lib.lsb_spill_create(C.byref(self._c), C.byref(self._L), r.c, C.c_int32(r.shape[0]), C.c_int32(r.shape[1]),
                                       C.c_int32(r.shape[2]), C.c_int32(1 if r.dtype == np.dtype("f8") else 0),
                                       C.byref(self.tables._c), C.c_int32(self.n_assn), C.c_int32(int(depth)))
        if not self._h:
            raise _abi.LsbError("lsb_spill_create failed: %s" % lib.lsb_last_error().decode())
        self._cap = 0
        self._stage = {}          # device / pinned staging buffers reused from spill to spill
        self._shared = None       # host tables shared by the ranks (host output, world > 1, one node); False: not available
        if self.world > 1 and os.environ.get("LSB_SPILL_SHARED_HOST", "1") == "0":
            self._shared = False
        f = self.dtype.fields
        self._seg = (f["segment_id"][1], _abi._DTYPE_CODE[np.dtype(f["segment_id"][0])]) if "segment_id" in f else (-1, 0)
        self._traj = (f["file_traj_id"][1], _abi._DTYPE_CODE[np.dtype(f["file_traj_id"][0])]) if "file_traj_id" in f else (-1, 0)

    def set_serial(self, on=True):
        """every stage of a unit on one stream (per-kernel timing with events)"""
        _l.check(_l.lib().lsb_spill_set_serial(C.c_void_p(self._h), C.c_int32(1 if on else 0)), "spill_set_serial")

    def close(self):
        if getattr(self, "_shared", None):
            for t in self._shared:
                t.close()
            self._shared = None
        if getattr(self, "_h", None) and _l is not None:
            _l.lib().lsb_spill_destroy(C.c_void_p(self._h))
            self._h = None

    __del__ = close

    # ------------------------------------------------------------------------------------------------------------------
    def _buf(self, name, nbytes, pinned=False):
        t = self._stage.get(name)
        if t is None or t.numel() < nbytes:
            n = int(nbytes * 1.25) + 4096
            t = torch.empty(n, dtype=torch.uint8, pin_memory=True) if pinned else torch.empty(n, dtype=torch.uint8, device="cuda")
            self._stage[name] = t
        return t

    def _shared_tables(self):
        """(packets, truth rows) host tables shared by the ranks, or None when the ranks are not on one node"""
        if self._shared is None:
            names = [None] * self.world
            dist.all_gather_object(names, socket.gethostname(), group=self.group)
            if len(set(names)) != 1:
                self._shared = False
            else:
                lib = _l.lib()

                def reg(ptr, n):
                    _l.check(lib.lsb_host_register(C.c_void_p(ptr), C.c_int64(n)), "host_register")

                def unreg(ptr):
                    lib.lsb_host_unregister(C.c_void_p(ptr))
                self._shared = (SharedHostTable(self.group, self.rank, reg, unreg), SharedHostTable(self.group, self.rank, reg, unreg))
        return self._shared or None

    def _event_level_packets(self, events, event_times):
        """Packets the reference writes between the batches of its loop (cli/simulate_pixels.py:872-887): the sync packets
        that have become due and the timestamp + trigger packets of every new event.  A few records per event: host code
        (fee.export_sync_to_hdf5 / export_timestamp_trigger_to_hdf5), as in the reference."""
        d, li = self._prov.detector, self._prov.light
        period = d.CLOCK_RESET_PERIOD * d.CLOCK_CYCLE
        t_first = float(event_times[0]) if len(event_times) else 0.0
        sync_start = t_first // period * period + period
        out = []
        for ev, t0 in zip(events, event_times):
            pk = []
            if t0 - sync_start >= 0:
                sync_times = np.arange(sync_start, t0 + 1, period)
                if len(sync_times):
                    pk.append(_fee.export_sync_to_hdf5(None, np.full(sync_times.shape, period))[0])
                    sync_start = sync_times[-1] + period
            if li.LIGHT_TRIG_MODE in (0, 1):
                pk.append(_fee.export_timestamp_trigger_to_hdf5(None, [t0])[0])
            out.append(np.concatenate(pk) if pk else np.zeros(0, dtype=_p.PACKET_DTYPE))
        return out

    def simulate(self, tracks, events=None, event_times=None, rand_seed=0, host_output=True, return_tracks=False):
        """``tracks``: structured host array (ideally a view of pinned memory) or device records of the whole file.
        ``events``: sorted unique values of the event field (computed when omitted).  ``event_times``: start time [us] per
        event (default: the reference's spill clock, ``(event % MAX_EVENTS_PER_FILE) * SPILL_PERIOD``).
        Returns a :class:`SpillOutput`; on rank 0 it holds every unit's packets in file order."""
        p = self._prov
        det, sim = p.detector, p.sim
        lib = _l.lib()
        st = _l.stream()
        out = SpillOutput()
        # (0) records on the device
        if isinstance(tracks, np.ndarray):
            if tracks.dtype != self.dtype:
                raise TypeError("tracks dtype differs from the dtype this runner was created for")
            raw = torch.from_numpy(tracks.reshape(-1).view(np.uint8))
            d_all = _l.DeviceRecords(dtype=self.dtype, n=len(tracks), buf=raw.to("cuda", non_blocking=True))
            if events is None:
                events = np.unique(tracks[self.sep])
        else:
            d_all = tracks
            if events is None:
                off, nb = self.dtype.fields[self.sep][1], np.dtype(self.dtype.fields[self.sep][0]).itemsize
                col = d_all.buf.view(len(d_all), self.dtype.itemsize)[:, off:off + nb].contiguous().view(-1)
                events = np.unique(col.cpu().numpy().view(self.dtype.fields[self.sep][0]))
        events = np.asarray(events)
        # (1) active volume cut (cli/simulate_pixels.py:667-671) -> compacted records
        first, idx, _ = _av.classify(d_all, det.TPC_BORDERS)
        S = int(idx.numel())
        if S == len(d_all):
            d_sel = d_all
        else:
            d_sel = _l.DeviceRecords(dtype=self.dtype, n=S)
            _l.check(lib.lsb_gather_records(C.c_void_p(d_all.buf.data_ptr()), C.c_void_p(idx.data_ptr()), C.c_int64(S),
                                            C.c_int32(self.dtype.itemsize), C.c_void_p(d_sel.buf.data_ptr()), st), "gather_records")
        out.n_segments = S
        # (2) quench, drift over the whole file (:732, :742)
        _q.quench[1, 1](d_sel, int(p.physics.BIRKS))
        _d.drift[1, 1](d_sel)
        # (3) every (event, TPC group) unit of the run in one device pass (:864, util/batching.py:40-67)
        batcher = _bt.TPCBatcher(d_sel, d_sel, self.sep, tpc_batch_size=self.tpc_batch_size, tpc_borders=det.TPC_BORDERS, events=events)
        sizes = batcher.unit_sizes
        offsets = batcher.unit_offsets
        nB = batcher.n_tpc_batches
        nU = len(sizes)
        out.unit_sizes = sizes
        if event_times is None:
            mx = int(getattr(sim, "MAX_EVENTS_PER_FILE", 1000))
            event_times = (events.astype(np.int64) % mx) * float(getattr(sim, "SPILL_PERIOD", 1.2e6))
        event_times = np.asarray(event_times, dtype=np.float64)
        # (4) partition
        plan = assign_units(sizes, self.world)
        mine = np.asarray(plan[self.rank], dtype=np.int64)
        n_mine = len(mine)
        begin = np.ascontiguousarray(offsets[mine], dtype=np.int64)
        count = np.ascontiguousarray(sizes[mine], dtype=np.int64)
        ev_of = np.ascontiguousarray(events[mine // nB].astype(np.int64))
        t0_of = np.ascontiguousarray(event_times[mine // nB], dtype=np.float64)
        seeds = np.ascontiguousarray((int(rand_seed) + mine).astype(np.uint64))
        upk = np.zeros(max(n_mine, 1), dtype=np.int64)
        res = _SpillResult()
        if self._cap == 0:
            self._cap = int(4 * count.sum()) + 65536

        def ptr(a):
            return a.ctypes.data_as(C.c_void_p)
        for attempt in range(3):
            rc = lib.lsb_spill_run(C.c_void_p(self._h), C.c_void_p(d_sel.buf.data_ptr()), C.c_void_p(batcher.order_dev.data_ptr()),
                                   C.c_int64(n_mine), ptr(begin), ptr(count), ptr(ev_of), ptr(t0_of), ptr(seeds),
                                   C.c_int32(self._seg[0]), C.c_int32(self._seg[1]), C.c_int32(self._traj[0]), C.c_int32(self._traj[1]),
                                   C.c_int64(self._cap), ptr(upk), C.byref(res), st)
            if rc == -2:                                   # output capacity too small: grow and run the share again
                self._cap = int(res.n_packets * 1.2) + 65536
                continue
            _l.check(rc, "spill_run")
            break
        else:
            raise _abi.LsbError("spill_run: packet buffer kept overflowing")
        n_local = int(res.n_packets)
        out.stats = dict(n_hits=int(res.n_hits), n_unique_pixels=int(res.n_unique_pixels), n_samples=int(res.n_samples),
                         n_fma=int(res.n_fma), pair_ticks=int(res.pair_ticks), pixel_ticks=int(res.pixel_ticks), n_units_here=n_mine, n_segments_here=int(count.sum()), n_packets_here=n_local)
        rowb = self.assn_dtype.itemsize
        pkb = _p.PACKET_DTYPE.itemsize
        # (5) packet counts of every unit, everywhere (one small collective)
        counts = torch.zeros(nU, dtype=torch.int64, device="cuda")
        if n_mine:
            counts[torch.from_numpy(mine).cuda()] = torch.from_numpy(upk[:n_mine]).cuda()
        if self.world > 1:
            dist.all_reduce(counts, group=self.group)
        counts_h = counts.cpu().numpy()
        out.unit_packets = counts_h
        if host_output and self.world > 1:
            tabs = self._shared_tables()
            if tabs is not None:
                try:
                    return self._finish_shared(out, tabs, res, counts_h, plan, events, event_times, nB, d_sel, S, return_tracks)
                except SharedHostUnavailable as e:          # collective: every rank lands here together
                    import warnings
                    warnings.warn("shared host table not available (%s): gathering through rank 0 instead" % e)
                    for t in tabs:
                        t.close()
                    self._shared = False
        # (6) the ranks' packet / truth-row blocks -> rank 0 (NCCL send / recv of the exact sizes)
        per_rank = [int(counts_h[plan[r]].sum()) if len(plan[r]) else 0 for r in range(self.world)]
        src_pk = {self.rank: int(res.packets or 0)}
        src_rw = {self.rank: int(res.assn_rows or 0)}
        if self.world > 1:
            ops, keep = [], []
            if self.rank == 0:
                for r in range(1, self.world):
                    if per_rank[r] == 0:
                        continue
                    bp, br = self._buf("recv_pk_%d" % r, per_rank[r] * pkb), self._buf("recv_rw_%d" % r, per_rank[r] * rowb)
                    ops += [dist.P2POp(dist.irecv, bp[:per_rank[r] * pkb], r, group=self.group),
                            dist.P2POp(dist.irecv, br[:per_rank[r] * rowb], r, group=self.group)]
                    src_pk[r], src_rw[r] = bp.data_ptr(), br.data_ptr()
            elif n_local:
                tp = _view_u8(res.packets, n_local * pkb)
                tr = _view_u8(res.assn_rows, n_local * rowb)
                keep += [tp, tr]
                ops += [dist.P2POp(dist.isend, tp, 0, group=self.group), dist.P2POp(dist.isend, tr, 0, group=self.group)]
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
        if self.rank != 0:
            return out
        # (7) file order on the device: per event its between-batch packets, then its units
        evp = self._event_level_packets(events, event_times)
        n_evp = sum(len(x) for x in evp)
        n_total = int(counts_h.sum()) + n_evp
        out.n_packets = n_total
        d_pk = self._buf("final_pk", max(n_total, 1) * pkb)
        d_rw = self._buf("final_rw", max(n_total, 1) * rowb)
        ev_blob = np.concatenate(evp) if n_evp else np.zeros(0, dtype=_p.PACKET_DTYPE)
        d_evp = torch.from_numpy(ev_blob.view(np.uint8).reshape(-1)).cuda() if n_evp else None
        ev_rows = _fee._no_truth_rows(1)
        blocks, pos = file_order_blocks(counts_h, plan, [len(x) for x in evp], nB)
        srcs_p, srcs_r, dsts, nbs, ev_positions = [], [], [], [], []
        for src, soff, dpos, n in blocks:
            if src < 0:
                srcs_p.append(d_evp.data_ptr() + soff * pkb); srcs_r.append(0)
                ev_positions.append((dpos, n))
            else:
                srcs_p.append(src_pk[src] + soff * pkb); srcs_r.append(src_rw[src] + soff * rowb)
            dsts.append(dpos); nbs.append(n)
        assert pos == n_total
        self._copy_blocks(srcs_p, dsts, nbs, pkb, d_pk)
        keep_r = [(s, dpos, n) for s, dpos, n in zip(srcs_r, dsts, nbs) if s]
        self._copy_blocks([k[0] for k in keep_r], [k[1] for k in keep_r], [k[2] for k in keep_r], rowb, d_rw)
        if ev_positions:                                   # truth rows of the between-batch packets: all -1 / 0 (fee.py:413-420)
            blank = torch.from_numpy(np.ascontiguousarray(ev_rows).view(np.uint8).reshape(-1)).cuda()
            rows2d = d_rw[:n_total * rowb].view(n_total, rowb)
            for dpos, ne in ev_positions:
                rows2d[dpos:dpos + ne] = blank
        if not host_output:
            out.packets, out.packets_mc_ds = d_pk[:n_total * pkb], d_rw[:n_total * rowb]
            if return_tracks:
                out.tracks = d_sel
            return out
        # (8) one D2H copy per table into pinned memory
        h_pk, h_rw = self._buf("host_pk", max(n_total, 1) * pkb, pinned=True), self._buf("host_rw", max(n_total, 1) * rowb, pinned=True)
        h_pk[:n_total * pkb].copy_(d_pk[:n_total * pkb], non_blocking=True)
        h_rw[:n_total * rowb].copy_(d_rw[:n_total * rowb], non_blocking=True)
        if return_tracks:
            h_tr = self._buf("host_tracks", max(S, 1) * self.dtype.itemsize, pinned=True)
            h_tr[:S * self.dtype.itemsize].copy_(d_sel.buf[:S * self.dtype.itemsize], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        out.packets = h_pk[:n_total * pkb].numpy().view(_p.PACKET_DTYPE)
        out.packets_mc_ds = h_rw[:n_total * rowb].numpy().view(self.assn_dtype)
        if return_tracks:
            out.tracks = h_tr[:S * self.dtype.itemsize].numpy().view(self.dtype)
        return out

    def _finish_shared(self, out, tabs, res, counts_h, plan, events, event_times, nB, d_sel, S, return_tracks):
        """Steps (6)-(8) for host output with several ranks on one node: every rank copies the blocks of its own units from its
        HBM to their file-order positions in two host tables shared by the ranks; rank 0 adds the between-batch packets."""
        lib = _l.lib()
        st = _l.stream()
        rowb, pkb = self.assn_dtype.itemsize, _p.PACKET_DTYPE.itemsize
        evp = self._event_level_packets(events, event_times)
        blocks, n_total = file_order_blocks(counts_h, plan, [len(x) for x in evp], nB)
        out.n_packets = n_total
        h_pk = tabs[0].ensure(max(n_total, 1) * pkb)
        h_rw = tabs[1].ensure(max(n_total, 1) * rowb)
        mine = [(soff, dpos, n) for src, soff, dpos, n in blocks if src == self.rank]
        if mine and res.packets:
            soff = np.asarray([m[0] for m in mine], dtype=np.int64)
            dpos = np.asarray([m[1] for m in mine], dtype=np.int64)
            cnt = np.asarray([m[2] for m in mine], dtype=np.int64)
            for base, item, table in ((int(res.packets), pkb, h_pk), (int(res.assn_rows), rowb, h_rw)):
                src = (C.c_void_p * len(mine))(*[base + int(o) * item for o in soff])
                d_off, nb = np.ascontiguousarray(dpos * item), np.ascontiguousarray(cnt * item)
                _l.check(lib.lsb_d2h_blocks(C.c_int64(len(mine)), src, d_off.ctypes.data_as(C.c_void_p), nb.ctypes.data_as(C.c_void_p),
                                            C.c_void_p(table.ctypes.data), st), "d2h_blocks")
        if self.rank == 0:
            ev_blob = np.concatenate(evp) if len(evp) else np.zeros(0, dtype=_p.PACKET_DTYPE)
            ev_u8 = ev_blob.view(np.uint8).reshape(-1)
            blank = np.ascontiguousarray(_fee._no_truth_rows(1)).view(np.uint8).reshape(-1)
            for src, soff, dpos, n in blocks:
                if src < 0:                                # between-batch packets; their truth rows are all -1 / 0 (fee.py:413-420)
                    h_pk[dpos * pkb:(dpos + n) * pkb] = ev_u8[soff * pkb:(soff + n) * pkb]
                    h_rw[dpos * rowb:(dpos + n) * rowb].reshape(n, rowb)[:] = blank
            if return_tracks:
                h_tr = self._buf("host_tracks", max(S, 1) * self.dtype.itemsize, pinned=True)
                h_tr[:S * self.dtype.itemsize].copy_(d_sel.buf[:S * self.dtype.itemsize], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        dist.barrier(group=self.group)                     # every rank's blocks have landed
        if self.rank != 0:
            return out
        out.packets = np.asarray(h_pk[:n_total * pkb]).view(_p.PACKET_DTYPE)
        out.packets_mc_ds = np.asarray(h_rw[:n_total * rowb]).view(self.assn_dtype)
        if return_tracks:
            out.tracks = h_tr[:S * self.dtype.itemsize].numpy().view(self.dtype)
        return out

    def _copy_blocks(self, srcs, dst_pos, counts, itemsize, dst):
        n = len(srcs)
        if n == 0:
            return
        lib = _l.lib()
        for i0 in range(0, n, 60000):
            i1 = min(n, i0 + 60000)
            a = (C.c_void_p * (i1 - i0))(*srcs[i0:i1])
            o = np.asarray(dst_pos[i0:i1], dtype=np.int64) * itemsize
            b = np.asarray(counts[i0:i1], dtype=np.int64) * itemsize
            _l.check(lib.lsb_copy_blocks(C.c_int64(i1 - i0), a, o.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                         C.c_void_p(dst.data_ptr()), _l.stream()), "copy_blocks")


def _view_u8(ptr, nbytes):
    class _H:
        pass
    h = _H()
    h.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3, "strides": None}
    return torch.as_tensor(h, device="cuda")
