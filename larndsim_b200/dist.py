"""Multi-GPU plumbing.  The chain has no exchange step: work units
(event x module / TPC pair) are independent, each rank runs whole batches, and the only communication is
the variable-length gather of hit packets to rank 0.  Uses torch.distributed (NCCL on GPUs, gloo in the
CPU tests); nothing here touches the arithmetic."""
import torch
import torch.distributed as dist


def assign_units(sizes, world_size):
    """Longest-processing-time-first assignment of work units (e.g. segments per (event, module)) to ranks.
    Returns a list of unit-index lists, one per rank; deterministic."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(sizes[i])
    for lst in out:
        lst.sort()
    return out


def hit_packets(unique_pix, adc_digit, adc_ticks, pedestal_adc):
    """Compact the per-pixel hit table into packet records [n_hits, 3] = (pixel id, ADC, timestamp);
    hits are entries above the pedestal code (fee.py:141 `if adc > digitize(0)`), in (pixel, hit) order."""
    idx = torch.nonzero(adc_digit > pedestal_adc)
    return torch.stack([unique_pix[idx[:, 0]].to(torch.float64), adc_digit[idx[:, 0], idx[:, 1]],
                        adc_ticks[idx[:, 0], idx[:, 1]]], dim=1).contiguous()


def gather_packets(rec, dst=0, group=None):
    """gatherv of packet records to `dst`: all_gather of the counts, then a padded gather.
    Returns the list of per-rank record tensors on `dst` (rank order = unit order), None elsewhere."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [rec]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = torch.tensor([rec.shape[0]], device=rec.device, dtype=torch.int64)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    pad = torch.zeros((mx,) + tuple(rec.shape[1:]), device=rec.device, dtype=rec.dtype)
    pad[: rec.shape[0]] = rec
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return [b[:c] for b, c in zip(bufs, counts)]


class HitTableGather:
    """Sync-free gather of the per-batch hits to `dst`.  Every rank compacts its hit table on the device into ONE
    fixed-size buffer [cap + 1, 3] of float64 -- row 0 holds the number of hits, rows 1.. hold (pixel id, ADC code,
    timestamp) of the hits in (pixel, hit) order -- with a prefix sum and a scatter (no `nonzero`, so no host
    synchronisation), and the buffers are gathered with one NCCL collective: no count exchange is needed.  The collective
    is issued asynchronously on NCCL's own stream with `depth` rotating send / receive buffers: the compute streams never
    wait for it, so ranks are not forced into lock-step every batch; a buffer is only waited for when it comes round again
    (or in :meth:`flush`)."""

    def __init__(self, cap, pedestal_adc, device, dst=0, group=None, depth=3):
        self.cap, self.ped, self.dst, self.group = int(cap), float(pedestal_adc), dst, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.depth = max(1, int(depth)) if self.world > 1 else 1
        self.sends = [torch.zeros((self.cap + 2, 3), dtype=torch.float64, device=device) for _ in range(self.depth)]   # last row: overflow dump
        self.recvs = [([torch.empty_like(self.sends[0]) for _ in range(self.world)] if self.rank == dst else None)
                      for _ in range(self.depth)]
        self.pending = [None] * self.depth
        self.i = 0

    def gather(self, unique_pix, adc_digit, adc_ticks):
        k = self.i % self.depth
        self.i += 1
        if self.pending[k] is not None:
            self.pending[k].wait()                      # the buffer is about to be overwritten
            self.pending[k] = None
        b = self.sends[k]
        hit = (adc_digit > self.ped).reshape(-1)
        pos = torch.cumsum(hit, 0)                       # 1-based row of every hit
        n = pos[-1:] if hit.numel() else torch.zeros(1, dtype=torch.int64, device=b.device)
        row = torch.where(hit & (pos <= self.cap), pos, torch.full_like(pos, self.cap + 1))     # non-hits / overflow -> dump row
        A = adc_digit.shape[1]
        pix = unique_pix.to(torch.float64).repeat_interleave(A)
        b.index_copy_(0, row, torch.stack([pix, adc_digit.reshape(-1), adc_ticks.reshape(-1)], dim=1))
        b[0, 0] = n[0].to(torch.float64) if hit.numel() else 0.0
        if self.world == 1:
            return [b]
        self.pending[k] = dist.gather(b, self.recvs[k], dst=self.dst, group=self.group, async_op=True)
        return self.recvs[k]

    def flush(self):
        """Make the current stream wait for every outstanding gather (call before reading the received tables)."""
        for k, w in enumerate(self.pending):
            if w is not None:
                w.wait()
                self.pending[k] = None

    def unpack(self, buf):
        """(pixel ids i4[n], ADC codes f8[n], timestamps f8[n]) of one received buffer"""
        n = int(buf[0, 0].item())
        if n > self.cap:
            raise OverflowError("HitTableGather: %d hits in a batch, capacity %d (size `cap` from U * MAX_ADC_VALUES to make this impossible)" % (n, self.cap))
        return buf[1:n + 1, 0].to(torch.int32), buf[1:n + 1, 1], buf[1:n + 1, 2]


def gather_unit_records(unit_ids, records, dtype, dst=0, group=None, device="cpu"):
    """Reassemble per-unit record arrays (e.g. the packets of every (event, TPC group) batch) on `dst` in UNIT order, i.e.
    the order in which the reference's sequential loop appends them to its output file ("ordered by (event,
    module)").  ``unit_ids``: the units this rank simulated (any order), ``records``: one NumPy structured array of
    ``dtype`` per unit.  Two collectives: the (unit id, record count) table, then one padded byte gather.
    Returns (unit ids ascending, list of arrays in that order) on `dst`, None elsewhere."""
    import numpy as np
    dtype = np.dtype(dtype)
    if len(unit_ids) != len(records):
        raise ValueError("one record array per unit")
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        order = sorted(range(len(unit_ids)), key=lambda i: int(unit_ids[i]))
        return [int(unit_ids[i]) for i in order], [np.asarray(records[i], dtype=dtype) for i in order]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    # (1) how many units / records every rank holds
    head = torch.tensor([len(unit_ids), sum(len(r) for r in records)], dtype=torch.int64, device=device)
    heads = [torch.zeros_like(head) for _ in range(world)]
    dist.all_gather(heads, head, group=group)
    max_units = max(max(int(h[0]) for h in heads), 1)
    max_bytes = max(max(int(h[1]) for h in heads) * dtype.itemsize, 1)
    # (2) per rank: [unit id, count] table and the concatenated record bytes, padded to the largest rank
    table = torch.full((max_units, 2), -1, dtype=torch.int64)
    for i, (u, r) in enumerate(zip(unit_ids, records)):
        table[i, 0], table[i, 1] = int(u), len(r)
    blob = np.zeros(max_bytes, dtype=np.uint8)
    if records and sum(len(r) for r in records):
        raw = np.concatenate([np.ascontiguousarray(r if r.dtype == dtype else r.astype(dtype)).view(np.uint8).reshape(-1)
                              for r in records])                   # byte-wise: padding of aligned records travels too
        blob[:raw.size] = raw
    table, payload = table.to(device), torch.from_numpy(blob).to(device)
    tabs = [torch.empty_like(table) for _ in range(world)] if rank == dst else None
    pays = [torch.empty_like(payload) for _ in range(world)] if rank == dst else None
    dist.gather(table, tabs, dst=dst, group=group)
    dist.gather(payload, pays, dst=dst, group=group)
    if rank != dst:
        return None
    found = {}
    for t, p in zip(tabs, pays):
        t, p = t.cpu().numpy(), p.cpu().numpy()
        off = 0
        for u, n in t:
            if u < 0:
                continue
            if int(u) in found:
                raise ValueError("unit %d was simulated by two ranks" % u)
            found[int(u)] = p[off:off + int(n) * dtype.itemsize].view(dtype).copy()
            off += int(n) * dtype.itemsize
    ids = sorted(found)
    return ids, [found[u] for u in ids]
