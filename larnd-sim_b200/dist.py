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
