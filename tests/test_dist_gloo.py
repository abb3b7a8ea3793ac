"""Host-side multi-rank logic on CPU: world size 2, gloo backend."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from larndsim_b200 import dist as ldist


def test_assign_units_lpt():
    sizes = [900, 10, 500, 400, 30, 20]
    a = ldist.assign_units(sizes, 2)
    assert sorted(a[0] + a[1]) == list(range(6))
    loads = [sum(sizes[i] for i in r) for r in a]
    assert abs(loads[0] - loads[1]) <= 100
    assert ldist.assign_units(sizes, 2) == a                      # deterministic
    assert ldist.assign_units([5, 5, 5], 4)[3] == []


def test_spill_partition_plan():
    """larndsim_b200.spill.assign_units: the (event, TPC pair) batches of a spill over the ranks, longest first; every rank
    computes the same plan, empty batches belong to nobody, loads are balanced to a few percent for an ND-LAr spill."""
    import numpy as np
    from larndsim_b200 import spill
    rng = np.random.default_rng(0)
    sizes = rng.integers(2400, 13600, 140)
    sizes[[3, 77]] = 0
    for world in (1, 2, 4, 8):
        plan = spill.assign_units(sizes, world)
        assert plan == spill.assign_units(sizes.copy(), world)
        flat = sorted(u for lst in plan for u in lst)
        assert flat == [u for u in range(140) if sizes[u] > 0]
        assert all(lst == sorted(lst) for lst in plan)
        loads = np.array([sizes[lst].sum() for lst in plan], dtype=np.float64)
        assert loads.max() / loads.mean() < 1.03
    assert spill.assign_units([5, 0, 5], 4) == [[0], [2], [], []]


def test_hit_packets_compaction():
    uniq = torch.tensor([7, 9, 11], dtype=torch.int32)
    digit = torch.tensor([[74., 90.], [74., 74.], [101., 74.]], dtype=torch.float64)
    ticks = torch.tensor([[0., 12.5], [0., 0.], [3.25, 0.]], dtype=torch.float64)
    rec = ldist.hit_packets(uniq, digit, ticks, 74.0)
    assert rec.tolist() == [[7.0, 90.0, 12.5], [11.0, 101.0, 3.25]]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 3 if rank == 0 else 5                                     # ragged
    rec = torch.arange(n * 3, dtype=torch.float64).reshape(n, 3) + 100 * rank
    out = ldist.gather_packets(rec, dst=0)
    if rank == 0:
        q.put([o.tolist() for o in out])
    else:
        assert out is None
    g = ldist.HitTableGather(cap=6, pedestal_adc=74.0, device="cpu")
    U = 2 + rank
    uniq = torch.arange(U, dtype=torch.int32) + 10 * rank
    digit = torch.full((U, 2), 74.0, dtype=torch.float64)
    digit[:, 0] = 80.0 + rank                                       # one hit per pixel, second slot at the pedestal
    ticks = torch.full((U, 2), 1.5 * (rank + 1), dtype=torch.float64)
    tabs = g.gather(uniq, digit, ticks)
    g.flush()                                                     # the collective is asynchronous
    if rank == 0:
        u1, d1, t1 = g.unpack(tabs[1])
        q.put((u1.tolist(), d1.tolist(), t1.tolist(), [int(t[0, 0].item()) for t in tabs]))
    empty = ldist.gather_packets(torch.zeros((0, 3), dtype=torch.float64), dst=0)     # no hits anywhere
    if rank == 0:
        q.put([tuple(e.shape) for e in empty])
    dist.barrier()
    dist.destroy_process_group()


def test_gather_packets_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    table = q.get(timeout=120)
    shapes = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert len(got) == 2 and len(got[0]) == 3 and len(got[1]) == 5
    assert got[1][0] == [100.0, 101.0, 102.0] and got[0][2] == [6.0, 7.0, 8.0]
    assert shapes == [(0, 3), (0, 3)]
    assert table[0] == [10, 11, 12] and table[1] == [81.0] * 3 and table[2] == [3.0] * 3 and table[3] == [2, 3]


# ------------------------------------------------------------------------------------------------
# units of a run spread over ranks, outputs reassembled in the reference's file order
# ------------------------------------------------------------------------------------------------
import numpy as np  # noqa: E402

REC = np.dtype([("unit", "i4"), ("k", "u1"), ("t", "f8")], align=True)
UNIT_SIZES = [40, 0, 7, 300, 12, 0, 90, 33, 5]


def _unit_records(u):
    """deterministic stand-in for 'simulate unit u': as many records as a third of its segments"""
    n = UNIT_SIZES[u] // 3
    r = np.zeros(n, dtype=REC)
    r["unit"], r["k"], r["t"] = u, np.arange(n) % 251, np.arange(n) * 0.5 + u
    return r


def _unit_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = ldist.assign_units(UNIT_SIZES, world)[rank]
    out = ldist.gather_unit_records(list(reversed(mine)), [_unit_records(u) for u in reversed(mine)], REC, dst=0)
    if rank == 0:
        ids, recs = out
        got = np.concatenate(recs)
        q.put((ids, [got[f].tolist() for f in REC.names]))
    else:
        assert out is None
    none = ldist.gather_unit_records([], [], REC, dst=0)               # a run without any unit
    if rank == 0:
        q.put(none)
    dist.barrier()
    dist.destroy_process_group()


def test_unit_outputs_come_back_in_file_order_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + os.getpid() % 300
    procs = [ctx.Process(target=_unit_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ids, blob = q.get(timeout=120)
    none = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ids == list(range(len(UNIT_SIZES)))
    want = np.concatenate([_unit_records(u) for u in range(len(UNIT_SIZES))])                        # = the sequential loop's output
    assert blob == [want[f].tolist() for f in REC.names]
    assert none == ([], [])
    # single process: same call, no process group
    ids1, recs1 = ldist.gather_unit_records([4, 2, 7], [_unit_records(u) for u in (4, 2, 7)], REC)
    assert ids1 == [2, 4, 7] and np.array_equal(np.concatenate(recs1), np.concatenate([_unit_records(u) for u in (2, 4, 7)]))


def _spill_worker(rank, world, port, q):
    """the exchange + file-order assembly of spill.SpillRunner.simulate, steps (5)-(7), on host tensors over gloo"""
    import numpy as np
    from larndsim_b200 import spill
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    nB, n_events = 5, 3
    sizes = rng.integers(0, 40, nB * n_events)
    sizes[[2, 9]] = 0
    n_pk = np.where(sizes > 0, sizes * 2 + 1, 0)                          # packets per unit (a function of the unit only)
    n_evp = [2, 0, 3]
    plan = spill.assign_units(sizes, world)
    mine = plan[rank]

    def unit_bytes(u):                                                    # recognisable content: (unit, running index)
        return np.stack([np.full(n_pk[u], u, dtype=np.int64), np.arange(n_pk[u], dtype=np.int64)], axis=1)
    local = np.concatenate([unit_bytes(u) for u in mine]) if len(mine) else np.zeros((0, 2), dtype=np.int64)
    counts = torch.zeros(len(sizes), dtype=torch.int64)
    counts[torch.tensor(mine, dtype=torch.int64)] = torch.from_numpy(n_pk[mine])
    dist.all_reduce(counts)
    assert np.array_equal(counts.numpy(), n_pk)
    per_rank = [int(n_pk[plan[r]].sum()) for r in range(world)]
    bufs = {rank: torch.from_numpy(local)}
    if rank == 0:
        works = []
        for r in range(1, world):
            bufs[r] = torch.zeros((per_rank[r], 2), dtype=torch.int64)
            works.append(dist.irecv(bufs[r], r))
        for w in works:
            w.wait()
    else:
        dist.isend(torch.from_numpy(local), 0).wait()
    if rank == 0:
        blocks, total = spill.file_order_blocks(counts.numpy(), plan, n_evp, nB)
        ev_blob = np.stack([np.full(sum(n_evp), -1, dtype=np.int64), np.arange(sum(n_evp), dtype=np.int64)], axis=1)
        out = np.zeros((total, 2), dtype=np.int64)
        for src, soff, dpos, n in blocks:
            out[dpos:dpos + n] = ev_blob[soff:soff + n] if src < 0 else bufs[src].numpy()[soff:soff + n]
        # the sequential single-process order
        want, e_off = [], 0
        for e in range(n_events):
            want.append(ev_blob[e_off:e_off + n_evp[e]]); e_off += n_evp[e]
            for b in range(nB):
                want.append(unit_bytes(e * nB + b))
        ok_gather = bool(np.array_equal(out, np.concatenate(want))) and total == int(n_pk.sum()) + sum(n_evp)
    # the same assembly without the exchange: one host table shared by the ranks, every rank writes its own units' blocks
    # at their file-order positions (SpillRunner._finish_shared; the device -> host copies are plain stores here)
    blocks, total = spill.file_order_blocks(counts.numpy(), plan, n_evp, nB)
    table = spill.SharedHostTable(None, rank)
    ok_shared = True
    for round_ in range(2):                                               # second round: the table has to grow
        rep = 1 + 3 * round_
        mm = table.ensure(total * 16 * rep)
        view = np.asarray(mm[:total * 16 * rep]).view(np.int64).reshape(rep, total, 2)
        for src, soff, dpos, n in blocks:
            if src == rank:
                view[:, dpos:dpos + n] = local[soff:soff + n]
            elif src < 0 and rank == 0:
                view[:, dpos:dpos + n] = ev_blob[soff:soff + n]
        dist.barrier()
        if rank == 0:
            ok_shared = ok_shared and all(np.array_equal(view[k], np.concatenate(want)) for k in range(rep))
        dist.barrier()
    table.close()
    # a rank that cannot page-lock the mapping must not leave the others waiting: the failure is raised everywhere

    def reg(ptr, n):
        if rank == 1:
            raise RuntimeError("no locked memory")
    bad = spill.SharedHostTable(None, rank, register=reg, unregister=lambda ptr: None)
    try:
        bad.ensure(1 << 16)
        ok_fail = False
    except spill.SharedHostUnavailable as e:
        ok_fail = "rank 1" in str(e) and bad.mm is None
    if rank == 0:
        q.put(ok_gather and ok_shared and ok_fail)
    else:
        assert ok_fail
    dist.barrier()
    dist.destroy_process_group()


def test_spill_exchange_and_file_order_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_spill_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
