"""Segment selection + batching (active_volume.select_active_volume, util.batching.TPCBatcher): oracle against the
reference's golden vectors on CPU, CUDA against both on the GPU.  Integer work: bit-exact."""
import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import batching_util as bu  # noqa: E402
from oracle import batching_oracle as orc  # noqa: E402

CASES = sorted(bu.CASES)


def golden(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", "batching_%s.npz" % name))
    all_seg, seg, borders, sizes = bu.case_inputs(name)
    digest = np.frombuffer(hashlib.sha256(all_seg.tobytes() + borders.tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, g["checksum"]), "seeded inputs changed: regenerate with tools/gen_golden_batching.py"
    return g, all_seg, seg, borders, sizes


def modules_in(g):
    return sorted(int(k.split("_")[-1]) for k in g.files if k.startswith("sel_module_"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    g, all_seg, seg, borders, sizes = golden(name)
    sep = bu.event_field(bu.CASES[name][2])
    assert np.array_equal(orc.select_active_volume(all_seg, borders), g["sel_all"])
    assert 0 < len(g["sel_all"]) < len(all_seg)
    for m in modules_in(g):
        assert np.array_equal(orc.select_active_volume(all_seg, borders, m), g["sel_module_%d" % m])
    for bs in sizes:
        batches = orc.tpc_batches(all_seg, seg, sep, bs, borders)
        assert np.array_equal(np.array([e for e, _ in batches]), g["events_bs%d" % bs])
        assert np.array_equal(orc.unit_of_segment(batches, len(seg)), g["unit_bs%d" % bs])


def test_oracle_unsorted_borders():
    g, all_seg, seg, borders, sizes = golden("2x2")
    assert np.array_equal(orc.select_active_volume(all_seg, borders[:, :, ::-1]), g["sel_all"])


# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cuda_mods():
    from larndsim_b200 import active_volume as av
    from larndsim_b200.util import batching as bt
    return av, bt


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_select_active_volume_golden(name, cuda_mods):
    av, _ = cuda_mods
    g, all_seg, seg, borders, sizes = golden(name)
    got = av.select_active_volume(all_seg, borders)
    assert isinstance(got, np.ndarray) and got.dtype == np.int64
    assert np.array_equal(got, g["sel_all"])
    assert np.array_equal(av.select_active_volume(all_seg, borders[:, :, ::-1].astype(np.float32).astype(np.float64)),
                          orc.select_active_volume(all_seg, borders.astype(np.float32).astype(np.float64)))
    for m in modules_in(g):
        assert np.array_equal(av.select_active_volume(all_seg, borders, m), g["sel_module_%d" % m])
    with pytest.raises(IndexError):
        av.select_active_volume(all_seg, borders, borders.shape[0] // 2 + 1)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_tpc_batcher_golden(name, cuda_mods):
    _, bt = cuda_mods
    g, all_seg, seg, borders, sizes = golden(name)
    sep = bu.event_field(bu.CASES[name][2])
    for bs in sizes:
        it = bt.TPCBatcher(all_seg, seg, sep, tpc_batch_size=bs, tpc_borders=borders)
        assert len(it) == len(g["events_bs%d" % bs])
        events, unit = [], np.full(len(seg), -1, dtype=np.int32)
        for u, (ev, mask) in enumerate(it):
            events.append(ev)
            assert mask.dtype == bool and mask.shape == seg.shape
            assert (unit[mask] == -1).all()
            unit[mask] = u
        assert np.array_equal(np.array(events), g["events_bs%d" % bs])
        assert np.array_equal(unit, g["unit_bs%d" % bs])
        with pytest.raises(StopIteration):
            next(it)
        # the same batches as index lists, ascending = tracks[mask] order
        it2 = bt.TPCBatcher(all_seg, seg, sep, tpc_batch_size=bs, tpc_borders=borders)
        sizes_u = it2.unit_sizes
        for u, (ev, idx) in enumerate(it2.units()):
            assert ev == g["events_bs%d" % bs][u]
            assert np.array_equal(idx, np.nonzero(g["unit_bs%d" % bs] == u)[0])
            assert sizes_u[u] == len(idx)


@pytest.mark.gpu
def test_batching_device_records_and_edges(cuda_mods):
    import torch
    av, bt = cuda_mods
    g, all_seg, seg, borders, sizes = golden("2x2")

    class DevRecords:                                   # records resident on the device (torch bytes + descr)
        def __init__(self, a):
            self.t = torch.from_numpy(a.view(np.uint8).copy()).cuda()
            self.dtype = a.dtype
            self.__cuda_array_interface__ = {"shape": a.shape, "typestr": "|V%d" % a.dtype.itemsize, "descr": a.dtype.descr,
                                             "data": (self.t.data_ptr(), False), "version": 3}
    got = av.select_active_volume(DevRecords(all_seg), borders)
    assert isinstance(got, torch.Tensor) and got.is_cuda
    assert np.array_equal(got.cpu().numpy(), g["sel_all"])
    # nothing to select / nothing selected
    empty = all_seg[:0]
    assert av.select_active_volume(empty, borders).shape == (0,)
    assert av.select_active_volume(all_seg, borders[:0]).shape == (0,)
    far = all_seg.copy()
    for a in "xyz":
        far[a + "_start"] += 1e5
        far[a + "_end"] += 1e5
    assert av.select_active_volume(far, borders).shape == (0,)
    it = bt.TPCBatcher(all_seg, far[:100], "event_id", tpc_batch_size=2, tpc_borders=borders)
    assert all(not m.any() for _, m in it)
    it = bt.TPCBatcher(all_seg, empty, "event_id", tpc_batch_size=2, tpc_borders=borders)
    assert [m.shape for _, m in it] == [(0,)] * (5 * 4)


@pytest.mark.gpu
def test_batching_full_size_properties(cuda_mods):
    """1e6 segments x 35 modules x 5 events: every segment the oracle's vectorised test keeps is handed out exactly once,
    in the unit of (its event, its first TPC group); indices ascend inside a unit."""
    av, bt = cuda_mods
    n = 1_000_000
    seg = bu.segments("ndlar", n, "f4", 99)
    borders = bu.borders_of("ndlar")
    b = np.sort(borders, axis=-1)
    first = np.full(n, -1, dtype=np.int64)
    for t in range(b.shape[0] - 1, -1, -1):
        first[orc.in_box(seg, "end", b[t]) | orc.in_box(seg, "start", b[t])] = t
    assert np.array_equal(av.select_active_volume(seg, borders), np.nonzero(first >= 0)[0])
    bs = 3
    nB = -(-b.shape[0] // bs)
    events = np.unique(seg["event_id"])
    want = np.where(first >= 0, np.searchsorted(events, seg["event_id"]) * nB + first // bs, -1)
    it = bt.TPCBatcher(seg, seg, "event_id", tpc_batch_size=bs, tpc_borders=borders)
    assert it.unit_sizes.sum() == (first >= 0).sum()
    seen = 0
    for u, (ev, idx) in enumerate(it.units()):
        assert ev == events[u // nB]
        assert (want[idx] == u).all() and (np.diff(idx) > 0).all()
        seen += len(idx)
    assert seen == (want >= 0).sum()


@pytest.mark.gpu
def test_example_batch_loop_end_to_end():
    """The batch loop (larndsim_b200.spill.SpillRunner, what examples/run_batches.py drives): selection -> batching -> chain ->
    packets.  Every hit above the pedestal of every batch becomes one data packet carrying its event; the loop is
    deterministic; the truth rows point at segments of the batch's event."""
    from larndsim_b200 import consts, synth, spill
    mod = consts.load_snapshot("2x2")
    tracks = synth.beam_spill_segments(3000, mod.detector, seed=5, n_events=3)
    tracks["segment_id"] = np.arange(len(tracks))
    runner = spill.SpillRunner(tracks.dtype, synth.response_lut(mod.detector), depth=2, tpc_batch_size=2)
    a = runner.simulate(tracks, rand_seed=1)
    pk, rows = a.packets.copy(), a.packets_mc_ds.copy()
    b = runner.simulate(tracks, rand_seed=1)
    assert pk.tobytes() == b.packets.tobytes() and rows.tobytes() == b.packets_mc_ds.tobytes()
    assert len(a.unit_sizes) == 3 * 4 and int(a.unit_sizes.sum()) == a.n_segments
    data = pk["packet_type"] == 0
    assert data.sum() > 500 and (pk["packet_type"] == 4).sum() >= 3
    ev_of_seg = dict(zip(tracks["segment_id"].tolist(), tracks["event_id"].tolist()))
    ev = rows["event_ids"][data, 0]
    first = rows["segment_ids"][data, 0]
    assert (first >= 0).all()
    assert all(ev_of_seg[int(s)] == int(e) for s, e in zip(first[::17], ev[::17]))
    frac = rows["fraction"][data]
    assert np.isfinite(frac).all() and (np.abs(frac).sum(axis=1) > 0).all()
    runner.close()
