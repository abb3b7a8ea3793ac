"""CudaDict drop-in (larndsim/util/cuda_dict.py): the reference's own test (tests/testCudaDict.py) restated for the
torch-backed version, plus a check against a plain Python dict on a pixel-threshold sized table."""
import os

import numpy as np
import pytest

KEYS = np.array([0, 1, 2, 3, 4, 10, 20, 30, 40, 100, 200, 300, 400])
VALUES = KEYS.astype(float) + 1.
DEFAULT = np.array([999.], dtype=float)
AVAIL = np.array([0, 0, 400, 2, 10, 30, 100])
UNAVAIL = np.array([5, 6, 7, 8, 9, 11, 21, 31, 41, 5000])


def _init():
    from larndsim_b200.util import CudaDict
    cd = CudaDict(default=DEFAULT, tpb=256, bpg=1)
    assert len(cd) == 0
    assert not cd.contains(KEYS).any()
    cd[KEYS] = VALUES
    return cd


@pytest.mark.gpu
def test_init(cuda):
    cd = _init()
    assert cd.contains(KEYS).all()
    assert (cd[KEYS].cpu().numpy() == VALUES).all()
    assert cd.contains(AVAIL).all()
    assert (cd[AVAIL].cpu().numpy() == AVAIL.astype(float) + 1.).all()
    assert (cd[UNAVAIL].cpu().numpy() == DEFAULT[0]).all()
    assert not cd.contains(UNAVAIL).any()
    with pytest.raises(NotImplementedError):
        cd[KEYS] = VALUES


@pytest.mark.gpu
def test_read_write(cuda, tmp_path):
    from larndsim_b200.util import CudaDict
    cd = _init()
    filename = os.path.join(tmp_path, "test_cd.npz")
    CudaDict.save(filename, cd)
    new_cd = CudaDict.load(filename, tpb=cd.tpb)
    assert len(new_cd) == len(cd)
    assert cd.contains(new_cd.keys()).all() and new_cd.contains(cd.keys()).all()
    assert (cd[new_cd.keys()] == new_cd.values()).all() and (new_cd[cd.keys()] == cd.values()).all()


@pytest.mark.gpu
def test_pixel_threshold_table_vs_python_dict(cuda):
    import torch
    from larndsim_b200.util import CudaDict
    rng = np.random.default_rng(5)
    keys = rng.choice(140 * 280 * 2, 30000, replace=False).astype(np.int32)
    vals = rng.uniform(3000, 9000, len(keys)).astype(np.float32)
    cd = CudaDict(default=np.array([7000.], dtype=np.float32))
    cd[keys] = vals
    ref = dict(zip(keys.tolist(), vals.tolist()))
    q = rng.integers(-5, 140 * 280 * 2 + 50, 100000).astype(np.int32)
    expect = np.array([ref.get(int(k), np.float32(7000.)) for k in q], dtype=np.float32)
    got = cd[torch.from_numpy(q).cuda()].cpu().numpy()          # device keys, like unique_pix.ravel()
    assert got.dtype == np.float32 and np.array_equal(got, expect)
    assert np.array_equal(cd.contains(q).cpu().numpy(), np.array([int(k) in ref for k in q]))
    assert cd[np.zeros(0, dtype=np.int32)].numel() == 0
