"""The reference's own unit tests (tests/testQuenching.py, testDrifting.py, testTracksCurrent.py, testTrackCharge.py)
restated against the CUDA drop-ins: same inputs (all-float64 records with the reference's field order), same calls
(``kernel[BPG, TPB](...)`` on host arrays), same assertions and tolerances.  testCudaDict.py is restated in
tests/test_cuda_dict.py.  ``detsim.rho`` is a device function of the kernel here, so testTrackCharge's normalisation of the
charge cloud is checked where the reference's testTracksCurrent checks it: through ``tracks_current`` with a response table
whose time integral is 1 under the pad (the shipped ``response_44.npy`` is not in the mount)."""
from math import ceil

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NAMES22 = ("eventID, dEdx, x_start, dE, t_start, z_end, trackID, x_end, y_end, n_electrons, n_photons, t, dx, pdgId, y, x, "
           "long_diff, z, z_start, y_start, tran_diff, t_end, pixel_plane").split(", ")
NAMES26 = NAMES22 + ["t0", "t0_start", "t0_end"]
INTS = ("eventID", "trackID", "pixel_plane")


def records(n, names):
    return np.zeros(n, dtype=[(k, "i8" if k in INTS else "f8") for k in names])


@pytest.fixture()
def consts(cuda):
    from larndsim_b200 import consts as lc
    return lc.load_snapshot("module0")


class TestQuenching:                                   # tests/testQuenching.py
    def _tracks(self):
        rng = np.random.default_rng(0)
        t = records(100, NAMES22)
        t["dE"], t["dEdx"] = rng.uniform(0.1, 100, 100), rng.uniform(1, 100, 100)
        return t

    def test_birksModel(self, consts):
        from larndsim_b200 import quenching
        ph, d = consts.physics, consts.detector
        t = self._tracks()
        de, dedx = t["dE"].copy(), t["dEdx"].copy()
        TPB = 128
        quenching.quench[ceil(len(t) / TPB), TPB](t, ph.BIRKS)
        recomb = ph.BIRKS_Ab / (1 + ph.BIRKS_kb * dedx / (d.E_FIELD * d.LAR_DENSITY))
        assert t["n_electrons"] == pytest.approx(recomb * de / ph.W_ION)

    def test_boxModel(self, consts):
        from larndsim_b200 import quenching
        ph, d = consts.physics, consts.detector
        t = self._tracks()
        de, dedx = t["dE"].copy(), t["dEdx"].copy()
        TPB = 128
        quenching.quench[ceil(len(t) / TPB), TPB](t, ph.BOX)
        csi = ph.BOX_BETA * dedx / (d.E_FIELD * d.LAR_DENSITY)
        recomb = np.maximum(0, np.log(ph.BOX_ALPHA + csi) / csi)
        assert t["n_electrons"] == pytest.approx(recomb * de / ph.W_ION)

    def test_extreme_values(self, consts):
        """dEdx = 0 and dEdx = 1e10 must not produce NaN / negative charge (the reference's track_zero / track_inf)"""
        from larndsim_b200 import quenching
        ph = consts.physics
        for mode in (ph.BIRKS, ph.BOX):
            zero, inf = records(1, NAMES22), records(1, NAMES22)
            zero["dE"] = 1
            inf["dE"], inf["dEdx"] = 1e10, 1e10
            quenching.quench[1, 128](inf, mode)
            assert np.isfinite(inf["n_electrons"]).all() and (inf["n_electrons"] >= 0).all()
            if mode == ph.BIRKS:
                quenching.quench[1, 128](zero, mode)
                assert zero["n_electrons"][0] == pytest.approx(ph.BIRKS_Ab * 1 / ph.W_ION)


class TestDrifting:                                    # tests/testDrifting.py
    def test_lifetime(self, consts):
        from larndsim_b200 import drifting
        d = consts.detector
        rng = np.random.default_rng(1)
        b = np.asarray(d.TPC_BORDERS)
        t = records(1, NAMES26)
        t["z"] = rng.uniform(b[0][2][0], b[0][2][1], 1)
        t["x"] = rng.uniform(b[0][0][0], b[0][0][1], 1)
        t["y"] = rng.uniform(b[0][1][0], b[0][1][1], 1)
        t["n_electrons"] = rng.uniform(1e6, 1e7, 1)
        z_anode = b[0][2][0]
        lifetime = np.exp(-np.abs(t["z"] - z_anode) / d.V_DRIFT / d.ELECTRON_LIFETIME)
        electrons_anode = t["n_electrons"] * lifetime
        drifting.drift[1, 128](t)
        assert t["n_electrons"] == pytest.approx(electrons_anode)
        assert t["pixel_plane"][0] == 0


def current_model_tracks(consts, n=10, seed=2):
    """testTracksCurrent.py's segments: random inside TPC 0, dE/dx = 2 MeV/cm.  Two deliberate differences: the segments
    stay 4 cm away from the pixel-plane edge (so the whole charge cloud lands on pixels) and extend at most 3 cm along the
    drift (the reference draws z_start over the whole TPC, which only makes the sampling error of the 40-point grid larger)."""
    d = consts.detector
    rng = np.random.default_rng(seed)
    b = np.asarray(d.TPC_BORDERS)
    t = records(n, NAMES26)
    t["z_end"] = rng.uniform(b[0][2][0], b[0][2][0] + 2, n)
    t["z_start"] = t["z_end"] + rng.uniform(0, 3, n)
    t["z"] = (t["z_end"] + t["z_start"]) / 2.
    for a, k in (("x", 0), ("y", 1)):
        t[a + "_start"] = rng.uniform(b[0][k][0] + 4, b[0][k][0] + 6, n)
        t[a + "_end"] = t[a + "_start"] + rng.uniform(-1, 1, n)
        t[a] = (t[a + "_end"] + t[a + "_start"]) / 2.
    t["dx"] = np.sqrt((t["x_end"] - t["x_start"]) ** 2 + (t["y_end"] - t["y_start"]) ** 2 + (t["z_end"] - t["z_start"]) ** 2)
    t["dEdx"] = 2
    t["dE"] = t["dEdx"] * t["dx"]
    return t


def unit_response(d):
    """response table with unit time integral under the pad (|dx|, |dy| < pitch / 2), nothing elsewhere"""
    half = int(round(d.PIXEL_PITCH / 2 / d.RESPONSE_BIN_SIZE))
    response = np.zeros((45, 45, 1950), dtype=np.float32)
    response[:half, :half, 1950 - 60] = 1.0 / d.TIME_SAMPLING
    return response


class TestTrackCurrent:                                # tests/testTracksCurrent.py + testTrackCharge.py
    def test_current_model(self, consts):
        from larndsim_b200 import quenching, drifting, pixels_from_track, detsim
        d, ph = consts.detector, consts.physics
        t = current_model_tracks(consts)
        n = len(t)
        TPB = 128
        BPG = ceil(n / TPB)
        quenching.quench[BPG, TPB](t, ph.BOX)
        drifting.drift[BPG, TPB](t)
        t["tran_diff"], t["long_diff"] = 1e-1, 1e-1    # as the reference's test: fixed 1 mm clouds
        MAX_PIXELS, MAX_ACTIVE_PIXELS = 110, 50
        active_pixels = np.full((n, MAX_ACTIVE_PIXELS), -1, dtype=np.int32)
        neighboring_pixels = np.full((n, MAX_PIXELS), -1, dtype=np.int32)
        neighboring_radius = np.full((n, MAX_PIXELS), -1, dtype=np.int32)
        n_pixels_list = np.zeros(shape=(n,))
        pixels_from_track.get_pixels[BPG, TPB](t, active_pixels, neighboring_pixels, neighboring_radius, n_pixels_list, 2)
        assert (neighboring_pixels >= 0).sum() > 25 * n / 2
        # waveform length as the simulation sizes it (simulate_pixels.py:996-1002); the reference's test uses
        # len(TIME_TICKS), which with module0's 190 us padding cuts the waveforms 15 us after the first arrival
        track_starts = np.zeros(n)
        max_length = np.array([0])
        detsim.time_intervals[BPG, TPB](track_starts, max_length, t)
        T = int(max_length[0])
        signals = np.zeros((n, MAX_PIXELS, T), dtype=np.float32)
        TPB3 = (1, 1, 64)
        BPG3 = (ceil(n / 1), ceil(MAX_PIXELS / 1), ceil(T / 64))
        detsim.tracks_current[BPG3, TPB3](signals, neighboring_pixels, t, unit_response(d))
        # signals are in electrons / us here (consts.units.e = 1); the shipped response carries the factor E_CHARGE the
        # reference's assertion divides out
        total = np.sum(signals, dtype=np.float64) * d.TIME_SAMPLING
        assert total == pytest.approx(np.sum(t["n_electrons"]), rel=0.05)
        per_segment = signals.astype(np.float64).sum(axis=(1, 2)) * d.TIME_SAMPLING
        assert np.all(np.abs(per_segment / t["n_electrons"] - 1) < 0.15)
