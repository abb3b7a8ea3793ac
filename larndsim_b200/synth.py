"""Synthetic inputs of the named shapes: the reference's example file and
response / light LUT blobs are not in the mount, so benchmarks and parity tests run on these.
Pure NumPy, seeded, host side; nothing here is on the product's compute path."""
import numpy as np

#: production segment record (cli/dumpTree.py:17-28 + cli/simulate_pixels.py:550-568), 152 bytes
segment_dtype = np.dtype([
    ("event_id", "u4"), ("vertex_id", "u8"), ("file_vertex_id", "u8"), ("segment_id", "u4"), ("z_end", "f4"),
    ("traj_id", "u4"), ("file_traj_id", "u4"), ("tran_diff", "f4"), ("z_start", "f4"), ("x_end", "f4"),
    ("y_end", "f4"), ("n_electrons", "u4"), ("pdg_id", "i4"), ("x_start", "f4"), ("y_start", "f4"),
    ("t_start", "f4"), ("t0_start", "f8"), ("t0_end", "f8"), ("t0", "f8"), ("dx", "f4"), ("long_diff", "f4"),
    ("pixel_plane", "i4"), ("t_end", "f4"), ("dEdx", "f4"), ("dE", "f4"), ("t", "f4"), ("y", "f4"), ("x", "f4"),
    ("z", "f4"), ("n_photons", "f4")], align=True)

#: the all-float64 record the reference's own unit tests build (tests/testQuenching.py:16-19)
test_dtype_f8 = np.dtype([(n, "f8") for n in (
    "eventID", "z_end", "trackID", "tran_diff", "z_start", "x_end", "y_end", "n_electrons", "pdgId", "x_start",
    "y_start", "t_start", "t0_start", "t0_end", "t0", "dx", "long_diff", "pixel_plane", "t_end", "dEdx", "dE", "t",
    "y", "x", "z", "n_photons")])


def _fill(seg, start, end, dedx, t0, event_id, dtype):
    n = start.shape[0]
    out = np.zeros(n, dtype=dtype)
    names = out.dtype.names
    d = end - start
    length = np.sqrt((d * d).sum(axis=1))
    for i, ax in enumerate("xyz"):
        out[ax + "_start"] = start[:, i]
        out[ax + "_end"] = end[:, i]
        out[ax] = 0.5 * (out[ax + "_start"].astype(np.float64) + out[ax + "_end"].astype(np.float64))
    out["dx"] = length
    out["dEdx"] = dedx
    out["dE"] = dedx * length
    for k in ("t0", "t0_start", "t0_end"):
        out[k] = t0
    if "event_id" in names:
        out["event_id"] = event_id
    if "eventID" in names:
        out["eventID"] = event_id
    if "segment_id" in names:
        out["segment_id"] = np.arange(n)
    if "traj_id" in names:
        out["traj_id"] = seg
    if "trackID" in names:
        out["trackID"] = seg
    return out


def _tpc_boxes(detector):
    b = np.asarray(detector.TPC_BORDERS, dtype=np.float64)
    lo = np.minimum(b[:, :, 0], b[:, :, 1])
    hi = np.maximum(b[:, :, 0], b[:, :, 1])
    return lo, hi


def _chop(rng, p0, direction, track_len, seg_lo, seg_hi):
    """Chop straight tracks into consecutive segments of length U(seg_lo, seg_hi)."""
    starts, ends, owner = [], [], []
    for it in range(p0.shape[0]):
        L = track_len[it]
        if L <= 0:
            continue
        n_est = int(L / (0.5 * (seg_lo + seg_hi)) * 1.5) + 4
        cuts = np.cumsum(rng.uniform(seg_lo, seg_hi, n_est))
        cuts = cuts[cuts < L]
        edges = np.concatenate([[0.0], cuts, [L]])
        if edges[-1] - edges[-2] < 1e-3 and len(edges) > 2:
            edges = np.delete(edges, -2)
        s = p0[it] + edges[:-1, None] * direction[it]
        e = p0[it] + edges[1:, None] * direction[it]
        starts.append(s); ends.append(e); owner.append(np.full(s.shape[0], it))
    if not starts:
        return np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0, dtype=np.int64)
    return np.concatenate(starts), np.concatenate(ends), np.concatenate(owner)


def _exit_length(p0, direction, lo, hi):
    """Distance along `direction` from p0 (inside the box) to the box boundary."""
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (lo - p0) / direction
        t2 = (hi - p0) / direction
    t = np.where(direction > 0, t2, np.where(direction < 0, t1, np.inf))
    return np.clip(t.min(axis=1), 0, None)


def cosmic_segments(n, detector, seed=12345, dtype=segment_dtype, seg_len=(0.05, 0.5), dedx=2.1):
    """~n segments of straight down-going muons (cos^2-like zenith) through randomly chosen TPCs, one event."""
    rng = np.random.default_rng(seed)
    lo, hi = _tpc_boxes(detector)
    margin = 0.05
    out_s, out_e, out_o = [], [], []
    total, base = 0, 0
    while total < n:
        nt = max(8, int((n - total) / 150) + 4)
        tpc = rng.integers(0, lo.shape[0], nt)
        l, h = lo[tpc] + margin, hi[tpc] - margin
        p0 = np.stack([rng.uniform(l[:, 0], h[:, 0]), h[:, 1], rng.uniform(l[:, 2], h[:, 2])], axis=1)
        ct = np.sqrt(rng.uniform(0.3, 1.0, nt))
        st = np.sqrt(1 - ct * ct)
        phi = rng.uniform(0, 2 * np.pi, nt)
        direction = np.stack([st * np.cos(phi), -ct, st * np.sin(phi)], axis=1)
        L = _exit_length(p0, direction, l, h)
        s, e, o = _chop(rng, p0, direction, L, *seg_len)
        out_s.append(s); out_e.append(e); out_o.append(o + base)
        total += s.shape[0]; base += nt
    s = np.concatenate(out_s)[:n]; e = np.concatenate(out_e)[:n]; o = np.concatenate(out_o)[:n]
    return _fill(o, s, e, dedx, 0.0, 0, dtype)


def beam_spill_segments(n, detector, seed=12345, dtype=segment_dtype, n_events=1, seg_len=(0.05, 0.5)):
    """~n segments of a beam-spill-like topology: vertices uniform in the active volume, 3-8 forward-peaked
    tracks each (exponential length, mean 50 cm, clipped to the TPC) plus 30% short high-dE/dx stubs;
    t0 ~ U(0, 10) us per vertex."""
    rng = np.random.default_rng(seed)
    lo, hi = _tpc_boxes(detector)
    margin = 0.05
    out_s, out_e, out_o, out_dedx, out_t0, out_ev = [], [], [], [], [], []
    total, base = 0, 0
    while total < n:
        nv = max(4, int((n - total) / 600) + 2)
        mult = rng.integers(3, 9, nv)
        nt = int(mult.sum())
        vtx_of = np.repeat(np.arange(nv), mult)
        tpc_v = rng.integers(0, lo.shape[0], nv)
        l, h = lo[tpc_v] + margin, hi[tpc_v] - margin
        vpos = np.stack([rng.uniform(l[:, k], h[:, k]) for k in range(3)], axis=1)
        t0_v = rng.uniform(0, 10.0, nv)
        ev_v = rng.integers(0, n_events, nv)
        p0 = vpos[vtx_of]
        lt, ht = l[vtx_of], h[vtx_of]
        # forward (+x in larnd-sim axes is the beam direction after the x<->z swap) peaked directions
        ct = 1 - rng.exponential(0.15, nt).clip(0, 2)
        st = np.sqrt(np.clip(1 - ct * ct, 0, 1))
        phi = rng.uniform(0, 2 * np.pi, nt)
        direction = np.stack([ct, st * np.cos(phi), st * np.sin(phi)], axis=1)
        stub = rng.uniform(0, 1, nt) < 0.3
        L = np.where(stub, rng.uniform(0.2, 2.0, nt), rng.exponential(50.0, nt))
        L = np.minimum(L, _exit_length(p0, direction, lt, ht))
        dedx_t = np.where(stub, rng.uniform(5, 20, nt), 2.1)
        s, e, o = _chop(rng, p0, direction, L, *seg_len)
        out_s.append(s); out_e.append(e); out_o.append(o + base)
        out_dedx.append(dedx_t[o]); out_t0.append(t0_v[vtx_of[o]]); out_ev.append(ev_v[vtx_of[o]])
        total += s.shape[0]; base += nt
    cat = lambda x: np.concatenate(x)[:n]
    return _fill(cat(out_o), cat(out_s), cat(out_e), cat(out_dedx), cat(out_t0), cat(out_ev), dtype)


def response_lut(detector, shape=None, dtype=np.float32):
    """Analytic induction-like field response [Rx, Ry, Rt]: unipolar collection pulse under the pad,
    bipolar induction pulse away from it; smooth and seed-free so results are checkable."""
    ratio = detector.TIME_SAMPLING / detector.RESPONSE_SAMPLING
    if shape is None:
        shape = (45, 45, int(round(1950 * ratio)))
    rx, ry, rt = shape
    dt = detector.RESPONSE_SAMPLING
    i = (np.arange(rx) + 0.5)[:, None, None] * detector.RESPONSE_BIN_SIZE
    j = (np.arange(ry) + 0.5)[None, :, None] * detector.RESPONSE_BIN_SIZE
    r = np.sqrt(i * i + j * j)
    k = np.arange(rt)[None, None, :]
    # the drifting charge reaches the anode TIME_WINDOW after the start of the tabulated waveform
    # (detsim.py:332: t0 = t_arrival - TIME_WINDOW); leave a 3 us tail and keep the table zero afterwards
    k_arr = min(rt - 1, int(round(detector.TIME_WINDOW / dt)) - int(round(3.0 / dt)))
    u = (k_arr - k) * dt                               # time before arrival at the anode [us]
    live = u >= 0
    u = np.where(live, u, 0.0)
    tau = 0.6 + 2.5 * r
    amp = np.exp(-(r / (0.6 * detector.PIXEL_PITCH)) ** 2)
    collect = amp * np.exp(-u / tau) / tau
    induct = (1 - amp) * 0.15 * np.exp(-r / detector.PIXEL_PITCH) * (np.exp(-u / (3 * tau)) / (3 * tau)
                                                                  - np.exp(-u / tau) / tau)
    collect = np.where(live, collect, 0.0)
    induct = np.where(live, induct, 0.0)
    return np.ascontiguousarray((collect + induct).astype(dtype))


def light_lut(shape=(14, 26, 8, 48), n_prof=16):
    """Light LUT records {vis, t0, t0_avg, time_dist[n_prof]} (SURVEY.md appendix B.3 formulas)."""
    dt = np.dtype([("vis", "f4"), ("t0", "f4"), ("t0_avg", "f4"), ("time_dist", "f4", (n_prof,))])
    lut = np.zeros(shape, dtype=dt)
    i, j, k, d = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    lut["vis"] = 1e-3 / (1 + 0.1 * i + 0.05 * j + 0.2 * k + 0.02 * d)
    lut["t0"] = 1 + 0.5 * k + 0.1 * (d % 6)
    lut["t0_avg"] = lut["t0"] + 3
    prof = np.exp(-np.arange(n_prof) / 4.0)
    lut["time_dist"] = (prof / prof.sum()).astype(np.float32)
    return lut
