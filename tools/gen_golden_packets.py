"""Golden vectors for the hit -> LArPix packet builder (tests/golden/packets_<config>.npz).

Runs ONLY in the build container.  Imports the unmodified reference from /root/reference and calls its own
``larndsim.fee.export_to_hdf5`` (fee.py:84-359) on synthetic hit tables.  ``larpix`` and ``h5py`` are absent
here, so they are replaced by *recording* stand-ins: every attribute the reference sets on a packet object and
the ``mc_packets_assn`` table it writes are captured and stored.  Nothing of the reference is copied; the
inputs, the readout tables the run used (so the GPU box can rebuild them) and the captured outputs are saved.

    python tools/gen_golden_packets.py
"""
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CAPTURE = {}


def install_recorders():
    class _Rec:
        kind = "?"

        def __init__(self, *a, **kw):
            self.__dict__["f"] = dict(kw)

        def __setattr__(self, k, v):
            self.f[k] = v

        def assign_parity(self):
            self.f["parity_assigned"] = 1

    def mk(name):
        return type(name, (_Rec,), {"kind": name})

    class Key:
        def __init__(self, io_group, io_channel, chip):
            self.t = (int(io_group), int(io_channel), int(chip))

    class PacketCollection(list):
        def __init__(self, packets, **kw):
            super().__init__(packets)

    class _H5:
        @staticmethod
        def to_file(filename, packet_list, workers=1):
            CAPTURE["packets"] = list(packet_list)

    class _Attrs(dict):
        pass

    class _Node:
        def __init__(self):
            self.attrs = _Attrs()

    class _File:
        store = {}

        def __init__(self, filename, mode):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def keys(self):
            return []                                   # every call creates its dataset afresh

        def create_dataset(self, name, data=None, maxshape=None):
            self.store[name] = data
            CAPTURE[name] = data

        def __getitem__(self, k):
            return self.store.setdefault(k, _Node())

    lp = types.ModuleType("larpix")
    pkt = types.ModuleType("larpix.packet")
    for n in ("Packet_v2", "TimestampPacket", "TriggerPacket", "SyncPacket"):
        setattr(pkt, n, mk(n))
    pkt.PacketCollection = PacketCollection
    key = types.ModuleType("larpix.key")
    key.Key = Key
    fmt = types.ModuleType("larpix.format")
    fmt.hdf5format = _H5
    lp.packet, lp.key, lp.format = pkt, key, fmt
    sys.modules.update({"larpix": lp, "larpix.packet": pkt, "larpix.key": key, "larpix.format": fmt})
    h5 = types.ModuleType("h5py")
    h5.File = _File
    sys.modules["h5py"] = h5


def readout_tables(consts):
    """The readout constants ``export_to_hdf5`` reads from larndsim.consts, in plain containers."""
    d, s, li, u = consts.detector, consts.sim, consts.light, consts.units
    from larndsim import fee
    return dict(clock_cycle=float(d.CLOCK_CYCLE), clock_reset_period=int(d.CLOCK_RESET_PERIOD), light_trig_mode=int(li.LIGHT_TRIG_MODE),
                n_pixels=[int(x) for x in d.N_PIXELS], n_pixels_per_tile=[int(x) for x in d.N_PIXELS_PER_TILE],
                module_to_io_groups={int(k): [int(x) for x in v] for k, v in d.MODULE_TO_IO_GROUPS.items()},
                tile_map=np.asarray(d.TILE_MAP).tolist(),
                tile_orientations={int(k): [int(x) for x in v] for k, v in d.TILE_ORIENTATIONS.items()},
                pixel_connection={(int(k[0]), int(k[1])): (int(v[0]), int(v[1])) for k, v in d.PIXEL_CONNECTION_DICT.items()},
                tile_chip_to_io={int(t): {int(c): int(x) for c, x in m.items()} for t, m in d.TILE_CHIP_TO_IO.items()},
                adc_pedestal=float(fee.digitize(0)), max_tracks_per_pixel=int(s.MAX_TRACKS_PER_PIXEL),
                association_count=int(s.ASSOCIATION_COUNT_TO_STORE), mus=float(u.mus), s=float(u.s))


def synth_hits(tables, U, A, K, seed, n_events, big_times=False):
    """Random hit tables with the shapes of cli/simulate_pixels.py:1264-1279: pixels spread over every tile,
    1..A hits per pixel, some pixels without hits, sparse fractions, -1 padded truth maps."""
    rng = np.random.default_rng(seed)
    npx, npy = tables["n_pixels"]
    n_planes = 2 * len(tables["module_to_io_groups"])
    pix = np.sort(rng.choice(npx * npy * n_planes, U, replace=False)).astype(np.int32)
    # one out-of-detector plane to exercise the "module not valid" skip
    pix[-1] = npx * npy * (n_planes + 1) + 5
    ped = tables["adc_pedestal"]
    adc = np.full((U, A), ped)
    ticks = np.zeros((U, A))
    ev = np.zeros((U, A), dtype=np.int64)
    events = np.sort(rng.choice(np.arange(3, 3 + 4 * n_events), n_events, replace=False))
    pix_event = np.sort(rng.integers(0, n_events, U))
    for i in range(U):
        nh = int(rng.integers(0, min(A, 4) + 1)) if rng.random() < 0.9 else A
        adc[i, :nh] = np.round(ped) + rng.integers(1, 120, nh)
        ticks[i, :nh] = np.sort(rng.uniform(0, 190.0, nh))
        ev[i, :] = events[pix_event[i]]
        if nh > 1 and rng.random() < 0.2 and pix_event[i] + 1 < n_events:
            ev[i, nh - 1] = events[pix_event[i] + 1]            # a pixel whose later hit belongs to the next event
    ticks = np.round(ticks * 10) / 10                             # multiples of the 0.1 us clock: equal timestamps happen
    trk = np.full((U, K), -1, dtype=np.int64)
    trj = np.full((U, K), -1, dtype=np.int64)
    cf = np.zeros((U, A, K))
    for i in range(U):
        nt = int(rng.integers(0, min(K, 30) + 1))
        trk[i, :nt] = rng.choice(100000, nt, replace=False)
        trj[i, :nt] = rng.integers(0, 6, nt) + 10 * pix_event[i]
        for a in range(A):
            if nt:
                w = rng.random(nt) * (rng.random(nt) < 0.7)
                w[rng.random(nt) < 0.1] *= -0.2                   # bipolar induction: negative fractions exist
                cf[i, a, :nt] = w / max(np.abs(w).sum(), 1e-9)
    period = tables["clock_reset_period"] * tables["clock_cycle"]
    if big_times:
        t0 = np.sort(rng.uniform(0.3 * period, 3.2 * period, n_events))      # several clock rollovers inside the batch
    else:
        t0 = np.sort(rng.uniform(1e3, 0.5 * period, n_events))
    trig_ev = np.repeat(events, 2)[: 2 * n_events - 1]
    trig_t = rng.uniform(0, 5.0, len(trig_ev))
    mods = sorted(tables["module_to_io_groups"])
    trig_mod = rng.choice(mods, len(trig_ev)).astype(np.float64)
    return dict(event_id=ev, adc=adc, ticks=ticks, unique_pix=pix, current_fractions=cf, track_ids=trk, traj_ids=trj,
                event_start_times=t0, trig_times=trig_t, trig_event=trig_ev.astype(np.int64), trig_modules=trig_mod)


FIELDS = ("dataword", "timestamp", "channel_id", "receipt_timestamp", "packet_type", "first_packet")


def capture_to_arrays(packets):
    """kind code, io_group, io_channel, chip, the scalar attributes, float timestamp, sub type byte"""
    kinds = {"Packet_v2": 0, "TimestampPacket": 4, "SyncPacket": 6, "TriggerPacket": 7}
    n = len(packets)
    out = dict(kind=np.zeros(n, np.int32), io_group=np.full(n, -1, np.int64), io_channel=np.full(n, -1, np.int64),
               chip=np.full(n, -1, np.int64), ts_float=np.full(n, np.nan), sub_type=np.full(n, -1, np.int64),
               parity_assigned=np.zeros(n, np.int32))
    for f in FIELDS:
        out[f] = np.full(n, -1, np.int64)
    for i, p in enumerate(packets):
        out["kind"][i] = kinds[p.kind]
        f = p.f
        ck = f.get("chip_key")
        if isinstance(ck, str):
            a, b, c = (int(x) for x in ck.split("-"))
            out["io_group"][i], out["io_channel"][i], out["chip"][i] = a, b, c
        elif ck is not None:
            out["io_group"][i], out["io_channel"][i], out["chip"][i] = ck.t
        if "io_group" in f:
            out["io_group"][i] = int(f["io_group"])
        if p.kind == "TimestampPacket":
            out["ts_float"][i] = float(f["timestamp"])
        else:
            for k in FIELDS:
                if k in f:
                    out[k][i] = int(f[k])
        for k in ("sync_type", "trigger_type"):
            if k in f:
                out["sub_type"][i] = f[k][0]
        out["parity_assigned"][i] = f.get("parity_assigned", 0)
    return out


def run_case(name, detprop, layout, simprop, U, seed, n_events, big_times, bad=False):
    import refharness as rh
    install_recorders()
    rh.load_reference()
    consts = rh.load_properties(detprop, layout, simprop)
    from larndsim import fee
    tables = readout_tables(consts)
    A, K = int(consts.sim.MAX_ADC_VALUES), int(consts.sim.MAX_TRACKS_PER_PIXEL)
    h = synth_hits(tables, U, A, K, seed, n_events, big_times)
    bad_file = None
    bad_dict = {}
    if bad:
        # disable a few channels that the batch actually hits
        import yaml
        import tempfile
        CAPTURE.clear()
        fee.export_to_hdf5(h["event_id"], h["adc"], h["ticks"], h["unique_pix"], h["current_fractions"], h["track_ids"], h["traj_ids"],
                           "unused.h5", h["event_start_times"].copy(), h["trig_times"], h["trig_event"], h["trig_modules"])
        first = capture_to_arrays(CAPTURE["packets"])
        sel = np.nonzero(first["kind"] == 0)[0][::7][:12]
        for i in sel:
            bad_dict.setdefault("%d-%d-%d" % (first["io_group"][i], first["io_channel"][i], first["chip"][i]), []).append(int(first["channel_id"][i]))
        tf = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
        yaml.safe_dump(bad_dict, tf)
        tf.close()
        bad_file = tf.name
    CAPTURE.clear()
    packets, ds = fee.export_to_hdf5(h["event_id"], h["adc"], h["ticks"], h["unique_pix"], h["current_fractions"], h["track_ids"],
                                     h["traj_ids"], "unused.h5", h["event_start_times"].copy(), h["trig_times"], h["trig_event"],
                                     h["trig_modules"], bad_channels=bad_file)
    cap = capture_to_arrays(packets)
    assert len(packets) == len(CAPTURE["packets"]) == len(ds)
    out = {("in_" + k): v for k, v in h.items()}
    out.update({("pk_" + k): v for k, v in cap.items()})
    for f in ds.dtype.names:
        out["assn_" + f] = ds[f]
    # the readout tables, flattened to arrays
    pc = np.array([[k[0], k[1], v[0], v[1]] for k, v in sorted(tables["pixel_connection"].items())], dtype=np.int32)
    tci = np.array([[t, c, x] for t, m in sorted(tables["tile_chip_to_io"].items()) for c, x in sorted(m.items())], dtype=np.int32)
    mio = np.array([[m, j, g] for m, v in sorted(tables["module_to_io_groups"].items()) for j, g in enumerate(v)], dtype=np.int32)
    tor = np.array([[t] + list(v) for t, v in sorted(tables["tile_orientations"].items())], dtype=np.int32)
    out.update(tab_pixel_connection=pc, tab_tile_chip_to_io=tci, tab_module_to_io_groups=mio, tab_tile_orientations=tor,
               tab_tile_map=np.asarray(tables["tile_map"], dtype=np.int32),
               tab_scalars=np.array([tables["clock_cycle"], tables["clock_reset_period"], tables["light_trig_mode"], tables["adc_pedestal"],
                                     tables["max_tracks_per_pixel"], tables["association_count"], tables["mus"],
                                     tables["n_pixels"][0], tables["n_pixels"][1], tables["n_pixels_per_tile"][0],
                                     tables["n_pixels_per_tile"][1], tables["s"]], dtype=np.float64),
               bad_keys=np.array(sorted(bad_dict), dtype="U32"),
               bad_channels=np.array([",".join(str(c) for c in bad_dict[k]) for k in sorted(bad_dict)], dtype="U64"))
    path = os.path.join(ROOT, "tests", "golden", "packets_%s.npz" % name)
    np.savez_compressed(path, **out)
    kinds, counts = np.unique(cap["kind"], return_counts=True)
    print(name, "packets", len(packets), dict(zip(kinds.tolist(), counts.tolist())), "->", path, "%.0f kB" % (os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    import subprocess
    if len(sys.argv) > 1:
        a = sys.argv[1:]
        run_case(a[0], a[1], a[2], a[3], int(a[4]), int(a[5]), int(a[6]), a[7] == "1", a[8] == "1")
    else:
        # one fresh process per configuration: larndsim.consts module globals persist across load_properties calls
        for case in (("module0", "module0.yaml", "multi_tile_layout-2.3.16.yaml", "singles_sim.yaml", "300", "11", "3", "0", "0"),
                     ("module0_rollover_bad", "module0.yaml", "multi_tile_layout-2.3.16.yaml", "singles_sim.yaml", "260", "12", "6", "1", "1"),
                     ("2x2", "2x2_no_modvar.yaml", "multi_tile_layout-2.4.16.yaml", "2x2_NuMI_sim_no_modvar.yaml", "300", "13", "4", "0", "0")):
            subprocess.check_call([sys.executable, os.path.abspath(__file__)] + list(case))
