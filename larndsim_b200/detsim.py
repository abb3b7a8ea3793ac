"""Drop-in for ``larndsim.detsim`` (reference: larndsim/detsim.py)."""
import ctypes as C

import numpy as np
import torch

from . import _launch as _l
from . import consts as _consts
from . import rng as _rng
from .pixels_from_track import id2pixel

#: RNG discipline of ``tracks_current_mc``: "cloud" (production: one sample cloud per
#: (segment,pixel), what a converged warp of the reference draws) or "replay" (the reference's
#: draw pattern with ticks consumed in order = the CUDA simulator with 1-thread blocks).
MC_MODE = "cloud"

_mc_workspace = None


def _workspace(nbytes):
    global _mc_workspace
    if _mc_workspace is None or _mc_workspace.numel() < nbytes:
        _mc_workspace = None
        _mc_workspace = torch.empty(int(nbytes), dtype=torch.uint8, device="cuda")
    return _mc_workspace


def get_pixel_coordinates(pixel_id):
    """Lower-left corner of the pixel pad (detsim.py:180-191); host helper."""
    i_x, i_y, plane_id = id2pixel(pixel_id)
    d = _consts.provider().detector
    b = d.TPC_BORDERS[int(plane_id)]
    return i_x * d.PIXEL_PITCH + b[0][0], i_y * d.PIXEL_PITCH + b[1][0]


@_l.kernel
def time_intervals(track_starts, time_max, tracks):
    """``time_intervals[BPG, TPB](track_starts, time_max, tracks)`` (detsim.py:18-40)."""
    c = _l.snapshot()
    t = _l.dev(tracks, name="tracks", records=True)
    L = _l.layout(t)
    ts = _l.dev(track_starts, want=np.float64, write=True, name="track_starts")
    tm = _l.dev(time_max, want=np.int64, write=True, name="time_max")
    _l.check(_l.lib().lsb_time_intervals(C.byref(c), C.byref(L), t.c, C.c_int64(t.shape[0]), ts.c, tm.c, _l.stream()),
             "time_intervals")
    _l.finish(ts, tm)


def _response(response):
    r = _l.dev(response, name="response")
    if len(r.shape) != 3 or r.dtype not in (np.dtype("f4"), np.dtype("f8")):
        raise TypeError("response must be a 3-D float32/float64 array")
    return r


@_l.kernel_with_config
def tracks_current_mc(signals, pixels, tracks, response, rng_states, _config=None):
    """``tracks_current_mc[BPG, TPB](signals, pixels, tracks, response, rng_states)``
    (detsim.py:258-348).  ``signals`` f4[S,P,T] must arrive zeroed like in the reference."""
    c = _l.snapshot()
    t = _l.dev(tracks, name="tracks", records=True)
    L = _l.layout(t)
    sg = _l.dev(signals, want=np.float32, write=True, name="signals")
    px = _l.dev(pixels, want=np.int32, name="pixels")
    r = _response(response)
    st, n_rng = _rng.states_dev(rng_states)
    S, P, T = sg.shape
    if px.shape != (S, P) or t.shape[0] < S:
        raise ValueError("tracks_current_mc: signals/pixels/tracks shapes disagree")
    ntrk = _l.grid_threads(_config, 0) or S
    lib = _l.lib()
    guess = max(S, 1) * 4000
    nbytes = lib.lsb_tracks_current_mc_workspace_bytes(C.c_int64(S), C.c_int32(P), C.c_int64(guess))
    ws = _workspace(nbytes)
    mode = {"cloud": 0, "replay": 1}[MC_MODE]
    _l.check(lib.lsb_tracks_current_mc(C.byref(c), C.byref(L), t.c, C.c_int64(S), px.c, C.c_int32(P), sg.c, C.c_int32(T),
                                       r.c, C.c_int32(r.shape[0]), C.c_int32(r.shape[1]), C.c_int32(r.shape[2]),
                                       C.c_int32(1 if r.dtype == np.dtype("f8") else 0), st.c, C.c_int64(n_rng),
                                       C.c_int64(ntrk), C.c_int32(mode), C.c_void_p(ws.data_ptr()), C.c_int64(ws.numel()),
                                       _l.stream()), "tracks_current_mc")
    _l.finish(sg, st)


@_l.kernel
def tracks_current(signals, pixels, tracks, response):
    """``tracks_current[BPG, TPB](signals, pixels, tracks, response)`` (detsim.py:351-453)."""
    c = _l.snapshot()
    t = _l.dev(tracks, name="tracks", records=True)
    L = _l.layout(t)
    sg = _l.dev(signals, want=np.float32, write=True, name="signals")
    px = _l.dev(pixels, want=np.int32, name="pixels")
    r = _response(response)
    S, P, T = sg.shape
    _l.check(_l.lib().lsb_tracks_current(C.byref(c), C.byref(L), t.c, C.c_int64(S), px.c, C.c_int32(P), sg.c, C.c_int32(T),
                                         r.c, C.c_int32(r.shape[0]), C.c_int32(r.shape[1]), C.c_int32(r.shape[2]),
                                         C.c_int32(1 if r.dtype == np.dtype("f8") else 0), _l.stream()), "tracks_current")
    _l.finish(sg)


@_l.kernel
def get_track_pixel_map(track_pixel_map, unique_pix, pixels):
    """``get_track_pixel_map[BPG, TPB](track_pixel_map, unique_pix, pixels)`` (detsim.py:529-562)."""
    m = _l.dev(track_pixel_map, want=np.int64, write=True, name="track_pixel_map")
    u = _l.dev(unique_pix, want=np.int32, name="unique_pix")
    px = _l.dev(pixels, want=np.int32, name="pixels")
    _l.check(_l.lib().lsb_get_track_pixel_map(m.c, C.c_int32(m.shape[1]), u.c, C.c_int64(u.shape[0]), px.c,
                                              C.c_int64(px.shape[0]), C.c_int32(px.shape[1]), _l.stream()), "get_track_pixel_map")
    _l.finish(m)


@_l.kernel
def get_track_pixel_map2(track_pixel_map, unique_pix, pixels, distances, max_distance):
    """``get_track_pixel_map2[BPG, TPB](track_pixel_map, unique_pix, pixels, distances, max_distance)``
    (detsim.py:564-607)."""
    m = _l.dev(track_pixel_map, want=np.int64, write=True, name="track_pixel_map")
    u = _l.dev(unique_pix, want=np.int32, name="unique_pix")
    px = _l.dev(pixels, want=np.int32, name="pixels")
    ds = _l.dev(distances, want=np.int32, name="distances")
    if ds.shape != px.shape:
        raise ValueError("get_track_pixel_map2: pixels and distances must have the same shape")
    _l.check(_l.lib().lsb_get_track_pixel_map2(m.c, C.c_int32(m.shape[1]), u.c, C.c_int64(u.shape[0]), px.c, ds.c,
                                               C.c_int64(px.shape[0]), C.c_int32(px.shape[1]), C.c_int32(int(max_distance)),
                                               _l.stream()), "get_track_pixel_map2")
    _l.finish(m)


@_l.kernel
def sum_pixel_signals(pixels_signals, signals, track_starts, pixel_index_map, track_pixel_map, pixels_tracks_signals,
                      overflow_flag):
    """``sum_pixel_signals[BPG, TPB](pixels_signals, signals, track_starts, pixel_index_map,
    track_pixel_map, pixels_tracks_signals, overflow_flag)`` (detsim.py:468-527)."""
    c = _l.snapshot()
    ps = _l.dev(pixels_signals, want=np.float64, write=True, name="pixels_signals")
    sg = _l.dev(signals, want=np.float32, name="signals")
    ts = _l.dev(track_starts, want=np.float64, name="track_starts")
    pim = _l.dev(pixel_index_map, want=np.int64, name="pixel_index_map")
    tpm = _l.dev(track_pixel_map, want=np.int64, name="track_pixel_map")
    pts = _l.dev(pixels_tracks_signals, want=np.float64, write=True, name="pixels_tracks_signals")
    of = _l.dev(overflow_flag, want=np.float64, write=True, name="overflow_flag")
    S, P, T = sg.shape
    U, Tt = ps.shape
    K = tpm.shape[1]
    if pts.shape != (U, Tt, K) or pim.shape != (S, P) or tpm.shape[0] != U or of.shape[0] < U:
        raise ValueError("sum_pixel_signals: array shapes disagree")
    _l.check(_l.lib().lsb_sum_pixel_signals(C.byref(c), ps.c, C.c_int64(U), C.c_int32(Tt), sg.c, C.c_int64(S), C.c_int32(P),
                                            C.c_int32(T), ts.c, pim.c, tpm.c, C.c_int32(K), pts.c, of.c, _l.stream()),
             "sum_pixel_signals")
    _l.finish(ps, pts, of)


def unique_pixels(neighboring_pixels):
    """Device replacement of ``cp.unique(neighboring_pixels)`` minus -1 plus the
    ``pixel_index_map`` loop (cli/simulate_pixels.py:953-956, 1021-1025).
    Returns ``(unique_pix i4[U], pixel_index_map i8[S,P])`` as torch CUDA tensors."""
    c = _l.snapshot()
    px = _l.dev(neighboring_pixels, want=np.int32, name="neighboring_pixels")
    n = px.size
    max_id = int(c.n_pixels[0]) * int(c.n_pixels[1]) * int(c.n_tpc) - 1
    lib = _l.lib()
    wsb = lib.lsb_unique_pixels_workspace_bytes(C.c_int64(max_id))
    ws = torch.empty(int(wsb), dtype=torch.uint8, device="cuda")
    uniq = torch.empty(min(n, max_id + 1), dtype=torch.int32, device="cuda")
    n_u = torch.zeros(1, dtype=torch.int64, device="cuda")
    _l.check(lib.lsb_unique_pixels(px.c, C.c_int64(n), C.c_int64(max_id), C.c_void_p(uniq.data_ptr()),
                                   C.c_void_p(n_u.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_int64(wsb), _l.stream()),
             "unique_pixels")
    pim = torch.empty(px.shape, dtype=torch.int64, device="cuda")
    _l.check(lib.lsb_pixel_index_map(px.c, C.c_int64(n), C.c_int64(max_id), C.c_void_p(ws.data_ptr()),
                                     C.c_void_p(pim.data_ptr()), _l.stream()), "pixel_index_map")
    return uniq[: int(n_u.item())], pim
