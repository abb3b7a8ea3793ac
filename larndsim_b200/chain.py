"""Device-resident batch driver: one call runs quench -> drift -> get_pixels -> tracks_current_mc ->
(unique / index maps) -> sum_pixel_signals -> get_adc_values -> digitize for one batch of segments,
i.e. the body of the reference's batch loop (cli/simulate_pixels.py:907-1117) without its host
round trips.  Thin wrapper over ``lsb_chain_*`` of the C ABI."""
import ctypes as C

import numpy as np
import torch

from . import _abi
from . import _launch as _l
from . import consts as _consts

STAGES = ("quench+drift", "max/get_pixels", "unique_pix", "time_intervals", "tracks_current_mc", "index_maps",
          "sum_pixel_signals", "get_adc_values", "digitize")


class ChainResult:
    def __init__(self, r, K, A, Tt, handle=None):
        self._handle = handle
        self.n_segments, self.n_unique_pixels = int(r.n_segments), int(r.n_unique_pixels)
        self.max_active, self.max_neighbors, self.n_ticks = int(r.max_active), int(r.max_neighbors), int(r.n_ticks)
        self.n_hits, self.n_samples, self.n_fma, self.n_pairs = int(r.n_hits), int(r.n_samples), int(r.n_fma), int(r.n_pairs)
        self.n_groups, self.n_edge, self.n_irregular = int(r.n_groups), int(r.n_edge), int(r.n_irregular)
        self.stage_ms = {STAGES[i]: float(r.stage_ms[i]) for i in range(len(STAGES))}
        #: async batches: ms since the library's reference event of (front begin, front end, MC begin, MC end, FEE begin, done)
        self.timeline = [float(r.stage_ms[i]) for i in range(6)]
        self._r, self._K, self._A, self._Tt = r, K, A, Tt

    def _view(self, ptr, shape, np_dtype):
        n = int(np.prod(shape))
        if not ptr or n == 0:
            return torch.empty(shape, dtype=_l._TORCH_OF[np.dtype(np_dtype)], device="cuda")

        class _Holder:
            pass
        h = _Holder()
        h.__cuda_array_interface__ = {"shape": tuple(int(v) for v in shape), "typestr": np.dtype(np_dtype).str,
                                      "data": (int(ptr), False), "version": 3, "strides": None}
        return torch.as_tensor(h, device="cuda")

    # device views valid until the next run / destruction of the chain
    @property
    def unique_pix(self):
        return self._view(self._r.unique_pix, (self.n_unique_pixels,), np.int32)

    @property
    def track_pixel_map(self):
        return self._view(self._r.track_pixel_map, (self.n_unique_pixels, self._K), np.int64)

    @property
    def adc_list(self):
        return self._view(self._r.adc_list, (self.n_unique_pixels, self._A), np.float64)

    @property
    def adc_digit(self):
        return self._view(self._r.adc_digit, (self.n_unique_pixels, self._A), np.float64)

    @property
    def adc_ticks_list(self):
        return self._view(self._r.adc_ticks_list, (self.n_unique_pixels, self._A), np.float64)

    @property
    def current_fractions(self):
        return self._view(self._r.current_fractions, (self.n_unique_pixels, self._A, self._K), np.float64)

    @property
    def signals(self):
        """f4[S, P, T] induced currents.  The fused chain stores the rows sparsely (only the ticks a pair's samples cover are
        written, nothing downstream reads the rest); the first access zero-fills the unwritten parts of this batch."""
        if self._handle:
            _l.check(_l.lib().lsb_chain_signals_dense(C.c_void_p(self._handle), _l.stream()), "chain_signals_dense")
        return self._view(self._r.signals, (self.n_segments, self.max_neighbors, self.n_ticks), np.float32)

    @property
    def pixels_signals(self):
        return self._view(self._r.pixels_signals, (self.n_unique_pixels, self._Tt), np.float64)


class Pipeline:
    """Round-robin over `depth` chains: batch i+1 is queued while batch i is still in flight, so the
    latency-bound FEE stage of one batch runs under the MC stage of the next (each chain has a
    high-priority stream for front/FEE work and a low-priority one for the MC kernels).

    RNG: the chains of a pipeline use the "fresh" policy (see :class:`Chain`) -- with private evolving state arrays two
    chains seeded alike would replay the same noise on consecutive batches.  Give every batch its own ``rng_seed``."""

    def __init__(self, track_dtype, response, depth=2, **kw):
        kw.setdefault("rng_fresh", True)
        self.chains = [Chain(track_dtype, response, **kw) for _ in range(depth)]
        self._inflight = []          # chains with a pending batch, oldest first
        self._next = 0

    def full(self):
        return len(self._inflight) == len(self.chains)

    def collect(self):
        """Wait for the oldest batch in flight and return its result.  The device views of a result stay valid
        until the chain that produced it is reused, i.e. until `depth` further submits."""
        ch = self._inflight.pop(0)
        return ch.wait()

    def _slot(self):
        if self.full():
            raise RuntimeError("pipeline full: collect() a result before submitting another batch")
        ch = self.chains[self._next]
        self._next = (self._next + 1) % len(self.chains)
        self._inflight.append(ch)
        return ch

    def submit(self, tracks_dev, **kw):
        self._slot().run_async(tracks_dev, **kw)

    def submit_host(self, tracks_host, unique_pix_out, adc_out, ticks_out, **kw):
        self._slot().run_host_async(tracks_host, unique_pix_out, adc_out, ticks_out, **kw)

    def drain(self):
        out = []
        while self._inflight:
            out.append(self.collect())
        return out

    def close(self):
        self.drain()
        for ch in self.chains:
            ch.close()


class Chain:
    """``Chain(track_dtype, response)``; ``run(tracks_dev)`` on device records, ``run_host(tracks)``
    on a (pinned) host structured array with H2D/D2H inside the call."""

    def __init__(self, track_dtype, response, rng_mode="cloud", stage_timing=False, dense=False, exact_fractions=False,
                 rng_fresh=False):
        """``rng_fresh=False``: the reference's ``maybe_create_rng_states`` -- one state array per chain that evolves from
        batch to batch (cli/simulate_pixels.py:92-104, 1015, 1079).  ``rng_fresh=True``: every batch starts from
        ``create_xoroshiro128p_states(n, seed=rng_seed)``, so its result depends on (records, rng_seed) only -- pass a
        different ``rng_seed`` per batch (e.g. ``rand_seed + batch number``)."""
        self._c = _consts.snapshot()
        self._L = _abi.track_layout(track_dtype)
        self.dtype = np.dtype(track_dtype)
        self._resp = _l.dev(response, name="response")
        if isinstance(response, np.ndarray):      # keep the staged copy alive
            self._resp_keep = self._resp.keep
        r = self._resp
        lib = _l.lib()
        self._h = lib.lsb_chain_create(C.byref(self._c), C.byref(self._L), r.c, C.c_int32(r.shape[0]), C.c_int32(r.shape[1]),
                                       C.c_int32(r.shape[2]), C.c_int32(1 if r.dtype == np.dtype("f8") else 0),
                                       C.c_int32({"cloud": 0, "replay": 1}[rng_mode]), C.c_int32(1 if stage_timing else 0))
        if not self._h:
            raise _abi.LsbError("lsb_chain_create failed: %s" % lib.lsb_last_error().decode())
        if dense:
            _l.check(lib.lsb_chain_set_dense(C.c_void_p(self._h), C.c_int32(1)), "chain_set_dense")
        if exact_fractions:
            _l.check(lib.lsb_chain_set_exact_fractions(C.c_void_p(self._h), C.c_int32(1)), "chain_set_exact_fractions")
        if rng_fresh:
            _l.check(lib.lsb_chain_set_rng_fresh(C.c_void_p(self._h), C.c_int32(1)), "chain_set_rng_fresh")
        self._K, self._A, self._Tt = int(self._c.max_tracks_per_pixel), int(self._c.max_adc_values), int(self._c.n_time_ticks)

    def close(self):
        if getattr(self, "_h", None) and _l is not None:          # `_l` is None during interpreter shutdown
            _l.lib().lsb_chain_destroy(C.c_void_p(self._h))
            self._h = None

    __del__ = close

    def run(self, tracks_dev, quench_mode=None, rng_seed=0, n_events=1):
        t = _l.dev(tracks_dev, name="tracks", records=True)
        if t.dtype != self.dtype:
            raise TypeError("tracks dtype differs from the dtype this chain was created for")
        r = _abi.ChainResult()
        qm = int(self._c.mode_birks if quench_mode is None else quench_mode)
        _l.check(_l.lib().lsb_chain_run(C.c_void_p(self._h), t.c, C.c_int64(t.shape[0]), C.c_int32(qm),
                                        C.c_uint64(int(rng_seed)), C.c_int32(int(n_events)), C.byref(r), _l.stream()), "chain_run")
        return ChainResult(r, self._K, self._A, self._Tt, self._h)

    def run_async(self, tracks_dev, quench_mode=None, rng_seed=0, n_events=1):
        """Queue one batch on this chain's own streams and return immediately; collect with :meth:`wait`."""
        t = _l.dev(tracks_dev, name="tracks", records=True)
        if t.dtype != self.dtype:
            raise TypeError("tracks dtype differs from the dtype this chain was created for")
        qm = int(self._c.mode_birks if quench_mode is None else quench_mode)
        self._keep = t
        _l.check(_l.lib().lsb_chain_run_async(C.c_void_p(self._h), t.c, C.c_int64(t.shape[0]), C.c_int32(qm),
                                              C.c_uint64(int(rng_seed)), C.c_int32(int(n_events)), _l.stream()), "chain_run_async")

    def run_host_async(self, tracks_host, unique_pix_out, adc_out, ticks_out, quench_mode=None, rng_seed=0, n_events=1):
        qm = int(self._c.mode_birks if quench_mode is None else quench_mode)

        def p(a):
            return C.c_void_p(a.data_ptr() if isinstance(a, torch.Tensor) else a.ctypes.data)
        n = tracks_host.shape[0] if not isinstance(tracks_host, torch.Tensor) else tracks_host.numel() // self.dtype.itemsize
        self._keep = (tracks_host, unique_pix_out, adc_out, ticks_out)
        _l.check(_l.lib().lsb_chain_run_host_async(C.c_void_p(self._h), p(tracks_host), C.c_int64(n), C.c_int32(qm),
                                                   C.c_uint64(int(rng_seed)), C.c_int32(int(n_events)), p(unique_pix_out),
                                                   p(adc_out), p(ticks_out), C.c_int64(unique_pix_out.shape[0])), "chain_run_host_async")

    def wait(self):
        r = _abi.ChainResult()
        _l.check(_l.lib().lsb_chain_wait(C.c_void_p(self._h), C.byref(r)), "chain_wait")
        self._keep = None
        return ChainResult(r, self._K, self._A, self._Tt, self._h)

    def run_host(self, tracks_host, unique_pix_out, adc_out, ticks_out, quench_mode=None, rng_seed=0, n_events=1):
        """tracks_host: structured host array (ideally pinned); outputs: preallocated (pinned) host
        buffers int32[U_cap], float64[U_cap, A], float64[U_cap, A]."""
        r = _abi.ChainResult()
        qm = int(self._c.mode_birks if quench_mode is None else quench_mode)

        def p(a):
            return C.c_void_p(a.data_ptr() if isinstance(a, torch.Tensor) else a.ctypes.data)
        n = tracks_host.shape[0] if not isinstance(tracks_host, torch.Tensor) else tracks_host.numel() // self.dtype.itemsize
        ucap = unique_pix_out.shape[0]
        _l.check(_l.lib().lsb_chain_run_host(C.c_void_p(self._h), p(tracks_host), C.c_int64(n), C.c_int32(qm),
                                             C.c_uint64(int(rng_seed)), C.c_int32(int(n_events)), p(unique_pix_out), p(adc_out),
                                             p(ticks_out), C.c_int64(ucap), C.byref(r), _l.stream()), "chain_run_host")
        return ChainResult(r, self._K, self._A, self._Tt, self._h)
