"""numba.cuda.random-compatible xoroshiro128+ states (third-party on the reference's path:
numba/cuda/random.py ``create_xoroshiro128p_states``; call sites cli/simulate_pixels.py:39-40,92-104).

State layout is Numba's: records ``{s0: u8, s1: u8}``, 16 bytes, so arrays made by Numba can be
passed to the kernels here and vice versa."""
import ctypes as C

import numpy as np
import torch

from . import _launch as _l

xoroshiro128p_dtype = np.dtype([("s0", np.uint64), ("s1", np.uint64)], align=True)


def create_xoroshiro128p_states_host(n, seed, subsequence_start=0):
    out = np.zeros(int(n), dtype=xoroshiro128p_dtype)
    if n:
        _l.check(_l.lib().lsb_rng_create_states_host(out.ctypes.data_as(C.c_void_p), C.c_int64(int(n)),
                                                     C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF),
                                                     C.c_uint64(int(subsequence_start))), "create_xoroshiro128p_states")
    return out


def create_xoroshiro128p_states(n, seed, subsequence_start=0):
    """Device array of ``n`` states, state i = jump^(subsequence_start+i)(splitmix64(seed)) -- made on the device in one
    launch (``lsb_rng_create_states``: the 2^64-step jump as a GF(2) matrix power), bit-identical to Numba's host loop."""
    out = _l.DeviceRecords(dtype=xoroshiro128p_dtype, n=int(n))
    if n:
        _l.check(_l.lib().lsb_rng_create_states(C.c_void_p(out.buf.data_ptr()), C.c_int64(int(n)),
                                                C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), C.c_uint64(int(subsequence_start)),
                                                _l.stream()), "create_xoroshiro128p_states")
    return out


def _state_bytes(rng_states):
    """uint8 CUDA tensor over any device state array (this package's, Numba's or CuPy's)."""
    if hasattr(rng_states, "buf"):
        return rng_states.buf
    d = _l.dev(rng_states, name="rng_states", records=True)
    nbytes = d.size * d.dtype.itemsize

    class _H:
        pass
    h = _H()
    h.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(d.ptr), False), "version": 3, "strides": None}
    t = torch.as_tensor(h, device="cuda")
    t._lsb_keep = rng_states
    return t


def maybe_create_rng_states(n, seed=0, rng_states=None):
    """cli/simulate_pixels.py:92-104: keep the evolved states, append fresh ones when growing."""
    if rng_states is None:
        return create_xoroshiro128p_states(n, seed=seed)
    if n > len(rng_states):
        fresh = create_xoroshiro128p_states(n - len(rng_states), seed=seed)
        buf = torch.cat([_state_bytes(rng_states), fresh.buf])
        return _l.DeviceRecords(dtype=xoroshiro128p_dtype, n=n, buf=buf)
    return rng_states


def states_dev(rng_states, name="rng_states"):
    d = _l.dev(rng_states, write=True, name=name, records=True)
    if d.dtype.itemsize == 16:
        n = d.size
    elif d.dtype.itemsize == 8 and len(d.shape) == 2 and d.shape[1] == 2:
        n = d.shape[0]
    else:
        raise TypeError("%s: expected xoroshiro128p states ({s0:u8,s1:u8} records or an (n,2) 8-byte array)" % name)
    return d, n
