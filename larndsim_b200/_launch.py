"""Numba-style launch protocol on top of the C ABI.

The reference calls its kernels as ``module.kernel[griddim, blockdim](*args)`` with a mix of host
NumPy arrays (structured records included), CuPy arrays and Numba device arrays
(cli/simulate_pixels.py:732-1176).  ``Kernel`` reproduces that protocol:

* ``kernel[grid, block]`` (also ``[grid, block, stream, shmem]``) is accepted and the launch
  configuration is ignored -- the CUDA kernels choose their own;
* host NumPy arrays are staged into device memory and, like Numba does, copied back after the
  call (only the arguments a kernel writes are copied back);
* anything exposing ``__cuda_array_interface__`` (torch / CuPy / Numba device arrays) is used in
  place, zero-copy;
* constants are read from ``larndsim.consts`` (or this package's snapshot) at call time.

PyTorch is used for device buffers and the current stream only.
"""
import ctypes as C

import numpy as np
import torch

from . import _abi, consts as _consts


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("larndsim_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


class DeviceRecords:
    """A structured (record) array living in device memory: raw bytes + the NumPy dtype.
    Exposes ``__cuda_array_interface__`` so it can be handed to every kernel of this package."""

    def __init__(self, host=None, dtype=None, n=None, buf=None):
        _require_cuda()
        if host is not None:
            host = np.ascontiguousarray(host)
            self.dtype = host.dtype
            self.shape = host.shape
            self.buf = torch.from_numpy(host.reshape(-1).view(np.uint8).copy()).cuda()
        else:
            self.dtype = np.dtype(dtype)
            self.shape = (int(n),)
            self.buf = buf if buf is not None else torch.zeros(int(n) * self.dtype.itemsize, dtype=torch.uint8, device="cuda")

    @property
    def __cuda_array_interface__(self):
        return {"shape": tuple(self.shape), "typestr": "|V%d" % self.dtype.itemsize, "descr": self.dtype.descr,
                "data": (self.buf.data_ptr(), False), "version": 3, "strides": None}

    def copy_to_host(self):
        return self.buf.cpu().numpy().view(self.dtype).reshape(self.shape)

    def __len__(self):
        return self.shape[0]


class Dev:
    """One marshalled array argument."""
    __slots__ = ("ptr", "shape", "dtype", "keep", "_back")

    def __init__(self, ptr, shape, dtype, keep=None, back=None):
        self.ptr, self.shape, self.dtype, self.keep, self._back = ptr, tuple(shape), dtype, keep, back

    def sync_back(self):
        if self._back is not None:
            self._back()

    @property
    def c(self):
        return C.c_void_p(self.ptr)

    @property
    def size(self):
        n = 1
        for s in self.shape:
            n *= int(s)
        return n


_TORCH_OF = {np.dtype("f4"): torch.float32, np.dtype("f8"): torch.float64, np.dtype("i4"): torch.int32,
             np.dtype("i8"): torch.int64, np.dtype("u1"): torch.uint8, np.dtype("i2"): torch.int16}


def _cai_dtype(obj, cai):
    dt = getattr(obj, "dtype", None)
    if isinstance(dt, np.dtype):
        return dt
    if isinstance(dt, torch.dtype):
        return np.dtype(str(dt).replace("torch.", ""))
    descr = cai.get("descr")
    if descr and not (len(descr) == 1 and descr[0][0] == ""):
        return np.dtype([tuple(d) for d in descr])
    return np.dtype(cai["typestr"])


def _contig(shape, strides, itemsize):
    if strides is None:
        return True
    exp = itemsize
    for n, s in zip(reversed(shape), reversed(strides)):
        if n != 1 and s != exp:
            return False
        exp *= n
    return True


def dev(arg, want=None, write=False, name="array", records=False):
    """Marshal ``arg``.  ``want``: required NumPy dtype (converted through a temporary when it
    differs, the way Numba would have typed the access); ``write``: the kernel modifies it."""
    _require_cuda()
    if isinstance(arg, torch.Tensor) and not arg.is_cuda:
        arg = arg.numpy()
    if isinstance(arg, np.ndarray):
        src = arg
        if records or want is None:
            host = np.ascontiguousarray(src)
        else:
            host = np.ascontiguousarray(src, dtype=want)
        dt = host.dtype
        if host.size:
            t = torch.from_numpy(host.reshape(-1).view(np.uint8)).cuda()
        else:
            t = torch.empty(0, dtype=torch.uint8, device="cuda")

        def back(t=t, src=src, dt=dt):
            if src.size:
                src[...] = t.cpu().numpy().view(dt).reshape(src.shape)   # casts if the caller's dtype differs
        return Dev(t.data_ptr() if t.numel() else 0, src.shape, dt, keep=t, back=back if write else None)
    cai = getattr(arg, "__cuda_array_interface__", None)
    if cai is None:
        raise TypeError("%s: expected a NumPy array or an object with __cuda_array_interface__, got %r" % (name, type(arg)))
    dt = _cai_dtype(arg, cai)
    shape = tuple(cai["shape"])
    if not _contig(shape, cai.get("strides"), dt.itemsize):
        raise ValueError("%s: device arrays must be C-contiguous" % name)
    ptr = cai["data"][0] or 0
    if want is not None and not records and dt != np.dtype(want):
        if dt not in _TORCH_OF or np.dtype(want) not in _TORCH_OF:
            raise TypeError("%s: dtype %s where %s is required" % (name, dt, np.dtype(want)))
        orig = torch.as_tensor(arg, device="cuda")
        conv = orig.to(_TORCH_OF[np.dtype(want)]).contiguous()
        return Dev(conv.data_ptr(), shape, np.dtype(want), keep=(arg, conv),
                   back=(lambda: orig.copy_(conv)) if write else None)
    return Dev(ptr, shape, dt, keep=arg)


class Kernel:
    """``kernel[grid, block](*args)`` -- launch-configuration subscripts are accepted and ignored."""

    def __init__(self, fn, name=None):
        self._fn = fn
        self.__name__ = name or fn.__name__
        self.__doc__ = fn.__doc__
        self.py_func = fn

    def __getitem__(self, config):
        if not isinstance(config, tuple):
            config = (config,)
        if len(config) not in (1, 2, 3, 4):
            raise ValueError("launch configuration must be [griddim, blockdim(, stream(, sharedmem))]")
        return _Configured(self, config)

    def __call__(self, *args):
        return self._fn(*args)

    def __repr__(self):
        return "<larndsim_b200 kernel %s>" % self.__name__


class _Configured:
    """A kernel with a launch configuration attached (only ``tracks_current_mc`` looks at it: the
    reference indexes its RNG states with the grid size, detsim.py:273,324)."""

    def __init__(self, k, config):
        self._k, self.config = k, config

    def __call__(self, *args):
        if getattr(self._k._fn, "wants_config", False):
            return self._k._fn(*args, _config=self.config)
        return self._k._fn(*args)


def grid_threads(config, axis):
    """gridDim[axis] * blockDim[axis] of a Numba launch configuration."""
    def comp(v):
        if isinstance(v, (tuple, list)):
            return int(v[axis]) if axis < len(v) else 1
        return int(v) if axis == 0 else 1
    if config is None or len(config) < 2:
        return None
    return comp(config[0]) * comp(config[1])


def kernel(fn):
    return Kernel(fn)


def kernel_with_config(fn):
    fn.wants_config = True
    return Kernel(fn)


def snapshot():
    return _consts.snapshot()


def layout(d):
    return _abi.track_layout(d.dtype)


def check(status, what):
    _abi.check(status, what)


def lib():
    return _abi.lib()


def finish(*devs):
    """Copy host-staged outputs back (Numba copies back synchronously after the launch)."""
    for d in devs:
        d.sync_back()
