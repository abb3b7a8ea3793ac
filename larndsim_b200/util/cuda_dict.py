"""Drop-in for ``larndsim.util.cuda_dict.CudaDict`` (reference: larndsim/util/cuda_dict.py:1-230): a static
integer-key table living on the GPU, used for the per-pixel thresholds and gains
(cli/simulate_pixels.py:439-449, 1080-1100: ``pixel_thresholds_lut[unique_pix.ravel()]``).

Same interface -- ``CudaDict(default, tpb, bpg)``, ``cd[keys] = values`` (once), ``cd[keys]``, ``contains``,
``keys / values / items``, ``len``, ``load / save`` (the reference's ``.npz`` files: keys, values, default).  The
reference is an open-addressing hash table filled with atomics; the keys never change after the fill, so this
version keeps them sorted and answers a query with one binary search (``lsb_table_lookup``): identical results,
no probing sequences.  Lookups return torch CUDA tensors (``__cuda_array_interface__``: usable by every kernel
of this package and by CuPy)."""
import ctypes as C

import numpy as np
import torch

from .. import _launch as _l


def _to_cuda(a, dtype=None):
    if isinstance(a, torch.Tensor):
        t = a.cuda()
    elif isinstance(a, np.ndarray) or np.isscalar(a) or isinstance(a, (list, tuple)):
        t = torch.from_numpy(np.ascontiguousarray(np.atleast_1d(a))).cuda()
    else:
        t = torch.as_tensor(a, device="cuda")
    return t.to(dtype).contiguous() if dtype is not None else t.contiguous()


class CudaDict(object):
    def __init__(self, default, tpb=256, bpg=1):
        self.tpb, self.bpg = tpb, bpg                       # accepted for compatibility; the lookup sizes its own grid
        self.default = _to_cuda(default).reshape(-1)[:1]
        self._keys = torch.empty(0, dtype=torch.int32, device="cuda")
        self._values = torch.empty(0, dtype=self.default.dtype, device="cuda")

    def __len__(self):
        return int(self._keys.numel())

    def keys(self):
        return self._keys

    def values(self):
        return self._values

    def items(self):
        return self.keys(), self.values()

    def __setitem__(self, key, value):
        if len(self) != 0:
            raise NotImplementedError('Trying to update CudaDict, not yet supported')
        k = _to_cuda(key, torch.int32).reshape(-1)
        v = _to_cuda(value).reshape(-1)
        if v.element_size() not in (4, 8):
            raise TypeError("CudaDict values must be 4 or 8 bytes wide")
        if k.numel() != v.numel():
            raise ValueError("keys and values differ in length")
        order = torch.argsort(k, stable=True)
        self._keys, self._values = k[order].contiguous(), v[order].contiguous()
        if self.default.dtype != self._values.dtype:
            self.default = self.default.to(self._values.dtype)

    def _lookup(self, key, want_values, want_exists):
        q = _to_cuda(key, torch.int32).reshape(-1)
        n = q.numel()
        out = torch.empty(n, dtype=self._values.dtype, device="cuda") if want_values else None
        ex = torch.empty(n, dtype=torch.uint8, device="cuda") if want_exists else None
        dflt = self.default.cpu().numpy()
        _l.check(_l.lib().lsb_table_lookup(C.c_void_p(self._keys.data_ptr()), C.c_void_p(self._values.data_ptr()), C.c_int64(len(self)),
                                           C.c_int32(self._values.element_size()), C.c_void_p(q.data_ptr()), C.c_int64(n),
                                           dflt.ctypes.data_as(C.c_void_p), C.c_void_p(out.data_ptr()) if out is not None else None,
                                           C.c_void_p(ex.data_ptr()) if ex is not None else None, _l.stream()), "table_lookup")
        return out, ex

    def __getitem__(self, key):
        return self._lookup(key, True, False)[0]

    def contains(self, key):
        return self._lookup(key, False, True)[1].bool()

    def __delitem__(self, key):
        raise NotImplementedError("CudaDict is static once filled")

    @staticmethod
    def load(filename, tpb=256):
        data = np.load(filename)
        cd = CudaDict(default=data["default"], tpb=tpb, bpg=max(1, -(-len(data["keys"]) // tpb)))
        cd[data["keys"]] = data["values"]
        return cd

    @staticmethod
    def save(filename, cdict):
        np.savez_compressed(filename, keys=cdict.keys().cpu().numpy(), values=cdict.values().cpu().numpy(),
                            default=cdict.default.cpu().numpy())
