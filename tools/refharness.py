"""Reference harness -- TEST INFRASTRUCTURE, runs ONLY in the build container.

Imports the unmodified reference package from /root/reference (read-only) and
offers two ways of executing its device kernels on the CPU:

* ``sim``  -- Numba's CUDA simulator (``NUMBA_ENABLE_CUDASIM=1``), i.e. the
  reference's own CI path.  Pure Python, seconds per segment; NumPy scalar
  arithmetic (f4 record fields are computed in float32 under NEP-50).
* ``njit`` -- the reference kernel's *own source text* is fetched with
  ``inspect`` at run time, ``cuda.grid`` / ``cuda.gridsize`` / ``cuda.atomic``
  are rewritten to loop indices, and the result is compiled by ``numba.njit``
  for the host.  Numba's type inference is shared between the CPU and CUDA
  targets, so this reproduces the *compiled* reference semantics (float64
  promotion through Python-float globals, float32 where both operands are
  float32, int64 ``round``) at JIT speed.  Threads are executed in grid order
  x-major / z-fastest with one thread at a time, i.e. the order of the
  simulator launched with 1-thread blocks.

Nothing from the reference is copied into this repository; the transformation
happens in memory.  Used by ``tools/gen_golden.py`` (fixtures under
``tests/golden``) and by ``tools/make_config_snapshots.py``.
"""
import importlib
import inspect
import os
import re
import sys
import textwrap
import types

REF_ROOT = os.environ.get("LARNDSIM_REFERENCE", "/root/reference")


def _install_shims():
    import numpy as np
    if "cupy" not in sys.modules:
        cp = types.ModuleType("cupy")
        cp.__dict__.update({k: getattr(np, k) for k in dir(np) if not k.startswith("__")})
        cp.get_array_module = lambda *a: np
        cp.asnumpy = lambda a: np.asarray(a)
        cuda = types.ModuleType("cupy.cuda")
        nvtx = types.ModuleType("cupy.cuda.nvtx")
        nvtx.RangePush = lambda *a, **k: None
        nvtx.RangePop = lambda *a, **k: None
        cuda.nvtx = nvtx
        cp.cuda = cuda
        sys.modules["cupy"] = cp
        sys.modules["cupy.cuda"] = cuda
        sys.modules["cupy.cuda.nvtx"] = nvtx
    for name in ("h5py",):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    if "larpix" not in sys.modules:
        try:
            importlib.import_module("larpix")
        except ImportError:
            lp = types.ModuleType("larpix")
            for sub, names in (("packet", ["Packet_v2", "TimestampPacket", "TriggerPacket",
                                           "SyncPacket", "PacketCollection"]),
                               ("key", ["Key"]), ("format", ["hdf5format"])):
                m = types.ModuleType("larpix." + sub)
                for n in names:
                    setattr(m, n, type(n, (), {}))
                setattr(lp, sub, m)
                sys.modules["larpix." + sub] = m
            sys.modules["larpix"] = lp


def load_reference(simulator=False):
    """Import ``larndsim`` from the read-only reference tree."""
    if simulator:
        os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
        os.environ["NUMBA_DISABLE_JIT"] = "1"
    _install_shims()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import larndsim  # noqa: F401
    from larndsim import consts  # noqa: F401
    mods = {}
    for name in ("quenching", "drifting", "pixels_from_track", "detsim", "fee",
                 "lightLUT", "light_sim"):
        mods[name] = importlib.import_module("larndsim." + name)
    mods["consts"] = consts
    return mods


def ref_path(*parts):
    return os.path.join(REF_ROOT, "larndsim", *parts)


def load_properties(detprop="module0.yaml", layout="multi_tile_layout-2.3.16.yaml",
                    simprop="singles_sim.yaml", i_module=-1):
    from larndsim import consts
    det = ref_path("detector_properties", detprop)
    pix = ref_path("pixel_layouts", layout)
    sim = ref_path("simulation_properties", simprop)
    if i_module < 0:
        consts.load_properties(det, pix, sim)
    else:
        consts.detector.set_detector_properties(det, pix, i_module)
        consts.light.set_light_properties(det)
        consts.sim.set_simulation_properties(sim)
    return consts


# --------------------------------------------------------------------------
# kernel source -> host njit transformation
# --------------------------------------------------------------------------
_ATOMIC_MAX = re.compile(r"cuda\.atomic\.max\(\s*(\w+)\s*,\s*0\s*,\s*(.+)\)\s*$")
_ATOMIC_ADD = re.compile(r"cuda\.atomic\.add\(")


def _kernel_source(kernel):
    fn = getattr(kernel, "py_func", None) or getattr(kernel, "fn", None) or kernel
    src = textwrap.dedent(inspect.getsource(fn))
    return fn, src


def _transform(src, name):
    lines = src.split("\n")
    out = []
    ndim = None
    i = 0
    # drop decorators
    while lines[i].lstrip().startswith("@"):
        i += 1
    header = lines[i]
    m = re.match(r"def\s+(\w+)\((.*)", header)
    assert m, header
    # the signature may span several lines
    sig = m.group(2)
    while "):" not in sig:
        i += 1
        sig += " " + lines[i].strip()
    args = sig[: sig.index("):")]
    body = lines[i + 1:]
    text = "\n".join(body)
    g = re.search(r"cuda\.grid\((\d)\)", text)
    ndim = int(g.group(1))
    idx = ", ".join("_g%d" % d for d in range(ndim))
    nn = ", ".join("_n%d" % d for d in range(ndim))
    if ndim == 1:
        text = re.sub(r"cuda\.grid\(1\)", "_g0", text)
        text = re.sub(r"cuda\.gridsize\(1\)", "_n0", text)
    else:
        text = re.sub(r"cuda\.grid\(%d\)" % ndim, "(" + idx + ")", text)
        text = re.sub(r"cuda\.gridsize\(%d\)" % ndim, "(" + nn + ")", text)
    new_lines = []
    for ln in text.split("\n"):
        s = ln.strip()
        mm = _ATOMIC_MAX.match(s)
        if mm:
            indent = ln[: len(ln) - len(ln.lstrip())]
            new_lines.append("%s%s[0] = max(%s[0], %s)" % (indent, mm.group(1), mm.group(1), mm.group(2)))
            continue
        new_lines.append(ln)
    text = "\n".join(new_lines)
    # cuda.atomic.add(arr, idx, val)  ->  arr[idx] += val   (multi-line aware)
    while True:
        m = _ATOMIC_ADD.search(text)
        if not m:
            break
        start = m.start()
        j = m.end()
        depth = 1
        parts = []
        cur = ""
        while depth:
            c = text[j]
            if c in "([":
                depth += 1
            elif c in ")]":
                depth -= 1
                if depth == 0:
                    break
            if c == "," and depth == 1:
                parts.append(cur)
                cur = ""
            else:
                cur += c
            j += 1
        parts.append(cur)
        arr, index, val = [" ".join(p.split()) for p in parts]
        text = text[:start] + "%s[%s] += %s" % (arr, index, val) + text[j + 1:]
    text = text.replace("cuda.random.", "")
    thread = "def %s__thread(%s, %s, %s):\n%s\n" % (name, idx, nn, args, text)
    loops = ""
    ind = "    "
    for d in range(ndim):
        loops += "%sfor _g%d in range(_n%d):\n" % (ind * (d + 1), d, d)
    call = "%s%s__thread(%s, %s, %s)\n" % (ind * (ndim + 1), name, idx, nn,
                                          ", ".join(a.split("=")[0].strip() for a in args.split(",")))
    grid = "def %s__grid(%s, %s):\n%s%s" % (name, nn, args, loops, call)
    return ndim, thread, grid


_cache = {}


def host_kernel(module, name):
    """Return ``run(grid_shape, *args)`` executing the reference kernel
    ``module.name`` on the host, compiled from its own source by ``numba.njit``.
    Must be called with the real JIT enabled (not under the simulator)."""
    key = (module.__name__, name, id(module))
    if key in _cache:
        return _cache[key]
    import numba as nb
    from numba.cuda import random as nbrandom
    kernel = getattr(module, name)
    fn, src = _kernel_source(kernel)
    ndim, thread_src, grid_src = _transform(src, name)
    ns = dict(module.__dict__)
    ns["xoroshiro128p_uniform_float32"] = nbrandom.xoroshiro128p_uniform_float32
    ns["xoroshiro128p_normal_float32"] = nbrandom.xoroshiro128p_normal_float32
    # device functions declared with cuda.jit(device=True) need host versions
    # error_model="numpy": no Python exceptions on x/0 etc., like the CUDA target (quenching.py:33 relies
    # on log(0.93)/0 = -inf).  @nb.njit helpers of the reference are re-jitted the same way.
    jit = nb.njit(error_model="numpy")
    for k, v in list(ns.items()):
        pf = getattr(v, "py_func", None)
        if pf is None or not callable(pf):
            continue
        tmod = type(v).__module__
        if tmod.startswith("numba.cuda") or tmod.startswith("numba.core.dispatcher") or tmod.startswith("numba.core.registry"):
            if getattr(pf, "__module__", "").startswith("larndsim"):
                ns[k] = jit(_rebuild(pf, ns))
    exec(compile(thread_src, "<ref:%s thread>" % name, "exec"), ns)
    ns[name + "__thread"] = jit(ns[name + "__thread"])
    exec(compile(grid_src, "<ref:%s grid>" % name, "exec"), ns)
    gridfn = jit(ns[name + "__grid"])

    def run(grid_shape, *args):
        if isinstance(grid_shape, int):
            grid_shape = (grid_shape,)
        assert len(grid_shape) == ndim, (grid_shape, ndim)
        gridfn(*[int(g) for g in grid_shape], *args)

    run.thread_source = thread_src
    _cache[key] = run
    return run


def _rebuild(pyfunc, ns):
    """Re-create a cuda device function as a plain function living in ``ns`` so
    that njit resolves its globals against the host versions."""
    src = textwrap.dedent(inspect.getsource(pyfunc))
    lines = src.split("\n")
    i = 0
    while lines[i].lstrip().startswith("@"):
        i += 1
    src = "\n".join(lines[i:])
    loc = {}
    exec(compile(src, "<ref:%s>" % pyfunc.__name__, "exec"), ns, loc)
    f = loc[pyfunc.__name__]
    return f
