"""ctypes mirror of ``include/larndsim_b200.h`` (POD structs + library loader).

Only interface declarations live here; the arithmetic is in ``csrc/*.cu``.
"""
import ctypes as C
import os

import numpy as np

LSB_MAX_TPC = 128

# dtype codes (enum lsb_dtype)
LSB_NONE, LSB_F32, LSB_F64, LSB_I32, LSB_U32, LSB_I64, LSB_U64 = range(7)
_DTYPE_CODE = {
    np.dtype("f4"): LSB_F32, np.dtype("f8"): LSB_F64, np.dtype("i4"): LSB_I32,
    np.dtype("u4"): LSB_U32, np.dtype("i8"): LSB_I64, np.dtype("u8"): LSB_U64,
}

# enum lsb_field, in header order
FIELDS = ("x", "y", "z", "x_start", "y_start", "z_start", "x_end", "y_end", "z_end",
          "t", "t_start", "t_end", "t0", "t0_start", "t0_end", "dEdx", "dE", "dx",
          "n_electrons", "n_photons", "long_diff", "tran_diff", "pixel_plane")
LSB_F_COUNT = len(FIELDS)


class TrackLayout(C.Structure):
    _fields_ = [("itemsize", C.c_int32),
                ("offset", C.c_int32 * LSB_F_COUNT),
                ("dtype", C.c_int32 * LSB_F_COUNT)]


class Consts(C.Structure):
    _fields_ = [
        ("box_alpha", C.c_double), ("box_beta", C.c_double), ("birks_ab", C.c_double),
        ("birks_kb", C.c_double), ("w_ion", C.c_double),
        ("mode_box", C.c_int32), ("mode_birks", C.c_int32),
        ("e_field", C.c_double), ("lar_density", C.c_double), ("v_drift", C.c_double),
        ("electron_lifetime", C.c_double), ("long_diff", C.c_double), ("tran_diff", C.c_double),
        ("time_sampling", C.c_double), ("time_padding", C.c_double), ("time_window", C.c_double),
        ("time_interval", C.c_double * 2),
        ("response_sampling", C.c_double), ("response_bin_size", C.c_double), ("pixel_pitch", C.c_double),
        ("n_time_ticks", C.c_int32), ("n_pixels", C.c_int32 * 2), ("n_tpc", C.c_int32),
        ("default_plane_index", C.c_int32), ("sampled_points", C.c_int32),
        ("max_neighbor_backtrack_distance", C.c_int32), ("pad0_", C.c_int32),
        ("discrimination_threshold", C.c_double), ("adc_hold_delay", C.c_double),
        ("adc_busy_delay", C.c_double), ("reset_cycles", C.c_double), ("clock_cycle", C.c_double),
        ("gain", C.c_double), ("buffer_risetime", C.c_double), ("v_cm", C.c_double),
        ("v_ref", C.c_double), ("v_pedestal", C.c_double), ("adc_counts", C.c_double),
        ("reset_noise_charge", C.c_double), ("uncorrelated_noise_charge", C.c_double),
        ("discriminator_noise", C.c_double),
        ("unit_e", C.c_double), ("unit_mV", C.c_double), ("unit_ns", C.c_double), ("unit_mus", C.c_double),
        ("w_ph", C.c_double), ("scint_prescale", C.c_double), ("light_tick_size", C.c_double),
        ("light_window", C.c_double * 2),
        ("singlet_fraction", C.c_double), ("tau_s", C.c_double), ("tau_t", C.c_double),
        ("light_response_time", C.c_double), ("light_oscillation_period", C.c_double),
        ("impulse_tick_size", C.c_double),
        ("sipm_response_model", C.c_int32), ("light_trig_mode", C.c_int32),
        ("enable_lut_smearing", C.c_int32), ("n_op_channel", C.c_int32),
        ("min_step_size", C.c_double), ("mc_truth_threshold", C.c_double),
        ("max_tracks_per_pixel", C.c_int32), ("mc_sample_multiplier", C.c_int32),
        ("max_adc_values", C.c_int32), ("pad1_", C.c_int32),
        ("tpc_borders", C.c_double * (LSB_MAX_TPC * 6)),
    ]


class LutLayout(C.Structure):
    _fields_ = [("itemsize", C.c_int32), ("off_vis", C.c_int32), ("off_t0", C.c_int32),
                ("off_t0_avg", C.c_int32), ("off_time_dist", C.c_int32), ("n_time_dist", C.c_int32),
                ("shape", C.c_int32 * 4)]


class LincLayout(C.Structure):
    _fields_ = [("itemsize", C.c_int32), ("off_n_photons_det", C.c_int32), ("off_t0_det", C.c_int32)]


class ChainResult(C.Structure):
    _fields_ = [("n_segments", C.c_int64), ("n_unique_pixels", C.c_int64), ("max_active", C.c_int64),
                ("max_neighbors", C.c_int64), ("n_ticks", C.c_int64), ("n_hits", C.c_int64),
                ("n_samples", C.c_int64), ("n_fma", C.c_int64), ("n_pairs", C.c_int64), ("n_groups", C.c_int64), ("n_edge", C.c_int64), ("n_irregular", C.c_int64),
                ("unique_pix", C.c_void_p), ("track_pixel_map", C.c_void_p), ("adc_list", C.c_void_p),
                ("adc_digit", C.c_void_p), ("adc_ticks_list", C.c_void_p), ("current_fractions", C.c_void_p),
                ("signals", C.c_void_p), ("pixels_signals", C.c_void_p),
                ("stage_ms", C.c_float * 12)]


def track_layout(dtype):
    """Resolve the record layout of a structured dtype by field NAME (SURVEY appendix A.1:
    offsets and dtypes vary per input file, so nothing is hard-coded)."""
    dtype = np.dtype(dtype)
    if dtype.fields is None:
        raise TypeError("tracks must be a NumPy structured array (got dtype %s)" % dtype)
    L = TrackLayout()
    L.itemsize = dtype.itemsize
    for i, name in enumerate(FIELDS):
        if name in dtype.fields:
            ft, off = dtype.fields[name][:2]
            code = _DTYPE_CODE.get(np.dtype(ft))
            if code is None:
                raise TypeError("field %r has unsupported dtype %s" % (name, ft))
            L.offset[i] = off
            L.dtype[i] = code
        else:
            L.offset[i] = -1
            L.dtype[i] = LSB_NONE
    return L


def lut_layout(lut_dtype, shape):
    dt = np.dtype(lut_dtype)
    LL = LutLayout()
    LL.itemsize = dt.itemsize
    def off(name):
        if dt.fields is None or name not in dt.fields:
            return -1
        ft, o = dt.fields[name][:2]
        base = np.dtype(ft).base
        if base != np.dtype("f4"):
            raise TypeError("light LUT field %r must be float32, got %s" % (name, base))
        return o
    LL.off_vis = off("vis")
    LL.off_t0 = off("t0")
    LL.off_t0_avg = off("t0_avg")
    LL.off_time_dist = off("time_dist")
    LL.n_time_dist = 0
    if LL.off_time_dist >= 0:
        sub = dt.fields["time_dist"][0]
        LL.n_time_dist = int(np.prod(sub.shape)) if sub.shape else 1
    for i in range(4):
        LL.shape[i] = int(shape[i])
    return LL


def linc_layout(dtype):
    dt = np.dtype(dtype)
    LI = LincLayout()
    LI.itemsize = dt.itemsize
    for name, attr in (("n_photons_det", "off_n_photons_det"), ("t0_det", "off_t0_det")):
        if name not in dt.fields:
            raise TypeError("light_incidence lacks field %r" % name)
        ft, o = dt.fields[name][:2]
        if np.dtype(ft) != np.dtype("f4"):
            raise TypeError("light_incidence field %r must be float32" % name)
        setattr(LI, attr, o)
    return LI


_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
#: LSB_LIB_PATH selects another build of the same library (kernel A/B variants, tools/ab_variants.sh)
LIB_PATH = os.environ.get("LSB_LIB_PATH") or os.path.join(_PKG_DIR, "csrc", "liblarndsim_b200.so")
_lib = None


class ExtensionMissing(RuntimeError):
    pass


def lib():
    """Load the CUDA extension.  There is NO CPU fallback: if the shared library is
    missing or cannot be loaded the product path fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ExtensionMissing(
                "CUDA extension %s not built; run `python -c 'import __graft_entry__ as g; g.build()'`"
                % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.lsb_last_error.restype = C.c_char_p
        _lib.lsb_launch_count.restype = C.c_int64
        _lib.lsb_unique_pixels_workspace_bytes.restype = C.c_int64
        _lib.lsb_tracks_current_mc_workspace_bytes.restype = C.c_int64
        _lib.lsb_tracks_current_mc_last_samples.restype = C.c_int64
        _lib.lsb_chain_create.restype = C.c_void_p
        if _lib.lsb_abi_version() != 2:
            raise ExtensionMissing("ABI version mismatch in %s" % LIB_PATH)
    return _lib


class LsbError(RuntimeError):
    pass


def check(status, what):
    if status != 0:
        msg = lib().lsb_last_error()
        raise LsbError("%s failed (status %d): %s" % (what, status, msg.decode() if msg else "?"))
