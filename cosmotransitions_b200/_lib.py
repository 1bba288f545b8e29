"""ctypes binding of libctbounce.so (C ABI in include/ctbounce.h).

There is NO fallback: if the shared library is missing or no sm_100 device is present, every
compute entry point raises.  Loading the library itself needs no GPU (symbol checks run on CPU).
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# CTB_LIB: tuning experiments may point at a differently-compiled build of the same sources
LIB_PATH = os.environ.get("CTB_LIB") or os.path.join(_PKG, "libctbounce.so")

# keep in sync with include/ctbounce.h
ABI_VERSION = 12
V_TREE, V_ONELOOP, V_THERMAL, V_TOTAL = 1, 2, 4, 7
KERNEL_AUTO, KERNEL_THREAD, KERNEL_WARP = 0, 1, 2
EXACT_X, EXACT_DX, EXACT_THETA = 0, 1, 2
POT_POLY4, POT_MODEL1_LINE, POT_MODEL1F, POT_SPLINE, POT_THERMAL1F = 0, 1, 2, 3, 4
SPLINE_MAX_PIECES = 255
THERMAL_KINDS = (POT_MODEL1_LINE, POT_MODEL1F, POT_THERMAL1F)   # need the Jb / Jf tables
NPARAM = {POT_POLY4: 5, POT_MODEL1_LINE: 13, POT_MODEL1F: 6, POT_SPLINE: 1280, POT_THERMAL1F: 64}
(ST_CONVERGED, ST_XTOL, ST_STABLE, ST_NO_BARRIER, ST_RMAX, ST_DRMIN, ST_STEP_UNDERFLOW, ST_NONFINITE,
 ST_BRENT_SIGN, ST_FIRST_SHOT, ST_NOT_RUN, ST_MAXSTEPS) = range(12)

EXPORTS = ["ctb_abi_version", "ctb_strerror", "ctb_status_string", "ctb_device_ok", "ctb_bounce_batch",
           "ctb_bounce_setup", "ctb_wall_batch", "ctb_find_action", "ctb_potential_batch", "ctb_minimize_1d",
           "ctb_jspline", "ctb_jseries", "ctb_jexact", "ctb_vtot_model1", "ctb_vtot_model1_grid", "ctb_vtot_thermal1f", "ctb_potential", "ctb_fp64_peak", "ctb_sm_count",
           "ctb_copy_blocks", "ctb_host_register", "ctb_host_unregister", "ctb_trace_minima", "ctb_exact_solution"]
MODEL_MODEL1, MODEL_THERMAL1F = 0, 1
MIN_OK, MIN_NOT_CONVERGED, MIN_NOT_A_MINIMUM, MIN_JUMPED, MIN_NOT_TRACED = range(5)


class JTable(C.Structure):
    _fields_ = [("d_brk", C.c_void_p), ("d_coef", C.c_void_p), ("nint", C.c_int32), ("reserved", C.c_int32),
                ("xmin", C.c_double), ("xmax", C.c_double), ("loc_scale", C.c_double), ("loc_u0", C.c_double),
                ("loc_inv_du", C.c_double)]


class Opts(C.Structure):
    _fields_ = [("xtol", C.c_double), ("phitol", C.c_double), ("thinCutoff", C.c_double), ("rmin", C.c_double),
                ("rmax", C.c_double), ("phi_eps", C.c_double), ("xguess", C.c_double), ("npoints", C.c_int32),
                ("max_interior_pts", C.c_int32), ("alpha", C.c_int32), ("block_threads", C.c_int32),
                ("kernel", C.c_int32), ("reserved", C.c_int32), ("alpha_real", C.c_double)]


class WallOpts(C.Structure):
    _fields_ = [("Fguess", C.c_double), ("Ftol", C.c_double), ("phitol", C.c_double), ("rmin", C.c_double),
                ("rmax", C.c_double), ("phi0_rel", C.c_double), ("phi_eps", C.c_double), ("npoints", C.c_int32),
                ("reserved", C.c_int32)]


class MinOpts(C.Structure):
    _fields_ = [("x_eps", C.c_double), ("xtol", C.c_double), ("jump_tol", C.c_double), ("max_iter", C.c_int32),
                ("reserved", C.c_int32)]


class BounceIn(C.Structure):
    _fields_ = [("d_params", C.c_void_p), ("d_phi_abs", C.c_void_p), ("d_phi_meta", C.c_void_p),
                ("d_phi_bar", C.c_void_p), ("d_rscale", C.c_void_p), ("d_order", C.c_void_p), ("d_scratch", C.c_void_p),
                ("n", C.c_int64)]


class BounceOut(C.Structure):
    _fields_ = [("d_S", C.c_void_p), ("d_phi0", C.c_void_p), ("d_r0", C.c_void_p), ("d_rf", C.c_void_p),
                ("d_x", C.c_void_p), ("d_phi_bar", C.c_void_p), ("d_rscale", C.c_void_p), ("d_status", C.c_void_p),
                ("d_n_shots", C.c_void_p), ("d_n_rkck", C.c_void_p), ("d_n_points", C.c_void_p),
                ("d_prof_R", C.c_void_p), ("d_prof_Phi", C.c_void_p), ("d_prof_dPhi", C.c_void_p),
                ("prof_stride", C.c_int64)]


class CtbError(RuntimeError):
    pass


_lib = None


def load():
    """Load libctbounce.so; raises CtbError if it has not been built (python -m cosmotransitions_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CtbError("libctbounce.so is missing (%s): build it with `python -m cosmotransitions_b200.build`; "
                       "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise CtbError("libctbounce.so does not export %s" % name)
    lib.ctb_abi_version.restype = C.c_int
    lib.ctb_strerror.restype = C.c_char_p
    lib.ctb_strerror.argtypes = [C.c_int]
    lib.ctb_status_string.restype = C.c_char_p
    lib.ctb_status_string.argtypes = [C.c_int]
    lib.ctb_device_ok.restype = C.c_int
    lib.ctb_sm_count.restype = C.c_int
    lib.ctb_bounce_batch.restype = C.c_int
    lib.ctb_bounce_batch.argtypes = [C.c_int, C.POINTER(BounceIn), C.POINTER(Opts), C.POINTER(JTable),
                                     C.POINTER(JTable), C.POINTER(BounceOut), C.c_void_p]
    lib.ctb_bounce_setup.restype = C.c_int
    lib.ctb_bounce_setup.argtypes = [C.c_int, C.POINTER(BounceIn), C.POINTER(JTable), C.POINTER(JTable),
                                     C.POINTER(BounceOut), C.c_void_p, C.c_void_p]
    lib.ctb_jspline.restype = C.c_int
    lib.ctb_jspline.argtypes = [C.POINTER(JTable), C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.ctb_jseries.restype = C.c_int
    lib.ctb_jseries.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p, C.c_void_p,
                                C.c_int64, C.c_void_p]
    lib.ctb_jexact.restype = C.c_int
    lib.ctb_jexact.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.ctb_vtot_model1.restype = C.c_int
    lib.ctb_vtot_model1.argtypes = [C.POINTER(C.c_double), C.POINTER(JTable), C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.c_int, C.c_void_p]
    lib.ctb_vtot_model1_grid.restype = C.c_int
    lib.ctb_vtot_model1_grid.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(JTable), C.c_void_p,
                                         C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.ctb_vtot_thermal1f.restype = C.c_int
    lib.ctb_vtot_thermal1f.argtypes = [C.POINTER(C.c_double), C.c_double, C.POINTER(JTable), C.POINTER(JTable), C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    lib.ctb_potential.restype = C.c_int
    lib.ctb_potential.argtypes = [C.c_int, C.c_void_p, C.POINTER(JTable), C.POINTER(JTable), C.c_void_p, C.c_void_p,
                                  C.c_int64, C.c_void_p]
    lib.ctb_potential_batch.restype = C.c_int
    lib.ctb_potential_batch.argtypes = [C.c_int, C.c_void_p, C.POINTER(JTable), C.POINTER(JTable), C.c_void_p,
                                        C.c_void_p, C.c_int64, C.c_void_p]
    lib.ctb_minimize_1d.restype = C.c_int
    lib.ctb_minimize_1d.argtypes = [C.c_int, C.c_void_p, C.POINTER(JTable), C.POINTER(JTable), C.c_void_p, C.c_void_p,
                                    C.c_double, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ctb_wall_batch.restype = C.c_int
    lib.ctb_wall_batch.argtypes = [C.c_int, C.POINTER(BounceIn), C.POINTER(WallOpts), C.POINTER(JTable),
                                   C.POINTER(JTable), C.POINTER(BounceOut), C.c_void_p]
    lib.ctb_find_action.restype = C.c_int
    lib.ctb_find_action.argtypes = [C.c_int, C.c_void_p, C.POINTER(JTable), C.POINTER(JTable), C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_double, C.c_void_p,
                                    C.c_void_p]
    lib.ctb_trace_minima.restype = C.c_int
    lib.ctb_trace_minima.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.POINTER(JTable), C.POINTER(JTable), C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.POINTER(MinOpts), C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ctb_exact_solution.restype = C.c_int
    lib.ctb_exact_solution.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.ctb_copy_blocks.restype = C.c_int
    lib.ctb_copy_blocks.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]
    lib.ctb_host_register.restype = C.c_int
    lib.ctb_host_register.argtypes = [C.c_void_p, C.c_size_t]
    lib.ctb_host_unregister.restype = C.c_int
    lib.ctb_host_unregister.argtypes = [C.c_void_p]
    lib.ctb_fp64_peak.restype = C.c_int
    lib.ctb_fp64_peak.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_float), C.c_void_p]
    if lib.ctb_abi_version() != ABI_VERSION:
        raise CtbError("libctbounce.so ABI %d != expected %d; rebuild" % (lib.ctb_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        raise CtbError("%s failed: %s (%d)" % (what, load().ctb_strerror(code).decode(), code))


def status_string(status):
    return load().ctb_status_string(int(status)).decode()


def require_cuda():
    """Returns torch after verifying that a usable device exists; raises otherwise (no CPU fallback)."""
    import torch
    if not torch.cuda.is_available():
        raise CtbError("no CUDA device: cosmotransitions_b200 runs on B200 (sm_100a) only; there is no CPU fallback")
    return torch
