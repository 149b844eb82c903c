"""ctypes loader for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.h): importable from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs — never from denseflow_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


class Tvl1Params(C.Structure):
    _fields_ = [("tau", C.c_double), ("lambda_", C.c_double), ("theta", C.c_double), ("nscales", C.c_int),
                ("warps", C.c_int), ("epsilon", C.c_double), ("iterations", C.c_int), ("scale_step", C.c_double)]


class FarnParams(C.Structure):
    _fields_ = [("num_levels", C.c_int), ("pyr_scale", C.c_double), ("win_size", C.c_int), ("num_iters", C.c_int),
                ("poly_n", C.c_int), ("poly_sigma", C.c_double), ("resize_convention", C.c_int)]


RESIZE_CUDA = 0
RESIZE_HALF_PIXEL = 1


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("tvl1_oracle.c", "farneback_oracle.c", "quantise_oracle.c", "oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _SO


_lib = None
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8 = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_tvl1_default_params.argtypes = [C.POINTER(Tvl1Params)]
        L.orc_farn_default_params.argtypes = [C.POINTER(FarnParams)]
        L.orc_u8_to_f32.argtypes = [_u8, C.c_int, C.c_int, _fp]
        L.orc_resize_linear.argtypes = [_fp, C.c_int, C.c_int, _fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
        L.orc_centered_gradient.argtypes = [_fp, C.c_int, C.c_int, _fp, _fp]
        L.orc_tvl1_warp_backward.argtypes = [_fp] * 6 + [C.c_int, C.c_int] + [_fp] * 5
        L.orc_tvl1_estimate_u.argtypes = [_fp] * 10 + [C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
        L.orc_tvl1_estimate_u.restype = C.c_double
        L.orc_tvl1_estimate_dual.argtypes = [_fp] * 6 + [C.c_int, C.c_int, C.c_float]
        L.orc_tvl1_level_sizes.argtypes = [C.c_int, C.c_int, C.POINTER(Tvl1Params), _i32, _i32]
        L.orc_tvl1_level_sizes.restype = C.c_int
        L.orc_tvl1_calc.argtypes = [_u8, _u8, C.c_int, C.c_int, C.POINTER(Tvl1Params), _fp, C.c_void_p]
        L.orc_tvl1_calc.restype = C.c_int
        L.orc_farn_poly_constants.argtypes = [C.c_int, C.c_double, _fp, _fp, _fp, _fp]
        L.orc_farn_levels.argtypes = [C.c_int, C.c_int, C.POINTER(FarnParams), _i32, _i32, _i32, _f64]
        L.orc_farn_levels.restype = C.c_int
        L.orc_farn_gaussian_blur.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_double, _fp]
        L.orc_farn_poly_exp.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_double, _fp]
        L.orc_farn_update_matrices.argtypes = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, _fp]
        L.orc_farn_box_filter5.argtypes = [_fp, C.c_int, C.c_int, C.c_int, _fp]
        L.orc_farn_update_flow.argtypes = [_fp, C.c_int, C.c_int, _fp, _fp]
        L.orc_farn_calc.argtypes = [_u8, _u8, C.c_int, C.c_int, C.POINTER(FarnParams), _fp]
        L.orc_farn_calc.restype = C.c_int
        L.orc_convert_flow_to_image.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_double, C.c_double, _u8, _u8]
        L.orc_quantise_flow_xy.argtypes = [_fp, C.c_int, C.c_int, C.c_int, _u8, _u8]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def tvl1_params(**kw):
    p = Tvl1Params()
    lib().orc_tvl1_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, "lambda_" if k == "lambda" else k, v)
    return p


def farn_params(**kw):
    p = FarnParams()
    lib().orc_farn_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def tvl1_calc(a, b, params=None, return_iters=False):
    """a, b: uint8 [H,W].  Returns flow float32 [H,W,2] (and iteration log [nscales, warps])."""
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    h, w = a.shape
    p = params or tvl1_params()
    flow = np.empty((h, w, 2), np.float32)
    log = np.zeros((p.nscales, p.warps), np.int32)
    rc = lib().orc_tvl1_calc(a, b, w, h, C.byref(p), flow, log.ctypes.data)
    if rc != 0:
        raise RuntimeError("orc_tvl1_calc failed: %d" % rc)
    return (flow, log) if return_iters else flow


def tvl1_level_sizes(w, h, params=None):
    p = params or tvl1_params()
    ws = np.zeros(16, np.int32)
    hs = np.zeros(16, np.int32)
    n = lib().orc_tvl1_level_sizes(w, h, C.byref(p), ws, hs)
    return [(int(ws[i]), int(hs[i])) for i in range(n)]


def farn_calc(a, b, params=None):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    h, w = a.shape
    p = params or farn_params()
    flow = np.empty((h, w, 2), np.float32)
    rc = lib().orc_farn_calc(a, b, w, h, C.byref(p), flow)
    if rc != 0:
        raise RuntimeError("orc_farn_calc failed: %d" % rc)
    return flow


def farn_levels(w, h, params=None):
    p = params or farn_params()
    ws = np.zeros(16, np.int32); hs = np.zeros(16, np.int32); sm = np.zeros(16, np.int32)
    sg = np.zeros(16, np.float64)
    n = lib().orc_farn_levels(w, h, C.byref(p), ws, hs, sm, sg)
    return [(int(ws[i]), int(hs[i]), int(sm[i]), float(sg[i])) for i in range(n)]


def quantise(flow, bound):
    flow = np.ascontiguousarray(flow, np.float32)
    h, w = flow.shape[:2]
    qx = np.empty((h, w), np.uint8)
    qy = np.empty((h, w), np.uint8)
    lib().orc_quantise_flow_xy(flow, w, h, int(bound), qx, qy)
    return qx, qy
