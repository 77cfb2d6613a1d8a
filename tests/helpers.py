"""Shared test plumbing: the oracle libraries (CHECKERS -- only tests/, smoke() and
bench.py's reference arm may load anything under oracle/), tolerance metric, device
buffers."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from bcnn_b200 import capi

ROOT = Path(__file__).resolve().parents[1]
ORACLE_DIR = ROOT / "oracle"
ORACLE_SO = ORACLE_DIR / "liboracle.so"
REF_SO = ORACLE_DIR / "_ref" / "libbcnn_ref.so"
GOLDEN = ROOT / "tests" / "golden"

FP32_TOL = 1e-5   # FP32 SIMT verification path (BASELINE.json north_star)
TC_TOL = 2e-2     # BF16 / TF32 tensor-core path


def rel_err(a: np.ndarray, b: np.ndarray):
    """(max-abs error / max-abs(b), L2 error / L2(b)): the normalised per-tensor metric of
    SURVEY.md section 8d (the reference itself sits 3e-7..9e-7 from an exact evaluation in
    this metric, but up to 1e-3 element-wise on small outputs)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    assert a.shape == b.shape, (a.shape, b.shape)
    scale_max = max(np.abs(b).max(initial=0.0), 1e-30)
    scale_l2 = max(np.linalg.norm(b), 1e-30)
    return float(np.abs(a - b).max(initial=0.0) / scale_max), float(np.linalg.norm(a - b) / scale_l2)


def assert_close(a, b, tol=FP32_TOL, what=""):
    assert np.all(np.isfinite(np.asarray(a))), f"{what}: non-finite values"
    e_max, e_l2 = rel_err(a, b)
    assert e_max <= tol and e_l2 <= tol, f"{what}: max-rel {e_max:.3e}, l2-rel {e_l2:.3e} > {tol}"


def assert_close_grad(a, b, l2_tol, what="", el_tol=2e-2, outlier_frac=5e-3):
    """Gradients behind (leaky-)ReLU masks: a pre-activation within rounding noise of zero changes
    sign between two correct FP32 evaluations, which multiplies that element's gradient by 1 vs the
    negative slope, and batch norm spreads the difference thinly over its channel. The max-abs
    metric then measures one flipped element; what is held instead: the L2 error of the tensor, and
    that at most `outlier_frac` of the elements are further than el_tol * max|b| from the reference."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    assert np.all(np.isfinite(a)), f"{what}: non-finite values"
    e_max, e_l2 = rel_err(a, b)
    assert e_l2 <= l2_tol, f"{what}: l2-rel {e_l2:.3e} > {l2_tol} (max-rel {e_max:.3e})"
    far = float(np.mean(np.abs(a - b) > el_tol * max(np.abs(b).max(initial=0.0), 1e-30)))
    assert far <= outlier_frac, f"{what}: {far:.2e} of the elements are off by more than {el_tol} of max"


# ---- oracle restatement (oracle/bcnn_oracle.c) ----------------------------------

_ORACLE = None


def oracle() -> C.CDLL:
    global _ORACLE
    if _ORACLE is not None:
        return _ORACLE
    if not ORACLE_SO.exists():
        subprocess.run(["make", "-C", str(ORACLE_DIR), "oracle"], check=True, capture_output=True)
    lib = C.CDLL(str(ORACLE_SO))
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    sigs = {
        "orc_vsum": (f, [i, vp]),
        "orc_dot": (f, [i, vp, vp]),
        "orc_shiftdot": (f, [i, vp, f, vp, f]),
        "orc_conv_out_dim": (i, [i, i, i, i]),
        "orc_conv_forward": (None, [vp, vp, vp] + [i] * 9),
        "orc_conv_backward": (None, [vp, vp, vp, vp, vp] + [i] * 9),
        "orc_add_bias": (None, [vp, vp, i, i, i]),
        "orc_grad_bias": (None, [vp, vp, i, i, i]),
        "orc_bn_forward": (None, [vp, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, i]),
        "orc_bn_backward": (None, [vp, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "orc_activation_forward": (None, [vp, i, vp, i, i, i]),
        "orc_activation_backward": (None, [vp, vp, i, vp, vp, i, i, i]),
        "orc_maxpool_out_dim": (i, [i, i, i, i]),
        "orc_maxpool_forward": (None, [vp, vp, vp] + [i] * 8),
        "orc_maxpool_backward": (None, [vp, vp, vp, i]),
        "orc_avgpool_forward": (None, [vp, vp, i, i, i]),
        "orc_avgpool_backward": (None, [vp, vp, i, i, i]),
        "orc_depthwise_forward": (None, [vp, vp, vp] + [i] * 7),
        "orc_depthwise_backward": (None, [vp, vp, vp, vp, vp] + [i] * 7),
        "orc_sgd_update": (None, [vp, vp, vp, vp, i, i, i, f, f, f]),
        "orc_adam_update": (None, [vp] * 6 + [i] * 4 + [f] * 5),
        "orc_yolo_forward": (None, [vp, vp, i, i, i, i, i]),
        "orc_fc_forward": (None, [vp, vp, vp, vp, i, i, i, i]),
        "orc_fc_backward": (None, [vp, vp, vp, vp, vp, vp, i, i, i]),
        "orc_softmax_forward": (None, [vp, vp, i, i, i]),
        "orc_eltwise_add": (None, [vp, vp, vp, i]),
        "orc_concat_forward": (None, [vp, vp, i, i, i, i]),
        "orc_concat_backward": (None, [vp, vp, i, i, i, i]),
        "orc_upsample_forward": (None, [vp, vp, i, i, i, i, i]),
        "orc_upsample_backward": (None, [vp, vp, i, i, i, i, i]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _ORACLE = lib
    return lib


def p(a: np.ndarray):
    """Raw pointer of a contiguous numpy array (None passes NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ---- the compiled reference (oracle/_ref) -----------------------------------------

_REF = None


def ref_available() -> bool:
    return REF_SO.exists()


def ref_lib() -> C.CDLL:
    """oracle/_ref/libbcnn_ref.so: the unmodified reference CPU path + ref_shim.c."""
    global _REF
    if _REF is None:
        import os
        lib = C.CDLL(str(REF_SO), mode=C.RTLD_LOCAL | getattr(os, "RTLD_DEEPBIND", 0))
        capi.bind_bcnn_api(lib, capi.TensorCPU)
        vp, i = C.c_void_p, C.c_int
        for name, (res, args) in {
            "bcnn_b200_num_nodes": (i, [vp]), "bcnn_b200_num_tensors": (i, [vp]),
            "bcnn_b200_node_type": (i, [vp, i]), "bcnn_b200_node_src": (i, [vp, i, i]),
            "bcnn_b200_node_dst": (i, [vp, i, i]),
            "bcnn_b200_maxpool_indexes": (i, [vp, i, vp]),
            "bcnn_b200_bn_saved_stats": (i, [vp, i, vp, vp]),
        }.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _REF = lib
    return _REF


def ref_net(mode=capi.MODE_TRAIN, threads=1) -> capi.Net:
    net = capi.Net(mode=mode, lib=ref_lib(), flavour="cpu")
    net.lib.bcnn_set_num_threads(net.handle, threads, None)
    return net


def dev(a: np.ndarray) -> capi.DeviceBuffer:
    return capi.DeviceBuffer(np.ascontiguousarray(a))


def dev_zeros(n_elems: int, itemsize: int = 4) -> capi.DeviceBuffer:
    return capi.DeviceBuffer(nbytes=n_elems * itemsize)


def check(err: int):
    if err:
        raise RuntimeError(capi.b200().bcnn_b200_error_string(err).decode())
