"""ctypes binding of ``libqcpinn_b200.so`` (the C-ABI in ``include/qcpinn_b200.h``).

There is deliberately no fallback: if the library is missing, or no CUDA device is usable, every
compute entry point raises.  ``load()`` only dlopen()s the library (works on a CPU-only box so the
ABI can be checked); ``require_cuda()`` is what the compute paths call.
"""

from __future__ import annotations

import ctypes
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libqcpinn_b200.so"

# symbol -> (restype, argtypes); must list every function declared in include/qcpinn_b200.h
_c_void_p = ctypes.c_void_p
_c_int = ctypes.c_int
_c_ll = ctypes.c_longlong
_dptr = ctypes.POINTER(ctypes.c_double)


class QcpMlp(ctypes.Structure):
    """``qcp_mlp_t``: the eight pre/post MLP tensors (device pointers)."""

    _fields_ = [(name, _c_void_p) for name in ("w1", "b1", "w2", "b2", "w3", "b3", "w4", "b4")]


SIGNATURES = {
    "qcp_last_error": (ctypes.c_char_p, []),
    "qcp_version": (_c_int, []),
    "qcp_device_count": (_c_int, []),
    "qcp_plan_create": (_c_int, [ctypes.POINTER(_c_void_p), _c_int, _c_int, _c_int, _c_int,
                                 ctypes.POINTER(ctypes.c_int32), _c_int, _dptr, _c_int, _c_int]),
    "qcp_plan_destroy": (_c_int, [_c_void_p]),
    "qcp_plan_num_features": (_c_int, [_c_void_p]),
    "qcp_plan_engine": (_c_int, [_c_void_p]),
    "qcp_plan_describe": (_c_int, [_c_void_p, ctypes.c_char_p, _c_int]),
    "qcp_plan_set_state_save": (_c_int, [_c_void_p, _c_int]),
    "qcp_plan_set_io_dtype": (_c_int, [_c_void_p, _c_int]),
    "qcp_prepare": (_c_int, [_c_void_p, _c_void_p, _c_void_p]),
    "qcp_feature_matrix": (_c_int, [_c_void_p, _dptr, _c_void_p]),
    "qcp_layer_forward": (_c_int, [_c_void_p, _c_void_p, _c_ll, _c_void_p, _c_void_p]),
    "qcp_layer_backward": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_void_p,
                                    _c_void_p, _c_void_p]),
    "qcp_solver_forward": (_c_int, [_c_void_p, ctypes.POINTER(QcpMlp), _c_void_p, _c_ll, _c_int,
                                    _dptr, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "qcp_solver_workspace_elems": (_c_ll, [_c_void_p, _c_ll, _c_int]),
    "qcp_solver_backward": (_c_int, [_c_void_p, ctypes.POINTER(QcpMlp), _c_void_p, _c_void_p,
                                     _c_void_p, _c_void_p, _c_ll, _c_int, _dptr, _c_void_p,
                                     ctypes.POINTER(QcpMlp), _c_void_p, _c_void_p, _c_void_p]),
    "qcp_solver_backward_streams": (_c_int, [_c_void_p, ctypes.POINTER(QcpMlp), _c_void_p, _c_void_p,
                                             _c_void_p, _c_ll, _c_void_p, ctypes.POINTER(QcpMlp),
                                             _c_void_p, _c_void_p]),
    "qcp_solver_backward_begin": (_c_int, [_c_void_p]),
    "qcp_solver_backward_add": (_c_int, [_c_void_p, ctypes.POINTER(QcpMlp), _c_void_p, _c_void_p,
                                         _c_void_p, _c_ll, _c_int, _dptr, _c_void_p, _c_void_p,
                                         _c_void_p]),
    "qcp_solver_backward_after_post": (_c_int, [_c_void_p, _c_void_p]),
    "qcp_solver_backward_finish": (_c_int, [_c_void_p, _c_void_p, ctypes.POINTER(QcpMlp),
                                            _c_void_p, _c_void_p]),
    "qcp_sample_targets": (_c_int, [_c_void_p, _c_ll, ctypes.POINTER(ctypes.c_float), _c_int,
                                    ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                    _c_void_p, _c_void_p, _c_void_p]),
    "qcp_mse_seed": (_c_int, [_c_void_p, _c_void_p, _c_ll, ctypes.c_double, _c_void_p, _c_void_p,
                              _c_void_p]),
    "qcp_clip_grads": (_c_int, [_c_void_p, _c_int, _c_int, ctypes.c_double, ctypes.c_double,
                                _c_void_p]),
    "qcp_pack_step": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p, ctypes.c_double, ctypes.c_double,
                               ctypes.c_double, _c_void_p, _c_void_p]),
    "qcp_plateau_step": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_int, _c_int,
                                  ctypes.c_double, ctypes.c_double, _c_ll, _c_ll, ctypes.c_double,
                                  ctypes.c_double, _c_void_p]),
    "qcp_peer_allreduce_floats": (_c_ll, [_c_int, _c_int]),
    "qcp_peer_allreduce_clip": (_c_int, [_c_void_p, _c_int, _c_int, ctypes.POINTER(_c_void_p), _c_int,
                                         _c_int, _c_void_p, ctypes.c_double, _c_void_p,
                                         ctypes.c_double, _c_void_p]),
    "qcp_debug_check_plan": (_c_int, [_c_int, _c_int, ctypes.POINTER(ctypes.c_int32), _c_int, _dptr,
                                      _c_int, _dptr, _c_int, _dptr, ctypes.POINTER(_c_int),
                                      ctypes.POINTER(_c_int)]),
    "qcp_bench_fma": (_c_int, [_c_int, _c_int, _dptr, _c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """dlopen the in-tree library and type every entry point.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("QCPINN_B200_LIB", LIB_PATH))
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: build it with `python {PKG_DIR.name}/build.py` "
            "(qcpinn_b200 has no CPU or PyTorch fallback)")
    lib = ctypes.CDLL(str(path))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError => ABI mismatch, surface it
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return load().qcp_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed: {last_error()}")


def require_cuda() -> ctypes.CDLL:
    """The product path: library present AND a CUDA device usable, else raise loudly."""
    lib = load()
    if lib.qcp_device_count() < 1:
        raise RuntimeError(
            "qcpinn_b200 needs a CUDA device (B200, sm_100a); none is usable and there is no CPU "
            "fallback")
    return lib
