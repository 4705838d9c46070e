"""ctypes binding of ``libpsbax_b200.so`` (the C ABI in ``include/psbax_b200.h``).

There is no CPU fallback: if the shared object is missing or a call fails, a
``PsbaxError`` is raised.  Build the library with ``python -m psbax_b200.build``
(or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PSBAX_LIB") or os.path.join(_HERE, "libpsbax_b200.so")

KERNEL_IDS = {"rbf": 0, "matern12": 1, "matern32": 2, "matern52": 3, "linear": 4}
ENGINE_IDS = {"auto": 0, "simt": 1, "tensor": 2, "tensor_bf16": 3, "tensor_proj": 4}

# every symbol include/psbax_b200.h declares
EXPORTS = [
    "psbax_version", "psbax_last_error", "psbax_debug_trace_buffer", "psbax_device_sms", "psbax_pack_bytes", "psbax_pack",
    "psbax_workspace_bytes", "psbax_eval_values", "psbax_sumsq_scratch_bytes", "psbax_eval_sumsq", "psbax_eval_sumsq_groups", "psbax_eval_levelset", "psbax_eval_levelset_ex", "psbax_eval_topk",
    "psbax_topk_merge", "psbax_subset_select", "psbax_subset_select_ex", "psbax_eval_grad", "psbax_ascend", "psbax_mlp_forward",
]


class PsbaxError(RuntimeError):
    pass


class PathsStruct(C.Structure):
    """``psbax_paths_t``."""

    _fields_ = [
        ("d", C.c_int32), ("L", C.c_int32), ("G", C.c_int32), ("S", C.c_int32),
        ("n", C.c_int32), ("kernel", C.c_int32), ("engine", C.c_int32), ("reserved", C.c_int32),
        ("omega", C.c_void_p), ("bias", C.c_void_p), ("W", C.c_void_p),
        ("Xtrain", C.c_void_p), ("inv_ls", C.c_void_p), ("V", C.c_void_p),
        ("mean_const", C.c_float), ("y_mean", C.c_float), ("y_std", C.c_float),
        ("reserved_f", C.c_float),
    ]


class MeshStruct(C.Structure):
    """``psbax_mesh_t``."""

    _fields_ = [("d", C.c_int32), ("reserved", C.c_int32), ("len", C.c_int64 * 4), ("axis", C.c_void_p * 4)]


ACTIVATION_IDS = {"relu": 0, "leaky_relu": 1, "swish": 2, "sigmoid": 3, "tanh": 4}
MLP_MAX_LAYERS = 8


class MlpStruct(C.Structure):
    """``psbax_mlp_t``."""

    _fields_ = [("n_layers", C.c_int32), ("activation", C.c_int32), ("width", C.c_int32 * (MLP_MAX_LAYERS + 1)),
                ("reserved", C.c_int32), ("W", C.c_void_p * MLP_MAX_LAYERS), ("b", C.c_void_p * MLP_MAX_LAYERS),
                ("negative_slope", C.c_float), ("reserved_f", C.c_float)]


_lib: Optional[C.CDLL] = None


def _declare(lib: C.CDLL) -> None:
    P = C.POINTER(PathsStruct)
    vp, i32, i64, sz, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t, C.c_float
    lib.psbax_version.restype = C.c_int
    lib.psbax_version.argtypes = []
    lib.psbax_last_error.argtypes = [C.c_char_p, sz]
    lib.psbax_debug_trace_buffer.argtypes = [vp]
    lib.psbax_device_sms.argtypes = [C.POINTER(C.c_int)]
    lib.psbax_pack_bytes.argtypes = [P, C.POINTER(sz)]
    lib.psbax_pack.argtypes = [P, vp, sz, vp]
    lib.psbax_workspace_bytes.argtypes = [P, i64, i32, C.POINTER(sz)]
    lib.psbax_eval_values.argtypes = [P, vp, vp, i64, vp, i64, vp]
    lib.psbax_sumsq_scratch_bytes.argtypes = [P, i64, C.POINTER(C.c_size_t)]
    lib.psbax_eval_sumsq.argtypes = [P, vp, vp, i64, vp, vp, C.c_size_t, vp]
    lib.psbax_eval_sumsq_groups.argtypes = [P, vp, vp, i64, i32, vp, vp]
    lib.psbax_eval_levelset.argtypes = [P, vp, vp, i64, i64, f32, vp, i64, vp, vp, vp, vp, sz, vp]
    lib.psbax_eval_levelset_ex.argtypes = [P, vp, vp, C.POINTER(MeshStruct), i64, i64, f32, vp, i64, vp, vp, vp, vp, vp,
                                           sz, vp]
    lib.psbax_eval_topk.argtypes = [P, vp, vp, i64, i64, i32, vp, vp, vp, sz, vp]
    lib.psbax_topk_merge.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    lib.psbax_subset_select.argtypes = [vp, i64, vp, i64, i32, i32, i32, i32, vp, vp]
    lib.psbax_subset_select_ex.argtypes = [vp, i64, vp, i64, i64, i32, i32, i32, i32, i32, vp, vp]
    lib.psbax_eval_grad.argtypes = [P, vp, vp, i64, vp, vp, vp, vp]
    lib.psbax_ascend.argtypes = [P, vp, vp, i64, vp, i32, f32, f32, vp, vp, vp, vp]
    lib.psbax_mlp_forward.argtypes = [C.POINTER(MlpStruct), vp, i64, vp, vp]
    for name in EXPORTS:
        getattr(lib, name).restype = C.c_int


def load() -> C.CDLL:
    """Load the library (once).  Raises ``PsbaxError`` when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PsbaxError(
                f"{LIB_PATH} not found: the CUDA library has not been built "
                "(run `python -m psbax_b200.build`); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        _declare(lib)
        _lib = lib
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    load().psbax_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise PsbaxError(f"{what} failed (code {rc}): {last_error()}")
