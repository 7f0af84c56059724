"""ctypes binding of ``libmsau_b200.so`` (the C ABI declared in ``include/msau_b200.h``).

There is deliberately no fallback: if the shared library is missing or a call fails, the caller gets an
exception.  PyTorch tensors are only buffer carriers here (``data_ptr()``), never compute.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MSAU_LIB_PATH") or os.path.join(_HERE, "lib", "libmsau_b200.so")   # (env override: kernel A/B builds)

SYMBOLS = [
    "msau_last_error", "msau_version", "msau_launch_count", "msau_launch_count_add",
    "msau_plan_create", "msau_plan_destroy", "msau_param_count", "msau_param_info", "msau_plan_set_feature_table",
    "msau_workspace_bytes",
    "msau_forward", "msau_loss_backward", "msau_loss_backward_ex", "msau_plan_error_flags", "msau_clip_adam_step",
    "msau_optimizer_step", "msau_onehot_argmax", "msau_confusion_counts", "msau_plan_set_option",
    "msau_raster_geometry", "msau_raster_features", "msau_raster_labels",
    "msau_raster_kv_geometry", "msau_raster_kv", "msau_one_hot",
    "msau_rect_filter", "msau_class_equals", "msau_class_closing_row", "msau_ccl4", "msau_kv_select_components", "msau_kv_char_range",
    "msau_debug_layout", "msau_debug_tensor", "msau_debug_c3_prof", "msau_profile_enable", "msau_profile_report",
    "msau_set_option",
    "msau_attention_scratch_bytes", "msau_attention_forward", "msau_attention_backward",
]


class MsauConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("channels", "n_class", "scale_space_num", "res_depth", "feat_root",
                                       "filter_size", "pool_size", "num_blocks")]


class MsauLossSpec(C.Structure):
    _fields_ = [("mode", C.c_int), ("weight_main", C.c_float), ("weight_aux", C.c_float), ("h_class_weights", C.POINTER(C.c_float))]


class MsauError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the library; raises if it has not been built (``make`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MsauError(f"{LIB_PATH} not found: build the CUDA extension first (make, or __graft_entry__.build()). "
                        "msau_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t
    L.msau_last_error.restype = C.c_char_p
    L.msau_last_error.argtypes = []
    L.msau_version.restype = i32
    L.msau_launch_count.restype = i64
    L.msau_launch_count_add.argtypes = [i64]
    L.msau_launch_count_add.restype = None
    L.msau_plan_create.argtypes = [C.POINTER(MsauConfig), i32, i32, i32, C.POINTER(vp)]
    L.msau_plan_destroy.argtypes = [vp]
    L.msau_plan_destroy.restype = None
    L.msau_param_count.argtypes = [vp]
    L.msau_param_count.restype = i64
    L.msau_param_info.argtypes = [vp, i32, C.POINTER(i64), C.POINTER(i64)]
    L.msau_plan_set_feature_table.argtypes = [vp, vp, i32]
    L.msau_workspace_bytes.argtypes = [vp, i32, C.POINTER(sz)]
    L.msau_forward.argtypes = [vp, vp, i32, vp, vp, sz, i32, vp, vp, vp, vp, vp]
    L.msau_loss_backward.argtypes = [vp, vp, i32, vp, i32, f32, vp, sz, vp, vp, vp]
    L.msau_clip_adam_step.argtypes = [vp, vp, vp, vp, i64, i32, f32, f32, f32, f32, f32, vp, vp, vp]
    L.msau_loss_backward_ex.argtypes = [vp, vp, i32, vp, vp, i32, C.POINTER(MsauLossSpec), f32, vp, sz, vp, vp, vp, vp, vp]
    L.msau_plan_error_flags.argtypes = [vp, C.POINTER(i32)]
    L.msau_optimizer_step.argtypes = [i32, vp, vp, vp, vp, i64, i32, vp, f32, f32, f32, f32, f32, f32, vp, vp, vp]
    L.msau_onehot_argmax.argtypes = [vp, i32, i32, i32, i64, i32, vp, vp]
    L.msau_confusion_counts.argtypes = [vp, vp, i32, i64, i32, vp, vp]
    L.msau_plan_set_option.argtypes = [vp, C.c_char_p, i32]
    L.msau_raster_geometry.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, vp]
    L.msau_raster_features.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, i32, vp, i32, i32, i32, i32, vp, vp, vp]
    L.msau_raster_labels.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp, i32, i32, vp, vp, vp]
    L.msau_raster_kv_geometry.argtypes = [vp, vp, i32, vp, vp]
    L.msau_raster_kv.argtypes = [vp, vp, i32, i32, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
    L.msau_one_hot.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp]
    L.msau_rect_filter.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]
    L.msau_class_equals.argtypes = [vp, vp, i64, i32, vp]
    L.msau_class_closing_row.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    L.msau_ccl4.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, vp, vp]
    L.msau_kv_select_components.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32, i32, vp, vp, vp]
    L.msau_kv_char_range.argtypes = [vp, vp, i32, i32, vp, i32, vp, vp]
    L.msau_debug_c3_prof.argtypes = [C.POINTER(C.c_ulonglong)]
    L.msau_debug_layout.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    L.msau_debug_tensor.argtypes = [vp, i32, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.msau_set_option.argtypes = [C.c_char_p, i32]
    L.msau_attention_scratch_bytes.argtypes = [i32, i32, i32]
    L.msau_attention_scratch_bytes.restype = sz
    L.msau_attention_forward.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, sz, vp]
    L.msau_attention_backward.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, sz, vp]
    L.msau_profile_enable.argtypes = [i32]
    L.msau_profile_report.argtypes = [C.c_char_p, sz]
    for name in SYMBOLS:
        getattr(L, name)   # every symbol the header declares must be exported
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().msau_last_error()
        raise MsauError(f"msau_b200 error {rc}: {msg.decode() if msg else '?'}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (or 0 for None)."""
    return 0 if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(lib().msau_launch_count())


def profile_enable(on: bool) -> None:
    check(lib().msau_profile_enable(int(on)))


def profile_report() -> dict:
    import json
    buf = C.create_string_buffer(1 << 16)
    check(lib().msau_profile_report(buf, len(buf)))
    return json.loads(buf.value.decode())


def set_option(name: str, value: int) -> None:
    check(lib().msau_set_option(name.encode(), int(value)))
