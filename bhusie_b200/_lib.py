"""ctypes binding of libbhray.so (include/bh_abi.h).  Fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None
ABI_VERSION = 2


class BhError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"libbhray error {code}: {text}")
        self.code = code


class PassStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("ray_steps", "px_traced", "px_copied", "px_interp", "node_visits",
                                          "tri_tests", "tex_samples", "rk_reject", "stack_overflow")]

    def as_dict(self) -> dict:
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class ModelInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("point_count", "normal_count", "triangle_count", "nodes_used",
                                         "max_depth", "leaf_count", "max_leaf_size")]

    def as_dict(self) -> dict:
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_VP, _U32, _I32 = C.c_void_p, C.c_uint32, C.c_int32


class FrameMultiDesc(C.Structure):
    _fields_ = [("base_width", C.c_uint32), ("base_height", C.c_uint32), ("levels", C.c_uint32), ("multiplier", C.c_uint32),
                ("band_rows", C.c_uint32), ("sky_format", C.c_int32)]


# every symbol include/bh_abi.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "bh_abi_version": (C.c_int, []),
    "bh_last_error": (C.c_char_p, []),
    "bh_ctx_create": (C.c_int, [C.c_int, C.POINTER(_VP)]),
    "bh_ctx_destroy": (None, [_VP]),
    "bh_ctx_set_numeric_mode": (C.c_int, [_VP, C.c_int]),
    "bh_ctx_get_numeric_mode": (C.c_int, [_VP]),
    "bh_ctx_set_texture": (C.c_int, [_VP, C.c_int, _VP, _U32, _U32]),
    "bh_ctx_generate_disk_texture": (C.c_int, [_VP, _U32, _U32, _VP, C.c_int]),
    "bh_ctx_upload_models": (C.c_int, [_VP, _VP, C.c_size_t]),
    "bh_ctx_upload_models_async": (C.c_int, [_VP, _VP, C.c_size_t, _VP]),
    "bh_ctx_set_model_header": (C.c_int, [_VP, _U32, C.POINTER(C.c_float), _I32]),
    "bh_ray_pipeline_create": (C.c_int, [_VP, _U32, _U32, _VP, C.POINTER(_VP)]),
    "bh_ray_pipeline_destroy": (None, [_VP]),
    "bh_ray_pipeline_set_tiling": (C.c_int, [_VP, _U32, _U32, _U32]),
    "bh_ray_pipeline_local_rows": (_U32, [_VP]),
    "bh_ray_pipeline_enable_aux": (C.c_int, [_VP, _U32]),
    "bh_ray_pipeline_bind_output": (C.c_int, [_VP, _VP]),
    "bh_ray_pipeline_bind_frame": (C.c_int, [_VP, _VP]),
    "bh_shared_frame_create": (C.c_int, [_VP, C.c_size_t, C.POINTER(_VP), _VP]),
    "bh_shared_frame_open": (C.c_int, [_VP, _VP, C.POINTER(_VP)]),
    "bh_shared_frame_release": (C.c_int, [_VP, _VP, C.c_int]),
    "bh_ray_pipeline_pass": (C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "bh_ray_pipeline_pass_to_host": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _U32, _VP]),
    "bh_ray_pipeline_sync": (C.c_int, [_VP]),
    "bh_ray_pipeline_output": (_VP, [_VP]),
    "bh_ray_pipeline_width": (_U32, [_VP]),
    "bh_ray_pipeline_height": (_U32, [_VP]),
    "bh_ray_pipeline_read": (C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "bh_ray_pipeline_stats": (C.c_int, [_VP, C.POINTER(PassStats)]),
    "bh_sky_pipeline_create": (C.c_int, [_VP, _VP, C.c_int, C.POINTER(_VP)]),
    "bh_sky_pipeline_destroy": (None, [_VP]),
    "bh_sky_pipeline_bind_output": (C.c_int, [_VP, _VP]),
    "bh_sky_pipeline_pass": (C.c_int, [_VP, _VP]),
    "bh_sky_pipeline_output": (_VP, [_VP]),
    "bh_sky_pipeline_read": (C.c_int, [_VP, _VP]),
    "bh_post_pass_create": (C.c_int, [_VP, C.c_int, _U32, _U32, _VP, _U32, _U32, _VP, C.POINTER(_VP)]),
    "bh_post_pass_destroy": (None, [_VP]),
    "bh_post_pass_run": (C.c_int, [_VP, _VP, _VP]),
    "bh_post_pass_output": (_VP, [_VP]),
    "bh_post_pass_read": (C.c_int, [_VP, _VP]),
    "bh_sky_pipeline_create_for_frame": (C.c_int, [_VP, _VP, _U32, _U32, C.c_int, C.POINTER(_VP)]),
    "bh_frame_multi_create": (C.c_int, [C.POINTER(_VP), _U32, C.POINTER(FrameMultiDesc), C.POINTER(_VP)]),
    "bh_frame_multi_destroy": (None, [_VP]),
    "bh_frame_multi_width": (_U32, [_VP]),
    "bh_frame_multi_height": (_U32, [_VP]),
    "bh_frame_multi_pass": (C.c_int, [_VP, _VP, _VP, _VP]),
    "bh_frame_multi_pass_to_host": (C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "bh_frame_multi_sync": (C.c_int, [_VP]),
    "bh_frame_multi_output": (_VP, [_VP]),
    "bh_frame_multi_sky_output": (_VP, [_VP]),
    "bh_frame_multi_read": (C.c_int, [_VP, _VP, _VP]),
    "bh_frame_multi_stats": (C.c_int, [_VP, C.POINTER(PassStats), C.POINTER(C.c_float)]),
    "bh_shared_frame_flags": (_VP, [_VP, C.c_size_t]),
    "bh_stream_signal": (C.c_int, [_VP, _VP, _U32, _VP]),
    "bh_stream_wait": (C.c_int, [_VP, _VP, _U32, _U32, _U32, _VP]),
    "bh_ctx_check_async": (C.c_int, [_VP]),
    "bh_host_frame_create": (C.c_int, [_VP, C.c_char_p, C.c_size_t, C.c_int, C.POINTER(_VP)]),
    "bh_host_frame_destroy": (None, [_VP, C.c_int]),
    "bh_host_frame_ptr": (_VP, [_VP]),
    "bh_host_frame_signal": (C.c_int, [_VP, _U32, _U32]),
    "bh_host_frame_wait": (C.c_int, [_VP, _U32, _U32, _U32, _U32]),
    "bh_ray_pipeline_pass_to_host_frame": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "bh_model_validate": (C.c_int, [_VP]),
    "bh_model_load_obj": (C.c_int, [C.c_char_p, _VP, C.POINTER(ModelInfo)]),
    "bh_model_from_arrays": (C.c_int, [_VP, _I32, _VP, _I32, _VP, _I32, C.POINTER(C.c_float), _I32, _VP, C.POINTER(ModelInfo)]),
    "bh_model_build_bvh": (C.c_int, [_VP, _I32, C.POINTER(ModelInfo)]),
    "bh_save_png": (C.c_int, [C.c_char_p, _VP, _U32, _U32, C.c_int]),
    "bh_ctx_math_probe": (C.c_int, [_VP, C.c_int, _VP, _VP, _VP, C.c_size_t]),
}


def lib_path() -> str:
    """bhusie_b200/lib/libbhray.so; BHRAY_LIB overrides it with another build of the same sources (tuning runs)."""
    return os.environ.get("BHRAY_LIB") or _build.LIB_PATH


def load() -> C.CDLL:
    """Loads libbhray.so from bhusie_b200/lib/ (built in-tree by bhusie_b200.build).  Raises if it is
    missing and cannot be built — nothing in this package computes without it."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            _build.build_library()
        lib = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)          # AttributeError here == ABI/header mismatch
            fn.restype = res
            fn.argtypes = args
        if lib.bh_abi_version() != ABI_VERSION:
            raise RuntimeError("libbhray.so ABI version mismatch")
        _LIB = lib
    return _LIB


def check(code: int) -> None:
    if code != 0:
        raise BhError(code, load().bh_last_error().decode(errors="replace"))
