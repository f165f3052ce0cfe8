"""Host-side mirror of the reference's pass objects for the ray path, on top of the C ABI.

    reference (Rust, wgpu)                                   here
    -------------------------------------------------------  ----------------------------------
    Device/Queue + Renderer-owned scene buffers (mod.rs:63-114)  Context
    RayPipeline::new / output_view / pass  (ray_pipeline.rs)     RayPipeline(...) / .output_ptr / .pass()
    SkyPipeline::new / output_view / pass  (sky_pipeline.rs)     SkyPipeline(...) / .output_ptr / .pass()
    4-level pyramid + sky in one compute pass (mod.rs:170-216,406-421)  RayPyramid

`pass_()` only enqueues on a CUDA stream, like recording into a ComputePass; `read()` synchronises.
All computation happens in libbhray.so; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .uniforms import (BLACK_HOLE_UNIFORM_SIZE, CAMERA_UNIFORM_SIZE, MODEL_UNIFORM_SIZE, RAY_DETAILS_SIZE,
                       BlackHole, Camera, RayDetails, pyramid_levels)

TEX_COLOR, TEX_DISK, TEX_SKY = 0, 1, 2
AUX_HIT, AUX_STEPS, AUX_CLASS = 1, 2, 4
SKY_RGBA16F, SKY_RGBA32F = 0, 1
NUMERIC_LITERAL, NUMERIC_FUSED = 0, 1
# which oracle flavour each kernel numeric mode is bit-comparable with (tests only; the product never imports the oracle)
ORACLE_FLAVOUR_OF_MODE = {NUMERIC_LITERAL: "contract", NUMERIC_FUSED: "fused"}


def _bytes(x, size: int) -> bytes:
    b = x.uniform() if hasattr(x, "uniform") else bytes(x)
    if len(b) != size:
        raise ValueError(f"uniform must be {size} bytes, got {len(b)}")
    return b


def _stream_ptr(stream) -> C.c_void_p:
    if stream is None:
        return C.c_void_p(0)
    if hasattr(stream, "cuda_stream"):          # torch.cuda.Stream
        return C.c_void_p(stream.cuda_stream)
    return C.c_void_p(int(stream))


class Context:
    """Owns the device copies of the scene: three textures and the ModelUniform array."""

    def __init__(self, device: int = 0, numeric_mode: int | None = None):
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.bh_ctx_create(device, C.byref(h)))
        self._h = h
        self.device = device
        self.model_count = 0
        if numeric_mode is not None:
            self.set_numeric_mode(numeric_mode)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bh_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_numeric_mode(self, mode: int):
        """LITERAL (one IEEE op per WGSL node) or FUSED (fma contraction + reciprocal-multiply; default)."""
        _lib.check(self._lib.bh_ctx_set_numeric_mode(self._h, int(mode)))

    @property
    def numeric_mode(self) -> int:
        return int(self._lib.bh_ctx_get_numeric_mode(self._h))

    def set_texture(self, slot: int, rgba8: np.ndarray):
        a = np.ascontiguousarray(rgba8, dtype=np.uint8)
        if a.ndim != 3 or a.shape[2] != 4:
            raise ValueError("texture must be (h, w, 4) uint8")
        _lib.check(self._lib.bh_ctx_set_texture(self._h, slot, a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[0]))

    def set_textures(self, tex: dict):
        self.set_texture(TEX_COLOR, tex["color"])
        self.set_texture(TEX_DISK, tex["disk"])
        self.set_texture(TEX_SKY, tex["sky"])

    def generate_disk_texture(self, w: int = 1000, h: int = 1000, install: bool = True) -> np.ndarray:
        """perlin/src/main.rs on the GPU: returns the (h, w, 4) RGBA8 texels and (install) binds them as the disk texture."""
        out = np.empty((h, w, 4), np.uint8)
        _lib.check(self._lib.bh_ctx_generate_disk_texture(self._h, w, h, out.ctypes.data_as(C.c_void_p), 1 if install else 0))
        return out

    def upload_models(self, blob: np.ndarray | None):
        if blob is None:
            self.model_count = 0
            return
        a = np.ascontiguousarray(blob, dtype=np.uint8).reshape(-1)
        _lib.check(self._lib.bh_ctx_upload_models(self._h, a.ctypes.data_as(C.c_void_p), a.size))
        self.model_count = a.size // MODEL_UNIFORM_SIZE

    def upload_models_async(self, host_ptr: int, nbytes: int, stream=None):
        _lib.check(self._lib.bh_ctx_upload_models_async(self._h, C.c_void_p(host_ptr), nbytes, _stream_ptr(stream)))
        self.model_count = nbytes // MODEL_UNIFORM_SIZE

    def set_model_header(self, index: int, position, visible: int):
        pos = (C.c_float * 3)(*map(float, position))
        _lib.check(self._lib.bh_ctx_set_model_header(self._h, index, pos, int(visible)))

    def shared_frame_create(self, nbytes: int) -> tuple[int, bytes]:
        """cudaMalloc + cudaIpcGetMemHandle: (device pointer, 64-byte handle) for bind_frame on other ranks."""
        ptr = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        _lib.check(self._lib.bh_shared_frame_create(self._h, nbytes, C.byref(ptr), handle))
        return int(ptr.value), bytes(handle)

    def shared_frame_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        _lib.check(self._lib.bh_shared_frame_open(self._h, buf, C.byref(ptr)))
        return int(ptr.value)

    def shared_frame_release(self, ptr: int, owner: bool):
        _lib.check(self._lib.bh_shared_frame_release(self._h, C.c_void_p(ptr), 1 if owner else 0))

    def shared_frame_flags(self, ptr: int, nbytes: int) -> int:
        """Device pointer of the BH_SHARED_FLAGS 32-bit flags that follow a shared frame of `nbytes` bytes."""
        return int(self._lib.bh_shared_frame_flags(C.c_void_p(ptr), nbytes) or 0)

    def stream_signal(self, flag_ptr: int, value: int, stream=None):
        """Stream-ordered `*flag = value` (release, system scope): everything enqueued before it on `stream` is visible first."""
        _lib.check(self._lib.bh_stream_signal(self._h, C.c_void_p(flag_ptr), value & 0xFFFFFFFF, _stream_ptr(stream)))

    def stream_wait(self, flags_ptr: int, n_flags: int, value: int, timeout_ms: int = 10000, stream=None):
        """Stream-ordered wait (on the device) until all `n_flags` flags have reached `value`; gives up after timeout_ms."""
        _lib.check(self._lib.bh_stream_wait(self._h, C.c_void_p(flags_ptr), n_flags, value & 0xFFFFFFFF, timeout_ms, _stream_ptr(stream)))

    def check_async(self):
        """Raises BhError(-110) if a stream_wait on this context has given up since the last call."""
        _lib.check(self._lib.bh_ctx_check_async(self._h))

    def math_probe(self, fn: str, a: np.ndarray, b: np.ndarray | None = None) -> np.ndarray:
        codes = {"pow": 0, "pow5": 1, "pow4": 2, "sin": 3, "cos": 4, "tan": 5, "atan2": 6, "acos": 7}
        a = np.ascontiguousarray(a, dtype=np.float32)
        bb = None if b is None else np.ascontiguousarray(b, dtype=np.float32)
        out = np.empty_like(a)
        _lib.check(self._lib.bh_ctx_math_probe(self._h, codes[fn], a.ctypes.data_as(C.c_void_p),
                                               None if bb is None else bb.ctypes.data_as(C.c_void_p),
                                               out.ctypes.data_as(C.c_void_p), a.size))
        return out


class RayPipeline:
    """RayPipeline::new (ray_pipeline.rs:36): `prev=None` is the 1x1 base texture (mod.rs:151-168)."""

    def __init__(self, ctx: Context, width: int, height: int, prev: "RayPipeline | None" = None, aux: int = 0):
        self._lib = ctx._lib
        self.ctx = ctx
        self.prev = prev
        self.width, self.height = int(width), int(height)
        h = C.c_void_p()
        _lib.check(self._lib.bh_ray_pipeline_create(ctx._h, self.width, self.height, prev._h if prev else None, C.byref(h)))
        self._h = h
        self._aux = 0
        if aux:
            self.enable_aux(aux)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bh_ray_pipeline_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def enable_aux(self, mask: int):
        _lib.check(self._lib.bh_ray_pipeline_enable_aux(self._h, mask))
        self._aux = mask

    def set_tiling(self, band_rows: int, rank: int, n_ranks: int):
        _lib.check(self._lib.bh_ray_pipeline_set_tiling(self._h, band_rows, rank, n_ranks))

    @property
    def local_rows(self) -> int:
        return int(self._lib.bh_ray_pipeline_local_rows(self._h))

    def bind_output(self, device_ptr: int | None):
        _lib.check(self._lib.bh_ray_pipeline_bind_output(self._h, C.c_void_p(device_ptr or 0)))

    def bind_frame(self, device_ptr: int | None):
        """Write rows at their GLOBAL index into a full frame (local or peer-mapped) instead of the compact local buffer."""
        _lib.check(self._lib.bh_ray_pipeline_bind_frame(self._h, C.c_void_p(device_ptr or 0)))

    @property
    def output_ptr(self) -> int:
        """RayPipeline::output_view: device pointer to row-major RGBA32F."""
        return int(self._lib.bh_ray_pipeline_output(self._h) or 0)

    def pass_(self, camera, black_hole, details, stream=None):
        """RayPipeline::pass (ray_pipeline.rs:301-309): enqueue on `stream`, return immediately."""
        cam = _bytes(camera, CAMERA_UNIFORM_SIZE)
        hole = _bytes(black_hole, BLACK_HOLE_UNIFORM_SIZE)
        det = _bytes(details, RAY_DETAILS_SIZE)
        _lib.check(self._lib.bh_ray_pipeline_pass(self._h, cam, hole, det, _stream_ptr(stream)))

    def pass_to_host(self, camera, black_hole, details, host_ptr: int, n_chunks: int = 8, stream=None):
        """pass_() followed by a chunk-overlapped copy of the RGBA32F output into pinned host memory (async)."""
        cam = _bytes(camera, CAMERA_UNIFORM_SIZE)
        hole = _bytes(black_hole, BLACK_HOLE_UNIFORM_SIZE)
        det = _bytes(details, RAY_DETAILS_SIZE)
        _lib.check(self._lib.bh_ray_pipeline_pass_to_host(self._h, cam, hole, det, C.c_void_p(host_ptr), n_chunks, _stream_ptr(stream)))

    def pass_to_host_frame(self, camera, black_hole, details, host_ptr: int, stream=None):
        """pass_() with this (tiled) pipeline's rows stored at their GLOBAL row into a full page-locked host frame (async)."""
        cam = _bytes(camera, CAMERA_UNIFORM_SIZE)
        hole = _bytes(black_hole, BLACK_HOLE_UNIFORM_SIZE)
        det = _bytes(details, RAY_DETAILS_SIZE)
        _lib.check(self._lib.bh_ray_pipeline_pass_to_host_frame(self._h, cam, hole, det, C.c_void_p(host_ptr), _stream_ptr(stream)))

    def sync(self):
        _lib.check(self._lib.bh_ray_pipeline_sync(self._h))

    def read(self, aux: bool = True) -> dict:
        rows, w = self.local_rows, self.width
        out = {"rgba": np.empty((rows, w, 4), np.float32)}
        hit = steps = cls = None
        if aux and self._aux & AUX_HIT:
            hit = out["hit"] = np.empty((rows, w), np.int32)
        if aux and self._aux & AUX_STEPS:
            steps = out["steps"] = np.empty((rows, w), np.uint32)
        if aux and self._aux & AUX_CLASS:
            cls = out["cls"] = np.empty((rows, w), np.uint8)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        _lib.check(self._lib.bh_ray_pipeline_read(self._h, p(out["rgba"]), p(hit), p(steps), p(cls)))
        return out

    def read_into(self, host_ptr: int):
        _lib.check(self._lib.bh_ray_pipeline_read(self._h, C.c_void_p(host_ptr), None, None, None))

    def stats(self, strict: bool = True) -> dict:
        """Totals of the last pass.  strict: raise if any RK step had an error norm > 1 (BH_ERR_NUMERIC) —
        the reference's accept loop would never terminate there (ray.wgsl:425-451)."""
        s = _lib.PassStats()
        rc = self._lib.bh_ray_pipeline_stats(self._h, C.byref(s))
        if rc != 0 and (strict or rc != -34):
            _lib.check(rc)
        return s.as_dict()


class SkyPipeline:
    """SkyPipeline::new (sky_pipeline.rs:18): resolves alpha==0 pixels of `prev` against the sky map."""

    def __init__(self, ctx: Context, prev: "RayPipeline | None", fmt: int = SKY_RGBA16F, frame: "tuple | None" = None):
        """`prev`: the ray level to resolve; or `frame=(device_ptr, width, height)`: a raw RGBA32F frame in device memory
        (the frame a multi-GPU render assembles)."""
        self._lib = ctx._lib
        self.ctx, self.prev, self.fmt = ctx, prev, fmt
        h = C.c_void_p()
        if frame is not None:
            ptr, w, hh = frame
            _lib.check(self._lib.bh_sky_pipeline_create_for_frame(ctx._h, C.c_void_p(ptr), w, hh, fmt, C.byref(h)))
            self._shape = (int(hh), int(w))
        else:
            _lib.check(self._lib.bh_sky_pipeline_create(ctx._h, prev._h, fmt, C.byref(h)))
            self._shape = None
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bh_sky_pipeline_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bind_output(self, device_ptr: int | None):
        _lib.check(self._lib.bh_sky_pipeline_bind_output(self._h, C.c_void_p(device_ptr or 0)))

    @property
    def output_ptr(self) -> int:
        return int(self._lib.bh_sky_pipeline_output(self._h) or 0)

    def pass_(self, stream=None):
        _lib.check(self._lib.bh_sky_pipeline_pass(self._h, _stream_ptr(stream)))

    def read(self) -> np.ndarray:
        rows, w = self._shape if self._shape else (self.prev.local_rows, self.prev.width)
        out = np.empty((rows, w, 4), np.float32 if self.fmt == SKY_RGBA32F else np.float16)
        _lib.check(self._lib.bh_sky_pipeline_read(self._h, out.ctypes.data_as(C.c_void_p)))
        return out


class RayPyramid:
    """The reference's adaptive grid: `iters` RayPipelines, level n reading level n-1, then the sky
    resolve — what Renderer::new builds (mod.rs:170-216) and Renderer::render dispatches in one
    compute pass (mod.rs:406-421)."""

    def __init__(self, ctx: Context, base=(72, 41), multiplier: int = 3, iters: int = 4, aux: int = 0,
                 sky_format: int = SKY_RGBA16F):
        self.ctx = ctx
        self.sizes = pyramid_levels(base, multiplier, iters)
        self.levels: list[RayPipeline] = []
        for (w, h) in self.sizes:
            self.levels.append(RayPipeline(ctx, w, h, self.levels[-1] if self.levels else None, aux=aux))
        self.sky = SkyPipeline(ctx, self.levels[-1], sky_format)

    def pass_(self, camera, black_hole, details, stream=None):
        for rp in self.levels:
            rp.pass_(camera, black_hole, details, stream)
        self.sky.pass_(stream)

    def close(self):
        self.sky.close()
        for rp in reversed(self.levels):
            rp.close()


class FrameMulti:
    """bh_frame_multi: the whole compute pass of Renderer::render (mod.rs:406-421) on N devices driven from ONE host
    thread of ONE process — coarse pyramid levels replicated per device, the last level tiled in cyclic row bands whose
    pixels every device stores straight into device 0's frame over NVLink, sky resolve on device 0.  No torch, no NCCL."""

    def __init__(self, ctxs: "list[Context]", base=(72, 41), multiplier: int = 3, iters: int = 4, band_rows: int = 8,
                 sky_format: "int | None" = SKY_RGBA16F):
        self._lib = ctxs[0]._lib
        self.ctxs = list(ctxs)
        self.sky_format = sky_format
        arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
        desc = _lib.FrameMultiDesc(int(base[0]), int(base[1]), int(iters), int(multiplier), int(band_rows),
                                   -1 if sky_format is None else int(sky_format))
        h = C.c_void_p()
        _lib.check(self._lib.bh_frame_multi_create(arr, len(ctxs), C.byref(desc), C.byref(h)))
        self._h = h
        self.width = int(self._lib.bh_frame_multi_width(h))
        self.height = int(self._lib.bh_frame_multi_height(h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bh_frame_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pass_(self, camera, black_hole, details):
        _lib.check(self._lib.bh_frame_multi_pass(self._h, _bytes(camera, CAMERA_UNIFORM_SIZE), _bytes(black_hole, BLACK_HOLE_UNIFORM_SIZE),
                                                 _bytes(details, RAY_DETAILS_SIZE)))

    def pass_to_host(self, camera, black_hole, details, host_ptr: int):
        _lib.check(self._lib.bh_frame_multi_pass_to_host(self._h, _bytes(camera, CAMERA_UNIFORM_SIZE), _bytes(black_hole, BLACK_HOLE_UNIFORM_SIZE),
                                                         _bytes(details, RAY_DETAILS_SIZE), C.c_void_p(host_ptr)))

    def sync(self):
        _lib.check(self._lib.bh_frame_multi_sync(self._h))

    @property
    def output_ptr(self) -> int:
        return int(self._lib.bh_frame_multi_output(self._h) or 0)

    @property
    def sky_output_ptr(self) -> int:
        return int(self._lib.bh_frame_multi_sky_output(self._h) or 0)

    def read(self, rgba: bool = True, sky: bool = True) -> dict:
        out = {}
        if rgba:
            out["rgba"] = np.empty((self.height, self.width, 4), np.float32)
        if sky and self.sky_format is not None:
            out["sky"] = np.empty((self.height, self.width, 4), np.float32 if self.sky_format == SKY_RGBA32F else np.float16)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        _lib.check(self._lib.bh_frame_multi_read(self._h, p(out.get("rgba")), p(out.get("sky"))))
        return out

    def stats(self, strict: bool = True) -> dict:
        s = _lib.PassStats()
        ms = C.c_float(0.0)
        rc = self._lib.bh_frame_multi_stats(self._h, C.byref(s), C.byref(ms))
        if rc != 0 and (strict or rc != -34):
            _lib.check(rc)
        d = s.as_dict()
        d["elapsed_ms"] = float(ms.value)
        return d


class HostFrame:
    """bh_host_frame: a page-locked frame in POSIX shared memory that every rank of a node maps and registers with CUDA, so
    each rank's kernel stores its bands straight into the caller's host frame over its own PCIe link."""

    def __init__(self, ctx: Context, name: str, nbytes: int, create: bool):
        self._lib = ctx._lib
        self.ctx, self.name, self.nbytes, self.owner = ctx, name, int(nbytes), bool(create)
        h = C.c_void_p()
        _lib.check(self._lib.bh_host_frame_create(ctx._h, name.encode(), self.nbytes, 1 if create else 0, C.byref(h)))
        self._h = h
        self.ptr = int(self._lib.bh_host_frame_ptr(h))

    def array(self, shape, dtype=np.float32) -> np.ndarray:
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        assert n <= self.nbytes
        buf = (C.c_uint8 * n).from_address(self.ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def signal(self, slot: int, value: int):
        _lib.check(self._lib.bh_host_frame_signal(self._h, slot, value & 0xFFFFFFFF))

    def wait(self, first_slot: int, n_slots: int, value: int, timeout_ms: int = 10000):
        _lib.check(self._lib.bh_host_frame_wait(self._h, first_slot, n_slots, value & 0xFFFFFFFF, timeout_ms))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bh_host_frame_destroy(self._h, 1 if self.owner else 0)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def validate_model(blob: np.ndarray):
    """bh_model_validate: raises BhError(-22) when the BVH of a ModelUniform blob leads the traversal out of its arrays."""
    a = np.ascontiguousarray(blob, dtype=np.uint8).reshape(-1)
    _lib.check(_lib.load().bh_model_validate(a.ctypes.data_as(C.c_void_p)))


def load_obj_model(path: str) -> tuple[np.ndarray, dict]:
    """load_model (model.rs:7-87) + build_bvh through the library's host code -> ModelUniform bytes."""
    lib = _lib.load()
    blob = np.zeros(MODEL_UNIFORM_SIZE, np.uint8)
    info = _lib.ModelInfo()
    _lib.check(lib.bh_model_load_obj(path.encode(), blob.ctypes.data_as(C.c_void_p), C.byref(info)))
    return blob, info.as_dict()


def model_from_arrays(points: np.ndarray, normals: np.ndarray, tris: np.ndarray, position=(-10.0, 0.0, 30.0),
                      visible: int = 1) -> tuple[np.ndarray, dict]:
    lib = _lib.load()
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    nrm = np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
    tr = np.ascontiguousarray(tris, np.int32).reshape(-1, 6)
    blob = np.zeros(MODEL_UNIFORM_SIZE, np.uint8)
    info = _lib.ModelInfo()
    pos = (C.c_float * 3)(*map(float, position))
    _lib.check(lib.bh_model_from_arrays(pts.ctypes.data_as(C.c_void_p), len(pts), nrm.ctypes.data_as(C.c_void_p), len(nrm),
                                        tr.ctypes.data_as(C.c_void_p), len(tr), pos, int(visible),
                                        blob.ctypes.data_as(C.c_void_p), C.byref(info)))
    return blob, info.as_dict()


__all__ = ["Context", "RayPipeline", "SkyPipeline", "RayPyramid", "FrameMulti", "HostFrame", "validate_model", "Camera", "BlackHole", "RayDetails",
           "load_obj_model", "model_from_arrays", "TEX_COLOR", "TEX_DISK", "TEX_SKY", "AUX_HIT", "AUX_STEPS",
           "AUX_CLASS", "SKY_RGBA16F", "SKY_RGBA32F", "NUMERIC_LITERAL", "NUMERIC_FUSED", "ORACLE_FLAVOUR_OF_MODE"]
