"""Host-side mirror of the reference's post chain pass objects, on top of the C ABI.

    reference (Rust, wgpu)                                              here
    ------------------------------------------------------------------  -----------------------------
    BloomPipeline::new/output_view/pass   (pipelines/bloom_pipline.rs)  BloomPipeline
    MixPipeline::new/output_view/pass     (pipelines/mix_pipeline.rs)   MixPipeline
    HDRPipeline::new/output_view/pass     (pipelines/hdr_pipeline.rs)   HDRPipeline
    FXAAPipeline::new/output_view/pass    (pipelines/fxaa_pipline.rs)   FXAAPipeline
    Renderer::new wiring + render order   (renderer/mod.rs:219-312,425-431)  PostChain

A "texture view" here is (device pointer, width, height) of an RGBA16F image.
"""
from __future__ import annotations

import ctypes as C
import struct
from dataclasses import dataclass

import numpy as np

from . import _lib

BLOOM_DOWN, BLOOM_UP, MIX, HDR, FXAA = 0, 1, 2, 3, 4


@dataclass
class MixDetails:
    """mix_pipeline.rs:5-7; default 0.7 (mod.rs:258-260)."""
    mix_ratio: float = 0.7

    def uniform(self) -> bytes:
        return struct.pack("<f", self.mix_ratio)


@dataclass
class FXAADetails:
    """FXAADetailsUniform (fxaa_pipline.rs:74-92); defaults = Ultra/Ultra, 12 iterations, 0.75 (mod.rs:290-295)."""
    edge_threshold_min: float = 0.0156
    edge_threshold_max: float = 0.063
    iterations: int = 12
    subpixel_quality: float = 0.75

    def uniform(self) -> bytes:
        return struct.pack("<ffif", self.edge_threshold_min, self.edge_threshold_max, self.iterations, self.subpixel_quality)


def bloom_sizes(resolution=(1918, 1081), count: int = 5, multiplier: float = 2.0):
    """mod.rs:219-229: resolution halves `count` times then doubles back, in f32, truncated to u32 at each use."""
    cw, ch = np.float32(resolution[0]), np.float32(resolution[1])
    m = np.float32(multiplier)
    out = []
    for i in range(2 * count):
        cw, ch = (np.float32(cw / m), np.float32(ch / m)) if i < count else (np.float32(cw * m), np.float32(ch * m))
        out.append((int(cw), int(ch)))
    return out


def save_png(path: str, rgba8: np.ndarray, force_opaque: bool = True):
    """Renderer's "Save Image" (mod.rs:460-486): RGBA8 frame -> PNG through the library's host code."""
    a = np.ascontiguousarray(rgba8, dtype=np.uint8)
    if a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("frame must be (h, w, 4) uint8")
    _lib.check(_lib.load().bh_save_png(path.encode(), a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[0], 1 if force_opaque else 0))


class _PostPass:
    def __init__(self, ctx, kind: int, resolution, view1, view2=None):
        self._lib = ctx._lib
        self.ctx, self.kind = ctx, kind
        self.width, self.height = int(resolution[0]), int(resolution[1])
        ptr1, w1, h1 = view1
        ptr2 = view2[0] if view2 is not None else None
        h = C.c_void_p()
        _lib.check(self._lib.bh_post_pass_create(ctx._h, kind, self.width, self.height, C.c_void_p(ptr1), int(w1), int(h1),
                                                 C.c_void_p(ptr2 or 0), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bh_post_pass_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def output_view(self):
        return (int(self._lib.bh_post_pass_output(self._h) or 0), self.width, self.height)

    def _run(self, details: bytes | None, stream):
        from .pipelines import _stream_ptr
        _lib.check(self._lib.bh_post_pass_run(self._h, details, _stream_ptr(stream)))

    def read(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), np.uint8 if self.kind == FXAA else np.float16)
        _lib.check(self._lib.bh_post_pass_read(self._h, out.ctypes.data_as(C.c_void_p)))
        return out


class BloomPipeline(_PostPass):
    def __init__(self, ctx, resolution, texture_view, direction: str):
        super().__init__(ctx, BLOOM_DOWN if direction == "down" else BLOOM_UP, resolution, texture_view)

    def pass_(self, stream=None):
        self._run(None, stream)


class MixPipeline(_PostPass):
    def __init__(self, ctx, resolution, texture_view_1, texture_view_2):
        super().__init__(ctx, MIX, resolution, texture_view_1, texture_view_2)

    def pass_(self, details: MixDetails, stream=None):
        self._run(details.uniform(), stream)


class HDRPipeline(_PostPass):
    def __init__(self, ctx, resolution, texture_view):
        super().__init__(ctx, HDR, resolution, texture_view)

    def pass_(self, stream=None):
        self._run(None, stream)


class FXAAPipeline(_PostPass):
    def __init__(self, ctx, resolution, texture_view):
        super().__init__(ctx, FXAA, resolution, texture_view)

    def pass_(self, details: FXAADetails, stream=None):
        self._run(details.uniform(), stream)


class PostChain:
    """What Renderer::new builds after the sky pipeline (mod.rs:219-312) and Renderer::render runs after the compute
    pass (mod.rs:425-431): 5 bloom downs, 5 bloom ups, mix(sky, bloom), ACES, FXAA."""

    def __init__(self, ctx, sky_pipeline, bloom_count: int = 5):
        from .pipelines import SKY_RGBA16F
        if sky_pipeline.fmt != SKY_RGBA16F:
            raise ValueError("the post chain consumes the sky pass's Rgba16Float output")
        res = (sky_pipeline.prev.width, sky_pipeline.prev.height)
        sky_view = (sky_pipeline.output_ptr, res[0], res[1])
        self.sizes = bloom_sizes(res, bloom_count)
        self.blooms: list[BloomPipeline] = []
        view = sky_view
        for i, sz in enumerate(self.sizes):
            bp = BloomPipeline(ctx, sz, view, "down" if i < bloom_count else "up")
            self.blooms.append(bp)
            view = bp.output_view()
        assert self.sizes[-1] == res
        self.mix = MixPipeline(ctx, res, sky_view, view)
        self.hdr = HDRPipeline(ctx, res, self.mix.output_view())
        self.fxaa = FXAAPipeline(ctx, res, self.hdr.output_view())
        self.mix_details, self.fxaa_details = MixDetails(), FXAADetails()

    def pass_(self, stream=None):
        for bp in self.blooms:
            bp.pass_(stream)
        self.mix.pass_(self.mix_details, stream)
        self.hdr.pass_(stream)
        self.fxaa.pass_(self.fxaa_details, stream)

    def read(self) -> np.ndarray:
        """The frame the reference hands to the screen pass: RGBA8, sRGB-encoded."""
        return self.fxaa.read()

    def close(self):
        for p in [self.fxaa, self.hdr, self.mix] + self.blooms[::-1]:
            p.close()
