"""ctypes front end of the CPU oracle (oracle/bh_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (bhusie_b200/) never does.

Flavours (see bho_math.h and the madd/vdivs helpers in bh_oracle.c):
  "strict"    glibc libm transcendentals, one IEEE op per WGSL expression node (neutral; also the CPU baseline)
  "contract"  det-math transcendentals, same literal op order -> bit-comparable with the kernel's LITERAL mode
  "fused"     det-math + explicit fma contraction + reciprocal-multiply for vec/scalar division
              -> bit-comparable with the kernel's FUSED mode (the fast default)
  "shadow"    float64 shadow of "strict": the same source with every arithmetic float widened to double (f32 inputs,
              constants and outputs).  |strict - shadow| > tolerance marks a pixel as ILL-CONDITIONED — decided by f32
              rounding — which is how the outliers of a device-vs-strict comparison are classified.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MU_SIZE = 48234572
MU_POINTS, MU_NORMALS, MU_TRIANGLES, MU_NODES, MU_LOOKUP = 48, 8388656, 16777264, 29360176, 46137392
MAX_MODEL_VERTICES = 524288

COUNTER_FIELDS = ("steps", "loop_iters", "node_visits", "tri_tests", "tex_samples", "stack_overflow",
                  "rk_reject", "px_traced", "px_copied", "px_interp", "bvh_calls")


class _Tex(C.Structure):
    _fields_ = [("rgba", C.c_void_p), ("w", C.c_int32), ("h", C.c_int32)]


class _Scene(C.Structure):
    _fields_ = [("color", _Tex), ("disk", _Tex), ("sky", _Tex), ("models", C.c_void_p), ("model_capacity", C.c_int32)]


class _Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in COUNTER_FIELDS]


FLAVOURS = ("strict", "contract", "fused", "shadow")


def build(force: bool = False) -> None:
    """Compile both flavours with oracle/Makefile (gcc, -ffp-contract=off, OpenMP)."""
    libs = [os.path.join(_HERE, "lib", f"libbh_oracle_{f}.so") for f in FLAVOURS]
    srcs = [os.path.join(_HERE, n) for n in ("bh_oracle.c", "bho_math.h", "bh_oracle_post.inc", "Makefile")]
    stale = force or any(not os.path.exists(l) for l in libs) or \
        max(os.path.getmtime(s) for s in srcs) > min(os.path.getmtime(l) for l in libs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)


_LIBS: dict[str, C.CDLL] = {}


def _lib(flavour: str) -> C.CDLL:
    if flavour not in FLAVOURS:
        raise ValueError(flavour)
    if flavour not in _LIBS:
        path = os.path.join(_HERE, "lib", f"libbh_oracle_{flavour}.so")
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        lib.bho_ray_pass.restype = C.c_int
        lib.bho_sky_pass.restype = C.c_int
        lib.bho_max_threads.restype = C.c_int
        if flavour == "shadow":                  # ray / sky pass only: no post chain, no leaf-function KAT entry points
            assert lib.bho_flavour() == 3
            _LIBS[flavour] = lib
            return lib
        for n in ("bho_bloom_down", "bho_bloom_up", "bho_mix", "bho_hdr", "bho_fxaa"):
            getattr(lib, n).restype = C.c_int
        lib.bho_build_bvh.restype = C.c_int64
        lib.bho_load_obj.restype = C.c_int64
        for n in ("pow",):
            getattr(lib, f"bho_kat_{n}").restype = C.c_float
            getattr(lib, f"bho_kat_{n}").argtypes = [C.c_float, C.c_float]
        lib.bho_kat_atan2.restype = C.c_float
        lib.bho_kat_atan2.argtypes = [C.c_float, C.c_float]
        for n in ("pow5", "pow4", "sin", "cos", "tan", "acos"):
            getattr(lib, f"bho_kat_{n}").restype = C.c_float
            getattr(lib, f"bho_kat_{n}").argtypes = [C.c_float]
        lib.bho_kat_f16.restype = C.c_uint16
        lib.bho_kat_f16.argtypes = [C.c_float]
        assert lib.bho_flavour() == FLAVOURS.index(flavour)
        _LIBS[flavour] = lib
    return _LIBS[flavour]


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


@dataclass
class OracleScene:
    """Host copies of everything the pass reads besides the three small uniforms."""
    color: np.ndarray   # (h, w, 4) uint8
    disk: np.ndarray
    sky: np.ndarray
    models: np.ndarray | None = None   # uint8, n * MU_SIZE verbatim ModelUniform bytes

    def _c(self) -> _Scene:
        s = _Scene()
        for name in ("color", "disk", "sky"):
            a = getattr(self, name)
            assert a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 4 and a.flags.c_contiguous
            t = getattr(s, name)
            t.rgba = a.ctypes.data
            t.h, t.w = a.shape[0], a.shape[1]
        if self.models is not None:
            assert self.models.dtype == np.uint8 and self.models.size % MU_SIZE == 0
            s.models = self.models.ctypes.data
            s.model_capacity = self.models.size // MU_SIZE
        else:
            s.models = None
            s.model_capacity = 0
        return s


@dataclass
class RayResult:
    rgba: np.ndarray      # (h, w, 4) float32
    hit: np.ndarray       # (h, w) int32, -1 = no triangle
    steps: np.ndarray     # (h, w) uint32
    cls: np.ndarray       # (h, w) uint8: 0 traced(base) 1 copied 2 interpolated 3 traced(fine)
    counters: dict


SHADOW_PERTURBATION = 1e-6     # radians: about the rounding error a binary32 integration accumulates on a direction (99th pct 8e-7)


def ray_pass(scene: OracleScene, w: int, h: int, camera: bytes, black_hole: bytes, details: bytes,
             prev: np.ndarray | None = None, rows: tuple[int, int] | None = None,
             flavour: str = "strict", nthreads: int = 0, out: RayResult | None = None, perturb: float = 0.0) -> RayResult:
    """One RayPipeline::pass (ray_pipeline.rs:301-309) on the CPU.  `perturb` (shadow flavour only): rotate every camera
    ray by that many radians first — the conditioning probe."""
    lib = _lib(flavour)
    if flavour == "shadow":
        lib.bho_shadow_set_perturbation(C.c_double(perturb))
    elif perturb:
        raise ValueError("perturb is a probe of the shadow flavour")
    assert len(camera) == 32 and len(black_hole) == 132 and len(details) == 32
    if out is None:
        out = RayResult(np.zeros((h, w, 4), np.float32), np.full((h, w), -1, np.int32),
                        np.zeros((h, w), np.uint32), np.zeros((h, w), np.uint8), {})
    r0, r1 = rows if rows is not None else (0, h)
    pw = ph = 0
    if prev is not None:
        prev = _f32(prev)
        ph, pw = prev.shape[0], prev.shape[1]
    cs = scene._c()
    cnt = _Counters()
    rc = lib.bho_ray_pass(C.byref(cs), C.c_int32(w), C.c_int32(h), _p(prev), C.c_int32(pw), C.c_int32(ph),
                          C.c_char_p(bytes(camera)), C.c_char_p(bytes(black_hole)), C.c_char_p(bytes(details)),
                          C.c_int32(r0), C.c_int32(r1), _p(out.rgba), _p(out.hit), _p(out.steps), _p(out.cls),
                          C.byref(cnt), C.c_int32(nthreads))
    if rc != 0:
        raise RuntimeError(f"bho_ray_pass failed: {rc}")
    out.counters = {n: int(getattr(cnt, n)) for n in COUNTER_FIELDS}
    return out


def sky_pass(scene: OracleScene, prev: np.ndarray, flavour: str = "strict", nthreads: int = 0,
             rows: tuple[int, int] | None = None):
    """One SkyPipeline::pass (sky_pipeline.rs:140-148).  Returns (rgba32f, rgba16f-as-uint16, counters)."""
    lib = _lib(flavour)
    prev = _f32(prev)
    h, w = prev.shape[0], prev.shape[1]
    o32 = np.zeros((h, w, 4), np.float32)
    o16 = np.zeros((h, w, 4), np.uint16)
    r0, r1 = rows if rows is not None else (0, h)
    cs = scene._c()
    cnt = _Counters()
    rc = lib.bho_sky_pass(C.byref(cs), C.c_int32(w), C.c_int32(h), _p(prev), C.c_int32(r0), C.c_int32(r1),
                          _p(o32), _p(o16), C.byref(cnt), C.c_int32(nthreads))
    if rc != 0:
        raise RuntimeError(f"bho_sky_pass failed: {rc}")
    return o32, o16, {n: int(getattr(cnt, n)) for n in COUNTER_FIELDS}


# ---------------------------------------------------------------- post chain
def _h16(a) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype == np.float16:
        a = a.view(np.uint16)
    assert a.dtype == np.uint16 and a.ndim == 3 and a.shape[2] == 4
    return np.ascontiguousarray(a)


def bloom(src, out_w: int, out_h: int, direction: str, flavour: str = "strict") -> np.ndarray:
    """bloom_down.wgsl / bloom_up.wgsl on an RGBA16F image (uint16 bits or float16); returns uint16 bits (h, w, 4)."""
    s = _h16(src)
    dst = np.zeros((out_h, out_w, 4), np.uint16)
    fn = getattr(_lib(flavour), "bho_bloom_down" if direction == "down" else "bho_bloom_up")
    rc = fn(_p(s), C.c_int32(s.shape[1]), C.c_int32(s.shape[0]), _p(dst), C.c_int32(out_w), C.c_int32(out_h))
    if rc != 0:
        raise RuntimeError(f"bho_bloom_{direction} failed: {rc}")
    return dst


def mix(in1, in2, mix_ratio: float = 0.7, flavour: str = "strict") -> np.ndarray:
    a, b = _h16(in1), _h16(in2)
    assert a.shape == b.shape
    dst = np.zeros_like(a)
    rc = _lib(flavour).bho_mix(_p(a), _p(b), C.c_int32(a.shape[1]), C.c_int32(a.shape[0]), C.c_float(mix_ratio), _p(dst))
    if rc != 0:
        raise RuntimeError(f"bho_mix failed: {rc}")
    return dst


def hdr(src, flavour: str = "strict") -> np.ndarray:
    s = _h16(src)
    dst = np.zeros_like(s)
    rc = _lib(flavour).bho_hdr(_p(s), C.c_int32(s.shape[1]), C.c_int32(s.shape[0]), _p(dst))
    if rc != 0:
        raise RuntimeError(f"bho_hdr failed: {rc}")
    return dst


def fxaa(src, details: bytes, flavour: str = "strict") -> np.ndarray:
    """fxaa.wgsl -> RGBA8 (sRGB-encoded rgb, linear alpha), uint8 (h, w, 4)."""
    s = _h16(src)
    assert len(details) == 16
    dst = np.zeros(s.shape, np.uint8)
    rc = _lib(flavour).bho_fxaa(_p(s), C.c_int32(s.shape[1]), C.c_int32(s.shape[0]), C.c_char_p(bytes(details)), _p(dst))
    if rc != 0:
        raise RuntimeError(f"bho_fxaa failed: {rc}")
    return dst


def post_chain(sky16, sizes, mix_ratio: float, fxaa_details: bytes, flavour: str = "strict") -> dict:
    """The whole chain of mod.rs:219-312 / 425-431: bloom down xN, up xN, mix, ACES, FXAA.  `sizes` = the 2N bloom
    resolutions [(w,h), ...].  Returns every stage's output."""
    out = {"bloom": []}
    cur = _h16(sky16)
    n = len(sizes) // 2
    for i, (w, h) in enumerate(sizes):
        cur = bloom(cur, w, h, "down" if i < n else "up", flavour)
        out["bloom"].append(cur)
    out["mix"] = mix(sky16, cur, mix_ratio, flavour)
    out["hdr"] = hdr(out["mix"], flavour)
    out["fxaa"] = fxaa(out["hdr"], fxaa_details, flavour)
    return out


def disk_texture(w: int = 1000, h: int = 1000, flavour: str = "strict") -> np.ndarray:
    """perlin/src/main.rs: the generator of disk.png -> (h, w, 4) uint8 with r=g=b=a."""
    out = np.zeros((h, w, 4), np.uint8)
    lib = _lib(flavour)
    lib.bho_disk_texture.restype = C.c_int
    rc = lib.bho_disk_texture(C.c_int32(w), C.c_int32(h), _p(out))
    if rc != 0:
        raise RuntimeError(f"bho_disk_texture failed: {rc}")
    return out


def parity_report(dev_rgba: np.ndarray, strict_rgba: np.ndarray, shadow_rgba: np.ndarray | None = None, tol: float = 1e-4,
                  shadow_perturbed_rgba: np.ndarray | None = None) -> dict:
    """SURVEY §8c protocol: device output against the strict flavour — fraction of pixels with any channel beyond `tol`,
    max abs difference, RMS — and, given the float64 shadow of the same frame, the share of those outliers that the
    shadow marks ILL-CONDITIONED, plus the outlier fraction that remains on well-conditioned pixels.  A pixel is
    ill-conditioned when (a) the strict binary32 evaluation itself is off the float64 one by more than `tol`, or (b) —
    given `shadow_perturbed_rgba`, the shadow traced from camera rays rotated by SHADOW_PERTURBATION — a perturbation of
    the size of binary32's accumulated rounding error moves the float64 result by more than `tol`."""
    d = np.abs(dev_rgba.astype(np.float64) - strict_rgba.astype(np.float64))
    d = np.where(np.isnan(d), np.inf, d)
    same_nan = np.isnan(dev_rgba) & np.isnan(strict_rgba)
    d = np.where(same_nan, 0.0, d)
    bad = (d > tol).any(axis=-1)
    fin = d[np.isfinite(d)]
    rep = {"pixels": int(bad.size), "outliers": int(bad.sum()), "outlier_frac": float(bad.mean()),
           "max_abs": float(fin.max()) if fin.size else 0.0, "rms": float(np.sqrt(np.mean(fin ** 2))) if fin.size else 0.0, "tol": tol}
    if shadow_rgba is not None:
        s = np.abs(strict_rgba.astype(np.float64) - shadow_rgba.astype(np.float64))
        s = np.where(np.isnan(s), np.inf, s)
        ill = (s > tol).any(axis=-1)
        rep["criterion"] = "|strict - shadow| > tol"
        if shadow_perturbed_rgba is not None:
            q = np.abs(shadow_perturbed_rgba.astype(np.float64) - shadow_rgba.astype(np.float64))
            q = np.where(np.isnan(q), np.inf, q)
            ill = ill | (q > tol).any(axis=-1)
            rep["criterion"] += f" or |shadow(rays rotated by {SHADOW_PERTURBATION:g} rad) - shadow| > tol"
        rep["ill_conditioned_frac"] = float(ill.mean())
        rep["outliers_ill_conditioned"] = int((bad & ill).sum())
        rep["outliers_ill_conditioned_share"] = float((bad & ill).sum() / bad.sum()) if bad.sum() else 1.0
        rep["outlier_frac_well_conditioned"] = float((bad & ~ill).sum() / bad.size)
    return rep


def new_model_blob() -> np.ndarray:
    return np.zeros(MU_SIZE, np.uint8)


def build_bvh(blob: np.ndarray, triangle_count: int, flavour: str = "strict") -> tuple[int, int]:
    """triangle.rs:143-259 on a ModelUniform blob in place.  Returns (nodes_used, max_depth)."""
    depth = C.c_int32(0)
    used = _lib(flavour).bho_build_bvh(_p(blob), C.c_int32(triangle_count), C.byref(depth))
    if used < 0:
        raise RuntimeError(f"bho_build_bvh failed: {used}")
    return int(used), int(depth.value)


def load_obj(path: str, flavour: str = "strict") -> tuple[np.ndarray, dict]:
    """model.rs:7-87: OBJ -> ModelUniform bytes (+BVH)."""
    blob = new_model_blob()
    pc, nu, md = C.c_int32(0), C.c_int64(0), C.c_int32(0)
    tc = _lib(flavour).bho_load_obj(C.c_char_p(path.encode()), _p(blob), C.byref(pc), C.byref(nu), C.byref(md))
    if tc < 0:
        raise RuntimeError(f"bho_load_obj({path}) failed: {tc}")
    return blob, {"triangles": int(tc), "points": int(pc.value), "nodes": int(nu.value), "max_depth": int(md.value)}


def blob_views(blob: np.ndarray) -> dict:
    """Typed numpy views into a ModelUniform blob (SURVEY App. B offsets)."""
    n = MAX_MODEL_VERTICES
    node_dt = np.dtype([("min", np.float32, 3), ("left_child", np.int32), ("max", np.float32, 3), ("obj_count", np.int32)])
    return {
        "position": blob[0:12].view(np.float32),
        "visible": blob[12:16].view(np.int32),
        "points": blob[MU_POINTS:MU_POINTS + 16 * n].view(np.float32).reshape(n, 4),
        "normals": blob[MU_NORMALS:MU_NORMALS + 16 * n].view(np.float32).reshape(n, 4),
        "triangles": blob[MU_TRIANGLES:MU_TRIANGLES + 24 * n].view(np.int32).reshape(n, 6),
        "nodes": blob[MU_NODES:MU_NODES + 32 * n].view(node_dt),
        "lookup": blob[MU_LOOKUP:MU_LOOKUP + 4 * n].view(np.int32),
    }


def math_array(fn: str, a, b=None, flavour: str = "contract") -> np.ndarray:
    """Vectorised access to the flavour's scalar math (for bit-exact comparison with the device)."""
    codes = {"pow": 0, "pow5": 1, "pow4": 2, "sin": 3, "cos": 4, "tan": 5, "atan2": 6, "acos": 7}
    a = _f32(a)
    b = _f32(b) if b is not None else np.zeros_like(a)
    out = np.empty_like(a)
    _lib(flavour).bho_kat_math_array(C.c_int(codes[fn]), _p(a), _p(b), _p(out), C.c_int64(a.size))
    return out


def max_threads() -> int:
    return int(_lib("strict").bho_max_threads())
