"""The product's device source on the CPU, against the oracle — no GPU needed.

tests/host_kernel/host_kernel.cpp compiles bhusie_b200/csrc/ray_impl.cuh + detmath.cuh with g++ (one lane per "warp",
warp votes / atomics / packed-FP32 intrinsics emulated one IEEE operation each) and runs trace_warp pixel by pixel;
tests/host_kernel/host_warp.cpp runs the kernels themselves as one warp of 32 host threads in lock-step.  What
this pins without a GPU is the LOGIC of the kernel's state machine — the speculative quiet step, the literal tail, the
event / shading / flat-space phases, the cold rows — and its arithmetic as written, bit for bit against the oracle flavour
of the numeric mode.  What it cannot see is anything ptxas or the hardware adds (contraction, MUFU seeds, scheduling): that
is the job of the `-m gpu` suite, which runs the same comparisons through the C ABI on the real kernels.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, bits
from bhusie_b200 import uniforms as U

SRC = os.path.join(ROOT, "tests", "host_kernel", "host_kernel.cpp")
OUT = os.path.join(ROOT, "tests", "host_kernel", "_build", "libbh_host_kernel.so")
DEPS = [SRC] + [os.path.join(ROOT, "bhusie_b200", "csrc", f) for f in ("ray_impl.cuh", "detmath.cuh", "bh_device.h")]
FLAVOUR = {0: "contract", 1: "fused"}
STAT_NAMES = ("steps", "px_traced", "px_copied", "px_interp", "node_visits", "tri_tests", "tex_samples", "rk_reject", "stack_overflow")


def _build(out, extra):
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.isdir(cuda_inc):
        pytest.skip("CUDA headers not found")
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in DEPS):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        cmd = [cxx, "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-march=x86-64-v3", *extra,
               "-I", cuda_inc, "-x", "c++", SRC, "-o", out]
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[-4000:]
    lib = C.CDLL(out)
    lib.bh_host_kernel_pass.restype = C.c_int
    return lib


@pytest.fixture(scope="session")
def host_kernel():
    return _build(OUT, [])


def host_pass(lib, mode, tex, blob, w, h, cam, hole, det):
    rgba = np.zeros((h, w, 4), np.float32)
    hit = np.zeros((h, w), np.int32)
    steps = np.zeros((h, w), np.uint32)
    stats = np.zeros(9, np.uint64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    t = {k: np.ascontiguousarray(tex[k]) for k in ("color", "disk", "sky")}
    rc = lib.bh_host_kernel_pass(C.c_int(mode), C.c_char_p(cam.uniform()), C.c_char_p(hole.uniform()), C.c_char_p(det.uniform()),
                                 p(t["color"]), C.c_int(t["color"].shape[1]), C.c_int(t["color"].shape[0]),
                                 p(t["disk"]), C.c_int(t["disk"].shape[1]), C.c_int(t["disk"].shape[0]),
                                 p(t["sky"]), C.c_int(t["sky"].shape[1]), C.c_int(t["sky"].shape[0]),
                                 p(blob), C.c_int(w), C.c_int(h), p(rgba), p(hit), p(steps), p(stats))
    assert rc == 0
    return rgba, hit, steps, dict(zip(STAT_NAMES, (int(x) for x in stats)))


CASES = {
    "default": (dict(), dict(), dict()),
    "outside_sphere": (dict(position=(0, 0, -45)), dict(), dict()),
    "oblique": (dict(position=(-28, 4, 24), forward=(0.6, -0.1, 0.2)), dict(), dict()),
    "near_hole": (dict(position=(3, 1, -6)), dict(), dict()),
    "moved_hole": (dict(), dict(position=(1.5, -0.5, 2.0)), dict()),
    "big_step": (dict(), dict(), dict(step_size=0.8)),
    "short": (dict(), dict(), dict(max_iterations=40)),
    "tiny_step": (dict(position=(3, 1, -6)), dict(), dict(step_size=2e-4, max_iterations=400)),
    "grazing_plane": (dict(), dict(accretion_disk_rotation=(0.0, 0.0, 0.0)), dict()),
}


@pytest.mark.parametrize("mode", [0, 1], ids=["literal", "fused"])
@pytest.mark.parametrize("method", [0, 1], ids=["euler", "rk"])
@pytest.mark.parametrize("case", list(CASES))
def test_device_source_on_cpu_matches_oracle(host_kernel, oracle, small_scene, small_oracle_scene, mode, method, case):
    tex, blob, _ = small_scene
    ck, hk, dk = CASES[case]
    cam, hole = U.Camera(**ck), U.BlackHole(**hk)
    det = U.RayDetails(integration_method=method, model_count=1, time=1.25, **dk)
    w, h = 33, 19                                     # odd: the centre pixel has zero angular momentum (sqrt operand exactly 0)
    rgba, hit, steps, st = host_pass(host_kernel, mode, tex, blob, w, h, cam, hole, det)
    ora = oracle.ray_pass(small_oracle_scene, w, h, cam.uniform(), hole.uniform(), det.uniform(), flavour=FLAVOUR[mode])
    same = bits(rgba) == bits(ora.rgba)
    assert same.all(), f"{case}: {(~same).mean():.3%} of RGBA words differ"
    assert np.array_equal(hit, ora.hit) and np.array_equal(steps, ora.steps)
    for k in ("steps", "px_traced", "tex_samples", "rk_reject", "stack_overflow"):
        assert st[k] == ora.counters[k], (k, st[k], ora.counters[k])
    for k in ("node_visits", "tri_tests"):              # the BVH walk is bounded by the sphere hit (trace_model): it may do LESS work
        assert st[k] <= ora.counters[k] and (st[k] > 0) == (ora.counters[k] > 0), (k, st[k], ora.counters[k])


# ---------------------------------------------------------------------------------------------------------------------
# The whole kernels as one warp of 32 host threads in lock-step (tests/host_kernel/host_warp.cpp): persistent work loop,
# phase sorting, batched disk shading, narrow work items, classification + ballot-compacted trace queue, sky resolve.
SRC_WARP = os.path.join(ROOT, "tests", "host_kernel", "host_warp.cpp")
OUT_WARP = os.path.join(ROOT, "tests", "host_kernel", "_build", "libbh_host_warp.so")


@pytest.fixture(scope="session")
def host_warp():
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.isdir(cuda_inc):
        pytest.skip("CUDA headers not found")
    deps = DEPS + [SRC_WARP]
    if not os.path.exists(OUT_WARP) or any(os.path.getmtime(d) > os.path.getmtime(OUT_WARP) for d in deps):
        os.makedirs(os.path.dirname(OUT_WARP), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        cmd = [cxx, "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-march=x86-64-v3",
               "-I", cuda_inc, "-x", "c++", SRC_WARP, "-o", OUT_WARP]
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[-4000:]
    lib = C.CDLL(OUT_WARP)
    lib.bh_host_warp_level.restype = C.c_int
    return lib


def warp_level(lib, mode, tex, blob, w, h, cam, hole, det, prev=None, tile_rows=4, want_sky=False, tiling=None, frame=None):
    """tiling = (band_rows, rank, n_ranks): this rank's cyclic bands only, band-major outputs of local_rows rows; with `frame`
    (a full h x w x 4 array) the pixels go to their global rows of it instead (the peer-store exchange)."""
    from bhusie_b200.multi import BandLayout
    rows = h if tiling is None else BandLayout(h, tiling[0], tiling[2]).local_rows(tiling[1])
    rgba = np.zeros((rows, w, 4), np.float32) if frame is None else frame
    hit = np.full((rows, w), -7, np.int32)
    steps = np.full((rows, w), 12345, np.uint32)
    cls = np.full((rows, w), 9, np.uint8)
    stats = np.zeros(9, np.uint64)
    sky = np.zeros((h, w, 4), np.float32) if want_sky else None
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    t = {k: np.ascontiguousarray(tex[k]) for k in ("color", "disk", "sky")}
    prev = None if prev is None else np.ascontiguousarray(prev, dtype=np.float32)
    ph, pw = (0, 0) if prev is None else prev.shape[:2]
    rc = lib.bh_host_warp_level(C.c_int(mode), C.c_char_p(cam.uniform()), C.c_char_p(hole.uniform()), C.c_char_p(det.uniform()),
                                p(t["color"]), C.c_int(t["color"].shape[1]), C.c_int(t["color"].shape[0]),
                                p(t["disk"]), C.c_int(t["disk"].shape[1]), C.c_int(t["disk"].shape[0]),
                                p(t["sky"]), C.c_int(t["sky"].shape[1]), C.c_int(t["sky"].shape[0]),
                                p(blob), C.c_int(w), C.c_int(h), p(prev), C.c_int(pw), C.c_int(ph), C.c_int(tile_rows),
                                p(rgba), p(hit), p(steps), p(cls), p(stats), p(sky),
                                *(C.c_int(v) for v in ((0, 0, 0, h, 0) if tiling is None else (tiling[0], tiling[1], tiling[2], rows, int(frame is not None)))))
    assert rc == 0
    return rgba, hit, steps, cls, dict(zip(STAT_NAMES, (int(x) for x in stats))), sky


def _assert_level(what, got, ora):
    rgba, hit, steps, cls, st = got
    same = bits(rgba) == bits(ora.rgba)
    assert same.all(), f"{what}: {(~same).mean():.3%} of RGBA words differ"
    assert np.array_equal(hit, ora.hit), what
    assert np.array_equal(steps, ora.steps), what
    assert np.array_equal(cls, ora.cls), what
    for k in STAT_NAMES[:7] + ("rk_reject", "stack_overflow"):
        if k in ("node_visits", "tri_tests"):       # the BVH walk is bounded by the sphere hit (trace_model): it may do LESS work
            assert st[k] <= ora.counters[k] and (st[k] > 0) == (ora.counters[k] > 0), (what, k, st[k], ora.counters[k])
        else:
            assert st[k] == ora.counters[k], (what, k, st[k], ora.counters[k])


@pytest.mark.parametrize("mode", [0, 1], ids=["literal", "fused"])
@pytest.mark.parametrize("method", [0, 1], ids=["euler", "rk"])
def test_kernels_as_a_lockstep_warp_pyramid_and_sky(host_warp, oracle, small_scene, small_oracle_scene, mode, method):
    """Three pyramid levels (S_n = 3 S_{n-1} - 2) + sky resolve through the real trace_kernel / classify_kernel / sky_kernel
    control flow, each level fed with the previous one's output, against the oracle: pixels, hit indices, step counts,
    pixel classes and pass statistics bit for bit."""
    tex, blob, _ = small_scene
    cam, hole = U.Camera(), U.BlackHole()
    det = U.RayDetails(integration_method=method, model_count=1, time=0.5, angle_division_threshold=0.08)
    prev_dev, prev_ora = None, None
    sizes = U.pyramid_levels(base=(12, 7), iters=3)
    for li, (w, h) in enumerate(sizes):
        last = li == len(sizes) - 1
        rgba, hit, steps, cls, st, sky = warp_level(host_warp, mode, tex, blob, w, h, cam, hole, det, prev=prev_dev, want_sky=last)
        ora = oracle.ray_pass(small_oracle_scene, w, h, cam.uniform(), hole.uniform(), det.uniform(), prev=prev_ora, flavour=FLAVOUR[mode])
        if last:                                      # the sky kernel added its texture samples to the same counters
            o32, _, sc = oracle.sky_pass(small_oracle_scene, ora.rgba, flavour=FLAVOUR[mode])
            st = dict(st, tex_samples=st["tex_samples"] - sc["tex_samples"])
            assert np.array_equal(bits(sky), bits(o32)), "sky resolve"
        _assert_level(f"level {li} {w}x{h}", (rgba, hit, steps, cls, st), ora)
        if li > 0:
            assert st["px_copied"] > 0 and st["px_interp"] + st["px_traced"] > 0
        prev_dev, prev_ora = rgba, ora.rgba


@pytest.mark.parametrize("thr", [0.02, 0.001, 0.0123, 0.5, 1.5707963, 3.0, 0.0005, 3.1, 0.0, -1.0, float("nan")])
def test_classification_angle_shortcut_equals_literal_acos(host_warp, thr):
    """classify_kernel decides `acos(c) < angle_division_threshold` from c alone unless c is within a few ulps of cos(threshold)
    (derive_pass_constants' cos_hi / cos_lo): same truth value as the literal acos for every probed cosine, and for thresholds in
    the supported range nearly all of them are decided without it."""
    host_warp.bh_host_angle_shortcut_mismatches.restype = C.c_long
    n_lit = C.c_long(0)
    bad = host_warp.bh_host_angle_shortcut_mismatches(C.c_float(thr), 4096, C.byref(n_lit))
    assert bad == 0
    if 1e-3 <= thr <= 0.6:
        assert n_lit.value <= 64                     # the bracket around cos(thr): 3 f32 values at the reference's 0.02
    elif not (1e-3 <= thr <= 3.0):
        assert n_lit.value > 200000                  # outside the supported range the shortcut is off


@pytest.mark.parametrize("tile_rows", [4, 2, 1])
@pytest.mark.parametrize("case", ["default", "outside_sphere", "near_hole", "grazing_plane"])
def test_kernels_as_a_lockstep_warp_single_level(host_warp, oracle, small_scene, small_oracle_scene, tile_rows, case):
    """Base level in tile mode with 32, 16 and 8 rays per warp (the narrow work items of launches that do not fill the GPU),
    ragged frame edges included."""
    tex, blob, _ = small_scene
    ck, hk, dk = CASES[case]
    cam, hole = U.Camera(**ck), U.BlackHole(**hk)
    det = U.RayDetails(integration_method=1, model_count=1, time=1.25, **dk)
    w, h = 21, 10
    rgba, hit, steps, cls, st, _ = warp_level(host_warp, 1, tex, blob, w, h, cam, hole, det, tile_rows=tile_rows)
    ora = oracle.ray_pass(small_oracle_scene, w, h, cam.uniform(), hole.uniform(), det.uniform(), flavour="fused")
    _assert_level(f"{case} rows {tile_rows}", (rgba, hit, steps, cls, st), ora)


@pytest.mark.parametrize("band,world", [(4, 3), (5, 2)])
def test_kernels_as_a_lockstep_warp_tiled_ranks(host_warp, oracle, small_scene, small_oracle_scene, band, world):
    """Image-space sharding inside the kernels: every rank traces only its cyclic row bands — once into its
    compact band-major buffer (the NCCL-gather exchange), once straight into the global rows of one shared frame (the
    peer-store exchange) — and the assembled frame is the single-rank frame, bit for bit, on a two-level pyramid."""
    from bhusie_b200.multi import BandLayout
    tex, blob, _ = small_scene
    cam, hole = U.Camera(), U.BlackHole()
    det = U.RayDetails(integration_method=1, model_count=1, angle_division_threshold=0.08)
    (w0, h0), (w1, h1) = U.pyramid_levels(base=(11, 8), iters=2)
    ora0 = oracle.ray_pass(small_oracle_scene, w0, h0, cam.uniform(), hole.uniform(), det.uniform(), flavour="fused")
    ora1 = oracle.ray_pass(small_oracle_scene, w1, h1, cam.uniform(), hole.uniform(), det.uniform(), prev=ora0.rgba, flavour="fused")
    for (w, h, prev, ora) in ((w0, h0, None, ora0), (w1, h1, ora0.rgba, ora1)):
        lay = BandLayout(h, band, world)
        shared = np.zeros((h, w, 4), np.float32)
        gathered = np.zeros((h, w, 4), np.float32)
        steps_total = 0
        for rank in range(world):
            rgba, hit, steps, cls, st, _ = warp_level(host_warp, 1, tex, blob, w, h, cam, hole, det, prev=prev, tiling=(band, rank, world))
            rows = lay.rows_of(rank)
            gathered[rows] = rgba
            assert np.array_equal(hit, ora.hit[rows]) and np.array_equal(steps, ora.steps[rows]) and np.array_equal(cls, ora.cls[rows])
            steps_total += st["steps"]
            warp_level(host_warp, 1, tex, blob, w, h, cam, hole, det, prev=prev, tiling=(band, rank, world), frame=shared)
        assert steps_total == ora.counters["steps"]
        assert np.array_equal(bits(gathered), bits(ora.rgba)) and np.array_equal(bits(shared), bits(ora.rgba))


# ---------------------------------------------------------------------------------------------------------------------
# Post chain: csrc/post_impl.cuh thread by thread over its launch grid (tests/host_kernel/host_post.cpp)
SRC_POST = os.path.join(ROOT, "tests", "host_kernel", "host_post.cpp")
OUT_POST = os.path.join(ROOT, "tests", "host_kernel", "_build", "libbh_host_post.so")


@pytest.fixture(scope="session")
def host_post():
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.isdir(cuda_inc):
        pytest.skip("CUDA headers not found")
    deps = DEPS + [SRC_POST, os.path.join(ROOT, "bhusie_b200", "csrc", "post_impl.cuh")]
    if not os.path.exists(OUT_POST) or any(os.path.getmtime(d) > os.path.getmtime(OUT_POST) for d in deps):
        os.makedirs(os.path.dirname(OUT_POST), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        cmd = [cxx, "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-march=x86-64-v3",
               "-I", cuda_inc, "-x", "c++", SRC_POST, "-o", OUT_POST]
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[-4000:]
    lib = C.CDLL(OUT_POST)
    lib.bh_host_post_pass.restype = C.c_int
    return lib


def post_pass(lib, mode, kind, in1, out_w, out_h, in2=None, mix_ratio=0.0, fxaa=None):
    in1 = np.ascontiguousarray(in1, dtype=np.uint16)
    in2 = None if in2 is None else np.ascontiguousarray(in2, dtype=np.uint16)
    out = np.zeros((out_h, out_w, 4), np.uint8 if kind == 4 else np.uint16)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = lib.bh_host_post_pass(C.c_int(mode), C.c_int(kind), p(in1), C.c_int(in1.shape[1]), C.c_int(in1.shape[0]), p(in2),
                               p(out), C.c_int(out_w), C.c_int(out_h), C.c_float(mix_ratio), None if fxaa is None else C.c_char_p(fxaa))
    assert rc == 0
    return out


@pytest.mark.parametrize("mode", [0, 1], ids=["literal", "fused"])
def test_post_chain_source_on_cpu_matches_oracle(host_post, host_warp, oracle, small_scene, small_oracle_scene, mode):
    """Every stage of the post chain — 10 bloom levels, mix, ACES, FXAA — computed by the product's kernels on the CPU from
    the oracle's previous stage, equal to the oracle's stage bit for bit (RGBA16F bits / RGBA8 bytes)."""
    from bhusie_b200.post import FXAADetails, bloom_sizes
    tex, blob, _ = small_scene
    cam, hole = U.Camera(), U.BlackHole()
    det = U.RayDetails(integration_method=1, model_count=1)
    w, h = 76, 43
    ray = oracle.ray_pass(small_oracle_scene, w, h, cam.uniform(), hole.uniform(), det.uniform(), flavour=FLAVOUR[mode])
    _, sky16, _ = oracle.sky_pass(small_oracle_scene, ray.rgba, flavour=FLAVOUR[mode])
    sizes = bloom_sizes((w, h))
    fx = FXAADetails().uniform()
    ora = oracle.post_chain(sky16, sizes, 0.7, fx, flavour=FLAVOUR[mode])
    cur = sky16
    n = len(sizes) // 2
    for i, (bw, bh_) in enumerate(sizes):
        got = post_pass(host_post, mode, 0 if i < n else 1, cur, bw, bh_)
        assert np.array_equal(got, ora["bloom"][i]), f"bloom level {i} ({bw}x{bh_})"
        cur = ora["bloom"][i]
    assert np.array_equal(post_pass(host_post, mode, 2, sky16, w, h, in2=cur, mix_ratio=0.7), ora["mix"]), "mix"
    assert np.array_equal(post_pass(host_post, mode, 3, ora["mix"], w, h), ora["hdr"]), "hdr"
    assert np.array_equal(post_pass(host_post, mode, 4, ora["hdr"], w, h, fxaa=fx), ora["fxaa"]), "fxaa"
