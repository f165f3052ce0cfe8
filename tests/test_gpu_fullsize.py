"""Full-frame parity at BASELINE.json's sizes (VERDICT r1 "next" #1): every pixel of

  C2(i)   1920x1080 Cash-Karp RK, disk, single level
  C2(ii)  the reference's adaptive grid 72x41 -> 214x121 -> 640x361 -> 1918x1081 (mod.rs:177-206) + the Rgba16Float sky resolve
  C3      3840x2160 RK, disk + relativity sphere + lucy.obj (99 970 triangles)
  C4      7680x4320 as 8 tiled ranks on one device, >= 40 sampled rows (the oracle would need minutes for the whole frame)

is compared with the oracle: BIT-EXACT against the flavour the kernel's numeric mode mirrors (RGBA, hit indices, step counts,
classes, statistics), and within the north star's 1e-4 against the neutral strict (glibc) flavour on all but a stated
fraction of pixels, the float64 shadow classifying those outliers.  Thresholds are 2x the values measured by
tools/parity_report.py (profiles/r2_parity_report.json).  Reports of this run are written to gpurun_out/parity_fullsize.json.
"""
import json
import os

import numpy as np
import pytest

from conftest import bits, ROOT
from bhusie_b200 import assets, pipelines as P, uniforms as U
from bhusie_b200.multi import BandLayout

pytestmark = pytest.mark.gpu

TOL = 1e-4
# 2x the measured flavour-vs-strict outlier fractions of profiles/r2_parity_report.json (tools/parity_report.py)
# (worst of the configurations tested here: LITERAL 1.16e-4 on the 214x121 level, FUSED 1.02e-3 on the 72x41 level)
MAX_OUTLIER_FRAC = {"literal": 2.4e-4, "fused": 2.1e-3}
# outliers on pixels the float64 shadow calls WELL-conditioned — |strict - shadow| <= 1e-4 AND a 1e-6 rad rotation of the camera
# ray moves the float64 result by <= 1e-4: what is left unexplained (measured: LITERAL <= 4.3e-6, FUSED <= 2.8e-5)
MAX_WELL_CONDITIONED_OUTLIER_FRAC = {"literal": 1e-5, "fused": 6e-5}
MODES = [pytest.param(P.NUMERIC_LITERAL, id="literal"), pytest.param(P.NUMERIC_FUSED, id="fused")]
MODE_NAME = {P.NUMERIC_LITERAL: "literal", P.NUMERIC_FUSED: "fused"}

_REPORT = {}


def _dump():
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_fullsize.json"), "w") as f:
        json.dump(_REPORT, f, indent=1)


@pytest.fixture(scope="module")
def scene(oracle):
    tex, src = assets.load_textures()
    if not assets.have_lucy():
        pytest.skip("lucy.obj not staged")
    blob, info = P.load_obj_model(assets.lucy_path())
    osc = oracle.OracleScene(tex["color"], tex["disk"], tex["sky"], blob)
    return tex, blob, osc


@pytest.fixture(scope="module", params=MODES)
def ctx(request, scene):
    tex, blob, _ = scene
    c = P.Context(0, numeric_mode=request.param)
    c.set_textures(tex)
    c.upload_models(blob)
    yield c
    c.close()


_NEUTRAL = {}


def neutral(oracle, osc, key, w, h, cam, hole, det, prev=None):
    """strict, float64 shadow and perturbed shadow of one frame, computed once per configuration (they do not depend on the
    kernel's mode)"""
    if key not in _NEUTRAL:
        _NEUTRAL[key] = (oracle.ray_pass(osc, w, h, cam, hole, det, prev=prev, flavour="strict"),
                         oracle.ray_pass(osc, w, h, cam, hole, det, prev=prev, flavour="shadow"),
                         oracle.ray_pass(osc, w, h, cam, hole, det, prev=prev, flavour="shadow", perturb=oracle.SHADOW_PERTURBATION))
    return _NEUTRAL[key]


def check_frame(oracle, osc, ctx, name, dev, st, w, h, cam, hole, det, prev=None, neutral_key=None):
    mode = MODE_NAME[ctx.numeric_mode]
    ora = oracle.ray_pass(osc, w, h, cam, hole, det, prev=prev, flavour=P.ORACLE_FLAVOUR_OF_MODE[ctx.numeric_mode])
    assert np.array_equal(bits(dev["rgba"]), bits(ora.rgba)), f"{name}: RGBA not bit-exact ({(bits(dev['rgba']) != bits(ora.rgba)).mean():.3%} words differ)"
    assert np.array_equal(dev["hit"], ora.hit), f"{name}: hit indices differ"
    assert np.array_equal(dev["steps"], ora.steps), f"{name}: step counts differ"
    if "cls" in dev:
        assert np.array_equal(dev["cls"], ora.cls), f"{name}: classes differ"
    c = ora.counters
    assert (st["ray_steps"], st["px_traced"], st["px_copied"], st["px_interp"]) == (c["steps"], c["px_traced"], c["px_copied"], c["px_interp"])
    assert st["tex_samples"] == c["tex_samples"]
    assert st["node_visits"] <= c["node_visits"] and st["tri_tests"] <= c["tri_tests"]      # sphere-bounded BVH walk: never more work
    assert st["rk_reject"] == 0 and st["stack_overflow"] == 0
    if neutral_key is not None:
        strict, shadow, probe = neutral(oracle, osc, neutral_key, w, h, cam, hole, det, prev)
        rep = oracle.parity_report(dev["rgba"], strict.rgba, shadow.rgba, TOL, probe.rgba)
        rep["hit_index_equal_frac"] = float((dev["hit"] == strict.hit).mean())
        rep["step_count_equal_frac"] = float((dev["steps"] == strict.steps).mean())
        _REPORT[f"{name}/{mode}"] = rep
        _dump()
        # (small levels: a handful of pixels is a large fraction of 3 000, so the fractions apply from ~100 000 pixels up and
        #  a fixed count below: 32 outliers, 8 of them on well-conditioned pixels)
        n_px = rep["pixels"]
        assert rep["outliers"] <= max(MAX_OUTLIER_FRAC[mode] * n_px, 32), (name, rep)
        assert rep["outlier_frac_well_conditioned"] * n_px <= max(MAX_WELL_CONDITIONED_OUTLIER_FRAC[mode] * n_px, 8), (name, rep)
        # north star: hit indices bit-exact — they are, vs the mode's flavour; vs libm a few edge pixels flip
        assert (1.0 - rep["hit_index_equal_frac"]) * n_px <= max(5e-4 * n_px, 16), (name, rep)
        assert (1.0 - rep["step_count_equal_frac"]) * n_px <= max(1e-3 * n_px, 48), (name, rep)
    return ora


def test_c2i_full_frame(ctx, scene, oracle):
    """BASELINE configs[1], sub-run (i): 1920x1080 RK, accretion disk, every pixel traced."""
    _, _, osc = scene
    cam, hole, det = U.Camera(), U.BlackHole(), U.RayDetails(integration_method=1, model_count=0)
    rp = P.RayPipeline(ctx, 1920, 1080, aux=P.AUX_HIT | P.AUX_STEPS)
    rp.pass_(cam, hole, det)
    check_frame(oracle, osc, ctx, "C2i_1920x1080", rp.read(), rp.stats(), 1920, 1080, cam.uniform(), hole.uniform(), det.uniform(), neutral_key="C2i")
    rp.close()


def test_euler_full_frame_with_mesh(ctx, scene, oracle):
    """1920x1080, the reference's default Euler integrator, disk + relativity sphere + lucy.obj, every pixel traced in tile mode:
    a launch large enough for launch_trace_mode to pick the Euler build with 5 resident CTAs per SM (BH_OCC_EULER) — every
    pixel, hit index, step count and counter against the mode's oracle flavour."""
    _, _, osc = scene
    cam, hole, det = U.Camera(), U.BlackHole(), U.RayDetails(integration_method=0, model_count=1)
    rp = P.RayPipeline(ctx, 1920, 1080, aux=P.AUX_HIT | P.AUX_STEPS)
    rp.pass_(cam, hole, det)
    check_frame(oracle, osc, ctx, "Euler_1920x1080_mesh", rp.read(), rp.stats(), 1920, 1080, cam.uniform(), hole.uniform(), det.uniform())
    rp.close()


@pytest.mark.parametrize("method", [1, 0])
def test_c2ii_reference_pyramid_all_levels(ctx, scene, oracle, method):
    """BASELINE configs[1], sub-run (ii): the reference's own frame — four levels up to 1918x1081 (mod.rs:177-206) and the
    Rgba16Float sky resolve (sky_pipeline.rs:34) — with the Cash-Karp integrator of the config and with Euler, the
    reference's default (mod.rs:116-121).  Every level is checked against the oracle run on the ORACLE's previous level, so
    the comparison of level n does not lean on the device's level n-1."""
    _, _, osc = scene
    # RK: the config as SURVEY §8d states it (disk only); Euler: what bhusie renders out of the box (lucy.obj loaded, scene/mod.rs:23-26)
    cam, hole, det = U.Camera(), U.BlackHole(), U.RayDetails(integration_method=method, model_count=0 if method == 1 else 1)
    pyr = P.RayPyramid(ctx, aux=P.AUX_HIT | P.AUX_STEPS | P.AUX_CLASS, sky_format=P.SKY_RGBA16F)
    assert pyr.sizes == [(72, 41), (214, 121), (640, 361), (1918, 1081)]
    pyr.pass_(cam, hole, det)
    prev = None
    for rp in pyr.levels:
        # (fine levels: the neutral flavours are run on the same previous level as the mode's flavour, so per mode)
        key = f"C2ii_{MODE_NAME[ctx.numeric_mode]}_{rp.width}x{rp.height}" if method == 1 else None
        ora = check_frame(oracle, osc, ctx, f"C2ii_m{method}_{rp.width}x{rp.height}", rp.read(), rp.stats(), rp.width, rp.height,
                          cam.uniform(), hole.uniform(), det.uniform(), prev=prev, neutral_key=key)
        prev = ora.rgba
    last = pyr.levels[-1].stats()
    assert last["px_copied"] + last["px_interp"] + last["px_traced"] == 1918 * 1081 and last["px_interp"] > 10 ** 6
    _, o16, _ = oracle.sky_pass(osc, prev, flavour=P.ORACLE_FLAVOUR_OF_MODE[ctx.numeric_mode])
    assert np.array_equal(pyr.sky.read().view(np.uint16), o16), "sky resolve (Rgba16Float) not bit-exact"
    pyr.close()


def test_c3_full_frame(ctx, scene, oracle):
    """BASELINE configs[2] — the headline workload: 3840x2160 RK, disk + relativity sphere + the 100k-triangle BVH."""
    _, _, osc = scene
    cam, hole, det = U.Camera(), U.BlackHole(), U.RayDetails(integration_method=1, model_count=1)
    rp = P.RayPipeline(ctx, 3840, 2160, aux=P.AUX_HIT | P.AUX_STEPS)
    rp.pass_(cam, hole, det)
    dev, st = rp.read(), rp.stats()
    check_frame(oracle, osc, ctx, "C3_3840x2160", dev, st, 3840, 2160, cam.uniform(), hole.uniform(), det.uniform(), neutral_key="C3")
    assert (dev["hit"] >= 0).sum() > 10000
    rp.close()


@pytest.mark.parametrize("method", [1, 0])
def test_c3_camera_outside_the_sphere_full_frame(ctx, scene, oracle, method):
    """SURVEY §8(d) C3's second camera, (0,0,-45), outside R_rel = 20: by Q3 every Cash-Karp ray ping-pongs between one step and
    one flat-space iteration ~170 times before it is inside (the exit-in-the-tail path and the sphere-bounded BVH walk at full
    scale); Euler enters once.  Whole 3840x2160 frame: pixels, hit indices, step counts and counters against the mode's flavour."""
    _, _, osc = scene
    cam, hole, det = U.Camera(position=(0.0, 0.0, -45.0)), U.BlackHole(), U.RayDetails(integration_method=method, model_count=1)
    rp = P.RayPipeline(ctx, 3840, 2160, aux=P.AUX_HIT | P.AUX_STEPS)
    rp.pass_(cam, hole, det)
    dev, st = rp.read(), rp.stats()
    check_frame(oracle, osc, ctx, f"C3_cam45_m{method}_3840x2160", dev, st, 3840, 2160, cam.uniform(), hole.uniform(), det.uniform())
    assert (dev["hit"] >= 0).sum() > 1000
    rp.close()


def test_c4_8k_as_eight_tiled_ranks(ctx, scene, oracle):
    """BASELINE configs[3]: 7680x4320, the frame cut into cyclic 8-row bands over 8 ranks — here the 8 ranks run one after
    the other on one device, each into its compact band buffer; 5 rows of every rank (40 in all) are compared with the
    oracle, and the per-rank statistics must add up to the per-pixel step counts."""
    _, _, osc = scene
    W, H, world, band = 7680, 4320, 8, 8
    cam, hole, det = U.Camera(), U.BlackHole(), U.RayDetails(integration_method=1, model_count=1)
    lay = BandLayout(H, band, world)
    fl = P.ORACLE_FLAVOUR_OF_MODE[ctx.numeric_mode]
    total_steps = 0
    rng = np.random.default_rng(8)
    for rank in range(world):
        rp = P.RayPipeline(ctx, W, H, aux=P.AUX_HIT | P.AUX_STEPS)
        rp.set_tiling(band, rank, world)
        rp.pass_(cam, hole, det)
        part, st = rp.read(), rp.stats()
        rows = lay.rows_of(rank)
        assert part["rgba"].shape[0] == len(rows) and abs(len(rows) - H // world) <= band
        assert st["ray_steps"] == int(part["steps"].sum(dtype=np.int64)) and st["px_traced"] == len(rows) * W
        total_steps += st["ray_steps"]
        # rows around the hole / disk / mesh (middle of the frame) and two random ones
        mid = np.searchsorted(rows, H // 2)
        picks = sorted({0, len(rows) - 1, int(mid) % len(rows), int(rng.integers(0, len(rows))), int(rng.integers(0, len(rows)))})
        for li in picks:
            y = int(rows[li])
            ora = oracle.ray_pass(osc, W, H, cam.uniform(), hole.uniform(), det.uniform(), rows=(y, y + 1), flavour=fl)
            assert np.array_equal(bits(part["rgba"][li]), bits(ora.rgba[y])), f"rank {rank} row {y}"
            assert np.array_equal(part["hit"][li], ora.hit[y]) and np.array_equal(part["steps"][li], ora.steps[y])
        rp.close()
    _REPORT[f"C4_7680x4320/{MODE_NAME[ctx.numeric_mode]}"] = {"ray_steps": total_steps, "rows_compared": "5 per rank, 8 ranks", "bit_exact": True}
    _dump()
    assert total_steps > 7.5e9
