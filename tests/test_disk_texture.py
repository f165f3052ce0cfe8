"""Disk-texture generator (perlin/src/main.rs) and frame dump (f3, mod.rs:460-486).

THE ONE PINNED ROW: the reference ships the generator's output, src/renderer/textures/disk.png.  The oracle's strict
(glibc) flavour reproduces it pixel for pixel; the committed digest freezes that, and when the asset is staged the
comparison is made against the file itself."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLD
from bhusie_b200 import assets, post


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_oracle_disk_texture_equals_reference_asset(oracle):
    gold = json.load(open(os.path.join(GOLD, "oracle_frames.json")))["disk_texture"]
    strict = oracle.disk_texture(1000, 1000, "strict")
    assert np.array_equal(strict[..., 0], strict[..., 1]) and np.array_equal(strict[..., 0], strict[..., 3])     # r=g=b=a
    assert sha(strict) == gold["strict_sha256"]
    if assets.have_reference_assets():
        ref = assets._decode_png(os.path.join(assets.ASSET_DIR, "disk.png"))
        assert ref.shape == (1000, 1000, 4)
        assert np.array_equal(strict, ref), "oracle (glibc flavour) must reproduce the reference's disk.png exactly"
    contract = oracle.disk_texture(1000, 1000, "contract")
    assert sha(contract) == gold["contract_sha256"]
    d = np.abs(contract.astype(int) - strict.astype(int)).max(axis=2)
    assert (d > 0).sum() <= 20 and d.max() <= 3          # det-math vs glibc: a handful of truncation-boundary texels
    assert sha(oracle.disk_texture(1000, 1000, "fused")) == gold["contract_sha256"]     # no contraction in this row (Rust semantics)


def test_oracle_disk_texture_small_and_errors(oracle):
    a = oracle.disk_texture(64, 64, "strict")
    assert a.shape == (64, 64, 4) and a.std() > 5           # noise, not a constant
    with pytest.raises(RuntimeError):
        oracle.disk_texture(0, 10)


def test_save_png_roundtrip(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(3)
    frame = rng.integers(0, 256, (41, 67, 4), dtype=np.uint8)
    p = str(tmp_path / "frame.png")
    post.save_png(p, frame)                                  # the reference forces alpha to 255 (mod.rs:479)
    back = np.asarray(Image.open(p))
    assert back.shape == (41, 67, 4) and np.array_equal(back[..., :3], frame[..., :3]) and np.all(back[..., 3] == 255)
    post.save_png(p, frame, force_opaque=False)
    assert np.array_equal(np.asarray(Image.open(p)), frame)
    with pytest.raises(Exception):
        post.save_png(str(tmp_path / "no" / "such" / "dir.png"), frame)
    with pytest.raises(ValueError):
        post.save_png(p, frame[..., :3])


@pytest.mark.gpu
def test_gpu_disk_texture(oracle, small_scene):
    from bhusie_b200 import pipelines as P, uniforms as U
    tex, blob, _ = small_scene
    ctx = P.Context(0)
    dev = ctx.generate_disk_texture(1000, 1000, install=False)
    assert np.array_equal(dev, oracle.disk_texture(1000, 1000, "contract")), "device generator != oracle (det-math flavour)"
    strict = oracle.disk_texture(1000, 1000, "strict")
    d = np.abs(dev.astype(int) - strict.astype(int)).max(axis=2)
    assert (d > 0).mean() < 1e-4 and d.max() <= 3
    if assets.have_reference_assets():
        ref = assets._decode_png(os.path.join(assets.ASSET_DIR, "disk.png"))
        assert (np.abs(dev.astype(int) - ref.astype(int)).max(axis=2) > 0).mean() < 1e-4     # vs the reference's own file
    for (w, h) in ((96, 96), (130, 70)):
        assert np.array_equal(ctx.generate_disk_texture(w, h, install=False), oracle.disk_texture(w, h, "contract"))
    # install: a scene rendered with the generated texture equals one rendered with the same texels uploaded by hand
    ctx.set_textures(tex); ctx.upload_models(blob)
    gen = ctx.generate_disk_texture(96, 96, install=True)
    cam, hole, det = U.Camera(), U.BlackHole(), U.RayDetails(integration_method=1, model_count=1)
    a = P.RayPipeline(ctx, 80, 45); a.pass_(cam, hole, det); ra = a.read()["rgba"]
    ctx.set_texture(P.TEX_DISK, gen)
    a.pass_(cam, hole, det)
    assert np.array_equal(a.read()["rgba"].view(np.uint32), ra.view(np.uint32))
    a.close(); ctx.close()
