"""Post chain: CPU tests pin the oracle restatement with analytic cases; GPU tests compare every stage
of the CUDA chain with the oracle (bit-exact against the mode's flavour; within quantisation against the libm flavour)."""
import numpy as np
import pytest

from bhusie_b200 import assets, uniforms as U
from bhusie_b200.post import FXAADetails, MixDetails, bloom_sizes


def h16(a):
    return np.ascontiguousarray(np.asarray(a, np.float16))


def test_bloom_sizes():
    assert bloom_sizes((1918, 1081)) == [(959, 540), (479, 270), (239, 135), (119, 67), (59, 33),
                                         (119, 67), (239, 135), (479, 270), (959, 540), (1918, 1081)]     # mod.rs:219-229
    assert len(MixDetails().uniform()) == 4 and len(FXAADetails().uniform()) == 16


@pytest.mark.parametrize("flavour", ["strict", "contract", "fused"])
def test_post_oracle_analytic(oracle, flavour):
    # constant image: every stage is a weighted average with unit weight sum (+ alpha 1)
    c = np.float16([0.25, 0.5, 0.125, 1.0])
    img = np.broadcast_to(c, (36, 64, 4)).copy()
    down = oracle.bloom(img, 32, 18, "down", flavour).view(np.float16)
    assert np.all(down == c)                                          # 0.125 + 4*0.03125 + 4*0.0625 + 4*0.125 = 1
    up = oracle.bloom(down, 64, 36, "up", flavour).view(np.float16)
    assert np.all(up == c)                                            # (4 + 4*2 + 4) / 16 = 1
    mx = oracle.mix(img, up, 0.7, flavour).view(np.float16)
    assert np.abs(mx.astype(np.float32) - c).max() <= 2e-4
    # single bright texel: downsample picks it up with the centre/inner weights only where a tap lands on it
    spike = np.zeros((8, 8, 4), np.float16); spike[..., 3] = 1; spike[3, 3, :3] = 8.0
    d = oracle.bloom(spike, 4, 4, "down", flavour).view(np.float16).astype(np.float32)
    # at an exact 2:1 ratio the centre/outer taps land on odd texels, the inner (+-1) taps on even ones, so texel (3,3) is
    # seen by e of pixel (1,1), by b/d/f/h of its 4-neighbours and by a/c/g/i of its diagonal neighbours
    want = np.zeros((4, 4), np.float32)
    want[1, 1] = 8 * 0.125
    want[0, 1] = want[2, 1] = want[1, 0] = want[1, 2] = 8 * 0.0625
    want[0, 0] = want[0, 2] = want[2, 0] = want[2, 2] = 8 * 0.03125
    assert np.array_equal(d[..., 0], want) and np.array_equal(d[..., 1], want) and np.all(d[..., 3] == 1)
    # ACES: black stays black-ish, large values saturate to 1, monotone in between; alpha passes through
    ramp = np.zeros((1, 64, 4), np.float16)
    ramp[0, :, :3] = np.linspace(0, 16, 64, dtype=np.float32)[:, None]
    ramp[0, :, 3] = 0.5
    t = oracle.hdr(ramp, flavour).view(np.float16).astype(np.float32)
    assert t[0, 0, 0] <= 0.01 and 0.98 < t[0, -1, 0] <= 1.0 and np.all(np.diff(t[0, :, 0]) >= -1e-3) and np.all(t[..., 3] == 0.5)
    # FXAA on a flat image is the identity followed by the sRGB encode
    flat = np.broadcast_to(np.float16([0.2, 0.2, 0.2, 1.0]), (16, 16, 4)).copy()
    f = oracle.fxaa(flat, FXAADetails().uniform(), flavour)
    want = round((1.055 * float(np.float16(0.2)) ** (1 / 2.4) - 0.055) * 255)
    assert np.all(f[..., :3] == want) and np.all(f[..., 3] == 255)
    # sRGB linear toe and clamping
    toe = np.broadcast_to(np.float16([0.002, 2.0, -1.0, 1.0]), (4, 4, 4)).copy()
    f = oracle.fxaa(toe, FXAADetails().uniform(), flavour)
    assert f[0, 0, 0] == round(12.92 * float(np.float16(0.002)) * 255) and f[0, 0, 1] == 255 and f[0, 0, 2] == 0
    # a vertical step edge gets blended by FXAA on the two columns next to it only
    edge = np.zeros((16, 16, 4), np.float16); edge[..., 3] = 1; edge[:, 8:, :3] = 1.0
    f = oracle.fxaa(edge, FXAADetails().uniform(), flavour).astype(int)
    assert np.all(f[:, :6, 0] == 0) and np.all(f[:, 10:, 0] == 255)
    assert (0 < f[8, 7, 0] < 255) or (0 < f[8, 8, 0] < 255)


def _chain_inputs(oracle, scene, w, h):
    cam, hole = U.Camera().uniform(), U.BlackHole().uniform()
    det = U.RayDetails(integration_method=1, model_count=1).uniform()
    return cam, hole, det


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1], ids=["literal", "fused"])
def test_post_chain_gpu_parity(oracle, small_scene, small_oracle_scene, mode):
    from bhusie_b200 import pipelines as P
    from bhusie_b200.post import PostChain
    tex, blob, _ = small_scene
    ctx = P.Context(0, numeric_mode=mode)
    ctx.set_textures(tex)
    ctx.upload_models(blob)
    flavour = P.ORACLE_FLAVOUR_OF_MODE[mode]
    cam, hole, det = U.Camera(), U.BlackHole(), U.RayDetails(integration_method=1, model_count=1)
    w, h = 322, 182
    rp = P.RayPipeline(ctx, w, h)
    sky = P.SkyPipeline(ctx, rp, P.SKY_RGBA16F)
    chain = PostChain(ctx, sky)
    rp.pass_(cam, hole, det)
    sky.pass_()
    chain.pass_()
    sky16 = sky.read()
    ora = oracle.post_chain(sky16, chain.sizes, 0.7, FXAADetails().uniform(), flavour=flavour)
    for i, bp in enumerate(chain.blooms):
        assert np.array_equal(bp.read().view(np.uint16), ora["bloom"][i]), f"bloom {i} {chain.sizes[i]}"
    assert np.array_equal(chain.mix.read().view(np.uint16), ora["mix"])
    assert np.array_equal(chain.hdr.read().view(np.uint16), ora["hdr"])
    dev = chain.read()
    assert np.array_equal(dev, ora["fxaa"])
    # against the neutral libm flavour: 8-bit output within 1 LSB on all but FXAA-branch-flip pixels
    strict = oracle.post_chain(sky16, chain.sizes, 0.7, FXAADetails().uniform(), flavour="strict")["fxaa"]
    d = np.abs(dev.astype(int) - strict.astype(int))
    assert (d > 1).any(axis=2).mean() < 0.02
    chain.close(); sky.close(); rp.close(); ctx.close()


@pytest.mark.gpu
def test_post_chain_reference_size(oracle):
    """The reference's own frame: 4-level pyramid -> sky -> bloom x10 -> mix -> ACES -> FXAA at 1918x1081."""
    from bhusie_b200 import pipelines as P
    from bhusie_b200.post import PostChain
    tex, src = assets.load_textures()
    blob, _ = P.load_obj_model(assets.lucy_path()) if assets.have_lucy() else P.model_from_arrays(*assets.uv_sphere())
    ctx = P.Context(0)
    ctx.set_textures(tex); ctx.upload_models(blob)
    pyr = P.RayPyramid(ctx)
    chain = PostChain(ctx, pyr.sky)
    assert chain.sizes[0] == (959, 540) and chain.sizes[4] == (59, 33)
    pyr.pass_(U.Camera(), U.BlackHole(), U.RayDetails(integration_method=1, model_count=1))
    chain.pass_()
    sky16 = pyr.sky.read()
    ora = oracle.post_chain(sky16, chain.sizes, 0.7, FXAADetails().uniform(), flavour=P.ORACLE_FLAVOUR_OF_MODE[ctx.numeric_mode])
    assert np.array_equal(chain.blooms[-1].read().view(np.uint16), ora["bloom"][-1])
    assert np.array_equal(chain.read(), ora["fxaa"])
    frame = chain.read()
    assert frame[..., 3].min() == 255 and frame[..., :3].max() > 200 and frame[..., :3].mean() < 80      # a mostly dark sky with a bright disk
    chain.close(); pyr.close(); ctx.close()
