"""The multi-GPU entry points of the C ABI (include/bh_abi.h "multi-GPU") exercised on whatever the box has: with ONE device
every code path still runs (bh_frame_multi over 1 device, the flag signal / wait pair on one stream, the tiled ranks of a
shared host frame one after the other); with >= 2 devices the same tests use them.  All results are compared bit for bit
with the plain single-pipeline render."""
import os

import numpy as np
import pytest

from conftest import bits
from bhusie_b200 import _lib, pipelines as P, uniforms as U

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def ctxs(small_scene):
    tex, blob, _ = small_scene
    out = []
    for d in range(min(_devices(), 4)):
        c = P.Context(d)
        c.set_textures(tex)
        c.upload_models(blob)
        out.append(c)
    yield out
    for c in out:
        c.close()


@pytest.mark.parametrize("method", [0, 1])
def test_frame_multi_pyramid_matches_single_device(ctxs, method):
    """bh_frame_multi (coarse levels replicated, last level tiled into device 0's frame, sky on device 0) == RayPyramid."""
    cam, hole = U.Camera(), U.BlackHole()
    det = U.RayDetails(integration_method=method, model_count=1, angle_division_threshold=0.05)
    pyr = P.RayPyramid(ctxs[0], base=(32, 18), iters=3, sky_format=P.SKY_RGBA16F)
    pyr.pass_(cam, hole, det)
    ref_rgba, ref_sky = pyr.levels[-1].read(aux=False)["rgba"], pyr.sky.read()
    ref_stats = [rp.stats() for rp in pyr.levels]
    for n in sorted({1, len(ctxs)}):
        fm = P.FrameMulti(ctxs[:n], base=(32, 18), iters=3, band_rows=4, sky_format=P.SKY_RGBA16F)
        assert (fm.width, fm.height) == (280, 154)
        for _ in range(2):                                    # twice: the second pass waits for the first frame's consumers
            fm.pass_(cam, hole, det)
        got = fm.read()
        assert np.array_equal(bits(got["rgba"]), bits(ref_rgba)), f"{n} device(s): assembled frame differs"
        assert np.array_equal(got["sky"].view(np.uint16), ref_sky.view(np.uint16)), f"{n} device(s): sky differs"
        st = fm.stats()
        for k in ("ray_steps", "px_traced", "px_copied", "px_interp", "tex_samples", "node_visits", "tri_tests"):
            assert st[k] == sum(s[k] for s in ref_stats), k
        assert st["elapsed_ms"] > 0
        fm.close()
    pyr.close()


def test_frame_multi_single_level_and_host_frame(ctxs):
    import torch
    cam, hole, det = U.Camera(), U.BlackHole(), U.RayDetails(integration_method=1, model_count=1)
    w, h = 120, 67
    single = P.RayPipeline(ctxs[0], w, h)
    single.pass_(cam, hole, det)
    ref = single.read(aux=False)["rgba"]
    single.close()
    fm = P.FrameMulti(ctxs, base=(w, h), iters=1, band_rows=5, sky_format=None)
    fm.pass_(cam, hole, det)
    assert np.array_equal(bits(fm.read(sky=False)["rgba"]), bits(ref))
    host = torch.zeros((h, w, 4), dtype=torch.float32).pin_memory()
    fm.pass_to_host(cam, hole, det, host.data_ptr())          # every device stores its bands straight into the host frame
    fm.sync()
    assert np.array_equal(bits(host.numpy()), bits(ref))
    with pytest.raises(_lib.BhError):
        fm.read(sky=False)                                    # the device frame holds nothing of a pass_to_host frame
    fm.pass_(cam, hole, det)
    assert np.array_equal(bits(fm.read(sky=False)["rgba"]), bits(ref))
    fm.close()


def test_frame_multi_rejects_bad_arguments(ctxs):
    with pytest.raises(_lib.BhError):
        P.FrameMulti([ctxs[0], ctxs[0]], base=(32, 18), iters=1)            # same device twice
    with pytest.raises(_lib.BhError):
        P.FrameMulti(ctxs[:1], base=(1, 18), iters=1)
    with pytest.raises(_lib.BhError):
        P.FrameMulti(ctxs[:1], base=(32, 18), iters=0)


def test_stream_flags_signal_wait_and_timeout(ctxs):
    """bh_stream_signal / bh_stream_wait on one device: a wait whose flags were signalled passes; one that is never
    signalled gives up after its timeout instead of hanging the GPU, and bh_ctx_check_async reports it once."""
    import torch
    ctx = ctxs[0]
    nbytes = 64 * 36 * 16
    ptr, _ = ctx.shared_frame_create(nbytes)
    flags = ctx.shared_frame_flags(ptr, nbytes)
    assert flags >= ptr + nbytes and flags % 256 == 0
    s = torch.cuda.Stream()
    for seq in (1, 2, 3):
        for slot in (1, 2, 3):
            ctx.stream_signal(flags + 4 * slot, seq, s)
        ctx.stream_wait(flags + 4, 3, seq, 2000, s)
    s.synchronize()
    ctx.check_async()
    ctx.stream_wait(flags + 4, 3, 2, 2000, s)                # already past: wrap-safe ">="
    s.synchronize()
    ctx.check_async()
    ctx.stream_wait(flags + 4 * 7, 1, 1, 50, s)              # flag 7 is never written: 50 ms, then the kernel gives up
    s.synchronize()
    with pytest.raises(_lib.BhError) as e:
        ctx.check_async()
    assert e.value.code == -110
    ctx.check_async()                                        # reported once
    ctx.shared_frame_release(ptr, owner=True)


def test_shared_host_frame_tiled_ranks(ctxs):
    """bh_host_frame + bh_ray_pipeline_pass_to_host_frame: the ranks of a tiled frame (here one after the other, on as many
    devices as the box has) store their bands at their global rows into ONE page-locked frame in POSIX shared memory."""
    cam, hole, det = U.Camera(), U.BlackHole(), U.RayDetails(integration_method=1, model_count=1)
    w, h, world, band = 97, 53, 3, 4
    single = P.RayPipeline(ctxs[0], w, h)
    single.pass_(cam, hole, det)
    ref = single.read(aux=False)["rgba"]
    name = f"/bhtest_{os.getpid()}"
    owner = P.HostFrame(ctxs[0], name, w * h * 16, create=True)
    frame = owner.array((h, w, 4))
    frame[:] = -1.0
    attached = []
    for rank in range(world):
        ctx = ctxs[rank % len(ctxs)]
        hf = owner if rank == 0 else P.HostFrame(ctx, name, w * h * 16, create=False)      # a second mapping of the same segment
        attached.append(hf)
        rp = P.RayPipeline(ctx, w, h)
        rp.set_tiling(band, rank, world)
        rp.pass_to_host_frame(cam, hole, det, hf.ptr)
        rp.sync()
        hf.signal(rank, 1)
        with pytest.raises(_lib.BhError):
            rp.read(aux=False)                                 # host-only pass: the device buffer was not written
        rp.close()
    owner.wait(0, world, 1, 1000)
    assert np.array_equal(bits(frame), bits(ref))
    with pytest.raises(_lib.BhError) as e:
        owner.wait(world, 1, 1, 20)                            # nobody signals slot `world`
    assert e.value.code == -110
    for hf in attached[1:]:
        hf.close()
    owner.close()
    single.close()
    with pytest.raises(_lib.BhError):
        P.HostFrame(ctxs[0], name, w * h * 16, create=False)   # unlinked by its owner


def test_two_frames_in_flight_on_one_device(ctxs, small_scene):
    """bench.py's e2e pattern: two frame slots on one device — each its own context (model buffer), stream and host frame —
    with slot B's upload + pass enqueued before slot A's frame is finished and read.  Every frame of both slots equals the
    plain device render (HostTiledFrame.enqueue / finish, bh_ctx_upload_models_async, bh_ray_pipeline_pass_to_host_frame)."""
    import torch
    from bhusie_b200.multi import HostTiledFrame
    tex, blob, _ = small_scene
    hole, det = U.BlackHole(), U.RayDetails(integration_method=1, model_count=1)
    cams = [U.Camera(), U.Camera(position=(0.5, 0.2, -17.0)), U.Camera(position=(-1.0, 0.0, -21.0)), U.Camera(position=(0, 1.5, -19.0))]
    w, h = 101, 47
    single = P.RayPipeline(ctxs[0], w, h)
    refs = []
    for cam in cams:
        single.pass_(cam, hole, det)
        refs.append(single.read(aux=False)["rgba"].copy())
    pinned = torch.from_numpy(np.ascontiguousarray(blob)).pin_memory()
    slots = []
    for i in range(2):
        c = P.Context(ctxs[0].device)
        c.set_textures(tex)
        c.upload_models(blob)
        slots.append((c, torch.cuda.Stream(), HostTiledFrame(c, w, h, 0, 1, name=f"/bhtest2_{os.getpid()}_{i}")))
    pending = []
    got = []
    for k, cam in enumerate(cams):
        c, st, hf = slots[k % 2]
        c.upload_models_async(pinned.data_ptr(), pinned.numel(), st)
        hf.enqueue(cam, hole, det, st)
        pending.append((k, hf))
        if len(pending) == 2:
            j, f = pending.pop(0)
            f.finish()
            got.append((j, f.frame_array().copy()))
            f.consumed()
    while pending:
        j, f = pending.pop(0)
        f.finish()
        got.append((j, f.frame_array().copy()))
        f.consumed()
    for j, frame in got:
        assert np.array_equal(bits(frame), bits(refs[j])), f"frame {j}"
    for c, _, hf in slots:
        hf.close()
        c.close()
    single.close()


def test_zero_copy_pass_marks_output_host_only(ctxs):
    """after bh_ray_pipeline_pass_to_host(n_chunks=0) the device buffer holds nothing of the pass, so read(),
    a child level and the sky pass must refuse it until a normal pass runs."""
    import torch
    ctx = ctxs[0]
    cam, hole, det = U.Camera(), U.BlackHole(), U.RayDetails(integration_method=1, model_count=1)
    l0 = P.RayPipeline(ctx, 32, 18)
    l1 = P.RayPipeline(ctx, 94, 52, l0)
    sky = P.SkyPipeline(ctx, l0, P.SKY_RGBA32F)
    host = torch.zeros((18, 32, 4), dtype=torch.float32).pin_memory()
    l0.pass_to_host(cam, hole, det, host.data_ptr(), 0)
    l0.sync()
    for bad in (lambda: l0.read(aux=False), lambda: l1.pass_(cam, hole, det), lambda: sky.pass_()):
        with pytest.raises(_lib.BhError) as e:
            bad()
        assert e.value.code == -1
    l0.pass_(cam, hole, det)
    assert np.array_equal(bits(l0.read(aux=False)["rgba"]), bits(host.numpy()))
    l1.pass_(cam, hole, det)
    sky.pass_()
    sky.close(); l1.close(); l0.close()


def test_upload_rejects_malformed_model(ctxs, small_scene):
    """bh_ctx_upload_models validates every index the kernel would follow."""
    _, blob, _ = small_scene
    P.validate_model(blob)
    from bhusie_b200.uniforms import MU_LOOKUP, MU_NODES, MU_TRIANGLES
    tri0 = int(blob[MU_LOOKUP:MU_LOOKUP + 4].view(np.int32)[0])      # a triangle the traversal can reach (first entry of the lookup)
    cases = {"root children beyond the node array": (MU_NODES + 12, 10 ** 6),
             "root is its own child (endless traversal)": (MU_NODES + 12, 0),
             "point index beyond the array": (MU_TRIANGLES + 24 * tri0, 600000),
             "negative normal index": (MU_TRIANGLES + 24 * tri0 + 12, -5),
             "lookup entry beyond the triangle array": (MU_LOOKUP, 1 << 20)}
    for what, (off, val) in cases.items():
        bad = blob.copy()
        bad[off:off + 4].view(np.int32)[0] = val
        with pytest.raises(_lib.BhError) as e:
            ctxs[0].upload_models(bad)
        assert e.value.code == -22, what
    ctxs[0].upload_models(blob)
