"""2-GPU, one-process-per-GPU test of the multi-GPU paths (needs >= 2 CUDA devices; skipped otherwise): the frame assembled
on rank 0 by (a) NCCL gather of band buffers, (b) direct peer stores into rank 0's frame ordered by stream flags, (c) every
rank's stores into one shared page-locked host frame, and (d) the tiled adaptive grid + sky resolve must all be bit-identical
to a single-GPU render.  (tests/test_gpu_frame_multi.py covers the same entry points on a single device.)"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, band, q):
    import torch
    import torch.distributed as dist
    from bhusie_b200 import assets, pipelines as P, uniforms as U
    from bhusie_b200.multi import HostTiledFrame, TiledFrame

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        tex = assets.small_textures()
        blob, _ = P.model_from_arrays(*assets.uv_sphere(12, 16, radius=4.0))
        ctx = P.Context(rank)
        ctx.set_textures(tex)
        ctx.upload_models(blob)
        cam, hole, det = U.Camera(), U.BlackHole(), U.RayDetails(integration_method=1, model_count=1)
        stream = torch.cuda.Stream()                             # NOT the current stream: the exchanges must follow the caller's stream
        frames = {}
        for exchange in ("nccl", "p2p"):
            tf = TiledFrame(ctx, w, h, rank, world, band_rows=band, exchange=exchange)
            for _ in range(3):                                   # several frames: the frame buffer is reused, the flags count up
                tf.render(cam, hole, det, stream)
                if rank == 0:
                    with torch.cuda.stream(stream):
                        frames[exchange] = tf.frame_tensor().clone()
                tf.consumed(stream)
            stream.synchronize()
            ctx.check_async()
            tf.close()
        # the reference's adaptive grid across the ranks: coarse levels replicated, last level tiled, sky on rank 0
        dete = U.RayDetails(integration_method=0, model_count=1, angle_division_threshold=0.05)
        tp = TiledFrame(ctx, rank=rank, world=world, band_rows=4, exchange="p2p",
                        pyramid=dict(base=(32, 18), iters=3, sky_format=P.SKY_RGBA16F))
        for _ in range(2):
            tp.render(cam, hole, dete, stream)
            tp.consumed(stream)
        stream.synchronize()
        ctx.check_async()
        sky_tiled = tp.sky.read() if rank == 0 else None
        tp.close()
        # end to end: every rank stores its bands into one host frame in shared memory
        hf = HostTiledFrame(ctx, w, h, rank, world, band_rows=band, name=f"/bhtest_multi_{port}")
        for _ in range(2):
            hf.render(cam, hole, det, stream)
            host = hf.frame_array().copy() if rank == 0 else None
            hf.consumed()
        hf.close()
        if rank == 0:
            single = P.RayPipeline(ctx, w, h)
            single.pass_(cam, hole, det)
            ref = single.read()["rgba"]
            pyr = P.RayPyramid(ctx, base=(32, 18), iters=3, sky_format=P.SKY_RGBA16F)
            pyr.pass_(cam, hole, dete)
            sky_ref = pyr.sky.read()
            same = lambda a: bool(np.array_equal(np.ascontiguousarray(a).view(np.uint32), ref.view(np.uint32)))
            q.put((same(frames["nccl"].cpu().numpy()), same(frames["p2p"].cpu().numpy()), same(host),
                   bool(np.array_equal(sky_tiled.view(np.uint16), sky_ref.view(np.uint16)))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("w,h,band", [(160, 90, 8), (97, 53, 5)])
def test_two_gpu_exchanges_match_single_gpu(w, h, band):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, w, h, band, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    nccl_ok, p2p_ok, host_ok, pyramid_ok = q.get(timeout=5)
    assert nccl_ok, "NCCL-gathered frame differs from the single-GPU frame"
    assert p2p_ok, "peer-store frame differs from the single-GPU frame"
    assert host_ok, "shared host frame differs from the single-GPU frame"
    assert pyramid_ok, "tiled adaptive grid + sky differs from the single-GPU pyramid"
