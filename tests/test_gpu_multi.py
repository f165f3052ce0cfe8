"""2-GPU test of both multi-GPU exchanges (needs >= 2 CUDA devices; skipped otherwise): the frame assembled on rank 0
by (a) NCCL gather of band buffers and (b) direct peer stores into rank 0's frame must be bit-identical to a
single-GPU render."""
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
    from bhusie_b200.multi import TiledFrame

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
        frames = {}
        for exchange in ("nccl", "p2p"):
            tf = TiledFrame(ctx, w, h, rank, world, band_rows=band, exchange=exchange)
            for _ in range(2):                                   # twice: the frame buffer is reused
                tf.render(cam, hole, det, torch.cuda.current_stream())
                if rank == 0:
                    frames[exchange] = tf.frame_tensor().cpu().numpy().copy()
                tf.consumed()
            torch.cuda.synchronize()
            tf.close()
        if rank == 0:
            single = P.RayPipeline(ctx, w, h)
            single.pass_(cam, hole, det)
            ref = single.read()["rgba"]
            q.put((bool(np.array_equal(frames["nccl"].view(np.uint32), ref.view(np.uint32))),
                   bool(np.array_equal(frames["p2p"].view(np.uint32), ref.view(np.uint32)))))
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
    nccl_ok, p2p_ok = q.get(timeout=5)
    assert nccl_ok, "NCCL-gathered frame differs from the single-GPU frame"
    assert p2p_ok, "peer-store frame differs from the single-GPU frame"
