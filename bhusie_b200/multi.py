"""Image-space sharding of one ray pass across the GPUs of a node.

The reference is single-GPU; pixels of a level are independent (ray.wgsl:167-243), so the frame
is cut into cyclic bands of `band_rows` rows — band b belongs to rank b mod N — which balances
the cost that concentrates around the hole, the disk and the mesh.  Every rank holds the whole
scene (textures + ModelUniform), renders its bands into a compact band-major buffer, and rank 0
receives all of them with ONE gather (NCCL over NVLink) and de-interleaves them into the frame.
There is no other collective on the data path.

Two exchanges are implemented:
  "nccl"  every rank renders into a compact band buffer; rank 0 gathers them (NCCL send/recv over NVLink) and
          de-interleaves with one copy.  The plain-library baseline.
  "p2p"   (default when CUDA IPC is available) rank 0 owns the frame and exports it with CUDA IPC; every other rank
          maps it and its ray kernel stores finished pixels STRAIGHT into rank 0's memory at their global row
          (bh_ray_pipeline_bind_frame) — 16-byte stores over NVLink while the warp keeps tracing, so the exchange is
          fused into the pass and overlaps it completely.  One 4-byte all-reduce per frame orders "all ranks done"
          before rank 0 consumes the frame.
"""
from __future__ import annotations

import numpy as np


class BandLayout:
    """Pure row bookkeeping: which global rows a rank renders, in its local (band-major) order.
    Must match bh_ray_pipeline_set_tiling / global_row() in csrc/ray_kernels.cu."""

    def __init__(self, height: int, band_rows: int, world: int):
        if height < 1 or band_rows < 1 or world < 1:
            raise ValueError("height, band_rows and world must be positive")
        self.height, self.band_rows, self.world = height, band_rows, world

    def rows_of(self, rank: int) -> np.ndarray:
        if not 0 <= rank < self.world:
            raise ValueError("rank out of range")
        y = np.arange(self.height)
        return y[(y // self.band_rows) % self.world == rank]

    def local_rows(self, rank: int) -> int:
        return int(self.rows_of(rank).size)

    @property
    def max_local_rows(self) -> int:
        return max(self.local_rows(r) for r in range(self.world))

    @property
    def uniform(self) -> bool:
        """True when every rank owns the same number of complete bands (single-copy de-interleave)."""
        return self.height % (self.band_rows * self.world) == 0


def gather_bands(local, layout: BandLayout, rank: int, frame=None, staging=None, row_index=None):
    """Gathers every rank's band buffer to rank 0 and scatters the rows into `frame` (H, W, C).

    `local` is (max_local_rows, W, C) on every rank.  Works with any torch.distributed backend
    (nccl on GPUs; gloo in the CPU tests).  Returns `frame` on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist

    world = layout.world
    if world == 1:
        if frame is not None and frame.data_ptr() != local.data_ptr():
            frame.copy_(local[: layout.height])
        return frame if frame is not None else local
    if rank == 0:
        if staging is None:
            staging = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
        dist.gather(local, list(staging.unbind(0)), dst=0)
        if layout.uniform:
            nb = layout.height // (layout.band_rows * world)
            w, c = local.shape[1], local.shape[2]
            frame.view(nb, world, layout.band_rows, w, c).copy_(
                staging[:, : nb * layout.band_rows].view(world, nb, layout.band_rows, w, c).permute(1, 0, 2, 3, 4))
        else:
            for r in range(world):
                idx = row_index[r] if row_index is not None else torch.as_tensor(layout.rows_of(r), device=local.device)
                frame.index_copy_(0, idx, staging[r, : idx.numel()])
        return frame
    dist.gather(local, None, dst=0)
    return None


class _DevicePointer:
    """Minimal __cuda_array_interface__ carrier so torch can view memory the library allocated."""

    def __init__(self, ptr: int, shape: tuple):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 2, "strides": None}


class TiledFrame:
    """One frame of the single-level ray pass on `world` ranks: RayPipeline with cyclic-band tiling, exchanged
    either by NCCL gather ("nccl") or by direct peer stores into rank 0's frame ("p2p")."""

    def __init__(self, ctx, width: int, height: int, rank: int = 0, world: int = 1, band_rows: int = 8, aux: int = 0,
                 exchange: str = "p2p"):
        import torch
        from .pipelines import RayPipeline

        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        self.ctx = ctx
        self.exchange = exchange if world > 1 else "none"
        self.rank, self.world, self.width, self.height = rank, world, width, height
        self.band_rows = band_rows if world > 1 else height
        self.layout = BandLayout(height, self.band_rows, world)
        self.pipeline = RayPipeline(ctx, width, height, aux=aux)
        if world > 1:
            self.pipeline.set_tiling(self.band_rows, rank, world)
        assert self.pipeline.local_rows == self.layout.local_rows(rank)
        dev = torch.device("cuda", ctx.device)
        self.frame = self.staging = self.row_index = self.local = None
        self._shared_ptr = 0
        if self.exchange == "p2p":
            import torch.distributed as dist
            nbytes = height * width * 16
            if rank == 0:
                self._shared_ptr, handle = ctx.shared_frame_create(nbytes)
                h = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
            else:
                h = torch.empty(64, dtype=torch.uint8, device=dev)
            dist.broadcast(h, 0)
            if rank != 0:
                self._shared_ptr = ctx.shared_frame_open(bytes(h.cpu().tolist()))
            self.pipeline.bind_frame(self._shared_ptr)
            self._flag = torch.zeros(1, dtype=torch.int32, device=dev)
            if rank == 0:
                self.frame = torch.as_tensor(_DevicePointer(self._shared_ptr, (height, width, 4)), device=dev)
            return
        self.local = torch.zeros((self.layout.max_local_rows, width, 4), dtype=torch.float32, device=dev)
        self.pipeline.bind_output(self.local.data_ptr())
        if rank == 0 and world > 1:
            self.frame = torch.zeros((height, width, 4), dtype=torch.float32, device=dev)
            self.staging = torch.empty((world,) + tuple(self.local.shape), dtype=torch.float32, device=dev)
            if not self.layout.uniform:
                self.row_index = [torch.as_tensor(self.layout.rows_of(r), device=dev) for r in range(world)]

    def render_local(self, camera, black_hole, details, stream=None):
        self.pipeline.pass_(camera, black_hole, details, stream)

    def gather(self, stream=None):
        """After this (stream-ordered) rank 0's frame holds every rank's rows."""
        if self.exchange == "p2p":
            import torch.distributed as dist
            dist.all_reduce(self._flag)          # 4 bytes: completes on a rank only after every rank's kernel has finished
        elif self.world > 1:
            gather_bands(self.local, self.layout, self.rank, self.frame, self.staging, self.row_index)

    def consumed(self, stream=None):
        """Call (on every rank) after rank 0 has finished reading the frame and before the next render: in p2p mode the
        other ranks write into rank 0's memory, so they must not start the next frame while it is still being read."""
        if self.exchange == "p2p":
            import torch.distributed as dist
            dist.all_reduce(self._flag)

    def close(self):
        self.pipeline.bind_frame(None)
        if self._shared_ptr:
            import torch
            torch.cuda.synchronize()
            self.frame = None
            self.ctx.shared_frame_release(self._shared_ptr, owner=self.rank == 0)
            self._shared_ptr = 0

    def render(self, camera, black_hole, details, stream=None):
        self.render_local(camera, black_hole, details, stream)
        self.gather(stream)

    def frame_tensor(self):
        """Rank 0: the assembled (H, W, 4) RGBA32F frame on the device."""
        return self.local[: self.height] if self.world == 1 else self.frame
