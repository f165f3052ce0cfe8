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
          fused into the pass and overlaps it completely.  Ordering needs no collective either: after its kernel a rank
          stores the frame number into its flag behind the frame (bh_stream_signal), rank 0 waits on the device until
          all flags reached it (bh_stream_wait), and the same pair tells the writers that rank 0 has consumed a frame.

With `pyramid=` the frame is the reference's adaptive grid (mod.rs:170-216): the coarse levels are replicated on every
rank (12.5 % of the pixels, no halo exchange), the last level is tiled, and rank 0 resolves the sky over the assembled frame.
`HostTiledFrame` is the end-to-end variant: every rank stores its bands straight into ONE page-locked host frame in POSIX
shared memory over its own PCIe link.
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


CONSUMED_SLOT = 48          # flag rank 0 bumps when it has finished reading a frame (ranks use slots 0..world-1)
WAIT_TIMEOUT_MS = 20000


class TiledFrame:
    """One frame of the ray pass on `world` ranks: RayPipeline(s) with cyclic-band tiling of the last level, exchanged
    either by NCCL gather ("nccl") or by direct peer stores into rank 0's frame ("p2p").

    pyramid=None: single level width x height, every pixel traced.  pyramid=dict(base=(72, 41), multiplier=3, iters=4,
    sky_format=...): the reference's adaptive grid; width/height are then the last level's."""

    def __init__(self, ctx, width: int = 0, height: int = 0, rank: int = 0, world: int = 1, band_rows: int = 8, aux: int = 0,
                 exchange: str = "p2p", pyramid: "dict | None" = None):
        import torch
        from .pipelines import RayPipeline, SkyPipeline
        from .uniforms import pyramid_levels

        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        self.ctx = ctx
        self.exchange = exchange if world > 1 else "none"
        sizes = [(width, height)]
        sky_format = None
        if pyramid is not None:
            sizes = pyramid_levels(pyramid.get("base", (72, 41)), pyramid.get("multiplier", 3), pyramid.get("iters", 4))
            sky_format = pyramid.get("sky_format")
            width, height = sizes[-1]
        self.rank, self.world, self.width, self.height = rank, world, width, height
        self.band_rows = band_rows if world > 1 else height
        self.layout = BandLayout(height, self.band_rows, world)
        self.levels = []
        for (w, h) in sizes:
            self.levels.append(RayPipeline(ctx, w, h, self.levels[-1] if self.levels else None, aux=aux))
        self.pipeline = self.levels[-1]                      # the tiled level
        if world > 1:
            self.pipeline.set_tiling(self.band_rows, rank, world)
        assert self.pipeline.local_rows == self.layout.local_rows(rank)
        dev = torch.device("cuda", ctx.device)
        self.frame = self.staging = self.row_index = self.local = self.sky = None
        self._shared_ptr = self._flags = 0
        self._seq = 0
        if self.exchange == "p2p":
            import torch.distributed as dist
            nbytes = height * width * 16
            if rank == 0:
                self._shared_ptr, handle = ctx.shared_frame_create(nbytes)
                h = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
            else:
                h = torch.empty(64, dtype=torch.uint8, device=dev)
            dist.broadcast(h, 0)                             # set-up only: the handle has to reach the other processes once
            if rank != 0:
                self._shared_ptr = ctx.shared_frame_open(bytes(h.cpu().tolist()))
            self._flags = ctx.shared_frame_flags(self._shared_ptr, nbytes)
            self.pipeline.bind_frame(self._shared_ptr)
            if rank == 0:
                self.frame = torch.as_tensor(_DevicePointer(self._shared_ptr, (height, width, 4)), device=dev)
        else:
            self.local = torch.zeros((self.layout.max_local_rows, width, 4), dtype=torch.float32, device=dev)
            self.pipeline.bind_output(self.local.data_ptr())
            if rank == 0 and world > 1:
                self.frame = torch.zeros((height, width, 4), dtype=torch.float32, device=dev)
                self.staging = torch.empty((world,) + tuple(self.local.shape), dtype=torch.float32, device=dev)
                if not self.layout.uniform:
                    self.row_index = [torch.as_tensor(self.layout.rows_of(r), device=dev) for r in range(world)]
        if sky_format is not None and rank == 0:
            t = self.frame_tensor()
            self.sky = SkyPipeline(ctx, None, sky_format, frame=(t.data_ptr(), width, height))

    def render_local(self, camera, black_hole, details, stream=None, events=None):
        """Enqueues this rank's share on `stream`: the replicated coarse levels, then its bands of the last level.
        `events` = (start, end): two events recorded around the ray passes proper (after the wait for rank 0)."""
        self._seq += 1
        if self.exchange == "p2p" and self.rank != 0 and self._seq > 1:
            # this rank's kernel writes into rank 0's frame: not before rank 0 has consumed the previous one
            self.ctx.stream_wait(self._flags + 4 * CONSUMED_SLOT, 1, self._seq - 1, WAIT_TIMEOUT_MS, stream)
        if events is not None:
            events[0].record(stream)
        for rp in self.levels:
            rp.pass_(camera, black_hole, details, stream)
        if events is not None:
            events[1].record(stream)
        if self.exchange == "p2p" and self.rank != 0:
            self.ctx.stream_signal(self._flags + 4 * self.rank, self._seq, stream)

    def gather(self, stream=None):
        """After this (stream-ordered on `stream`) rank 0's frame holds every rank's rows of the frame just rendered."""
        if self.exchange == "p2p":
            if self.rank == 0:
                self.ctx.stream_wait(self._flags + 4, self.world - 1, self._seq, WAIT_TIMEOUT_MS, stream)
        elif self.world > 1:
            import torch
            ts = _torch_stream(stream, self.ctx.device)
            with torch.cuda.stream(ts):                       # the collective runs on the stream the kernels were enqueued on
                gather_bands(self.local, self.layout, self.rank, self.frame, self.staging, self.row_index)

    def resolve_sky(self, stream=None):
        """Rank 0, pyramid mode: the sky resolve over the assembled frame (sky.wgsl, after the gather)."""
        if self.sky is not None:
            self.sky.pass_(stream)

    def consumed(self, stream=None):
        """Call (on every rank) after rank 0 has finished reading the frame (stream-ordered) and before the next render: in
        p2p mode the other ranks write into rank 0's memory, so they must not start the next frame while it is being read."""
        if self.exchange == "p2p" and self.rank == 0:
            self.ctx.stream_signal(self._flags + 4 * CONSUMED_SLOT, self._seq, stream)

    def close(self):
        if self.sky is not None:
            self.sky.close()
            self.sky = None
        self.pipeline.bind_frame(None)
        if self._shared_ptr:
            import torch
            import torch.distributed as dist
            torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier()                               # nobody unmaps / frees while a peer may still touch the frame
            self.frame = None
            self.ctx.shared_frame_release(self._shared_ptr, owner=self.rank == 0)
            self._shared_ptr = 0
        for rp in reversed(self.levels):
            rp.close()
        self.levels = []

    def render(self, camera, black_hole, details, stream=None):
        self.render_local(camera, black_hole, details, stream)
        self.gather(stream)
        self.resolve_sky(stream)

    def frame_tensor(self):
        """Rank 0: the assembled (H, W, 4) RGBA32F frame on the device."""
        return self.local[: self.height] if self.world == 1 else self.frame


def _torch_stream(stream, device: int):
    import torch
    if stream is None:
        return torch.cuda.current_stream(device)
    if hasattr(stream, "cuda_stream"):
        return stream
    return torch.cuda.ExternalStream(int(stream), device=device)


def _create_host_frame(HostFrame, ctx, name: str, nbytes: int):
    """(name used, error or None, HostFrame or None): the POSIX shared-memory object `name`, else a file of that name under the
    temp directory (bh_host_frame_create maps either)."""
    import os
    import tempfile
    errors = []
    for candidate in (name, os.path.join(tempfile.gettempdir(), name.lstrip("/"))):
        try:
            return candidate, None, HostFrame(ctx, candidate, nbytes, create=True)
        except Exception as ex:              # noqa: BLE001 - the message travels to the other ranks
            errors.append(f"{candidate}: {ex}")
    return None, "; ".join(errors), None


class HostTiledFrame:
    """End-to-end frame on `world` ranks with HOST buffers: one page-locked frame in POSIX shared memory that every rank
    maps (bh_host_frame); each rank's ray kernel stores its cyclic bands straight into it over its own PCIe link — no
    NVLink hop, no D2H copy, all links busy at once.  Ordering is host-side flags in the same segment."""

    def __init__(self, ctx, width: int, height: int, rank: int, world: int, band_rows: int = 8, name: "str | None" = None):
        import os
        from .pipelines import HostFrame, RayPipeline
        self.ctx, self.rank, self.world, self.width, self.height = ctx, rank, world, width, height
        self.band_rows = band_rows if world > 1 else height
        self.pipeline = RayPipeline(ctx, width, height)
        if world > 1:
            self.pipeline.set_tiling(self.band_rows, rank, world)
        self.name = name or f"/bhframe_{os.environ.get('MASTER_PORT', '0')}_{width}x{height}"
        nbytes = width * height * 16
        if world > 1:
            import torch.distributed as dist
            # Rank 0 creates the segment and tells the others its name — or why it could not — so that every rank attaches or
            # raises together (a rank that raised alone would leave the others waiting in a collective).  A /dev/shm too small
            # for the frame (container default: 64 MB) falls back to a file under the temp directory.
            verdict = [None, None]
            if rank == 0:
                verdict = list(_create_host_frame(HostFrame, ctx, self.name, nbytes))
                self.host = verdict.pop()
            dist.broadcast_object_list(verdict, src=0)
            chosen, error = verdict[0], verdict[1]
            attach_error = None
            if error is None and rank != 0:
                try:
                    self.host = HostFrame(ctx, chosen, nbytes, create=False)
                except Exception as ex:      # noqa: BLE001 - reported to every rank below
                    attach_error = f"rank {rank}: {ex}"
            errors = [None] * world
            dist.all_gather_object(errors, error or attach_error)
            errors = [e for e in errors if e]
            if errors:
                if getattr(self, "host", None) is not None:
                    self.host.close()
                self.pipeline.close()
                raise RuntimeError("HostTiledFrame: " + "; ".join(errors))
            self.name = chosen
        else:
            chosen, error, self.host = _create_host_frame(HostFrame, ctx, self.name, nbytes)
            if error is not None:
                self.pipeline.close()
                raise RuntimeError("HostTiledFrame: " + error)
            self.name = chosen
        self._seq = 0

    def enqueue(self, camera, black_hole, details, stream=None):
        """First half of render(): wait until rank 0 has let go of this buffer's previous frame, then enqueue the pass.  Returns
        without waiting for the GPU, so the caller can enqueue the next frame on ANOTHER HostTiledFrame (its own context,
        stream and host buffer) before finish()ing this one: two frames in flight, the upload of one under the pass of the other."""
        self._seq += 1
        if self.rank != 0 and self._seq > 1:
            self.host.wait(CONSUMED_SLOT, 1, self._seq - 1, WAIT_TIMEOUT_MS)        # rank 0 still reads the previous frame
        self.pipeline.pass_to_host_frame(camera, black_hole, details, self.host.ptr, stream)

    def finish(self):
        """Second half of render(): wait for the own kernel, publish; rank 0 returns once the whole frame is in host memory."""
        self.pipeline.sync()
        self.host.signal(self.rank, self._seq)
        if self.rank == 0 and self.world > 1:
            self.host.wait(1, self.world - 1, self._seq, WAIT_TIMEOUT_MS)

    def render(self, camera, black_hole, details, stream=None):
        """Every rank: enqueue, wait for the own kernel, publish.  Rank 0 returns once the whole frame is in host memory."""
        self.enqueue(camera, black_hole, details, stream)
        self.finish()

    def consumed(self):
        if self.rank == 0:
            self.host.signal(CONSUMED_SLOT, self._seq)

    def frame_array(self) -> np.ndarray:
        return self.host.array((self.height, self.width, 4))

    def close(self):
        self.pipeline.close()
        if self.world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.barrier()
        self.host.close()
