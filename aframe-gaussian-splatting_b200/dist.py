"""Multi-GPU frame sharding: one process per GPU, torch.distributed for the plumbing.

The reference has no multi-GPU path (one worker + one GL context).  Here the FRAME is sharded, not the splat
table: rank r rasters the BIN columns bx with bx % world == r (a bin = 6x6 tiles of 16x16 pixels = 96 px by default, the
granularity splats are binned at).  Every rank keeps the full 36 B/splat table
in its own HBM (80 M splats = 2.9 GB of 180 GB) and computes the same global draw order, so every pixel is
composited on exactly one GPU in exactly the reference's order - the sharded frame is bit-identical to the
single-GPU frame.  The only exchange step is one all-gather of finished RGBA tiles per frame
(width*height*4 bytes in total, 8.3 MB at 1080p), followed by an un-tiling kernel.

`TileSharding` is pure host arithmetic (also used by the gloo CPU tests); `ShardedRenderer` drives a
SplatContext per rank.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

TILE = 16


def bin_tiles() -> int:
    """tile columns per bin column: gs_bin_size() / 16 (6 for the default 96-pixel bins)"""
    from . import _lib
    return int(_lib.load().gs_bin_size()) // TILE



@dataclass(frozen=True)
class TileSharding:
    width: int
    height: int
    world: int

    @property
    def tiles_x(self) -> int:
        return (self.width + TILE - 1) // TILE

    @property
    def tiles_y(self) -> int:
        return (self.height + TILE - 1) // TILE

    def owner(self, tx: int, ty: int) -> int:
        """rank r owns the bin columns bx = tx // bin_tiles() with bx % world == r (csrc/gs_common.cuh owned_*)"""
        return (tx // bin_tiles()) % self.world

    def owned_cols(self, rank: int) -> int:
        """owned TILE columns: 4 per owned bin column, fewer in a partial last bin column (owned_tile_cols)"""
        bt = bin_tiles()
        full, rem = divmod(self.tiles_x, bt)
        n = bt * ((full - 1 - rank) // self.world + 1 if rank < full else 0)
        if rem and full % self.world == rank:
            n += rem
        return n

    def slot(self, tx: int, ty: int, rank: int) -> int:
        """index of tile (tx, ty) inside rank's packed tile buffer (mirrors owned_slot)."""
        bt = bin_tiles()
        bx = tx // bt
        return ty * self.owned_cols(rank) + (bx - rank) // self.world * bt + tx % bt

    def owned_tiles(self, rank: int) -> int:
        return self.tiles_y * self.owned_cols(rank)

    @property
    def tiles_per_rank(self) -> int:
        """every rank pads its buffer to the largest share so the all-gather is uniform"""
        return max(self.owned_tiles(r) for r in range(self.world))

    def pack_owned(self, frame: np.ndarray, rank: int) -> np.ndarray:
        """full (H, W, C) frame -> (tiles_per_rank, 256, C) buffer holding this rank's tiles (host reference)."""
        c = frame.shape[2]
        out = np.zeros((self.tiles_per_rank, TILE * TILE, c), frame.dtype)
        for ty in range(self.tiles_y):
            for tx in range(self.tiles_x):
                if self.owner(tx, ty) != rank:
                    continue
                blk = np.zeros((TILE, TILE, c), frame.dtype)
                src = frame[ty * TILE:(ty + 1) * TILE, tx * TILE:(tx + 1) * TILE]
                blk[: src.shape[0], : src.shape[1]] = src
                out[self.slot(tx, ty, rank)] = blk.reshape(TILE * TILE, c)
        return out

    def assemble(self, gathered: np.ndarray) -> np.ndarray:
        """(world, tiles_per_rank, 256, C) gathered buffers -> (H, W, C) frame (host reference of k_assemble)."""
        c = gathered.shape[-1]
        frame = np.zeros((self.height, self.width, c), gathered.dtype)
        for ty in range(self.tiles_y):
            for tx in range(self.tiles_x):
                r = self.owner(tx, ty)
                blk = gathered[r, self.slot(tx, ty, r)].reshape(TILE, TILE, c)
                dst = frame[ty * TILE:(ty + 1) * TILE, tx * TILE:(tx + 1) * TILE]
                dst[...] = blk[: dst.shape[0], : dst.shape[1]]
        return frame


def rank_frame(step: int, rank: int, world: int) -> int:
    """Frame-parallel mode: index in the frame stream of the frame `rank` renders at its local step `step`
    (rank r takes frames r, r + world, r + 2*world, ...: whole frames are the independent units, no exchange)."""
    return step * world + rank


class ShardedRenderer:
    """One rank of a frame-sharded render.

    render_tiles(frame_inputs) must return this rank's packed tiles as a torch tensor of shape
    (tiles_per_rank * 256 * C,) on the rank's device; `assemble(gathered, out)` un-tiles.  On GPUs both are the
    C ABI (gs_render with GS_RENDER_OUT_TILED, gs_assemble_tiles); the gloo tests inject host versions.
    """

    def __init__(self, sharding: TileSharding, rank: int, render_tiles: Callable, assemble: Callable, process_group=None):
        self.sharding, self.rank = sharding, rank
        self._render_tiles, self._assemble, self.pg = render_tiles, assemble, process_group

    def render(self, frame_inputs, out=None):
        import torch.distributed as dist
        import torch
        mine = self._render_tiles(frame_inputs)
        gathered = torch.empty((self.sharding.world * mine.numel(),), dtype=mine.dtype, device=mine.device)
        if self.sharding.world > 1:
            dist.all_gather_into_tensor(gathered, mine, group=self.pg)
        else:
            gathered.copy_(mine)
        return self._assemble(gathered, out)


def make_gpu_sharded_renderer(ctx, frame_like, rank: int, world: int, fmt: int = 0, process_group=None):
    """Wire a SplatContext into a ShardedRenderer (RGBA8 by default).  Returns (renderer, frame tensor)."""
    import torch
    from ._lib import GS_RENDER_OUT_DEVICE, GS_RENDER_OUT_TILED
    w, h = frame_like.width, frame_like.height
    sh = TileSharding(w, h, world)
    px = 4 if fmt == 0 else 16
    dev = torch.device("cuda", ctx.device)
    ctx.set_shard(rank, world)
    stream = torch.cuda.ExternalStream(ctx._lib.gs_stream(ctx._h), device=dev)
    with torch.cuda.stream(stream):
        tiles = torch.zeros(sh.tiles_per_rank * 256 * px, dtype=torch.uint8, device=dev)
        frame = torch.zeros(h * w * px, dtype=torch.uint8, device=dev)
    stream.synchronize()

    def render_tiles(fi):
        p = ctx.make_params(fi, fmt=fmt, flags=GS_RENDER_OUT_DEVICE | GS_RENDER_OUT_TILED)
        ctx.render_raw(p, tiles.data_ptr())
        return tiles

    def assemble(gathered, out):
        torch.cuda.current_stream(dev).synchronize()
        dst = frame if out is None else out
        ctx.assemble_tiles(gathered.data_ptr(), sh.tiles_per_rank, world, w, h, fmt, dst.data_ptr())
        ctx.synchronize()  # gs_assemble_tiles is stream-ordered
        return dst

    return ShardedRenderer(sh, rank, render_tiles, assemble, process_group), frame
