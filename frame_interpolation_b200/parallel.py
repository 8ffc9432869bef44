"""Multi-GPU sharding of the FILM hot path (one process per GPU, torch.distributed).

The path shards into independent units:

* frame pairs  -- `interpolate_pairs`: contiguous block partition of N independent
  (x0, x1) pairs over ranks, weights replicated, ONE all-gather of the outputs so every
  rank (or a single writer) holds the full sequence.
* tiles        -- `interpolate_tiled`: the reference's `--block_height/--block_width`
  tiles (eval/interpolator.py:192-206) are independent by construction (no halo, each
  tile padded on its own), so tile t goes to rank t % world; ONE all-gather reassembles
  the stitched frame (NCCL over NVLink on GPUs; gloo on CPU for the host-logic tests).
* recursion    -- `interpolate_recursively`: eval/util.py:62-91's binary dependency tree
  scheduled level-synchronously: the 2^(k-1) calls of level k are sharded, new mid-frames
  are all-gathered so every rank holds the parents of level k+1; the output order is the
  generator's in-order traversal.

There is no model parallelism (34 M parameters) and no collective inside the network.
`engine` is any callable `(x0, x1, dt) -> mid` on numpy NHWC batches -- the real
`Interpolator` in production, a stand-in in the CPU (gloo) tests of this host logic.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

Engine = Callable[[np.ndarray, np.ndarray, np.ndarray], np.ndarray]


def block_partition(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) share of `rank`; the first n % world ranks get one extra."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def round_robin(n_items: int, world: int, rank: int) -> List[int]:
    return list(range(rank, n_items, world))


def _dist():
    import torch.distributed as dist
    return dist


def _world_rank(group=None) -> Tuple[int, int]:
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _all_gather_padded(local: np.ndarray, counts: Sequence[int], device=None, group=None) -> np.ndarray:
    """All-gathers `local` (n_local, ...) where ranks hold different n_local (`counts`).
    One collective: every rank pads to max(counts) rows, all_gather_into_tensor, then the
    padding is dropped. Returns the concatenation in rank order."""
    import torch
    dist = _dist()
    world, rank = _world_rank(group)
    if world == 1:
        return local
    item_shape = local.shape[1:]
    m = max(counts)
    buf = torch.zeros((m,) + tuple(item_shape), dtype=torch.float32, device=device or "cpu")
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(buf.device)
    out = torch.empty((world * m,) + tuple(item_shape), dtype=torch.float32, device=buf.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.cpu().numpy().reshape((world, m) + tuple(item_shape))
    return np.concatenate([out[r, : counts[r]] for r in range(world)], axis=0)


def interpolate_pairs(engine: Engine, x0: np.ndarray, x1: np.ndarray, device=None, group=None,
                      gather: bool = True) -> np.ndarray:
    """Mid-frames of N independent pairs (x0[i], x1[i]), sharded over ranks.

    x0, x1: (N, H, W, 3). Returns (N, H, W, 3) on every rank when `gather`, else only the
    local block (N_local, H, W, 3)."""
    world, rank = _world_rank(group)
    n = x0.shape[0]
    lo, hi = block_partition(n, world, rank)
    outs = [engine(x0[i:i + 1], x1[i:i + 1], np.full((1,), 0.5, np.float32)) for i in range(lo, hi)]
    local = np.concatenate(outs, axis=0) if outs else np.zeros((0,) + x0.shape[1:], np.float32)
    if not gather or world == 1:
        return local
    counts = [block_partition(n, world, r)[1] - block_partition(n, world, r)[0] for r in range(world)]
    return _all_gather_padded(local, counts, device, group)


def interpolate_tiled(engine: Engine, x0: np.ndarray, x1: np.ndarray, block_shape: Sequence[int],
                      device=None, group=None) -> np.ndarray:
    """The reference's tiled path with tiles sharded round-robin over ranks and one
    all-gather to reassemble (eval/interpolator.py:192-206 runs them sequentially)."""
    from .interpolator import image_to_patches, patches_to_image
    world, rank = _world_rank(group)
    bh, bw = int(block_shape[0]), int(block_shape[1])
    p0 = image_to_patches(x0, [bh, bw])
    p1 = image_to_patches(x1, [bh, bw])
    nt = bh * bw
    mine = round_robin(nt, world, rank)
    dt = np.full((1,), 0.5, np.float32)
    outs = [engine(p0[t][np.newaxis], p1[t][np.newaxis], dt) for t in mine]
    local = np.concatenate(outs, axis=0) if outs else np.zeros((0,) + p0.shape[1:], np.float32)
    if world > 1:
        counts = [len(round_robin(nt, world, r)) for r in range(world)]
        gathered = _all_gather_padded(local, counts, device, group)
        # rank-major order -> tile order
        order = [t for r in range(world) for t in round_robin(nt, world, r)]
        tiles = np.empty_like(gathered)
        tiles[order] = gathered
    else:
        tiles = local
    return patches_to_image(tiles, [bh, bw])


def interpolate_recursively(engine: Engine, frame0: np.ndarray, frame1: np.ndarray,
                            times_to_interpolate: int, device=None, group=None) -> List[np.ndarray]:
    """All 2^n - 1 mid-frames between two (H, W, 3) frames plus the end points, in display
    order -- the sequence `_recursive_generator` (eval/util.py:62-91) yields, followed by
    frame1 (eval/util.py:118-123). Level-synchronous over ranks: at level k the 2^(k-1)
    independent calls are block-partitioned and their results all-gathered."""
    world, rank = _world_rank(group)
    frames = [np.asarray(frame0, np.float32), np.asarray(frame1, np.float32)]
    for _ in range(times_to_interpolate):
        a = np.stack(frames[:-1])
        b = np.stack(frames[1:])
        mids = interpolate_pairs(engine, a, b, device, group, gather=True)
        nxt = []
        for i in range(len(frames) - 1):
            nxt.append(frames[i])
            nxt.append(mids[i])
        nxt.append(frames[-1])
        frames = nxt
    return frames
