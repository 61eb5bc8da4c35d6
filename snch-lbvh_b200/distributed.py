"""Multi-GPU plumbing: one process per GPU, the tree is built once and replicated, query batches are
sharded contiguously by rank, results are gathered to the host.  torch.distributed is used only as plumbing
(NCCL broadcast of the arena over NVLink / NVSwitch; gloo in the CPU tests).

The path has exactly ONE exchange step — the broadcast of the pointer-free scene arena.  Queries never communicate.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of a batch of n queries owned by `rank` (the reference has no multi-GPU path; this is
    the partitioning SURVEY 8(e) specifies: [r*Q/G, (r+1)*Q/G))."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (n * rank) // world, (n * (rank + 1)) // world


def replicate_scene(scene, rank: int, world: int, device: int, dist=None, src: int = 0):
    """Rank `src` passes its built Scene3; every rank returns a Scene3 living on its own GPU.

    The arena is broadcast as one uint8 tensor (ncclBroadcast under torch.distributed) and adopted with
    snch_scene_adopt_arena, which re-patches the raw pointers embedded in the reference-layout structs."""
    import torch
    from .binding import Scene3

    if world == 1:
        return scene
    if dist is None:
        import torch.distributed as dist  # noqa: PLW0642
    size = torch.zeros(1, dtype=torch.int64, device=f"cuda:{device}")
    if rank == src:
        view = scene.arena_tensor()
        size[0] = view.numel()
    dist.broadcast(size, src=src)
    nbytes = int(size.item())
    if rank == src:
        dist.broadcast(view, src=src)
        return scene
    buf = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{device}")
    dist.broadcast(buf, src=src)
    torch.cuda.current_stream().synchronize()
    return Scene3.adopt_arena(buf, device=device)


def sharded_query(query_fn, inputs: list, n: int, rank: int, world: int):
    """Run `query_fn(*shard_of_inputs)` on this rank's contiguous shard; returns (lo, hi, result)."""
    lo, hi = shard_range(n, rank, world)
    return lo, hi, query_fn(*[None if a is None else a[lo:hi] for a in inputs])


def gather_to_host(local: np.ndarray, n: int, rank: int, world: int, dist=None, dst: int = 0):
    """Assemble per-rank result shards (host arrays) into one host array of length n on rank `dst` (None elsewhere).
    Shards are the contiguous ranges of shard_range(); uses gather_object-free tensor collectives so it works with
    gloo (CPU tests) and NCCL alike."""
    import torch

    if world == 1:
        return local
    if dist is None:
        import torch.distributed as dist  # noqa: PLW0642
    backend = dist.get_backend()
    dev = "cpu" if backend == "gloo" else f"cuda:{torch.cuda.current_device()}"
    lo, hi = shard_range(n, rank, world)
    assert len(local) == hi - lo
    maxlen = max(shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world))
    flat = np.ascontiguousarray(local).reshape(len(local), -1)
    width = flat.shape[1]
    pad = np.zeros((maxlen, width), dtype=flat.dtype)
    pad[: len(flat)] = flat
    t = torch.from_numpy(pad.view(np.uint8).reshape(-1)).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    if rank != dst:
        return None
    full = np.zeros((n, width), dtype=flat.dtype)
    for r in range(world):
        a, b = shard_range(n, r, world)
        piece = outs[r].cpu().numpy().view(flat.dtype).reshape(maxlen, width)
        full[a:b] = piece[: b - a]
    return full.reshape((n,) + local.shape[1:])
