"""Multi-GPU plumbing: one process per GPU, the tree is built once and replicated, query batches are
sharded contiguously by rank, results are gathered to the host.  The exchange itself is the library's (C-ABI
snch_comm_* / snch_scene_broadcast on libnccl.so.2); torch.distributed only bootstraps the communicator's id and gathers
host-side results (gloo in the CPU tests).

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


def make_comm(rank: int, world: int, device: int, dist=None, src: int = 0):
    """The library's own NCCL communicator (binding.Comm -> snch_comm_create: libnccl.so.2, no torch on the data path).
    torch.distributed — any backend — is only the bootstrap that carries rank `src`'s 128-byte unique id to the others."""
    import torch
    from .binding import Comm

    if dist is None:
        import torch.distributed as dist  # noqa: PLW0642
    dev = "cpu" if dist.get_backend() == "gloo" else f"cuda:{device}"
    uid = torch.zeros(Comm.ID_BYTES, dtype=torch.uint8)
    if rank == src:
        uid = torch.frombuffer(bytearray(Comm.unique_id()), dtype=torch.uint8).clone()
    uid = uid.to(dev)
    dist.broadcast(uid, src=src)
    return Comm(bytes(uid.cpu().numpy().tobytes()), rank, world, device)


def replicate_scene(scene, rank: int, world: int, device: int, dist=None, src: int = 0, comm=None):
    """Rank `src` passes its built Scene3; every rank returns a Scene3 living on its own GPU.

    The arena is broadcast by the library itself (snch_scene_broadcast: ncclBroadcast over NVLink / NVSwitch, received
    directly into the replica's arena) and its embedded pointers are re-patched on arrival.  `comm`: a binding.Comm to
    reuse (one is made — and kept on the returned scene as `.comm` — otherwise)."""
    if world == 1:
        return scene
    if comm is None:
        comm = make_comm(rank, world, device, dist, src)
    out = comm.broadcast(scene if rank == src else None, root=src)
    out.comm = comm
    return out


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
