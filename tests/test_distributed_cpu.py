"""Host-side multi-GPU logic on CPU: two gloo ranks shard a query batch contiguously, each "answers" its
shard, and the shards are gathered to the host of rank 0.  No CUDA involved: the traversal itself is covered by the
gpu-marked tests, the arena broadcast by tests/test_gpu_replication.py."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    import snch_lbvh_b200  # noqa: F401
    from snch_lbvh_b200.distributed import shard_range
    for n in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    sys.path.insert(0, {root!r})
    import torch.distributed as dist
    import snch_lbvh_b200
    from snch_lbvh_b200.distributed import shard_range, sharded_query, gather_to_host
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 10007  # ragged: not divisible by the world size
    q = np.random.default_rng(5).random((n, 3)).astype(np.float32)  # every rank holds the same batch description
    def answer(pts):  # stands in for one rank's traversal of its shard
        return np.stack([pts.sum(axis=1), pts[:, 0] * 2.0], axis=1).astype(np.float32)
    lo, hi, local = sharded_query(answer, [q], n, rank, world)
    assert (lo, hi) == shard_range(n, rank, world) and len(local) == hi - lo
    full = gather_to_host(local, n, rank, world, dist)
    idx = gather_to_host(np.arange(lo, hi, dtype=np.uint32), n, rank, world, dist)
    if rank == 0:
        assert np.array_equal(full, answer(q)), "gathered result differs from the single-rank answer"
        assert np.array_equal(idx, np.arange(n, dtype=np.uint32))
        print("GATHER_OK")
    else:
        assert full is None and idx is None
    dist.barrier()
    dist.destroy_process_group()
""")


def test_two_rank_gloo_shard_and_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "GATHER_OK" in out.stdout
