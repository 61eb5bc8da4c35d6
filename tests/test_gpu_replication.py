"""Multi-GPU plumbing on one GPU: the arena is pointer-free, so a byte copy adopted as a second scene answers
queries identically (this is what every non-zero rank does after the NCCL broadcast)."""
import numpy as np
import pytest

from parity import bits

pytestmark = pytest.mark.gpu


def test_arena_roundtrip(pkg, meshes):
    import torch
    v, f = meshes.bumpy_torus(80, 60)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    view = sc.arena_tensor()
    assert view.numel() == sc.stats()["arena_bytes"]
    copy = view.clone()          # stands in for the broadcast receive buffer
    rep = pkg.Scene3.adopt_arena(copy, device=0)
    del copy                     # the replica owns its own copy
    torch.cuda.empty_cache()
    K = pkg.ExportKind
    for k in (K.NODES, K.AABBS, K.CONES, K.MORTON_SORTED, K.SORTED_INDEX, K.RANGES, K.EDGES, K.TRI_OWNED, K.TRI_EDGES):
        a, b = sc.export(k), rep.export(k)
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    lo, hi = meshes.mesh_bounds(v)
    q = meshes.points_in_box(8000, lo, hi, 1.3, seed=51)
    d = meshes.unit_directions(8000, seed=52)
    assert np.array_equal(bits(sc.closest_point(q)[1]), bits(rep.closest_point(q)[1]))
    assert np.array_equal(bits(sc.closest_silhouette(q)), bits(rep.closest_silhouette(q)))
    assert np.array_equal(bits(sc.intersect(q, d)[1]["t"]), bits(rep.intersect(q, d)[1]["t"]))
    # the reference-layout view of the replica points into ITS arena (embedded pointers were re-patched)
    pa, pb = sc.get_bvh_device_ptr(), rep.get_bvh_device_ptr()
    assert pa.nodes != pb.nodes and pa.num_nodes == pb.num_nodes
    with pytest.raises(pkg.SnchError):
        pkg.Scene3.adopt_arena(torch.zeros(4096, dtype=torch.uint8, device="cuda"), device=0)


def _same_answers(pkg, meshes, sc, rep, v):
    K = pkg.ExportKind
    for k in (K.NODES, K.AABBS, K.CONES, K.RANGES, K.EDGES, K.TRI_OWNED):
        assert np.array_equal(sc.export(k).view(np.uint8), rep.export(k).view(np.uint8))
    lo, hi = meshes.mesh_bounds(v)
    q = meshes.points_in_box(8000, lo, hi, 1.3, seed=53)
    assert np.array_equal(bits(sc.closest_point(q)[1]), bits(rep.closest_point(q)[1]))
    a, b = sc.closest_silhouette(q, with_edge=True), rep.closest_silhouette(q, with_edge=True)
    assert all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(a, b))


def test_replicate_local_peer_copy(pkg, meshes):
    """snch_scene_replicate_local: replicas made by the library's own peer-copy fan-out — onto the scene's own device, and onto a
    second GPU when the box has one."""
    import torch
    v, f = meshes.bumpy_torus(80, 60)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    devices = [0] + ([1] if torch.cuda.device_count() > 1 else [])
    reps = sc.replicate_local(devices)
    assert [r.device for r in reps] == devices
    for r in reps:
        _same_answers(pkg, meshes, sc, r, v)
        assert r.get_bvh_device_ptr().nodes != sc.get_bvh_device_ptr().nodes
        with pytest.raises(pkg.SnchError):
            r.build_bvh()  # a replica cannot be rebuilt
    with pytest.raises(pkg.SnchError):
        sc.replicate_local([99])


def test_comm_single_rank_broadcast(pkg, meshes):
    """The library's own NCCL path (libnccl.so.2 through snch_comm_*) with a world of one: id, init, broadcast, rebroadcast."""
    v, f = meshes.icosphere(3)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    comm = pkg.Comm(pkg.Comm.unique_id(), 0, 1, 0)
    assert comm.broadcast(sc, root=0) is sc
    comm.rebroadcast(sc, root=0)
    _same_answers(pkg, meshes, sc, sc, v)
    comm.close()


_BCAST_WORKER = """
import os, sys, time
import numpy as np
sys.path.insert(0, {root!r})
import snch_lbvh_b200 as pkg
rank, world, idfile = int(sys.argv[1]), 2, sys.argv[2]
m = pkg.meshes
v, f = m.bumpy_torus(80, 60)
if rank == 0:
    open(idfile + ".tmp", "wb").write(pkg.Comm.unique_id())
    os.rename(idfile + ".tmp", idfile)
while not os.path.exists(idfile):
    time.sleep(0.05)
comm = pkg.Comm(open(idfile, "rb").read(), rank, world, rank)
sc = pkg.Scene3(v, f, device=0).compute_silhouettes().build_bvh() if rank == 0 else None
rep = comm.broadcast(sc, root=0)
lo, hi = m.mesh_bounds(v)
q = m.points_in_box(4000, lo, hi, 1.3, seed=54)
d1, e1, p1 = rep.closest_silhouette(q, with_edge=True)
np.save(idfile + f".rank{{rank}}.npy", np.concatenate([d1, rep.closest_point(q)[1], p1.ravel()]))
if rank == 0:  # move the mesh, rebuild on the root, refresh the replica in place
    sc.update_vertices((v * 1.25).astype(np.float32)).build_bvh()
comm.rebroadcast(rep, root=0)
np.save(idfile + f".rank{{rank}}.b.npy", rep.closest_point((q * 1.25).astype(np.float32))[1])
comm.close()
print("BCAST_OK", rank)
"""


def test_broadcast_two_ranks(pkg, meshes, tmp_path):
    """snch_scene_broadcast / snch_scene_rebroadcast between two processes on two GPUs: the replica answers bit-identically."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (covered by bench.py --gpus 2)")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(_BCAST_WORKER.format(root=root))
    idfile = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), idfile], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs) and all("BCAST_OK" in o for o in outs), "\n".join(o[-2000:] for o in outs)
    a, b = np.load(idfile + ".rank0.npy"), np.load(idfile + ".rank1.npy")
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    a, b = np.load(idfile + ".rank0.b.npy"), np.load(idfile + ".rank1.b.npy")
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
