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
