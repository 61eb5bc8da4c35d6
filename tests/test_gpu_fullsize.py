"""BASELINE.json full-size configurations through size-independent properties (the oracle is sampled, not run in full)."""
import numpy as np
import pytest

from oracle import OracleScene
from parity import bits, check_closest, check_rays, check_silhouette, rel_close

pytestmark = pytest.mark.gpu

NQ = 1 << 22  # 4M queries per property run (the 16M bench batch is the same kernel; memory kept modest for the test box)


@pytest.fixture(scope="module")
def big(pkg, meshes):
    import torch
    v, f = meshes.bumpy_torus(708, 708)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    orc = OracleScene(v, f)
    lo, hi = meshes.mesh_bounds(v)
    q = meshes.points_in_box(NQ, lo, hi, 1.1, seed=2025)
    d = meshes.unit_directions(NQ, seed=77)
    return sc, orc, q, d, torch.from_numpy(q).cuda(), torch.from_numpy(d).cuda()


def test_closest_full_size(big):
    import torch
    sc, orc, q, d, qd, dd = big
    idx, dist = sc.closest_point(qd)
    torch.cuda.synchronize()
    idx, dist = idx.cpu().numpy().view(np.uint32), dist.cpu().numpy()
    assert np.all(idx < orc.n) and np.all(np.isfinite(dist))
    sel = np.random.default_rng(1).choice(NQ, 3000, replace=False)
    check_closest(q[sel], idx[sel], dist[sel], orc)
    # self-consistency on ALL queries: distance to the returned triangle equals the returned distance
    d_at = orc.point_triangle_distance(q, idx)
    assert rel_close(d_at, dist).all()
    # permutation invariance (results do not depend on which thread/warp a query lands in)
    perm = torch.randperm(NQ, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    _, dist_p = sc.closest_point(qd[perm].contiguous())
    assert torch.equal(dist_p, torch.from_numpy(dist).cuda()[perm])


def test_silhouette_full_size(big, meshes):
    import torch
    sc, orc, q, d, qd, dd = big
    _, dcp = sc.closest_point(qd)
    rmax = dcp * torch.from_numpy(meshes.star_radius_scale(NQ)).cuda()
    bnd = sc.closest_silhouette(qd, r_max=rmax)
    unb = sc.closest_silhouette(qd)
    torch.cuda.synchronize()
    inf = torch.full_like(unb, float("inf"))
    assert torch.equal(bnd, torch.where(unb <= rmax, unb, inf)), "bounded search != filtered unbounded search"
    assert torch.all((bnd <= rmax) | torch.isinf(bnd))
    assert torch.all(unb >= dcp * (1 - 1e-5) - 1e-6), "a silhouette point cannot be closer than the closest point"
    sel = np.random.default_rng(2).choice(NQ, 3000, replace=False)
    check_silhouette(unb.cpu().numpy()[sel], orc.silhouette(q[sel], nthreads=8))
    check_silhouette(bnd.cpu().numpy()[sel], orc.silhouette(q[sel], r_max=rmax.cpu().numpy()[sel], nthreads=8))


def test_silhouette_compact_records_full_size(big, meshes):
    """"query.sil_nodes": walking the 64 B compact records (48-bit cone codes + exact fallback) gives bit-identical distances
    to walking the 96 B records, bounded and unbounded, on the C3 mesh.  (Opt-in: measured slower)"""
    import torch
    sc, orc, q, d, qd, dd = big
    _, dcp = sc.closest_point(qd)
    rmax = dcp * torch.from_numpy(meshes.star_radius_scale(NQ)).cuda()
    try:
        sc.set_option("build.compact_nodes", 1).build_bvh()  # re-lays the arena with the optional CNode array
        sc.set_option("query.sil_nodes", 0)
        full_b, full_u = sc.closest_silhouette(qd, r_max=rmax), sc.closest_silhouette(qd)
        sc.set_option("query.sil_nodes", 1)
        comp_b, comp_u = sc.closest_silhouette(qd, r_max=rmax), sc.closest_silhouette(qd)
        flip = (torch.arange(NQ, device="cuda") % 2).to(torch.uint8)
        comp_f = sc.closest_silhouette(qd, flip=flip)
        sc.set_option("query.sil_nodes", 0)
        full_f = sc.closest_silhouette(qd, flip=flip)
    finally:
        sc.set_option("query.sil_nodes", 0).set_option("build.compact_nodes", 0).build_bvh()
    torch.cuda.synchronize()
    for a, b in ((full_b, comp_b), (full_u, comp_u), (full_f, comp_f)):
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), f"{int((a.view(torch.int32) != b.view(torch.int32)).sum())} distances differ"


def test_rays_full_size(big):
    import torch
    sc, orc, q, d, qd, dd = big
    found, hits = sc.intersect(qd, dd)
    any_found, _ = sc.intersect(qd, dd, any_hit=True)
    torch.cuda.synchronize()
    assert (found != any_found).float().mean().item() < 2e-4
    found, hits = found.cpu().numpy(), hits.cpu().numpy()
    t, prim = hits[:, 0].copy(), hits[:, 3].copy().view(np.uint32)
    sel = np.random.default_rng(3).choice(NQ, 3000, replace=False)
    check_rays(found[sel], t[sel], prim[sel], q[sel], d[sel], None, orc)
    # shrinking t_max to just below the hit removes it; just above keeps it (monotonicity in max_dist)
    hit = found.astype(bool)
    tm_lo = np.where(hit, t * 0.999, 1.0).astype(np.float32)
    f2, h2 = sc.intersect(qd, dd, t_max=torch.from_numpy(tm_lo).cuda())
    h2 = h2.cpu().numpy()
    assert np.all(h2[:, 0][f2.cpu().numpy().astype(bool)] < tm_lo[f2.cpu().numpy().astype(bool)])
    tm_hi = np.where(hit, t * 1.001 + 1e-6, 1.0).astype(np.float32)
    f3, h3 = sc.intersect(qd, dd, t_max=torch.from_numpy(tm_hi).cuda())
    assert np.array_equal(bits(h3.cpu().numpy()[:, 0][hit]), bits(t[hit]))
