"""BASELINE.json full-size configurations through size-independent properties (the oracle is sampled, not run in full)."""
import numpy as np
import pytest

from oracle import OracleScene
from parity import bits, check_closest, check_rays, check_silhouette, check_silhouette_edges, rel_close

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
    assert np.array_equal(bits(d_at), bits(dist))
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
    assert torch.equal(bnd, torch.where((unb <= rmax) & (rmax * rmax > 0), unb, inf)), "bounded search != filtered unbounded search"
    assert torch.all((bnd <= rmax) | torch.isinf(bnd))
    assert torch.all(unb >= dcp * (1 - 1e-5) - 1e-6), "a silhouette point cannot be closer than the closest point"
    sel = np.random.default_rng(2).choice(NQ, 3000, replace=False)
    check_silhouette(unb.cpu().numpy()[sel], orc.silhouette(q[sel], nthreads=8))
    check_silhouette(bnd.cpu().numpy()[sel], orc.silhouette(q[sel], r_max=rmax.cpu().numpy()[sel], nthreads=8))


def test_silhouette_edges_full_size(big, meshes):
    """out_edge / out_point on the C3 batch shape: distances identical to the plain call, every returned edge attains its
    distance, sampled queries against the oracle's silhouette_ex."""
    import torch
    sc, orc, q, d, qd, dd = big
    _, dcp = sc.closest_point(qd)
    rmax = dcp * torch.from_numpy(meshes.star_radius_scale(NQ)).cuda()
    plain = sc.closest_silhouette(qd, r_max=rmax)
    dist, edge, pt = sc.closest_silhouette(qd, r_max=rmax, with_edge=True)
    torch.cuda.synchronize()
    assert torch.equal(plain.view(torch.int32), dist.view(torch.int32))
    dist, edge, pt = dist.cpu().numpy(), edge.cpu().numpy().view(np.uint32), pt.cpu().numpy()
    fin = np.isfinite(dist)
    assert np.all(edge[~fin] == 0xFFFFFFFF) and np.all(edge[fin] < orc.num_edges)
    d_at, p_at = orc.point_edge_distance(q[fin], edge[fin].astype(np.int32))
    assert np.array_equal(bits(d_at), bits(dist[fin])) and np.array_equal(bits(p_at), bits(pt[fin]))
    sel = np.random.default_rng(5).choice(NQ, 3000, replace=False)
    check_silhouette_edges(q[sel], dist[sel], edge[sel], pt[sel], orc, False, r_max=rmax.cpu().numpy()[sel])


def test_rays_full_size(big):
    import torch
    sc, orc, q, d, qd, dd = big
    found, hits = sc.intersect(qd, dd)
    any_found, _ = sc.intersect(qd, dd, any_hit=True)
    torch.cuda.synchronize()
    assert (found != any_found).float().mean().item() < 2e-4  # any-hit takes the first hit in walk order (Q4 ties aside, the same flag)
    found, hits = found.cpu().numpy(), hits.cpu().numpy()
    t, prim = hits[:, 0].copy(), hits[:, 3].copy().view(np.uint32)
    sel = np.random.default_rng(3).choice(NQ, 3000, replace=False)
    check_rays(found[sel], {"t": t[sel], "u": hits[sel, 1].copy(), "v": hits[sel, 2].copy(), "prim": prim[sel]}, q[sel], d[sel], None, orc)
    # shrinking t_max to just below the hit removes it; just above keeps it (monotonicity in max_dist)
    hit = found.astype(bool)
    tm_lo = np.where(hit, t * 0.999, 1.0).astype(np.float32)
    f2, h2 = sc.intersect(qd, dd, t_max=torch.from_numpy(tm_lo).cuda())
    h2 = h2.cpu().numpy()
    assert np.all(h2[:, 0][f2.cpu().numpy().astype(bool)] < tm_lo[f2.cpu().numpy().astype(bool)])
    tm_hi = np.where(hit, t * 1.001 + 1e-6, 1.0).astype(np.float32)
    f3, h3 = sc.intersect(qd, dd, t_max=torch.from_numpy(tm_hi).cuda())
    assert np.array_equal(bits(h3.cpu().numpy()[:, 0][hit]), bits(t[hit]))


@pytest.mark.parametrize("name,nu", [("c4", 1416), ("c5", 2240)])
def test_c4_c5_mesh_sizes(pkg, meshes, name, nu):
    """BASELINE.json configs C4 (4 010 112 triangles) and C5 (10 035 200 triangles, 15 052 800 silhouette edges): every build
    product against the oracle's (bit-exact integer pipeline and boxes, cones within parity.check_cones), then 4 096 sampled
    closest-point / silhouette (unbounded, star radius, with edge + point) / ray / sphere-sampling queries and one wavefront
    step, all bit-identical to the oracle's answers."""
    import torch
    from parity import check_build_vs_oracle
    v, f = meshes.bumpy_torus(nu, nu)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    orc = OracleScene(v, f)
    check_build_vs_oracle(sc, orc, pkg)
    lo, hi = meshes.mesh_bounds(v)
    n = 4096
    q = meshes.points_in_box(n, lo, hi, 1.1, seed=91)
    d = meshes.unit_directions(n, seed=92)
    idx, dist = sc.closest_point(q)
    check_closest(q, idx, dist, orc)
    rmax = (dist * meshes.star_radius_scale(n, seed=93)).astype(np.float32)
    check_silhouette(sc.closest_silhouette(q), orc.silhouette(q, nthreads=8))
    check_silhouette(sc.closest_silhouette(q, flip=True, r_max=rmax), orc.silhouette(q, True, r_max=rmax, nthreads=8))
    check_silhouette_edges(q, *sc.closest_silhouette(q, r_max=rmax, with_edge=True), orc, False, r_max=rmax)
    found, hits = sc.intersect(q, d)
    check_rays(found, hits, q, d, None, orc)
    sph = np.concatenate([q, rmax[:, None]], axis=1).astype(np.float32)
    rnd = meshes.uniforms(n, 3, seed=94)
    sidx, pdf, _ = sc.sample_in_sphere(sph, rnd)
    sidx_o, pdf_o = orc.sample(sph, rnd[:, 0].copy())
    assert np.array_equal(sidx, sidx_o) and np.array_equal(bits(pdf), bits(pdf_o))
    # a device-resident batch large enough for the packet / per-lane kernels (the 4 096 above take the one-query-per-warp
    # kernels): sampled against the oracle, and the wavefront step equals the four calls
    nb = 1 << 21
    qb = meshes.points_in_box(nb, lo, hi, 1.1, seed=95)
    db = meshes.unit_directions(nb, seed=96)
    qd, dd = torch.from_numpy(qb).cuda(), torch.from_numpy(db).cuda()
    sc.set_option("query.wide_max_n", 0)  # packets even though the batch is sparse relative to the mesh
    _, dist_p = sc.closest_point(qd)
    sc.set_option("query.wide_max_n", 2097152)
    ib, dist_b = sc.closest_point(qd)
    assert torch.equal(dist_p.view(torch.int32), dist_b.view(torch.int32))
    sil_b, edge_b, pt_b = sc.closest_silhouette(qd, r_max=dist_b, with_edge=True)
    fb, hb = sc.intersect(qd, dd)
    torch.cuda.synchronize()
    sel = np.random.default_rng(6).choice(nb, 3000, replace=False)
    dist_h = dist_b.cpu().numpy()
    check_closest(qb[sel], ib.cpu().numpy().view(np.uint32)[sel], dist_h[sel], orc)
    check_silhouette_edges(qb[sel], sil_b.cpu().numpy()[sel], edge_b.cpu().numpy().view(np.uint32)[sel], pt_b.cpu().numpy()[sel], orc, False,
                           r_max=dist_h[sel])
    hb = hb.cpu().numpy()
    check_rays(fb.cpu().numpy()[sel], {"t": hb[sel, 0].copy(), "u": hb[sel, 1].copy(), "v": hb[sel, 2].copy(), "prim": hb[sel, 3].copy().view(np.uint32)},
               qb[sel], db[sel], None, orc)
    w = sc.wost_step(qd, dd, torch.from_numpy(meshes.uniforms(nb, 3, seed=97)).cuda(), with_edge=True)
    torch.cuda.synchronize()
    assert torch.equal(w["closest_distance"].view(torch.int32), dist_b.view(torch.int32))
    assert torch.equal(w["silhouette_distance"].view(torch.int32), sil_b.view(torch.int32))
    assert torch.equal(w["silhouette_point"].view(torch.int32), pt_b.view(torch.int32))
