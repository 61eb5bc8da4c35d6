"""Query parity on the GPU, through the C-ABI, against the CPU oracle on the same seeded inputs."""
import os

import numpy as np
import pytest

from conftest import small_cases
from oracle import OracleScene
from parity import check_rays_exact, bits, check_closest, check_rays, check_silhouette, check_silhouette_edges, rel_close

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _mesh(meshes, name):
    cases = small_cases(meshes)
    if name in cases:
        return cases[name]
    def moved(vf, offset, scale=1.0):  # a structured mesh far from the origin: every coordinate-sized quantity rounds at the offset's ulp
        return (vf[0] * np.float32(scale) + np.asarray(offset, np.float32)).astype(np.float32), vf[1]
    return {"ico5": lambda: meshes.icosphere(5), "grid40": lambda: meshes.open_grid(40),
            "torus200": lambda: meshes.bumpy_torus(200, 200),
            "torus128_far": lambda: moved(meshes.bumpy_torus(128, 128), (1000.0, -2000.0, 500.0)),
            "ico5_far_small": lambda: moved(meshes.icosphere(5), (100.0, 100.0, -100.0), 0.01)}[name]()


@pytest.fixture(scope="module", params=["tet", "ico2", "grid6", "torus24x16", "ico5", "grid40", "torus200", "torus128_far", "ico5_far_small"])
def scene(request, pkg, meshes):
    v, f = _mesh(meshes, request.param)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    orc = OracleScene(v, f)
    lo, hi = meshes.mesh_bounds(v)
    n = 20000 if len(f) > 1000 else 4000
    q = meshes.points_in_box(n, lo, hi, 1.5, seed=41)
    d = meshes.unit_directions(n, seed=42)
    return request.param, sc, orc, q, d


def test_closest_point(scene):
    _, sc, orc, q, _ = scene
    idx, dist = sc.closest_point(q)
    check_closest(q, idx, dist, orc)


@pytest.mark.parametrize("flip", [False, True])
def test_closest_silhouette_unbounded(scene, flip):
    _, sc, orc, q, _ = scene
    dist = sc.closest_silhouette(q, flip=flip)
    check_silhouette(dist, orc.silhouette(q, flip, nthreads=8))


def test_closest_silhouette_star_radius(scene, meshes):
    """config C3 semantics: r_max = s * closest-point distance; equals the filtered unbounded answer (Q5)."""
    _, sc, orc, q, _ = scene
    _, dcp = orc.closest(q, nthreads=8)
    rmax = (dcp * meshes.star_radius_scale(len(q))).astype(np.float32)
    dist = sc.closest_silhouette(q, r_max=rmax)
    check_silhouette(dist, orc.silhouette(q, False, r_max=rmax, nthreads=8))
    unb = sc.closest_silhouette(q)
    assert np.array_equal(bits(dist), bits(np.where((unb <= rmax) & (rmax * rmax > 0), unb, np.inf).astype(np.float32)))


@pytest.mark.parametrize("flip", [False, True])
def test_closest_silhouette_edge_and_point(scene, meshes, flip):
    """out_edge / out_point of snch_closest_silhouette_batch (what the reference computes at scene.cuh:796-799 and drops),
    unbounded and with star radii, against the oracle's silhouette_ex; asking for them never changes a distance."""
    _, sc, orc, q, _ = scene
    dist, edge, pt = sc.closest_silhouette(q, flip=flip, with_edge=True)
    assert np.array_equal(bits(dist), bits(sc.closest_silhouette(q, flip=flip)))
    check_silhouette_edges(q, dist, edge, pt, orc, flip)
    rmax = (orc.closest(q, nthreads=8)[1] * meshes.star_radius_scale(len(q))).astype(np.float32)
    dist, edge, pt = sc.closest_silhouette(q, flip=flip, r_max=rmax, with_edge=True)
    check_silhouette_edges(q, dist, edge, pt, orc, flip, r_max=rmax)


@pytest.mark.parametrize("ray_kernel", [0, 1, 2])  # 0: k_intersect (leaves where they are met); 1, 2: the reference-order walk (flush at 1 / 8 parked lanes)
def test_rays_closest_hit(scene, ray_kernel):
    _, sc, orc, q, d = scene
    sc.set_option("query.ray_kernel", ray_kernel)
    found, hits = sc.intersect(q, d)
    sc.set_option("query.ray_kernel", 1)
    both = check_rays(found, hits, q, d, None, orc)
    if ray_kernel:
        check_rays_exact(found, hits, q, d, None, orc)  # the triangle too, ties included
    # the reported primitive really is hit at the reported t (ties between triangles sharing an edge are legal, Q4)
    f_b, t_b, _, _ = orc.ray(q[both], d[both], brute=True)
    assert rel_close(hits["t"][both], t_b).all()
    assert np.all(hits["prim"][~found.astype(bool)] == 0xFFFFFFFF)
    u, v = hits["u"][both], hits["v"][both]
    assert np.all(u >= 0) and np.all(v >= 0) and np.all(u + v <= 1 + 1e-6)


@pytest.mark.parametrize("ray_kernel", [1, 2])
def test_rays_tmax_and_any_hit(scene, ray_kernel):
    _, sc, orc, q, d = scene
    tm = np.full(len(q), 0.7, np.float32)
    sc.set_option("query.ray_kernel", ray_kernel)
    found, hits = sc.intersect(q, d, t_max=tm)
    check_rays(found, hits, q, d, tm, orc)
    assert np.all(hits["t"][found.astype(bool)] < 0.7)
    any_found, _ = sc.intersect(q, d, t_max=tm, any_hit=True)
    sc.set_option("query.ray_kernel", 1)
    assert np.array_equal(any_found.astype(bool), orc.ray(q, d, tm, any_hit=True, nthreads=8)[0].astype(bool))
    assert np.mean(any_found.astype(bool) == found.astype(bool)) > 0.9998  # (any-hit stops at the first hit in walk order: Q4 ties aside, same flag)


def test_sample_in_sphere(scene, meshes):
    _, sc, orc, q, _ = scene
    _, dcp = orc.closest(q, nthreads=8)
    sph = np.concatenate([q, (dcp * 1.5 + 0.05)[:, None]], axis=1).astype(np.float32)
    rnd = meshes.uniforms(len(q), 3, seed=43)
    idx, pdf, pt = sc.sample_in_sphere(sph, rnd)
    idx_o, pdf_o = orc.sample(sph, rnd[:, 0].copy())
    assert np.array_equal(idx, idx_o), f"sampled primitive differs on {np.count_nonzero(idx != idx_o)} of {len(idx)}"  # one deterministic descent
    hit = idx >= 0
    assert np.array_equal(bits(pdf[hit]), bits(pdf_o[hit])), "sampling pdf differs"
    assert np.all(pdf[idx < 0] == 0)
    pt_o = orc.sample_on_object(idx, rnd[:, 1].copy(), rnd[:, 2].copy())
    assert np.array_equal(bits(pt[hit]), bits(pt_o[hit])), "sampled point differs"


@pytest.mark.parametrize("name", ["tet", "ico2", "grid6", "torus24x16"])
def test_queries_match_golden(pkg, name):
    """Against vectors produced by the reference itself (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    sc = pkg.Scene3(g["verts"], g["tris"]).compute_silhouettes().build_bvh()
    q, d = g["q"], g["d"]
    _, dist = sc.closest_point(q)
    assert np.array_equal(bits(dist), bits(g["closest_dist"]))
    check_silhouette(sc.closest_silhouette(q, flip=False), g["sil_noflip"])
    check_silhouette(sc.closest_silhouette(q, flip=True), g["sil_flip"])
    found, hits = sc.intersect(q, d)
    assert np.array_equal(found.astype(bool), g["ray_found"].astype(bool))
    both = found.astype(bool)
    assert np.array_equal(bits(hits["t"][both]), bits(g["ray_t"][both]))
    idx, pdf, _ = sc.sample_in_sphere(g["sph"], np.stack([g["u"], g["u"], g["u"]], 1))
    assert np.array_equal(idx, g["sample_idx"])


def test_edge_cases(pkg, meshes):
    # empty scene: sentinels
    sc = pkg.Scene3(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32)).compute_silhouettes().build_bvh()
    q = np.zeros((5, 3), np.float32)
    idx, dist = sc.closest_point(q)
    assert np.all(idx == 0xFFFFFFFF) and np.all(np.isinf(dist))
    assert np.all(np.isinf(sc.closest_silhouette(q)))
    sd, se, sp_ = sc.closest_silhouette(q, with_edge=True)
    assert np.all(np.isinf(sd)) and np.all(se == 0xFFFFFFFF) and np.all(sp_ == 0)
    found, hits = sc.intersect(q, np.ones((5, 3), np.float32))
    assert not found.any() and np.all(np.isinf(hits["t"]))
    sidx, pdf, _ = sc.sample_in_sphere(np.ones((5, 4), np.float32), np.zeros((5, 3), np.float32))
    assert np.all(sidx == -1) and np.all(pdf == 0)
    # zero queries
    v, f = meshes.icosphere(1)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    idx, dist = sc.closest_point(np.zeros((0, 3), np.float32))
    assert len(idx) == 0
    # single triangle (the reference reads out of bounds here, Q6; defined behaviour: test the only leaf)
    v1 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    sc = pkg.Scene3(v1, np.array([[0, 1, 2]], np.int32)).compute_silhouettes().build_bvh()
    orc = OracleScene(v1, np.array([[0, 1, 2]], np.int32))
    q = meshes.points_in_box(500, [-1, -1, -1], [2, 2, 1], 1.0, seed=44)
    idx, dist = sc.closest_point(q)
    assert np.all(idx == 0) and rel_close(dist, orc.closest(q)[1]).all()
    assert np.array_equal(bits(sc.closest_silhouette(q)), bits(orc.silhouette(q)))
    check_silhouette_edges(q, *sc.closest_silhouette(q, with_edge=True), orc)
    d = meshes.unit_directions(500, seed=45)
    found, hits = sc.intersect(q, d)
    f_o, t_o, _, _ = orc.ray(q, d)
    assert np.array_equal(found.astype(bool), f_o.astype(bool)) and rel_close(hits["t"], t_o).all()
    # query points ON the surface and far away
    v, f = meshes.icosphere(3)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    orc = OracleScene(v, f)
    q = np.concatenate([v[:200], v[:200] * 1000.0, np.zeros((1, 3), np.float32)]).astype(np.float32)
    check_closest(q, *sc.closest_point(q), orc)
    check_silhouette(sc.closest_silhouette(q), orc.silhouette(q))
    # axis-aligned rays (zero direction components -> infinite inverse directions)
    o = np.tile(np.array([[0.1, 0.2, -3.0]], np.float32), (4, 1))
    dd = np.array([[0, 0, 1], [0, 0, -1], [1, 0, 0], [0, 1, 0]], np.float32)
    found, hits = sc.intersect(o, dd)
    f_o, t_o, _, _ = orc.ray(o, dd)
    assert np.array_equal(found.astype(bool), f_o.astype(bool)) and rel_close(hits["t"], t_o).all()


def test_nan_inf_and_huge_inputs(pkg, meshes):
    """Query points, radii, directions and spheres that are NaN, infinite, huge, denormal or zero, mixed into an ordinary batch:
    every kernel terminates and answers what the reference's comparisons answer (a NaN compares false everywhere), in the
    one-query-per-warp kernels and in the large-batch ones (packets, leaf queue, parked rays), with and without batch ordering."""
    from parity import same_bits
    v, f = meshes.icosphere(3)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    orc = OracleScene(v, f)
    nan, inf = np.nan, np.inf
    qs = np.array([[nan, 0, 0], [0, nan, 0], [nan, nan, nan], [inf, 0, 0], [-inf, 0, 0], [inf, inf, inf], [1e30, 0, 0], [3e38, 3e38, 3e38],
                   [0, 0, 0], [1e-30, 0, 0], [1, 0, 0], [1e-45, -1e-45, 0]], np.float32)
    ds = np.array([[0, 0, 0], [nan, 0, 1], [inf, 0, 0], [1, 0, 0], [0, 0, 1], [0, 1, 0], [-1, 0, 0], [1, 1, 1], [0, 0, inf], [nan, nan, nan], [-1, 0, 0],
                   [1e-45, 0, 1]], np.float32)
    rs = np.array([inf, nan, 1, 0, inf, 1, inf, inf, nan, 1, -1, 1e-30], np.float32)
    n, k = 4096, len(qs)
    q = meshes.points_in_box(n, *meshes.mesh_bounds(v), 1.5, seed=61)
    d = meshes.unit_directions(n, seed=62)
    rmax = (orc.closest(q, nthreads=8)[1] * meshes.star_radius_scale(n)).astype(np.float32)
    at = np.arange(k) * 331 % n  # scattered through the batch
    q[at], rmax[at] = qs, rs
    d[at] = ds
    q[at[3:6] + 1] = 0.3  # ordinary points with special directions / radii next to them
    d[at[3:6] + 1], rmax[at[3:6] + 1] = ds[:3], rs[:3]
    sph = np.concatenate([q, np.abs(rmax)[:, None] + 0.05], axis=1).astype(np.float32)
    sph[at, 3] = np.array([1, 1, 1, 1, inf, 1, 1e30, inf, nan, 0, -1, 1e-30], np.float32)
    rnd = meshes.uniforms(n, 3, seed=63)
    o_dist = orc.closest(q, nthreads=8)[1]
    o_sil = [orc.silhouette(q, False, nthreads=8), orc.silhouette(q, True, nthreads=8), orc.silhouette(q, False, r_max=rmax, nthreads=8)]
    o_f, o_t, _, o_p = orc.ray(q, d, nthreads=8)
    o_ft, o_tt, _, _ = orc.ray(q, d, rmax, nthreads=8)
    o_si, o_pdf = orc.sample(sph, rnd[:, 0].copy())
    try:
        for knobs in ({}, {"query.wide_max_n": 0, "query.wide_max_n_sil": 0, "query.sort_min_n": 1, "query.ray_kernel": 2},
                      {"query.wide_max_n": 0, "query.wide_max_n_sil": 0, "query.sort_min_n": 0, "query.ray_kernel": 0}):
            for kk, val in {"query.wide_max_n": 2097152, "query.wide_max_n_sil": 262144, "query.sort_min_n": 16384, "query.ray_kernel": 1, **knobs}.items():
                sc.set_option(kk, val)
            _, dist = sc.closest_point(q)
            assert same_bits(dist, o_dist).all(), knobs
            for got, want in zip([sc.closest_silhouette(q), sc.closest_silhouette(q, flip=True), sc.closest_silhouette(q, r_max=rmax)], o_sil):
                assert same_bits(got, want).all(), knobs
            found, hits = sc.intersect(q, d)
            assert np.array_equal(found.astype(bool), o_f.astype(bool)) and same_bits(hits["t"], o_t).all(), knobs
            if knobs.get("query.ray_kernel", 1):
                assert np.array_equal(hits["prim"], o_p.astype(np.uint32)), knobs
            found, hits = sc.intersect(q, d, t_max=rmax)
            assert np.array_equal(found.astype(bool), o_ft.astype(bool)) and same_bits(hits["t"], o_tt).all(), knobs
            si, pdf, _ = sc.sample_in_sphere(sph, rnd)
            assert np.array_equal(si, o_si) and same_bits(pdf, o_pdf).all(), knobs
            r = sc.wost_step(q, d, rnd)  # the fused step on the same inputs: terminates, and its closest distances are the same
            assert same_bits(r["closest_distance"], o_dist).all(), knobs
    finally:
        for kk, val in {"query.wide_max_n": 2097152, "query.wide_max_n_sil": 262144, "query.sort_min_n": 16384, "query.ray_kernel": 1}.items():
            sc.set_option(kk, val)


def test_host_and_device_pointer_paths_agree(pkg, meshes):
    import torch
    v, f = meshes.bumpy_torus(64, 48)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    lo, hi = meshes.mesh_bounds(v)
    q = meshes.points_in_box(10000, lo, hi, 1.3, seed=46)
    d = meshes.unit_directions(10000, seed=47)
    qd, dd = torch.from_numpy(q).cuda(), torch.from_numpy(d).cuda()
    ih, dh = sc.closest_point(q)
    it, dt = sc.closest_point(qd)
    torch.cuda.synchronize()
    assert np.array_equal(ih, it.cpu().numpy().view(np.uint32)) and np.array_equal(bits(dh), bits(dt.cpu().numpy()))
    assert np.array_equal(bits(sc.closest_silhouette(q)), bits(sc.closest_silhouette(qd).cpu().numpy()))
    (sdh, seh, sph_), (sdt, set_, spt) = sc.closest_silhouette(q, with_edge=True), sc.closest_silhouette(qd, with_edge=True)
    torch.cuda.synchronize()
    assert np.array_equal(bits(sdh), bits(sdt.cpu().numpy())) and np.array_equal(bits(sph_), bits(spt.cpu().numpy()))
    fh, hh = sc.intersect(q, d)
    ft, ht = sc.intersect(qd, dd)
    torch.cuda.synchronize()
    assert np.array_equal(fh, ft.cpu().numpy()) and np.array_equal(bits(hh["t"]), bits(ht.cpu().numpy()[:, 0]))
    # mixed pointer kinds are rejected, not guessed
    st = pkg.lib().snch_closest_point_batch(sc._h, qd.data_ptr(), 10, ih.ctypes.data, dh.ctypes.data, None)
    assert st == -5 and b"mixed" in pkg.lib().snch_last_error()


def test_results_do_not_depend_on_scheduling_knobs(pkg, meshes):
    """Query ordering, the guard-banded cone test, seeding, the tail hand-off, the ray kernel and grid size only change
    scheduling: bit-identical results for every knob setting, and the defaults' results are the oracle's."""
    v, f = meshes.bumpy_torus(120, 90)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    orc = OracleScene(v, f)
    lo, hi = meshes.mesh_bounds(v)
    n = 60000
    q = meshes.points_in_box(n, lo, hi, 1.3, seed=48)
    d = meshes.unit_directions(n, seed=49)
    flip = (np.arange(n) % 3 == 0).astype(np.uint8)
    rmax = (orc.closest(q, nthreads=8)[1] * meshes.star_radius_scale(n)).astype(np.float32)
    # the baseline forces the per-lane / packet kernels (the wide ones would take a batch this small); variants re-enable them
    defaults = {"query.sort_min_n": 16384, "query.sort_bits": 24, "query.sort_rays": 0, "query.cone_filter": 1, "query.seed": 1,
                "query.blocks_per_sm": 0, "query.host_chunk": 1 << 23, "query.host_first": 0, "query.host_split_min": 3 << 20, "query.sort_radius": 0, "query.sil_tail": 8, "query.sil_flush": 24, "query.sil_chunk": 0,
                "query.wide_max_n": 0, "query.wide_max_n_sil": 0, "query.ray_kernel": 2, "query.ray_flush": 8, "query.ray_refill": 8}

    def run():
        idx, dist = sc.closest_point(q)
        found, hits = sc.intersect(q, d)
        sidx, pdf, _ = sc.sample_in_sphere(np.concatenate([q, rmax[:, None] + 0.05], 1).astype(np.float32), meshes.uniforms(n, 3, seed=50))
        sile, edge, pt = sc.closest_silhouette(q, flip=flip, with_edge=True)
        return dict(dist=dist, sil=sc.closest_silhouette(q, flip=flip), silr=sc.closest_silhouette(q, r_max=rmax), sile=sile, found=found,
                    t=hits["t"].copy(), u=hits["u"].copy(), v=hits["v"].copy(), prim=hits["prim"].copy(), sidx=sidx, pdf=pdf), idx, edge, pt

    for k, val in defaults.items():
        sc.set_option(k, val)
    base, idx0, edge0, pt0 = run()
    check_closest(q[:5000], idx0[:5000], base["dist"][:5000], orc)
    assert np.array_equal(bits(base["sil"]), bits(base["sile"])), "asking for the edge changed a distance"
    for fl in (0, 1):  # per-query flip bytes: each subset must match the oracle run with that scalar flag
        sel = np.nonzero(flip[:6000] == fl)[0]
        check_silhouette_edges(q[sel], base["sile"][sel], edge0[sel], pt0[sel], orc, bool(fl))
    check_rays(base["found"][:5000], {k: base[k][:5000] for k in ("t", "u", "v", "prim")}, q[:5000], d[:5000], None, orc)
    variants = [{"query.sort_min_n": 0}, {"query.cone_filter": 0}, {"query.cone_filter": 0, "query.sort_min_n": 0}, {"query.seed": 3}, {"query.seed": 2},
                {"query.seed": 0}, {"query.wide_max_n": 1 << 30}, {"query.wide_max_n": 1 << 30, "query.sort_min_n": 0}, {"query.wide_max_n": 1 << 30, "query.seed": 0},
                {"query.wide_max_n": 1 << 30, "query.seed": 3}, {"query.wide_max_n_sil": 1 << 30}, {"query.wide_max_n_sil": 1 << 30, "query.cone_filter": 0},
                {"query.wide_max_n_sil": 1 << 30, "query.sort_min_n": 0},
                {"query.sil_flush": 1}, {"query.sil_flush": 32}, {"query.sil_chunk": 8}, {"query.sil_chunk": 64}, {"query.sil_chunk": 32, "query.sil_tail": 0}, {"query.sil_flush": 12, "query.cone_filter": 0}, {"query.sil_tail": 0}, {"query.sil_tail": 31}, {"query.sil_tail": 31, "query.sort_min_n": 0}, {"query.sil_tail": 31, "query.cone_filter": 0},
                {"query.sil_tail": 16, "query.blocks_per_sm": 1}, {"query.sort_radius": 1}, {"query.sort_radius": 2}, {"query.sort_radius": 3},
                {"query.sort_rays": 1}, {"query.sort_rays": 2}, {"query.sort_rays": -1}, {"query.ray_kernel": 0}, {"query.ray_kernel": 1}, {"query.ray_kernel": 0, "query.sort_rays": 1},
                {"query.ray_flush": 1, "query.ray_refill": 1}, {"query.ray_flush": 32, "query.ray_refill": 32}, {"query.ray_flush": 16, "query.sort_rays": 1},
                {"query.sort_bits": 12}, {"query.blocks_per_sm": 1},
                {"query.host_chunk": 7001}, {"query.host_chunk": 0}, {"query.host_chunk": 500, "query.sort_min_n": 0},  # host-pointer pipeline
                {"query.host_first": 1000, "query.host_split_min": 1}, {"query.host_first": 20000, "query.host_split_min": 1, "query.host_chunk": 17000},
                {"query.host_first": -1, "query.host_chunk": 25000}, {"query.host_first": 1, "query.host_chunk": 3000},
                {"query.sort_min_n": 0, "query.cone_filter": 0, "query.seed": 0, "query.sil_tail": 0, "query.ray_kernel": 0}]
    for kv in variants:
        for k, val in {**defaults, **kv}.items():
            sc.set_option(k, val)
        other, idx1, edge1, pt1 = run()
        for k in base:
            a, b = base[k], other[k]
            assert np.array_equal(bits(a) if a.dtype == np.float32 else a, bits(b) if b.dtype == np.float32 else b), (kv, k)
        # indices may differ only between triangles / edges at the same distance (documented tie rule, Q3)
        assert np.array_equal(bits(orc.point_triangle_distance(q, idx1)), bits(base["dist"])), kv
        fin = np.isfinite(base["sile"])
        d_at, p_at = orc.point_edge_distance(q[fin], edge1[fin].astype(np.int32))
        assert np.array_equal(bits(d_at), bits(base["sile"][fin])) and np.array_equal(bits(p_at), bits(pt1[fin])), kv
    with pytest.raises(pkg.SnchError):
        sc.set_option("query.no_such_knob", 1)


@pytest.mark.parametrize("case", ["offset", "stretched", "near_surface", "tiny"])
def test_triangle_lower_bound_never_changes_a_distance(pkg, meshes, case):
    """The plane-and-reach rejection in front of the point-triangle distance (query.cu tri_cannot_improve) must be invisible:
    bit-identical distances with it on ("query.seed" 1) and off (3), in the packet and the one-query-per-warp kernels, on
    geometry chosen to stress its rounding margins — coordinates far from the origin (absolute rounding of the reference's
    closest point grows with the coordinate magnitude), sliver triangles (ill-conditioned normals), queries a few ulps off
    the surface, and a mesh scaled to 1e-3."""
    import torch
    v, f = meshes.bumpy_torus(160, 120)
    rng = np.random.default_rng(17)
    n = 400000
    lo, hi = meshes.mesh_bounds(v)
    q = meshes.points_in_box(n, lo, hi, 1.2, seed=18)
    if case == "offset":
        shift = np.array([1000.0, -2000.0, 500.0], np.float32)
        v, q = (v + shift).astype(np.float32), (q + shift).astype(np.float32)
    elif case == "stretched":
        scale = np.array([300.0, 1.0, 0.01], np.float32)
        v, q = (v * scale).astype(np.float32), (q * scale).astype(np.float32)
    elif case == "near_surface":
        tri = v[f[rng.integers(0, len(f), n)]]
        w = rng.dirichlet([1.0, 1.0, 1.0], n).astype(np.float32)
        q = (np.einsum("nk,nkd->nd", w, tri) + rng.normal(0.0, 2e-6, (n, 3))).astype(np.float32)
    elif case == "tiny":
        v, q = (v * 1e-3).astype(np.float32), (q * 1e-3).astype(np.float32)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    qd = torch.from_numpy(q).cuda()
    out = {}
    for wide in (0, 1 << 30):
        for seed in (1, 3):
            sc.set_option("query.wide_max_n", wide).set_option("query.seed", seed)
            idx, dist = sc.closest_point(qd)
            out[(wide, seed)] = dist.clone()
    ref = out[(0, 3)]
    for k, d in out.items():
        assert torch.equal(d.view(torch.int32), ref.view(torch.int32)), (case, k)
