"""Seeded fuzz of the whole path on the GPU against the CPU oracle: random triangle SOUPS — non-manifold edges shared by many
triangles, duplicated triangles, inconsistent orientation, boundary edges, self-intersections, slivers, far-from-origin
and tiny coordinates — the inputs the structured meshes of the other tests never produce.  Build arrays, adjacency and every
query kind must match the oracle the same way they do on the structured meshes (bit for bit where parity.py says so)."""
import numpy as np
import pytest

from oracle import OracleScene
from parity import bits, check_build_vs_oracle, check_closest, check_rays_exact, check_silhouette, check_silhouette_edges, same_bits

pytestmark = pytest.mark.gpu

# (seed, vertices, triangles, offset, scale, sliver): few vertices and many triangles = heavily non-manifold
CASES = [
    (1, 4, 7, 0.0, 1.0, False),
    (2, 6, 40, 0.0, 1.0, False),
    (3, 12, 100, 0.0, 1.0, False),
    (4, 50, 60, 0.0, 1.0, False),
    (5, 200, 1000, 0.0, 1.0, False),
    (6, 30, 3000, 0.0, 1.0, False),
    (7, 300, 500, 1000.0, 1.0, False),   # far from the origin: absolute rounding of every coordinate-sized quantity
    (8, 300, 500, 0.0, 1e-3, False),     # tiny geometry
    (9, 100, 400, 0.0, 1.0, True),       # slivers: one coordinate squeezed by 1e-4
    (10, 2000, 20000, 0.0, 1.0, False),  # enough triangles for several sort tiles and refit CTAs
]


def _extra_cases():
    """SNCH_FUZZ_EXTRA=N adds N more soups with drawn parameters (seeds SNCH_FUZZ_FIRST = 101 ...): the long sweep of tools/gpu_fuzz_sweep.sh"""
    import os
    n, first = int(os.environ.get("SNCH_FUZZ_EXTRA", "0")), int(os.environ.get("SNCH_FUZZ_FIRST", "101"))
    out = []
    for seed in range(first, first + n):
        rng = np.random.default_rng(seed)
        nv = int(rng.choice([5, 12, 40, 150, 600, 2500]))
        nt = int(nv * rng.choice([0.7, 1.5, 2.0, 4.0, 20.0])) + 4
        offset = float(rng.choice([0.0, 0.0, 3.0, 100.0, 1000.0, -5000.0]))
        scale = float(rng.choice([1.0, 1.0, 1e-3, 50.0]))
        out.append((seed, nv, nt, offset, scale, bool(rng.random() < 0.2)))
    return out


CASES += _extra_cases()


def soup(seed, nv, nt, offset, scale, sliver):
    rng = np.random.default_rng(seed)
    v = rng.random((nv, 3))
    if sliver:
        v[:, 2] *= 1e-4
    v = (v * scale + offset).astype(np.float32)
    f = rng.integers(0, nv, (nt, 3))
    for _ in range(64):  # no repeated vertex inside a triangle (exactly degenerate input is test_gpu_adjacency's business)
        bad = (f[:, 0] == f[:, 1]) | (f[:, 1] == f[:, 2]) | (f[:, 0] == f[:, 2])
        if not bad.any():
            break
        f[bad] = rng.integers(0, nv, (int(bad.sum()), 3))
    f = f[~((f[:, 0] == f[:, 1]) | (f[:, 1] == f[:, 2]) | (f[:, 0] == f[:, 2]))]
    if nt >= 40:
        f = np.concatenate([f, f[:5], f[:3, ::-1]])  # exact duplicates, and the same triangles with the opposite orientation
    return v, f.astype(np.int32)


@pytest.fixture(scope="module", params=CASES, ids=lambda c: f"seed{c[0]}_v{c[1]}_t{c[2]}")
def fuzz_scene(request, pkg, meshes):
    v, f = soup(*request.param)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    orc = OracleScene(v, f)
    lo, hi = meshes.mesh_bounds(v)
    n = 6000
    q = meshes.points_in_box(n, lo, hi, 1.5, seed=1000 + request.param[0])
    d = meshes.unit_directions(n, seed=2000 + request.param[0])
    return sc, orc, q, d, pkg


def test_fuzz_build(fuzz_scene):
    sc, orc, _, _, pkg = fuzz_scene
    check_build_vs_oracle(sc, orc, pkg)


def test_fuzz_closest(fuzz_scene):
    sc, orc, q, _, _ = fuzz_scene
    idx, dist = sc.closest_point(q)
    check_closest(q, idx, dist, orc)


@pytest.mark.parametrize("flip", [False, True])
def test_fuzz_silhouette(fuzz_scene, meshes, flip):
    sc, orc, q, _, _ = fuzz_scene
    dist = sc.closest_silhouette(q, flip=flip)
    check_silhouette(dist, orc.silhouette(q, flip, nthreads=8))
    rmax = (orc.closest(q, nthreads=8)[1] * meshes.star_radius_scale(len(q))).astype(np.float32)
    dist_r, edge, pt = sc.closest_silhouette(q, flip=flip, r_max=rmax, with_edge=True)
    check_silhouette(dist_r, orc.silhouette(q, flip, r_max=rmax, nthreads=8))
    check_silhouette_edges(q, dist_r, edge, pt, orc, flip, r_max=rmax)
    # bounded == filtered unbounded (Q5) — except that a radius whose square is zero finds nothing (scene.cuh:791: min_r2 >= max_r2)
    assert np.array_equal(bits(dist_r), bits(np.where((dist <= rmax) & (rmax * rmax > 0), dist, np.inf).astype(np.float32)))


@pytest.mark.parametrize("ray_kernel", [1, 2])
def test_fuzz_rays(fuzz_scene, ray_kernel):
    sc, orc, q, d, _ = fuzz_scene
    sc.set_option("query.ray_kernel", ray_kernel)
    found, hits = sc.intersect(q, d)
    tm = np.full(len(q), 0.4 * float(np.ptp(orc.verts, axis=0).max()), np.float32)
    found_t, hits_t = sc.intersect(q, d, t_max=tm)
    sc.set_option("query.ray_kernel", 1)
    check_rays_exact(found, hits, q, d, None, orc)
    check_rays_exact(found_t, hits_t, q, d, tm, orc)


@pytest.mark.parametrize("reverse", [False, True])
def test_zero_area_face_is_not_a_missing_face(pkg, meshes, reverse):
    """A ZERO-AREA triangle D (three collinear vertices) whose three edges are each shared with a proper triangle: no edge of D
    is a boundary edge, but each has a face whose normal is normalize(0) = NaN.  The reference's silhouette test then answers
    "no" (every comparison with NaN is false), whereas an edge with a MISSING face is always a silhouette.  `reverse` flips D, which
    moves the NaN between the edge's first and second face."""
    v = np.array([[0, 0, 0], [1, 0, 0], [1, 2, 0], [2, 0, 0], [1.5, -1, 0.5], [0.5, -1, -0.5]], np.float32)
    d = [0, 1, 3] if reverse else [0, 3, 1]
    f = np.array([d, [3, 0, 2], [1, 3, 4], [0, 1, 5]], np.int32)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    orc = OracleScene(v, f)
    e4, _, _ = orc.adjacency()
    if not reverse:
        assert sum(1 for e in e4 if e[0] != -1 and e[3] != -1) >= 3  # D's three edges have two faces each (reversed: same-direction half-edges overwrite a slot, Q18)
    check_build_vs_oracle(sc, orc, pkg)
    q = meshes.points_in_box(6000, np.array([-0.2, -0.3, -0.3], np.float32), np.array([2.2, 0.3, 0.3], np.float32), 1.0, seed=5)  # around the line
    for flip in (False, True):
        check_silhouette(sc.closest_silhouette(q, flip=flip), orc.silhouette(q, flip, nthreads=4))
        sc.set_option("query.wide_max_n_sil", 0)
        check_silhouette(sc.closest_silhouette(q, flip=flip), orc.silhouette(q, flip, nthreads=4))
        sc.set_option("query.wide_max_n_sil", 262144)
    check_closest(q, *sc.closest_point(q), orc)


def test_fuzz_sample(fuzz_scene, meshes):
    sc, orc, q, _, _ = fuzz_scene
    _, dcp = orc.closest(q, nthreads=8)
    sph = np.concatenate([q, (dcp * 1.5 + 0.05 * float(dcp.max()))[:, None]], axis=1).astype(np.float32)
    rnd = meshes.uniforms(len(q), 3, seed=43)
    idx, pdf, pt = sc.sample_in_sphere(sph, rnd)
    idx_o, pdf_o = orc.sample(sph, rnd[:, 0].copy())
    assert np.array_equal(idx, idx_o), f"sampled primitive differs on {np.count_nonzero(idx != idx_o)} of {len(idx)}"
    hit = idx >= 0
    assert np.array_equal(bits(pdf[hit]), bits(pdf_o[hit])), "sampling pdf differs"
    pt_o = orc.sample_on_object(idx, rnd[:, 1].copy(), rnd[:, 2].copy())
    assert np.array_equal(bits(pt[hit]), bits(pt_o[hit])), "sampled point differs"


def test_fuzz_wost_step(fuzz_scene, meshes):
    """the fused wavefront step equals its four stages on these inputs too"""
    sc, _, q, d, _ = fuzz_scene
    rnd = meshes.uniforms(len(q), 3, seed=44)
    r = sc.wost_step(q, d, rnd)
    idx, dist = sc.closest_point(q)
    assert np.array_equal(bits(r["closest_distance"]), bits(dist))
    sd = sc.closest_silhouette(q, r_max=dist)
    assert np.array_equal(bits(r["silhouette_distance"]), bits(sd))
    star = np.minimum(dist, sd)
    assert np.array_equal(bits(r["star_radius"]), bits(star))
    found, hits = sc.intersect(q, d, t_max=star)
    assert np.array_equal(r["found"].astype(bool), found.astype(bool))
    assert np.array_equal(bits(r["hits"]["t"]), bits(hits["t"]))


def test_fuzz_large_batch_kernels(fuzz_scene, meshes):
    """A 6 000-query batch takes the one-query-per-warp kernels; the kernels of the 16M-query batches (packets, per-lane walks
    with the warp leaf queue and its tail launch, parked rays) must give the same bits on the same soups."""
    sc, orc, q, d, _ = fuzz_scene
    _, dist = sc.closest_point(q)
    rmax = (dist * meshes.star_radius_scale(len(q))).astype(np.float32)
    ref = [dist, sc.closest_silhouette(q), sc.closest_silhouette(q, flip=True), sc.closest_silhouette(q, r_max=rmax)]
    ref_e = sc.closest_silhouette(q, r_max=rmax, with_edge=True)
    kernels = set()
    try:
        sc.set_option("query.wide_max_n", 0).set_option("query.wide_max_n_sil", 0).set_option("query.sort_min_n", 1)
        got = [sc.closest_point(q)[1]]
        kernels.add(sc.last_kernel())
        got += [sc.closest_silhouette(q), sc.closest_silhouette(q, flip=True), sc.closest_silhouette(q, r_max=rmax)]
        kernels.add(sc.last_kernel())
        got_e = sc.closest_silhouette(q, r_max=rmax, with_edge=True)
    finally:
        sc.set_option("query.wide_max_n", 2097152).set_option("query.wide_max_n_sil", 262144).set_option("query.sort_min_n", 16384)
    assert {"k_closest_packet", "k_silhouette_coop"} <= kernels, kernels
    for a, b in zip(ref, got):
        assert np.array_equal(bits(a), bits(b))
    assert np.array_equal(bits(ref_e[0]), bits(got_e[0]))
    check_silhouette_edges(q, got_e[0], got_e[1], got_e[2], orc, False, r_max=rmax)


# ---- 2-D: random segment soups (vertices of any valence, duplicated and reversed segments, crossings) -------------------------
# (seed, vertices, segments, offset, scale)
CASES2 = [
    (1, 4, 5, 0.0, 1.0),
    (2, 8, 40, 0.0, 1.0),
    (3, 100, 300, 0.0, 1.0),
    (4, 300, 500, 1000.0, 1.0),  # far from the origin
    (5, 300, 500, 0.0, 1e-3),    # tiny geometry
    (6, 3000, 20000, 0.0, 1.0),
]


def _extra_cases2():
    import os
    n, first = int(os.environ.get("SNCH_FUZZ_EXTRA", "0")), int(os.environ.get("SNCH_FUZZ_FIRST", "101")) + 100
    out = []
    for seed in range(first, first + n):
        rng = np.random.default_rng(seed)
        nv = int(rng.choice([4, 10, 60, 400, 3000]))
        ns = int(nv * rng.choice([0.8, 1.0, 2.0, 6.0])) + 3
        out.append((seed, nv, ns, float(rng.choice([0.0, 0.0, 3.0, 1000.0, -5000.0])), float(rng.choice([1.0, 1.0, 1e-3, 50.0]))))
    return out


CASES2 += _extra_cases2()


def soup2(seed, nv, ns, offset, scale):
    rng = np.random.default_rng(seed)
    v = (rng.random((nv, 2)) * scale + offset).astype(np.float32)
    s = rng.integers(0, nv, (ns, 2))
    s = s[s[:, 0] != s[:, 1]]
    if ns >= 40:
        s = np.concatenate([s, s[:5], s[:3, ::-1]])
    return v, s.astype(np.int32)


@pytest.mark.parametrize("case", CASES2, ids=lambda c: f"seed{c[0]}_v{c[1]}_s{c[2]}")
def test_fuzz_2d(pkg, meshes, case):
    """Every product of a 2-D scene against the 2-D oracle, bit for bit, with the one-query-per-warp kernels (the default for a
    batch this small) and with the per-lane kernels of the large batches."""
    from oracle import OracleScene2
    v, s = soup2(*case)
    sc = pkg.Scene2(v, s).compute_silhouettes().build_bvh()
    orc = OracleScene2(v, s)
    nodes, aabbs, cones, q1 = orc.tree()
    assert np.array_equal(sc.export(pkg.ExportKind.NODES), nodes) and np.array_equal(bits(sc.export(pkg.ExportKind.AABBS)), bits(aabbs))
    mine = sc.export(pkg.ExportKind.CONES)
    keep = (cones[:, 2] >= 0) & ~q1.astype(bool)
    assert np.array_equal(mine[:, 2] >= 0, cones[:, 2] >= 0) and same_bits(mine[keep], cones[keep]).all(), "2-D cones differ from the oracle's"
    n = 6000
    q = meshes.points_in_box2(n, v.min(0), v.max(0), 1.3, seed=3000 + case[0])
    d = meshes.unit_directions2(n, seed=4000 + case[0])
    _, odist = orc.closest(q)
    rmax = (odist * meshes.star_radius_scale(n)).astype(np.float32)
    osil = [orc.silhouette(q, False), orc.silhouette(q, True), orc.silhouette(q, False, rmax)]
    of, ot, _, op = orc.ray(q, d)
    tm = np.full(n, 0.4 * case[4], np.float32)
    of_t, ot_t, _, op_t = orc.ray(q, d, tm)
    sph = np.concatenate([q, (odist * 1.5 + 0.05 * float(odist.max()))[:, None]], axis=1).astype(np.float32)
    u = meshes.uniforms(n, 2, seed=45)
    oi, opdf = orc.sample(sph, u[:, 0].copy())
    for wide in (131072, 0):
        sc.set_option("query.wide_max_n", wide).set_option("query.sort_min_n", 16384 if wide else 1)
        _, dist = sc.closest_point(q)
        assert np.array_equal(bits(dist), bits(odist)), f"closest distance (wide_max_n={wide})"
        sil = [sc.closest_silhouette(q), sc.closest_silhouette(q, flip=True), sc.closest_silhouette(q, r_max=rmax)]
        for a, b in zip(sil, osil):
            assert np.array_equal(bits(a), bits(b)), f"silhouette distance (wide_max_n={wide})"
        dv, vid, pt = sc.closest_silhouette(q, r_max=rmax, with_vertex=True)
        assert np.array_equal(bits(dv), bits(osil[2]))
        fin = np.isfinite(dv)
        assert np.all(vid[~fin] == 0xFFFFFFFF)
        assert np.allclose(np.linalg.norm(pt[fin].astype(np.float64) - q[fin], axis=1), dv[fin], rtol=1e-5, atol=1e-6 * case[4] + 1e-7 * abs(case[3]))
        found, hits = sc.intersect(q, d)
        assert np.array_equal(found.astype(bool), of.astype(bool)) and np.array_equal(bits(hits["t"]), bits(ot)), f"rays (wide_max_n={wide})"
        assert np.array_equal(hits["prim"], op.astype(np.uint32)), "ray segment (the walk follows the reference's order: ties included)"
        found_t, hits_t = sc.intersect(q, d, t_max=tm)
        assert np.array_equal(found_t.astype(bool), of_t.astype(bool)) and np.array_equal(bits(hits_t["t"]), bits(ot_t))
        assert np.array_equal(hits_t["prim"], op_t.astype(np.uint32))
        si, pdf, _ = sc.sample_in_sphere(sph, u)
        assert np.array_equal(si, oi) and np.array_equal(bits(pdf), bits(opdf)), f"sample (wide_max_n={wide})"
