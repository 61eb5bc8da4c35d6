"""Shared comparison helpers for the parity tests (test infrastructure)."""
import numpy as np

REL_TOL = 1e-5  # BASELINE.json north_star: distances / silhouette distances / hit t within 1e-5 relative


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def rel_close(a, b, rtol=REL_TOL, atol=1e-7):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    with np.errstate(invalid="ignore"):
        ok = np.abs(a - b) <= atol + rtol * np.maximum(np.abs(a), np.abs(b))
    return ok | both_inf


def check_build_vs_oracle(sc, orc, pkg, cone_rtol=REL_TOL, cone_atol=2e-6):
    """Integer pipeline bit-exact; AABBs bit-exact; cones within tolerance outside the Q1-tainted set."""
    K = pkg.ExportKind
    morton_o, sidx_o = orc.morton()
    assert np.array_equal(sc.export(K.MORTON_SORTED), morton_o), "Morton codes differ"
    assert np.array_equal(sc.export(K.SORTED_INDEX), sidx_o), "sort order differs"
    nodes_o, aabbs_o, cones_o = orc.tree()
    assert np.array_equal(sc.export(K.NODES), nodes_o), "tree topology differs"
    assert np.array_equal(sc.export(K.RANGES), orc.ranges()), "Karras leaf ranges differ"
    assert np.array_equal(bits(sc.export(K.AABBS)), bits(aabbs_o)), "AABBs differ"
    e_o, te_o, to_o = orc.adjacency()
    assert np.array_equal(sc.export(K.EDGES), e_o)
    assert np.array_equal(sc.export(K.TRI_EDGES), te_o)
    assert np.array_equal(sc.export(K.TRI_OWNED), to_o)
    st = sc.stats()
    assert bool(st["morton_collision"]) == orc.collision
    check_cones(sc.export(K.CONES), cones_o, sc.export(K.Q1_TAINT).astype(bool), orc.q1_taint(), cone_rtol, cone_atol)


def angle_close(a, b, rtol=REL_TOL, atol=2e-6, cos_tol=6e-7):
    """Half-angles are acos() of a float dot product: near 0 (and pi) acos amplifies one ulp of the dot product by
    1/sin(angle), so angles are accepted when they agree to rtol OR their cosines agree to a few ulps."""
    a64, b64 = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return rel_close(a, b, rtol, atol) | (np.abs(np.cos(a64) - np.cos(b64)) <= cos_tol)


def same_bits(a, b):
    """element-wise bit equality of two float32 arrays, any NaN equal to any NaN (x86 and the GPU have different default NaNs)"""
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def check_cones(cones, cones_o, taint, taint_o, rtol=None, atol=None):
    """Cones BIT-IDENTICAL to the oracle's (= the reference's CPU build): axis, half-angle and radius of every valid cone outside
    the Q1-tainted set.  The build evaluates acosf / sinf / cosf as the HOST's libm does (include/snch_lbvh/core/host_libm.cuh,
    tests/test_gpu_host_libm.py), so the rotation of merged axes — which amplifies one ulp by 1 / angle and hands it to every
    ancestor — starts from the same bits at every node.  (Until round 2 the device libm was used: 99.3 % of the half-angles within
    1e-5 at 1M triangles, the rest within 2.6e-4 rad.)  An axis can be NaN in both (opposite normals cancel: normalize(0))."""
    assert np.array_equal(taint, taint_o), "Q1 taint sets differ"
    valid_o = cones_o[:, 3] >= 0
    assert np.array_equal(cones[:, 3] >= 0, valid_o), "cone validity differs"
    assert np.array_equal(bits(cones[~valid_o, 3]), bits(cones_o[~valid_o, 3])), "invalid cones: half-angle marker differs"
    ok = ~taint_o & valid_o
    eq = same_bits(cones[ok], cones_o[ok]).all(axis=1)
    assert eq.all(), (f"{np.count_nonzero(~eq)} of {len(eq)} cones differ from the oracle's, e.g. node {np.nonzero(ok)[0][~eq][:3]}: "
                      f"{cones[ok][~eq][:3]} vs {cones_o[ok][~eq][:3]}")
    # tainted nodes: the reference leaves half_angle uninitialised there and the product defines it as pi; axis and
    # radius are defined in both
    t = taint_o & valid_o
    assert np.all(cones[t, 3] >= np.float32(np.pi / 2)), "tainted cones must stay non-pruning"
    assert same_bits(cones[t][:, [0, 1, 2, 4]], cones_o[t][:, [0, 1, 2, 4]]).all(), "tainted cones: axis / radius differ"


def check_closest(q, idx, dist, orc):
    """Distances BIT-IDENTICAL to the oracle's (same float sequence); the returned triangle attains the minimum (any member of
    the argmin set, Q3): the oracle's own point-triangle distance to it is that minimum, bit for bit."""
    idx = np.asarray(idx).astype(np.uint32)
    _, dist_o = orc.closest(q, nthreads=8)
    bad = bits(dist) != bits(dist_o)
    assert not bad.any(), f"closest distance differs from the oracle on {np.count_nonzero(bad)} of {len(bad)} queries"
    d_at = orc.point_triangle_distance(q, idx)
    assert np.array_equal(bits(d_at), bits(dist_o)), "returned triangle does not attain the minimum distance"


def check_silhouette(dist, dist_o, max_outlier_frac=0.0):
    """Silhouette distances.  Against the ORACLE the bar is bit equality on every query (max_outlier_frac = 0: the kernels
    take the reference's prune decisions and run its float sequence).  Against the reference's CUDA build (nvcc contracts its
    dot products into FMAs) a borderline cone test can flip, so callers pass the fraction the reference's own two builds
    disagree on; the rest must agree to 1e-5."""
    dist = np.asarray(dist)
    if max_outlier_frac == 0.0:
        bad = bits(dist) != bits(np.asarray(dist_o, np.float32))
        assert not bad.any(), f"silhouette distance differs from the oracle on {np.count_nonzero(bad)} of {len(bad)} queries"
        return 0.0
    ok = rel_close(dist, dist_o)
    frac = 1.0 - ok.mean() if len(ok) else 0.0
    assert frac <= max_outlier_frac, f"silhouette mismatch fraction {frac:.2e} ({np.count_nonzero(~ok)} of {len(ok)})"
    return frac


def check_silhouette_edges(q, dist, edge, point, orc, flip=False, r_max=None):
    """The optional outputs of a silhouette query against the oracle's silhouette_ex(): the distance is the oracle's bit for
    bit; the edge is ANY silhouette edge attaining it (exact ties are common: edges meeting at a vertex) — checked by recomputing
    the oracle's point-edge distance to the returned edge; the point is the oracle's closest point on the returned edge bit
    for bit, hence bit-identical to the oracle's own point whenever both chose the same edge."""
    edge = np.asarray(edge).astype(np.uint32)
    d_o, e_o, p_o = orc.silhouette_ex(q, flip, r_max=r_max, nthreads=8)
    assert np.array_equal(bits(dist), bits(d_o)), "silhouette distance differs from the oracle"
    fin = np.isfinite(d_o)
    assert np.all(edge[~fin] == 0xFFFFFFFF) and np.all(np.asarray(point)[~fin] == 0), "sentinels for queries without an answer"
    assert np.all(edge[fin] < orc.num_edges)
    d_at, p_at = orc.point_edge_distance(q[fin], edge[fin].astype(np.int32))
    assert np.array_equal(bits(d_at), bits(d_o[fin])), "returned edge does not attain the silhouette distance"
    assert np.array_equal(bits(p_at), bits(np.asarray(point)[fin])), "returned point is not the closest point on the returned edge"
    scale = np.maximum(np.abs(p_o[fin]).max(axis=1), 1e-30)
    same = edge[fin] == e_o[fin].astype(np.uint32)
    assert np.array_equal(bits(np.asarray(point)[fin][same]), bits(p_o[fin][same])), "same edge, different point"
    # A DIFFERENT attaining edge is an exact float tie.  Most ties are edges meeting at the closest point itself (a shared
    # vertex: same point).  The rest are a property of the problem, not of the kernel: the distance is flat to second order
    # around its minimiser, so points up to sqrt(2 d ulp(d)) ~ 5e-4 d apart along the silhouette round to the SAME float
    # distance (measured on the 1M-triangle torus: ~0.5% of queries have such a tie, tools/parity_report.py).  The point
    # of either edge is then an equally valid answer; it must stay inside that conditioning bound.
    # (Unrelated edges can also tie bit for bit by coincidence — about one query in 1e5 — so the bar is a fraction, not a maximum.)
    dev = np.abs(np.asarray(point)[fin] - p_o[fin]).max(axis=1)
    inside = dev <= 2e-3 * np.maximum(d_o[fin], scale * 1e-3) + 1e-6
    assert (inside.mean() >= 0.999) if fin.any() else True, f"silhouette point beyond the tie conditioning bound on {np.count_nonzero(~inside)} queries (max {dev.max()})"
    return float(same.mean()) if fin.any() else 1.0


def check_rays_exact(found, hits, q, d, tmax, orc):
    """The reference-order kernels ("query.ray_kernel" 1, 2): hit flag, t, (u, v) AND the triangle are the oracle's bit for bit —
    also where several triangles are hit at the same t (duplicates, shared edges, t = +0 / -0): the walk meets them in the
    reference's order and keeps the first."""
    f_o, t_o, uv_o, p_o = orc.ray(q, d, tmax, nthreads=8)
    assert np.array_equal(found.astype(bool), f_o.astype(bool)), "ray hit flags differ"
    assert np.array_equal(bits(hits["t"]), bits(t_o)), f"ray t differs on {np.count_nonzero(bits(hits['t']) != bits(t_o))} rays"
    assert np.array_equal(hits["prim"].astype(np.uint32), p_o.astype(np.uint32)), f"ray triangle differs on {np.count_nonzero(hits['prim'].astype(np.uint32) != p_o.astype(np.uint32))} rays"
    assert np.array_equal(bits(hits["u"]), bits(uv_o[:, 0])) and np.array_equal(bits(hits["v"]), bits(uv_o[:, 1])), "ray (u, v) differ"


def check_rays(found, hits, q, d, tmax, orc, max_tie_frac=1e-3):
    """Hit flags and t BIT-IDENTICAL to the oracle's walk; the triangle and (u, v) too, except on exact ties (Q4: two
    triangles sharing an edge are hit at the same t — the kernel that tests leaves where it meets them reaches them in another
    order than the reference's stack): there the returned triangle must attain that t, recomputed here in double precision."""
    found = np.asarray(found).astype(bool)
    f_o, t_o, uv_o, p_o = orc.ray(q, d, tmax, nthreads=8)
    f_o = f_o.astype(bool)
    assert np.array_equal(found, f_o), f"ray hit flags differ on {np.count_nonzero(found != f_o)} of {len(found)} rays"
    assert np.array_equal(bits(hits["t"]), bits(t_o)), f"ray t differs on {np.count_nonzero(bits(hits['t']) != bits(t_o))} rays"
    prim = np.asarray(hits["prim"]).astype(np.uint32)
    same = prim == p_o
    assert np.array_equal(bits(hits["u"])[same], bits(uv_o[:, 0])[same]) and np.array_equal(bits(hits["v"])[same], bits(uv_o[:, 1])[same]), "ray (u, v) differ"
    tie = np.nonzero(~same)[0]
    assert len(tie) <= max(1, int(len(found) * max_tie_frac)), f"ray triangle differs on {len(tie)} rays"  # (soups of duplicated triangles: any)
    if len(tie):
        tri = orc.verts[orc.tris[prim[tie]]].astype(np.float64)
        o64, d64 = np.asarray(q, np.float64)[tie], np.asarray(d, np.float64)[tie]
        e1, e2 = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
        h = np.cross(d64, e2)
        det = np.einsum("ij,ij->i", e1, h)
        sv = o64 - tri[:, 0]
        qv = np.cross(sv, e1)
        t64 = np.einsum("ij,ij->i", e2, qv) / det
        assert rel_close(t64, t_o[tie], 1e-5, 1e-6).all(), "a differing ray triangle does not attain the hit distance"
    assert np.all(np.isinf(np.asarray(hits["t"])[~found]))
    return found
