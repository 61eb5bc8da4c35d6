"""Size-independent properties of the oracle itself (the same properties the GPU tests use at full size)."""
import numpy as np

from oracle import OracleScene
from parity import bits, rel_close


def test_bruteforce_agreement(meshes):
    m = meshes
    v, f = m.bumpy_torus(40, 32)
    o = OracleScene(v, f)
    lo, hi = m.mesh_bounds(v)
    q = m.points_in_box(1500, lo, hi, 1.4, seed=31)
    d = m.unit_directions(1500, seed=32)
    _, dist = o.closest(q)
    _, dist_b = o.closest(q, brute=True)
    assert rel_close(dist, dist_b, 1e-6).all()  # the minmaxdist pruning never loses the minimum
    f1, t1, _, _ = o.ray(q, d)
    f2, t2, _, _ = o.ray(q, d, brute=True)
    assert np.array_equal(f1, f2) and np.array_equal(bits(t1), bits(t2))
    fa, _, _, _ = o.ray(q, d, any_hit=True)
    assert np.array_equal(fa, f1)


def test_bounded_silhouette_equals_filtered_unbounded(meshes):
    m = meshes
    v, f = m.bumpy_torus(40, 32)
    o = OracleScene(v, f)
    lo, hi = m.mesh_bounds(v)
    q = m.points_in_box(2000, lo, hi, 1.1, seed=33)
    _, dcp = o.closest(q)
    rmax = (dcp * m.star_radius_scale(2000)).astype(np.float32)
    unb = o.silhouette(q)
    bnd = o.silhouette(q, r_max=rmax)
    assert np.array_equal(bits(bnd), bits(np.where(unb <= rmax, unb, np.inf).astype(np.float32)))


def test_sample_pdf_is_path_probability_over_area(meshes):
    m = meshes
    v, f = m.icosphere(3)
    o = OracleScene(v, f)
    q = m.points_in_box(500, [-1, -1, -1], [1, 1, 1], 1.0, seed=34)
    sph = np.concatenate([q, np.full((500, 1), 0.8, np.float32)], axis=1)
    u = m.uniforms(500, seed=35)
    idx, pdf = o.sample(sph, u)
    hit = idx >= 0
    assert hit.any() and np.all(pdf[hit] > 0) and np.all(pdf[~hit] == 0)
    pts = o.sample_on_object(idx, m.uniforms(500, seed=36), m.uniforms(500, seed=37))
    d = o.point_triangle_distance(pts[hit], idx[hit].astype(np.uint32))
    assert np.all(d < 1e-5)  # sampled points lie on the chosen triangle
