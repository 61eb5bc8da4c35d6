"""2-D scenes (lbvh::scene<2>) without a GPU: the committed golden vectors are what the UNMODIFIED reference headers produce
on the CPU (oracle/_ref, when present), and the host-side logic of the C-ABI scene (silhouette records, argument checks,
no fallback) works on a CPU-only box."""
import os

import numpy as np
import pytest

from conftest import small_cases2
from oracle import ref_available

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold(name):
    return np.load(os.path.join(GOLD, f"{name}.npz"))


@pytest.mark.parametrize("name", ["poly_circle", "poly_open", "poly_soup"])
def test_goldens_are_the_reference(pkg, name):
    if not ref_available("cpu"):
        pytest.skip("oracle/_ref/libsnch_ref_cpu.so not built (needs /root/reference)")
    from oracle import RefScene2
    g = _gold(name)
    v, s = small_cases2(pkg.meshes)[name]
    assert np.array_equal(v, g["verts"]) and np.array_equal(s, g["segs"])  # generators still produce the fixture's inputs
    r = RefScene2(v, s, "cpu")
    nodes, aabbs, cones = r.tree()
    assert np.array_equal(nodes, g["nodes"])
    assert np.array_equal(aabbs.view(np.uint32), g["aabbs"].view(np.uint32))
    assert np.array_equal(cones.view(np.uint32), g["cones"].view(np.uint32))
    v4, owned = r.adjacency()
    assert np.array_equal(v4, g["vert4"]) and np.array_equal(owned, g["owned"])
    ci, cd = r.closest(g["q"])
    assert np.array_equal(cd.view(np.uint32), g["closest_dist"].view(np.uint32)) and np.array_equal(ci, g["closest_idx"])
    assert np.array_equal(r.silhouette(g["q"], False).view(np.uint32), g["sil_noflip"].view(np.uint32))
    assert np.array_equal(r.silhouette(g["q"], True).view(np.uint32), g["sil_flip"].view(np.uint32))
    f, t, s_, p = r.ray(g["q"], g["d"])
    assert np.array_equal(f, g["ray_found"]) and np.array_equal(t.view(np.uint32), g["ray_t"].view(np.uint32)) and np.array_equal(p, g["ray_prim"])
    si, sp = r.sample(g["sph"], g["u"])
    assert np.array_equal(si, g["sample_idx"]) and np.array_equal(sp.view(np.uint32), g["sample_pdf"].view(np.uint32))


@pytest.mark.parametrize("name", ["poly_circle", "poly_open", "poly_soup"])
def test_silhouette_records_match_reference(pkg, name):
    """compute_silhouettes (scene.cuh:629-656) runs on the host: one {previous, self, next} record per vertex, later segments
    overwrite earlier ones — identical to the reference's on inconsistently oriented input too."""
    g = _gold(name)
    sc = pkg.Scene2(g["verts"], g["segs"]).compute_silhouettes()
    assert np.array_equal(sc.export(pkg.ExportKind.EDGES), g["vert4"])


def _same_cones(a, b, q1=None):
    """bit-equal where defined: invalid cones carry only their half-angle (axis / radius are never read, the reference leaves
    them uninitialised), Q1-tainted ones have an indeterminate half-angle in the reference"""
    valid = a[:, 2] >= 0
    assert np.array_equal(valid, b[:, 2] >= 0) and np.array_equal(a[~valid, 2], b[~valid, 2])
    keep = valid if q1 is None else valid & ~q1.astype(bool)
    assert np.array_equal(a[keep].view(np.uint32), b[keep].view(np.uint32))


@pytest.mark.parametrize("name", ["poly_circle", "poly_open", "poly_soup"])
def test_oracle2_matches_goldens(name):
    """The plain-C restatement of lbvh::scene<2> (oracle/snch_oracle2.c) reproduces what the reference produced, bit for bit."""
    from oracle import OracleScene2
    g = _gold(name)
    o = OracleScene2(g["verts"], g["segs"])
    nodes, aabbs, cones, q1 = o.tree()
    assert np.array_equal(nodes, g["nodes"]) and np.array_equal(aabbs.view(np.uint32), g["aabbs"].view(np.uint32))
    _same_cones(cones, g["cones"], q1)
    v4, owned = o.adjacency()
    assert np.array_equal(v4, g["vert4"]) and np.array_equal(owned, g["owned"])
    ci, cd = o.closest(g["q"])
    assert np.array_equal(cd.view(np.uint32), g["closest_dist"].view(np.uint32)) and np.array_equal(ci, g["closest_idx"])
    assert np.array_equal(o.silhouette(g["q"], False).view(np.uint32), g["sil_noflip"].view(np.uint32))
    assert np.array_equal(o.silhouette(g["q"], True).view(np.uint32), g["sil_flip"].view(np.uint32))
    f, t, s_, p = o.ray(g["q"], g["d"])
    assert np.array_equal(f, g["ray_found"]) and np.array_equal(t.view(np.uint32), g["ray_t"].view(np.uint32)) and np.array_equal(p, g["ray_prim"])
    assert np.array_equal(s_.view(np.uint32), g["ray_s"].view(np.uint32))
    f, t, _, p = o.ray(g["q"], g["d"], g["tmax"])
    assert np.array_equal(f, g["ray_found_tmax"]) and np.array_equal(t.view(np.uint32), g["ray_t_tmax"].view(np.uint32)) and np.array_equal(p, g["ray_prim_tmax"])
    si, sp = o.sample(g["sph"], g["u"])
    assert np.array_equal(si, g["sample_idx"]) and np.array_equal(sp.view(np.uint32), g["sample_pdf"].view(np.uint32))


def test_oracle2_is_the_reference(pkg):
    """... and the reference itself, run here on larger polylines (closed, open, shuffled / inconsistently oriented)."""
    if not ref_available("cpu"):
        pytest.skip("oracle/_ref/libsnch_ref_cpu.so not built (needs /root/reference)")
    from oracle import OracleScene2, RefScene2
    m = pkg.meshes
    for v, s in (m.wavy_circle(3000, 7, 0.25), m.open_polyline(449), m.polyline_soup(3, 5, 200)):
        o, r = OracleScene2(v, s), RefScene2(v, s, "cpu")
        on, oa, oc, q1 = o.tree()
        rn, ra, rc = r.tree()
        assert np.array_equal(on, rn) and np.array_equal(oa.view(np.uint32), ra.view(np.uint32))
        _same_cones(oc, rc, q1)
        q = m.points_in_box2(4000, v.min(0), v.max(0), 1.4, seed=3)
        d = m.unit_directions2(4000, seed=4)
        (oi, od), (ri, rd) = o.closest(q), r.closest(q)
        assert np.array_equal(od.view(np.uint32), rd.view(np.uint32)) and np.array_equal(oi, ri)
        for fl in (False, True):
            assert np.array_equal(o.silhouette(q, fl).view(np.uint32), r.silhouette(q, fl).view(np.uint32))
        for tm in (None, np.full(4000, 0.3, np.float32)):
            (of, ot, os_, op), (rf, rt, rs, rp) = o.ray(q, d, tm), r.ray(q, d, tm)
            assert np.array_equal(of, rf) and np.array_equal(ot.view(np.uint32), rt.view(np.uint32)) and np.array_equal(op, rp)
        sph = np.concatenate([q, (od * 1.5 + 0.05)[:, None]], 1).astype(np.float32)
        u = m.uniforms(4000, seed=5)
        (oi, op), (ri, rp) = o.sample(sph, u), r.sample(sph, u)
        assert np.array_equal(oi, ri) and np.array_equal(op.view(np.uint32), rp.view(np.uint32))


def test_oracle2_is_the_reference_on_soups(pkg):
    """... and on the random segment soups of test_gpu_fuzz.py (vertices of any valence, duplicates, far / tiny coordinates)."""
    if not ref_available("cpu"):
        pytest.skip("oracle/_ref/libsnch_ref_cpu.so not built (needs /root/reference)")
    from oracle import OracleScene2, RefScene2
    from test_gpu_fuzz import CASES2, soup2
    m = pkg.meshes
    for case in CASES2[:5]:
        v, s = soup2(*case)
        o, r = OracleScene2(v, s), RefScene2(v, s, "cpu")
        on, oa, oc, q1 = o.tree()
        rn, ra, rc = r.tree()
        assert np.array_equal(on, rn) and np.array_equal(oa.view(np.uint32), ra.view(np.uint32))
        _same_cones(oc, rc, q1)
        q = m.points_in_box2(2000, v.min(0), v.max(0), 1.3, seed=3000 + case[0])
        d = m.unit_directions2(2000, seed=4000 + case[0])
        (oi, od), (ri, rd) = o.closest(q), r.closest(q)
        assert np.array_equal(od.view(np.uint32), rd.view(np.uint32))
        if not q1.any():  # (a Q1 node's half-angle is whatever the reference's stack held)
            for fl in (False, True):
                assert np.array_equal(o.silhouette(q, fl).view(np.uint32), r.silhouette(q, fl).view(np.uint32))
        (of, ot, _, _), (rf, rt, _, _) = o.ray(q, d), r.ray(q, d)
        assert np.array_equal(of, rf) and np.array_equal(ot.view(np.uint32), rt.view(np.uint32))
        sph = np.concatenate([q, (od * 1.5 + 0.05 * float(od.max()))[:, None]], 1).astype(np.float32)
        u = m.uniforms(2000, seed=5)
        (oi, op), (ri, rp) = o.sample(sph, u), r.sample(sph, u)
        assert np.array_equal(oi, ri) and np.array_equal(op.view(np.uint32), rp.view(np.uint32))


def test_argument_errors_2d(pkg):
    v = np.zeros((3, 2), np.float32)
    with pytest.raises(pkg.SnchError) as e:
        pkg.Scene2(v, np.array([[0, 3]], np.int32))
    assert e.value.status == -1 and "out of range" in str(e.value)
    sc = pkg.Scene2(v, np.array([[0, 1], [1, 2]], np.int32))
    with pytest.raises(pkg.SnchError) as e:
        sc.get_bvh_device_ptr()
    assert e.value.status == -2 and str(e.value) == "BVH is not built yet."  # scene.cuh:686
    with pytest.raises(pkg.SnchError) as e:
        sc.closest_point(np.zeros((1, 2), np.float32))
    assert e.value.status == -2
    with pytest.raises(pkg.SnchError):
        sc.set_option("query.no_such_knob", 1)


def test_no_cpu_fallback_2d(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    v, s = pkg.meshes.wavy_circle(64)
    sc = pkg.Scene2(v, s).compute_silhouettes()
    with pytest.raises(pkg.SnchError) as e:
        sc.build_bvh()
    assert e.value.status == -3 and "no CPU fallback" in str(e.value)
