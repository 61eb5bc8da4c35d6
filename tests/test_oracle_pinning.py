"""Pins the C restatement against the reference's own code executed on the CPU (oracle/_ref/libsnch_ref_cpu.so).
Skipped where the reference build is not present (it cannot be rebuilt without /root/reference)."""
import numpy as np
import pytest

from oracle import OracleScene, RefScene, ref_available
from parity import bits

pytestmark = pytest.mark.skipif(not ref_available("cpu"), reason="oracle/_ref/libsnch_ref_cpu.so not built")


def _cases(m):
    return {"ico5": m.icosphere(5), "grid16": m.open_grid(16), "torus128": m.bumpy_torus(128, 128), "torus97x61": m.bumpy_torus(97, 61)}


@pytest.mark.parametrize("name", ["ico5", "grid16", "torus128", "torus97x61"])
def test_build_bit_exact(meshes, name):
    v, f = _cases(meshes)[name]
    o, r = OracleScene(v, f), RefScene(v, f, "cpu")
    on, oa, oc = o.tree()
    rn, ra, rc = r.tree()
    assert np.array_equal(on, rn) and np.array_equal(bits(oa), bits(ra))
    for a, b in zip(o.adjacency(), r.adjacency()):
        assert np.array_equal(a, b)
    om, osi = o.morton()
    rm, rsi = r.morton()
    assert np.array_equal(om, rm) and np.array_equal(osi, rsi)
    taint = o.q1_taint()
    valid = oc[:, 3] >= 0
    eq = np.where(valid, (bits(oc) == bits(rc)).all(axis=1), bits(oc)[:, 3] == bits(rc)[:, 3])
    assert eq[~taint].all()
    # where the reference's half_angle is indeterminate it still came out >= pi/2 here, i.e. non-pruning like our pi
    assert np.all(rc[taint & (np.arange(len(taint)) != 0), 3] >= np.float32(np.pi / 2))


@pytest.mark.parametrize("name", ["ico5", "grid16", "torus128"])
def test_queries_bit_exact(meshes, name):
    m = meshes
    v, f = _cases(m)[name]
    o, r = OracleScene(v, f), RefScene(v, f, "cpu")
    lo, hi = m.mesh_bounds(v)
    n = 3000
    q = m.points_in_box(n, lo, hi, 1.5, seed=21)
    d = m.unit_directions(n, seed=22)
    oi, od = o.closest(q, nthreads=4)
    ri, rd = r.closest(q, nthreads=4)
    assert np.array_equal(oi, ri) and np.array_equal(bits(od), bits(rd))
    for flip in (False, True):
        assert np.array_equal(bits(o.silhouette(q, flip, nthreads=4)), bits(r.silhouette(q, flip, nthreads=4)))
    of, ot, ouv, op = o.ray(q, d, nthreads=4)
    rf, rt, ruv, rp = r.ray(q, d, nthreads=4)
    assert np.array_equal(of, rf) and np.array_equal(bits(ot), bits(rt)) and np.array_equal(bits(ouv), bits(ruv)) and np.array_equal(op, rp)
    tm = np.full(n, 0.6, np.float32)
    of, ot, _, op = o.ray(q, d, tm, nthreads=4)
    rf, rt, _, rp = r.ray(q, d, tm, nthreads=4)
    assert np.array_equal(of, rf) and np.array_equal(bits(ot), bits(rt)) and np.array_equal(op, rp)
    sph = np.concatenate([q, (od * 1.5 + 0.05)[:, None]], axis=1).astype(np.float32)
    u = m.uniforms(n, seed=23)
    oi2, op2 = o.sample(sph, u)
    ri2, rp2 = r.sample(sph, u)
    assert np.array_equal(oi2, ri2) and np.array_equal(bits(op2), bits(rp2))


def _soup_cases():
    from test_gpu_fuzz import CASES
    return CASES


@pytest.mark.parametrize("case", _soup_cases()[:9], ids=lambda c: f"seed{c[0]}_v{c[1]}_t{c[2]}")
def test_soups_bit_exact(meshes, case):
    """The random triangle soups of test_gpu_fuzz.py (non-manifold, duplicated, flipped, self-intersecting, far / tiny / sliver
    coordinates): the restatement equals the reference's own code on them too, so the GPU fuzz is checked against the
    reference's behaviour, not merely against this port.  Silhouette distances are compared where the reference is defined:
    a scene with Q1-tainted nodes reads an uninitialised half-angle there (cone.cuh:454-459)."""
    from test_gpu_fuzz import soup
    m = meshes
    v, f = soup(*case)
    o, r = OracleScene(v, f), RefScene(v, f, "cpu")
    on, oa, oc = o.tree()
    rn, ra, rc = r.tree()
    assert np.array_equal(on, rn) and np.array_equal(bits(oa), bits(ra))
    for a, b in zip(o.adjacency(), r.adjacency()):
        assert np.array_equal(a, b)
    taint = o.q1_taint()
    valid = oc[:, 3] >= 0
    eq = np.where(valid, (bits(oc) == bits(rc)).all(axis=1), bits(oc)[:, 3] == bits(rc)[:, 3])
    assert eq[~taint].all()
    lo, hi = m.mesh_bounds(v)
    n = 2000
    q = m.points_in_box(n, lo, hi, 1.5, seed=1000 + case[0])
    d = m.unit_directions(n, seed=2000 + case[0])
    assert np.array_equal(bits(o.closest(q, nthreads=4)[1]), bits(r.closest(q, nthreads=4)[1]))
    if not taint.any():
        for flip in (False, True):
            assert np.array_equal(bits(o.silhouette(q, flip, nthreads=4)), bits(r.silhouette(q, flip, nthreads=4)))
    of, ot, _, op = o.ray(q, d, nthreads=4)
    rf, rt, _, rp = r.ray(q, d, nthreads=4)
    assert np.array_equal(of, rf) and np.array_equal(bits(ot), bits(rt)) and np.array_equal(op, rp)
