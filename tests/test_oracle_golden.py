"""The C oracle against the committed golden vectors (generated from the reference itself by tests/golden/make_golden.py).
Both run on the CPU with the same libm, so everything is compared bit-for-bit."""
import os

import numpy as np
import pytest

from oracle import OracleScene
from parity import bits

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["tet", "ico2", "grid6", "torus24x16"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_golden(name):
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    o = OracleScene(g["verts"], g["tris"])
    nodes, aabbs, cones = o.tree()
    assert np.array_equal(nodes, g["nodes"])
    assert np.array_equal(bits(aabbs), bits(g["aabbs"]))
    morton, sidx = o.morton()
    assert np.array_equal(morton, g["morton"]) and np.array_equal(sidx, g["sorted_idx"])
    e, te, to = o.adjacency()
    assert np.array_equal(e, g["edges"]) and np.array_equal(te, g["tri_edges"]) and np.array_equal(to, g["tri_owned"])
    taint = o.q1_taint()
    valid = cones[:, 3] >= 0
    full = (bits(cones) == bits(g["cones"])).all(axis=1)
    half = bits(cones)[:, 3] == bits(g["cones"])[:, 3]
    assert np.where(valid, full, half)[~taint].all(), "cones differ outside the Q1-tainted set"
    q, d = g["q"], g["d"]
    ci, cd = o.closest(q)
    assert np.array_equal(ci, g["closest_idx"]) and np.array_equal(bits(cd), bits(g["closest_dist"]))
    assert np.array_equal(bits(o.silhouette(q, False)), bits(g["sil_noflip"]))
    assert np.array_equal(bits(o.silhouette(q, True)), bits(g["sil_flip"]))
    f, t, uv, p = o.ray(q, d)
    assert np.array_equal(f, g["ray_found"]) and np.array_equal(bits(t), bits(g["ray_t"]))
    assert np.array_equal(bits(uv), bits(g["ray_uv"])) and np.array_equal(p, g["ray_prim"])
    f, t, _, p = o.ray(q, d, g["tmax"])
    assert np.array_equal(f, g["ray_found_tmax"]) and np.array_equal(bits(t), bits(g["ray_t_tmax"])) and np.array_equal(p, g["ray_prim_tmax"])
    si, sp = o.sample(g["sph"], g["u"])
    assert np.array_equal(si, g["sample_idx"]) and np.array_equal(bits(sp), bits(g["sample_pdf"]))
