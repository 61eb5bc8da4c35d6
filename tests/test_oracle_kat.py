"""Known-answer tests for the CPU oracle (SURVEY 8(c) KATs, extracted from the reference run on the CPU)."""
import numpy as np

from oracle import OracleScene, oracle_lib


def test_morton_kat():
    L = oracle_lib()
    assert L.orc_morton3(0.5, 0.25, 0.75) == 721420288
    assert L.orc_morton3(1.0, 1.0, 1.0) == 0x3FFFFFFF
    assert L.orc_morton3(0.0, 0.0, 0.0) == 0
    assert L.orc_morton3(-1.0, 2.0, float("nan")) == L.orc_expand_bits(1023) * 2  # clamp; fmaxf(NaN,0)=0 (Q12)
    assert L.orc_expand_bits(1023) == 0x09249249
    assert L.orc_expand_bits(1) == 1 and L.orc_expand_bits(2) == 8


def test_tetrahedron_kat(meshes):
    v, f = meshes.tetrahedron()
    o = OracleScene(v, f)
    idx, dist = o.closest(np.array([[2.0, 2.0, 2.0]], np.float32))
    assert abs(dist[0] - 1.7320508) < 1e-6  # vertex (1,1,1) is the closest point
    found, t, uv, prim = o.ray(np.array([[2.0, 2.0, 2.0]], np.float32), np.array([[-1, -1, -1]], np.float32) / np.sqrt(3.0))
    assert found[0] == 1 and abs(t[0] - 1.7320508) < 1e-5
    # a ray from far along -(1,1,1) hits the far face (1,3,2) first when shot from the other side
    found, t, _, prim = o.ray(np.array([[-2.0, -2.0, -2.0]], np.float32), np.array([[1, 1, 1]], np.float32) / np.sqrt(3.0))
    assert found[0] == 1 and prim[0] == 3 and abs(t[0] - (2 * np.sqrt(3.0) - 1 / np.sqrt(3.0))) < 1e-5
    assert o.num_edges == 6 and o.num_nodes == 7


def test_empty_and_single(meshes):
    o = OracleScene(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32))
    assert o.num_nodes == 0
    idx, dist = o.closest(np.zeros((2, 3), np.float32))
    assert np.all(idx == 0xFFFFFFFF) and np.all(np.isinf(dist))
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    o = OracleScene(v, np.array([[0, 1, 2]], np.int32))
    idx, dist = o.closest(np.array([[0.25, 0.25, 1.0]], np.float32))
    assert idx[0] == 0 and abs(dist[0] - 1.0) < 1e-6
    assert np.isfinite(o.silhouette(np.array([[0.25, 0.25, 1.0]], np.float32))[0])  # boundary edges are silhouettes
