"""Host-side scene preparation of the product library (snch_scene_compute_silhouettes: edge ids, silhouette int4,
first-owner rule) against the oracle's restatement of scene.cuh:1135-1229 — no GPU involved."""
import numpy as np
import pytest

from conftest import small_cases
from oracle import OracleScene


@pytest.mark.parametrize("name", ["tet", "ico2", "grid6", "torus24x16", "ico6", "torus708"])
def test_adjacency_matches_oracle(pkg, meshes, name):
    cases = small_cases(meshes)
    if name == "ico6":
        v, f = meshes.icosphere(6)
    elif name == "torus708":
        v, f = meshes.bumpy_torus(708, 708)  # the 1 002 528-triangle mesh of configs C2/C3
    else:
        v, f = cases[name]
    sc = pkg.Scene3(v, f).compute_silhouettes()
    o = OracleScene(v, f)
    e, te, to = o.adjacency()
    K = pkg.ExportKind
    assert np.array_equal(sc.export(K.EDGES), e)
    assert np.array_equal(sc.export(K.TRI_EDGES), te)
    assert np.array_equal(sc.export(K.TRI_OWNED), to)
    assert sc.stats()["num_edges"] == o.num_edges
    if name.startswith(("ico", "torus", "tet")):  # closed manifolds: E = 3N/2, every edge has two faces and one owner
        assert o.num_edges * 2 == 3 * len(f)
        assert np.all(e[:, 0] >= 0) and np.all(e[:, 3] >= 0)
        assert np.count_nonzero(to >= 0) == o.num_edges


def test_non_manifold_and_shuffled_input(pkg, meshes):
    """Q17/Q18: ownership follows input order; a third face on an edge overwrites a slot (last writer wins)."""
    v, f = meshes.icosphere(1)
    rng = np.random.default_rng(5)
    f2 = f[rng.permutation(len(f))]
    extra = np.array([[f2[0, 0], f2[0, 1], (f2[0, 2] + 1) % len(v)]], np.int32)
    f3 = np.concatenate([f2, extra])
    sc = pkg.Scene3(v, f3).compute_silhouettes()
    o = OracleScene(v, f3)
    for a, b in zip((sc.export(pkg.ExportKind.EDGES), sc.export(pkg.ExportKind.TRI_EDGES), sc.export(pkg.ExportKind.TRI_OWNED)),
                    o.adjacency()):
        assert np.array_equal(a, b)


def test_owned_edge_ids_are_consecutive(meshes):
    """The traversal records index a leaf's silhouette edges by EDGE ID (build.cu refit_leaf): that relies on the reference's
    numbering — edges are numbered in first-seen order and owned by the first triangle that touches them (scene.cuh:1135-1229),
    so triangle i owns exactly the next cnt_i ids, in slot order.  Checked on the oracle's adjacency (bit-identical to the
    product's, tests above) for closed, open, non-manifold, duplicated and degenerate input."""
    import numpy as np
    from oracle import OracleScene
    v = np.random.default_rng(0).random((30, 3)).astype(np.float32)
    nasty = np.array([[0, 1, 2], [0, 1, 3], [0, 1, 4], [1, 0, 5], [2, 2, 3], [0, 1, 2], [6, 7, 8], [8, 7, 6], [9, 9, 9]], np.int32)
    cases = [meshes.tetrahedron(), meshes.icosphere(2), meshes.open_grid(6), meshes.bumpy_torus(24, 16), meshes.bumpy_torus(120, 90), (v, nasty)]
    for vv, ff in cases:
        _, _, owned = OracleScene(vv, ff).adjacency()
        nxt = 0
        for row in owned:
            ids = [int(x) for x in row if x != -1]
            assert list(row[:len(ids)]) == ids, "owned slots are not compacted to the front"
            assert ids == list(range(nxt, nxt + len(ids))), "owned edge ids are not the next consecutive ids"
            nxt += len(ids)
