"""GPU silhouette adjacency (adjacency.cu) against the oracle's restatement of scene.cuh:1135-1229 and against the
library's own host passes: edge ids in first-seen order, silhouette int4 (last writer wins, Q18), first-owner rule (Q17)
— all bit-exact, before the build (device arrays fetched) and after it (unpacked from the arena)."""
import numpy as np
import pytest

from conftest import small_cases
from oracle import OracleScene

pytestmark = pytest.mark.gpu


def _products(sc, pkg):
    K = pkg.ExportKind
    return sc.export(K.EDGES), sc.export(K.TRI_EDGES), sc.export(K.TRI_OWNED)


def _cases(meshes):
    c = dict(small_cases(meshes))
    c["ico5"] = meshes.icosphere(5)
    c["grid40"] = meshes.open_grid(40)
    c["torus97x61"] = meshes.bumpy_torus(97, 61)
    v, f = meshes.icosphere(2)
    rng = np.random.default_rng(11)
    c["shuffled"] = (v, f[rng.permutation(len(f))])
    # non-manifold fan: four extra faces on one edge, both orientations (Q18)
    a, b = f[0, 0], f[0, 1]
    fan = np.array([[a, b, 5], [b, a, 7], [a, b, 9], [b, a, 11]], np.int32)
    c["fan"] = (v, np.concatenate([f, fan]))
    # degenerate faces: repeated vertices give edges (i,i) and the same edge twice inside one triangle
    deg = np.array([[3, 3, 8], [4, 9, 4], [6, 6, 6]], np.int32)
    c["degenerate"] = (v, np.concatenate([f[:40], deg, f[40:]]))
    c["dups"] = (v, np.concatenate([f, f, f[:9]]).astype(np.int32))
    c["single"] = (np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.int32))
    return c


@pytest.mark.parametrize("name", ["tet", "ico2", "grid6", "torus24x16", "ico5", "grid40", "torus97x61", "shuffled", "fan", "degenerate",
                                  "dups", "single"])
def test_device_adjacency_matches_oracle_and_host(pkg, meshes, name):
    v, f = _cases(meshes)[name]
    want = OracleScene(v, f).adjacency()
    dev = pkg.Scene3(v, f).set_option("adjacency.device", 1).compute_silhouettes()
    host = pkg.Scene3(v, f).set_option("adjacency.device", 0).compute_silhouettes()
    assert dev.stats()["num_edges"] == host.stats()["num_edges"] == len(want[0])
    for got in (_products(dev, pkg), _products(host, pkg)):
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
    # after the build the arena is the only copy: same answers, and both builds are the same tree
    dev.build_bvh()
    host.build_bvh()
    for a, b in zip(_products(dev, pkg), want):
        assert np.array_equal(a, b)
    K = pkg.ExportKind
    for kind in (K.NODES, K.AABBS, K.CONES, K.SORTED_INDEX):
        assert np.array_equal(dev.export(kind).view(np.uint32), host.export(kind).view(np.uint32))


def test_device_adjacency_at_full_size(pkg, meshes):
    """Config C2/C3 mesh (1 002 528 triangles, 1 503 792 edges): bit-exact against the host passes, and much faster."""
    v, f = meshes.bumpy_torus(708, 708)
    dev = pkg.Scene3(v, f).set_option("adjacency.device", 1).compute_silhouettes()
    dev_ms = []
    for _ in range(3):  # later calls: context, lazily loaded kernels and the allocator are warm
        dev.compute_silhouettes()
        dev_ms.append(dev.stats()["adjacency_ms"])
    host = pkg.Scene3(v, f).set_option("adjacency.device", 0).compute_silhouettes()
    for a, b in zip(_products(dev, pkg), _products(host, pkg)):
        assert np.array_equal(a, b)
    assert dev.stats()["num_edges"] == 3 * len(f) // 2
    print("adjacency ms: device", dev_ms, "host", host.stats()["adjacency_ms"])
    assert min(dev_ms) < host.stats()["adjacency_ms"]


def test_rebuild_keeps_topology(pkg, meshes):
    """A second build of the same scene re-uses the topology already in the arena."""
    v, f = meshes.bumpy_torus(40, 30)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    first = [sc.export(k).copy() for k in (pkg.ExportKind.NODES, pkg.ExportKind.AABBS, pkg.ExportKind.CONES, pkg.ExportKind.EDGES)]
    sc.build_bvh()
    second = [sc.export(k) for k in (pkg.ExportKind.NODES, pkg.ExportKind.AABBS, pkg.ExportKind.CONES, pkg.ExportKind.EDGES)]
    for a, b in zip(first, second):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
