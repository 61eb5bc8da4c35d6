"""Build parity on the GPU, through the C-ABI: Morton codes, sort order, topology, leaf ranges and AABBs bit-exact
against the oracle and the committed golden vectors; normal cones within 1e-5."""
import os

import numpy as np
import pytest

from conftest import small_cases
from oracle import OracleScene
from parity import bits, check_build_vs_oracle, check_cones

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _mesh(meshes, name):
    cases = small_cases(meshes)
    if name in cases:
        return cases[name]
    return {
        "ico5": lambda: meshes.icosphere(5),               # 20 480 tris: config C1, unique Morton codes
        "grid40": lambda: meshes.open_grid(40),            # open surface: boundary edges, flat-ish, collisions
        "torus97x61": lambda: meshes.bumpy_torus(97, 61),  # ragged (non power of two, partial sort tile)
        "torus300": lambda: meshes.bumpy_torus(300, 300),  # 180 000 tris, many sort tiles
        "torus708": lambda: meshes.bumpy_torus(708, 708),  # 1 002 528 tris: config C2 (collision path)
        "flat": lambda: _flat(meshes),                     # zero extent on z: NaN normalisation (quirk Q12)
        "dups": lambda: _dups(meshes),                     # identical triangles: equal Morton codes, index tie-break
    }[name]()


def _flat(meshes):
    v, f = meshes.open_grid(12)
    v = v.copy()
    v[:, 2] = 0.0
    return v, f


def _dups(meshes):
    v, f = meshes.icosphere(1)
    return v, np.concatenate([f, f, f[:7]]).astype(np.int32)


@pytest.mark.parametrize("name", ["tet", "ico2", "grid6", "torus24x16", "ico5", "grid40", "torus97x61", "torus300", "flat", "dups",
                                  "torus708"])
def test_build_matches_oracle(pkg, meshes, name):
    v, f = _mesh(meshes, name)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    orc = OracleScene(v, f)
    check_build_vs_oracle(sc, orc, pkg)
    st = sc.stats()
    assert st["num_nodes"] == 2 * len(f) - 1 and st["build_ms"] > 0
    # scene box == root AABB == what the Morton normalisation used
    aabbs = sc.export(pkg.ExportKind.AABBS)
    assert np.allclose(aabbs[0, :3], st["scene_upper"], rtol=0, atol=0) and np.allclose(aabbs[0, 3:], st["scene_lower"], rtol=0, atol=0)


@pytest.mark.parametrize("name", ["tet", "ico2", "grid6", "torus24x16"])
def test_build_matches_golden(pkg, name):
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    sc = pkg.Scene3(g["verts"], g["tris"]).compute_silhouettes().build_bvh()
    K = pkg.ExportKind
    assert np.array_equal(sc.export(K.MORTON_SORTED), g["morton"])
    assert np.array_equal(sc.export(K.SORTED_INDEX), g["sorted_idx"])
    assert np.array_equal(sc.export(K.NODES), g["nodes"])
    assert np.array_equal(bits(sc.export(K.AABBS)), bits(g["aabbs"]))
    assert np.array_equal(sc.export(K.EDGES), g["edges"]) and np.array_equal(sc.export(K.TRI_OWNED), g["tri_owned"])
    taint = sc.export(K.Q1_TAINT).astype(bool)
    check_cones(sc.export(K.CONES), np.where(taint[:, None], sc.export(K.CONES), g["cones"]), taint, taint)


def test_rebuild_is_deterministic(pkg, meshes):
    v, f = meshes.bumpy_torus(120, 90)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    K = pkg.ExportKind
    first = [sc.export(k).copy() for k in (K.NODES, K.AABBS, K.CONES, K.MORTON_SORTED, K.SORTED_INDEX)]
    sc.build_bvh()
    for a, k in zip(first, (K.NODES, K.AABBS, K.CONES, K.MORTON_SORTED, K.SORTED_INDEX)):
        assert np.array_equal(bits(a) if a.dtype == np.float32 else a, bits(sc.export(k)) if a.dtype == np.float32 else sc.export(k))


@pytest.mark.parametrize("name", ["tet", "one", "ico2", "grid6", "grid40", "torus97x61", "flat", "dups", "torus300", "torus708"])
def test_refit_kernels_produce_identical_arenas(pkg, meshes, name):
    """"build.refit_kernel": the block-cooperative refit (v2, rounds in shared memory) and the per-thread climb (v1) apply the
    same merges to the same children; "sort.onesweep" / "sort.lookback": both radix sorts are stable and the look-back window
    changes no offset.  The whole arena — reference-layout
    arrays and traversal records — is bit-identical for every combination."""
    if name == "one":
        v, f = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.int32)
    else:
        v, f = _mesh(meshes, name)
    sc = pkg.Scene3(v, f).compute_silhouettes()
    arenas = []
    import torch
    try:
        # "sort.onesweep" / "sort.lookback" are process-wide: restored below
        for kern, onesweep, lookback in ((0, 0, 8), (1, 1, 8), (1, 0, 8), (1, 1, 1)):
            sc.set_option("build.refit_kernel", kern).set_option("sort.onesweep", onesweep).set_option("sort.lookback", lookback).build_bvh()
            arenas.append(sc.arena_tensor().clone())
    finally:
        sc.set_option("build.refit_kernel", 1).set_option("sort.onesweep", 1).set_option("sort.lookback", 8)
    for a in arenas[1:]:
        assert a.shape == arenas[0].shape
        assert torch.equal(arenas[0], a), f"{int((arenas[0] != a).sum())} arena bytes differ between refit kernels / radix sorts"


def test_tree_invariants_full_size(pkg, meshes):
    """Size-independent structure checks on the 1M-triangle build: sortedness, permutation, parent/child consistency,
    every internal box encloses its children, ranges partition."""
    v, f = meshes.bumpy_torus(708, 708)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    K = pkg.ExportKind
    n = len(f)
    morton, sidx = sc.export(K.MORTON_SORTED), sc.export(K.SORTED_INDEX)
    key = (morton.astype(np.uint64) << np.uint64(32)) | sidx.astype(np.uint64)
    assert np.all(key[1:] > key[:-1]), "augmented keys must be strictly increasing (stable sort)"
    assert np.array_equal(np.sort(sidx), np.arange(n, dtype=np.uint32))
    nodes, aabbs, ranges = sc.export(K.NODES), sc.export(K.AABBS), sc.export(K.RANGES)
    ni = n - 1
    L, R = nodes[:ni, 1], nodes[:ni, 2]
    assert np.all(nodes[L, 0] == np.arange(ni)) and np.all(nodes[R, 0] == np.arange(ni)) and nodes[0, 0] == 0xFFFFFFFF
    assert np.array_equal(np.sort(np.concatenate([L, R])), np.arange(1, 2 * n - 1, dtype=np.uint32))
    assert np.all(aabbs[:ni, :3] == np.maximum(aabbs[L, :3], aabbs[R, :3])) and np.all(aabbs[:ni, 3:] == np.minimum(aabbs[L, 3:], aabbs[R, 3:]))
    assert np.all(nodes[ni:, 3] == sidx) and np.all(nodes[:ni, 3] == 0xFFFFFFFF)
    assert ranges[0, 0] == 0 and ranges[0, 1] == n - 1 and np.all(ranges[:, 0] < ranges[:, 1])
