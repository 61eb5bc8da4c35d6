"""Against the reference's own CUDA path compiled for sm_100a (oracle/_ref/libsnch_ref_cuda.so, "Oracle A").
Skipped where that prebuilt library did not travel."""
import numpy as np
import pytest

from oracle import RefScene, ref_available
from parity import angle_close, bits, check_silhouette, rel_close

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_available("cuda"), reason="oracle/_ref/libsnch_ref_cuda.so not built")]


@pytest.mark.parametrize("name", ["ico5", "torus300", "grid40"])
def test_against_reference_cuda(pkg, meshes, name):
    m = meshes
    v, f = {"ico5": lambda: m.icosphere(5), "torus300": lambda: m.bumpy_torus(300, 300), "grid40": lambda: m.open_grid(40)}[name]()
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    ref = RefScene(v, f, "cuda")
    K = pkg.ExportKind
    rn, ra, rc = ref.tree()
    assert np.array_equal(sc.export(K.NODES), rn), "topology differs from the reference CUDA build"
    assert np.array_equal(bits(sc.export(K.AABBS)), bits(ra))
    rm, rsi = ref.morton()
    assert np.array_equal(sc.export(K.MORTON_SORTED), rm) and np.array_equal(sc.export(K.SORTED_INDEX), rsi)
    taint = sc.export(K.Q1_TAINT).astype(bool)
    cones = sc.export(K.CONES)
    ok = ~taint & (cones[:, 3] >= 0)
    assert np.array_equal(cones[:, 3] >= 0, rc[:, 3] >= 0) or np.array_equal((cones[:, 3] >= 0)[~taint], (rc[:, 3] >= 0)[~taint])
    # Cones are acos()-built and nvcc contracts the reference's dot products into FMAs (we compile with -fmad=false to stay
    # bit-identical to the C oracle), so half-angles are compared in cos-space and statistically; the bit-level bar is on
    # topology/AABBs above and the 1e-5 bar on the query results below.
    close = angle_close(cones[ok, 3], rc[ok, 3], 1e-5, 2e-6, cos_tol=5e-6)
    absd = np.abs(cones[ok, 3].astype(np.float64) - rc[ok, 3].astype(np.float64))
    assert close.mean() >= 0.97, f"only {close.mean():.4f} of the half-angles agree with the reference CUDA build"
    assert np.percentile(absd, 99.9) <= 5e-2, f"half-angle p99.9 abs diff {np.percentile(absd, 99.9)}"
    assert np.mean(rel_close(cones[ok, 4], rc[ok, 4], 1e-5, 2e-6)) >= 0.999
    lo, hi = m.mesh_bounds(v)
    n = 200000
    q = m.points_in_box(n, lo, hi, 1.5, seed=61)
    d = m.unit_directions(n, seed=62)
    _, dist = sc.closest_point(q)
    _, rdist = ref.closest(q)
    assert rel_close(dist, rdist).all()
    # The reference's cone pruning is numerically chaotic: its own CPU and CUDA builds (same headers, without / with FMA
    # contraction) disagree on a fraction of queries because a borderline cone test flips and one side misses its closest
    # silhouette.  That fraction is MEASURED here — the reference's CPU build (oracle/_ref/libsnch_ref_cpu.so where it
    # travelled, else the C oracle pinned bit-identical to it by tests/test_oracle_pinning.py) against its CUDA build on the
    # same queries — and this library must not disagree with the CUDA build on more queries than that (+1e-4 of slack for the
    # sample size); beyond those flips every distance agrees to 1e-5.
    sil_ref_cuda = ref.silhouette(q)
    if ref_available("cpu") and len(f) <= 200000:
        sil_ref_cpu = RefScene(v, f, "cpu").silhouette(q)
    else:
        from oracle import OracleScene
        sil_ref_cpu = OracleScene(v, f).silhouette(q, nthreads=8)
    ref_vs_ref = 1.0 - rel_close(sil_ref_cpu, sil_ref_cuda).mean()
    ours = check_silhouette(sc.closest_silhouette(q), sil_ref_cuda, ref_vs_ref + 1e-4)
    print(f"{name}: silhouette mismatch vs reference CUDA {ours:.2e}; reference CPU vs reference CUDA {ref_vs_ref:.2e}")
    found, hits = sc.intersect(q, d)
    rf, rt, _, _ = ref.ray(q, d)
    assert np.mean(found.astype(bool) == rf.astype(bool)) > 0.9998
    both = found.astype(bool) & rf.astype(bool)
    assert np.mean(rel_close(hits["t"][both], rt[both])) > 0.9998
