"""The drop-in C++ API (include/snch_lbvh/*.cuh) used the way the reference is used: per-thread query_device() /
sample_object_in_sphere() calls from a user kernel with the scene's functors.  tests/cpp/dropin_test.cu does the work and
prints CHECK lines; this test builds the inputs, runs it on the GPU and asserts the bars of DESIGN.md "Parity rules"."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "dropin_test")


def _write_obj(path, v, f):
    with open(path, "w") as fh:
        for p in v:
            fh.write(f"v {p[0]:.9g} {p[1]:.9g} {p[2]:.9g}\n")
        fh.write("# faces\n")
        for t in f:
            fh.write(f"f {t[0] + 1} {t[1] + 1} {t[2] + 1}\n")


def _run(obj):
    out = subprocess.run([EXE, obj], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DROPIN_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    return {l.split()[1]: float(l.split()[2]) for l in out.stdout.splitlines() if l.startswith("CHECK ")}


@pytest.mark.gpu
@pytest.mark.parametrize("mesh", ["torus", "open_grid"])
def test_cpp_dropin_api(meshes, tmp_path, mesh):
    if not os.path.exists(EXE):
        pytest.fail("tests/cpp/dropin_test is not built: run __graft_entry__.build() (make -C tests/cpp)")
    v, f = meshes.bumpy_torus(80, 60) if mesh == "torus" else meshes.open_grid(30)
    obj = str(tmp_path / "mesh.obj")
    _write_obj(obj, v, f)
    c = _run(obj)
    assert c["not_built_throws"] == 1 and c["loader_throws"] == 1          # the reference's exception messages
    assert c["tris"] == len(f) and c["nodes"] == 2 * len(f) - 1
    # per-thread header traversals vs the batched kernels of the same scene, and vs brute force with the same functors
    assert c["closest_vs_batched_worst_rel"] <= 1e-5 and c["closest_vs_brute_worst_rel"] <= 1e-5
    assert c["silhouette_vs_batched_mismatch_frac"] <= 1e-3 and c["silhouette_edge_point_bad"] == 0
    assert c["ray_found_diff"] <= 2e-4 * 20000 and c["anyhit_diff"] <= 2e-4 * 20000
    assert c["ray_t_vs_batched_mismatch_frac"] <= 2e-4 and c["ray_t_vs_brute_mismatch_frac"] <= 2e-4
    assert c["sample_idx_diff"] <= 1e-3 * 20000 and c["sample_pdf_mismatch_frac"] <= 2e-3
    # generic lbvh::bvh<...> (snch_lbvh_build) == the fused scene build: topology and boxes bit-equal, cones to rounding
    assert c["generic_node_diff"] == 0 and c["generic_aabb_diff"] == 0
    assert c["generic_cone_bad"] <= 1e-3 * c["nodes"]
    # 2-D scene through the same headers
    # snch_scene_replicate_local from C++: the replicas answer bit-identically
    assert c["replicas"] >= 1 and c["replica_silhouette_diff"] == 0
    assert c["seg2d"] == 449 and c["nodes2d"] == 2 * 449 - 1
    assert c["closest2d_vs_brute_worst_rel"] <= 1e-5 and c["ray2d_vs_brute_mismatch_frac"] <= 1e-3
    assert c["silhouette2d_vs_brute_mismatch_frac"] <= 2e-2  # cone pruning is only approximately conservative in the reference too
    assert c["sample2d_hits"] > 0
    # query_device(line_intersect): the BVH walk finds exactly the segments the functor accepts (a crossing lies inside the leaf box)
    assert c["line2d_crossings"] > 5000 and c["line2d_count_diff"] <= 5
    # batched 2-D entry points (snch_*_batch2) == the per-thread header traversals of the same scene
    assert c["closest2d_vs_batched_worst_rel"] <= 1e-5
    assert c["silhouette2d_vs_batched_mismatch_frac"] <= 1e-3 and c["ray2d_vs_batched_mismatch_frac"] <= 1e-3
