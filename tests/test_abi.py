"""The C-ABI library loads on a CPU-only box, exports exactly what include/snch_b200.h declares, keeps its host-side
logic (creation, adjacency, argument checks) working without a GPU and fails loudly — never falls back — for compute."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "snch_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(snch_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(pkg):
    declared = _declared_symbols()
    assert sorted(pkg.ABI_SYMBOLS) == declared
    out = subprocess.check_output(["nm", "-D", "--defined-only", pkg.lib_path()], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in snch_b200.h but not exported: {missing}"
    L = pkg.lib()
    assert L.snch_abi_version() == 2
    for s in declared:
        assert hasattr(L, s)


def test_library_is_sm100a_cuda(pkg):
    out = subprocess.run(["cuobjdump", "-lelf", pkg.lib_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout


def test_argument_errors(pkg):
    v = np.zeros((3, 3), np.float32)
    with pytest.raises(pkg.SnchError) as e:
        pkg.Scene3(v, np.array([[0, 1, 3]], np.int32))
    assert e.value.status == -1 and "out of range" in str(e.value)
    sc = pkg.Scene3(v, np.array([[0, 1, 2]], np.int32))
    with pytest.raises(pkg.SnchError) as e:
        sc.get_bvh_device_ptr()
    assert e.value.status == -2 and str(e.value) == "BVH is not built yet."  # scene.cuh:1250
    with pytest.raises(pkg.SnchError) as e:
        sc.closest_point(np.zeros((1, 3), np.float32))
    assert e.value.status == -2


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    v, f = pkg.meshes.tetrahedron()
    sc = pkg.Scene3(v, f).compute_silhouettes()
    with pytest.raises(pkg.SnchError) as e:
        sc.build_bvh()
    assert e.value.status == -3 and "no CPU fallback" in str(e.value)


def test_product_never_touches_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pk = os.path.join(ROOT, "snch-lbvh_b200")
    bad = re.compile(r"(^\s*(import|from)\s+oracle\b)|liboracle|snch_oracle|oracle/|oracle\.loader|_ref/", re.M)
    for dirpath, _, files in os.walk(pk):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or fn == "Makefile":
                txt = open(os.path.join(dirpath, fn), errors="replace").read()
                assert not bad.search(txt), f"{fn} references the oracle"
            if fn.endswith((".cu", ".cuh", ".h", ".cpp")):  # the one library resolved at run time is NCCL itself (replicate.cu)
                txt = open(os.path.join(dirpath, fn), errors="replace").read()
                for m in re.finditer(r"dlopen\(([^,]*),", txt):
                    assert m.group(1).strip() in ("env", '"libnccl.so.2"'), f"{fn} dlopens {m.group(1)}"
    out = subprocess.check_output(["ldd", os.path.join(pk, "libsnch_b200.so")], text=True)
    assert "oracle" not in out and "snch_ref" not in out


def test_cpp_dropin_headers_compile(tmp_path):
    """include/snch_lbvh/*.cuh (the reference's header names) compile for sm_100a and instantiate both scene types."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    src = tmp_path / "tu.cu"
    src.write_text('#include <snch_lbvh/lbvh.cuh>\n#include <snch_lbvh/scene.cuh>\n#include <snch_lbvh/scene_loader.cuh>\n'
                   'template class lbvh::scene_loader<2>;\ntemplate class lbvh::scene_loader<3>;\n'
                   '__global__ void k(lbvh::bvh_device<float, 3, lbvh::scene<3>::triangle> b, float3 p, float *o)\n'
                   '{ *o = lbvh::query_device(b, lbvh::nearest(p), lbvh::scene<3>::distance_calculator()).second; }\n'
                   'int main() { lbvh::scene<2> a; lbvh::scene<3> b; return a.lines.size() + b.triangles.size(); }\n')
    out = subprocess.run([nvcc, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                          "-o", str(tmp_path / "tu.o")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]


def test_replication_entry_points_reject_bad_arguments(pkg):
    """The replication part of the C-ABI validates its arguments before it touches CUDA or NCCL (so this runs on a CPU box)."""
    L = pkg.lib()
    out = C.c_void_p()
    buf = C.create_string_buffer(128)
    assert L.snch_comm_unique_id(buf, 16) == -1 and b"SNCH_COMM_ID_BYTES" in L.snch_last_error()
    assert L.snch_comm_create(buf, 128, 3, 2, 0, C.byref(out)) == -1          # rank outside the world
    assert L.snch_comm_create(buf, 64, 0, 2, 0, C.byref(out)) == -1           # id buffer too short
    assert L.snch_comm_adopt(None, 0, 1, 0, C.byref(out)) == -1
    assert L.snch_comm_destroy(None) == 0
    assert L.snch_scene_broadcast(None, 0, None, None, C.byref(out)) == -1
    v, f = pkg.meshes.tetrahedron()
    sc = pkg.Scene3(v, f).compute_silhouettes()                                 # not built
    devs = (C.c_int * 1)(0)
    reps = (C.c_void_p * 1)()
    assert L.snch_scene_replicate_local(sc._h, devs, 1, reps) == -2 and L.snch_last_error() == b"BVH is not built yet."
    assert L.snch_scene_last_kernel(None) == b"" and L.snch_scene_last_kernel(sc._h) == b""
    with pytest.raises(pkg.SnchError):
        sc.set_option("query.sil_seed", 1)                                      # a knob removed in round 2 is an error, not a no-op
