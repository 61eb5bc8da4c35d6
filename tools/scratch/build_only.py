"""Builds the C2 scene a few times (for ncu captures of the build kernels) and prints build_ms."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import snch_lbvh_b200 as pkg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 708
v, f = pkg.meshes.bumpy_torus(n, n)
sc = pkg.Scene3(v, f).compute_silhouettes()
for k in range(6):
    sc.build_bvh()
    print(k, "build_ms", sc.stats()["build_ms"], flush=True)
