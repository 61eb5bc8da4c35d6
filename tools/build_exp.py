"""Build time under build knobs (GPU).  usage: build_exp.py [nu]  — prints median build_ms per setting; under ncu prints the launch list"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import snch_lbvh_b200 as pkg  # noqa: E402
from snch_lbvh_b200 import meshes  # noqa: E402

nu = int(sys.argv[1]) if len(sys.argv) > 1 else 708
reps = int(os.environ.get("REPS", 15))
v, f = meshes.bumpy_torus(nu, nu)
sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
for knobs in ({}, {"build.refit_kernel": 0}):
    for k, val in {"build.refit_kernel": 1, **knobs}.items():
        sc.set_option(k, val)
    ms = []
    for _ in range(reps):
        sc.build_bvh()
        ms.append(sc.stats()["build_ms"])
    print(nu, knobs, "median", round(float(np.median(ms)), 4), "min", round(min(ms), 4), flush=True)
