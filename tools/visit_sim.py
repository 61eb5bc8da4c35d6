"""CPU simulation on the oracle's tree: how many nodes does a near-first silhouette walk open when leaf results tighten the
bound immediately, and when they arrive `defer` nodes later (the kernels' leaf queue)?  C3 workload, sampled."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import snch_lbvh_b200.meshes as m  # noqa: E402
from oracle import OracleScene, oracle_lib  # noqa: E402

nu = int(os.environ.get("NU", 708))
v, f = m.bumpy_torus(nu, nu)
lo, hi = m.mesh_bounds(v)
n = int(os.environ.get("N", 20000))
q = m.points_in_box(n, lo, hi, 1.1, seed=2025)
orc = OracleScene(v, f)
rmax = (orc.closest(q, nthreads=16)[1] * m.star_radius_scale(n, seed=4242)).astype(np.float32)
L = oracle_lib()
fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
dp = C.POINTER(C.c_double)
L.orc_silhouette_nearfirst_visits.argtypes = [C.c_void_p, fp, C.c_long, C.c_int, C.c_void_p, C.c_int, dp, dp]
mi, ml = C.c_double(), C.c_double()
L.orc_silhouette_must_visit(orc.h, q, n, 0, rmax.ctypes.data_as(C.c_void_p), C.byref(mi), C.byref(ml))
out = {"mesh": f"torus{nu}", "n": n, "must_visit": [mi.value, ml.value], "near_first": {}}
print("must visit", mi.value, ml.value)
for bounded in (True, False):
    for defer in (0, 1, 2, 4, 8, 16, 32):
        L.orc_silhouette_nearfirst_visits(orc.h, q, n, 0, rmax.ctypes.data_as(C.c_void_p) if bounded else None, defer, C.byref(mi), C.byref(ml))
        out["near_first"][f"{'bounded' if bounded else 'unbounded'}_defer{defer}"] = [mi.value, ml.value]
        print("bounded" if bounded else "unbounded", "defer", defer, "internal", round(mi.value, 1), "leaves", round(ml.value, 2), flush=True)
json.dump(out, open(os.path.join(ROOT, "profiles", f"visit_sim_torus{nu}.json"), "w"), indent=1)
