"""Which quantity of test_fuzz_large_batch_kernels differs, on which queries, under which knob (GPU; writes gpurun_out/fuzz_debug.json)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import snch_lbvh_b200 as pkg  # noqa: E402
from oracle import OracleScene  # noqa: E402
from snch_lbvh_b200 import meshes  # noqa: E402
from test_gpu_fuzz import CASES, soup  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


out = {}
for case in CASES:
    v, f = soup(*case)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    orc = OracleScene(v, f)
    lo, hi = meshes.mesh_bounds(v)
    q = meshes.points_in_box(6000, lo, hi, 1.5, seed=1000 + case[0])
    _, dist = sc.closest_point(q)
    rmax = (dist * meshes.star_radius_scale(len(q))).astype(np.float32)
    orc_v = [orc.closest(q, nthreads=8)[1], orc.silhouette(q, False, nthreads=8), orc.silhouette(q, True, nthreads=8),
             orc.silhouette(q, False, r_max=rmax, nthreads=8)]
    names = ["closest", "sil", "sil_flip", "sil_rmax"]

    def run():
        return [sc.closest_point(q)[1], sc.closest_silhouette(q), sc.closest_silhouette(q, flip=True), sc.closest_silhouette(q, r_max=rmax)]

    ref = run()
    rec = {}
    for knobs in ({}, {"query.seed": 0}, {"query.seed": 2}, {"query.cone_filter": 0}, {"query.sil_tail": 0}):
        sc.set_option("query.wide_max_n", 0).set_option("query.wide_max_n_sil", 0).set_option("query.sort_min_n", 1)
        for k, val in knobs.items():
            sc.set_option(k, val)
        got = run()
        sc.set_option("query.wide_max_n", 2097152).set_option("query.wide_max_n_sil", 262144).set_option("query.sort_min_n", 16384)
        sc.set_option("query.seed", 1).set_option("query.cone_filter", 1).set_option("query.sil_tail", 8)
        r = {}
        for nm, a, b, o in zip(names, ref, got, orc_v):
            bad = np.nonzero(bits(a) != bits(b))[0]
            r[nm] = {"n_diff": int(len(bad)), "small_vs_oracle": int(np.count_nonzero(bits(a) != bits(o))),
                     "large_vs_oracle": int(np.count_nonzero(bits(b) != bits(o))),
                     "examples": [{"i": int(i), "q": q[i].tolist(), "small": float(a[i]), "large": float(b[i]), "oracle": float(o[i])} for i in bad[:5]]}
        rec[json.dumps(knobs)] = r
    out[f"seed{case[0]}"] = rec
    print(case[0], {k: {n: (x["n_diff"], x["small_vs_oracle"], x["large_vs_oracle"]) for n, x in r.items()} for k, r in rec.items()}, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fuzz_debug.json"), "w"), indent=1)
