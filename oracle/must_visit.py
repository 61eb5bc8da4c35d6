"""Freezes the must-visit statistics V*, L*  that bench.py's roofline uses as ALGORITHMIC work per query.

Counting pass of the CPU oracle on a sample of each benchmark workload:
    V* = internal nodes any exact traversal of this tree must open,  L* = leaves it must test.
Run in the build container:   python oracle/must_visit.py      -> profiles/must_visit.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import snch_lbvh_b200 as pkg  # noqa: E402  (meshes only; nothing is computed on a GPU here)
from oracle import OracleScene  # noqa: E402

NS = 20000


def main():
    m = pkg.meshes
    out = {"sample_queries": NS, "configs": {}}
    for cfg, nu in (("torus708", 708), ("ico5", 0), ("torus1416", 1416), ("torus2240", 2240)):  # C3, C1, C4, C5 meshes
        v, f = m.bumpy_torus(nu, nu) if nu else m.icosphere(5)
        t0 = time.time()
        o = OracleScene(v, f)
        lo, hi = m.mesh_bounds(v)
        scale = 1.1 if nu else 1.5
        q = m.points_in_box(NS, lo, hi, 1.0 if nu == 2240 else scale, seed=2025)  # (bench.py's C5 walkers start inside the bounding box)
        d = m.unit_directions(NS, seed=77)
        _, dcp = o.closest(q, nthreads=8)
        rmax = (dcp * m.star_radius_scale(NS)).astype(np.float32)
        res = {"n_tris": int(len(f))}
        res["closest"] = dict(zip(("V", "L"), o.must_visit("closest", q)))
        res["silhouette_unbounded"] = dict(zip(("V", "L"), o.must_visit("silhouette", q)))
        res["silhouette_star_radius"] = dict(zip(("V", "L"), o.must_visit("silhouette", q, r_max=rmax)))
        res["ray"] = dict(zip(("V", "L"), o.must_visit("ray", q, dirs=d)))
        res["finite_fraction_star_radius"] = float(np.isfinite(o.silhouette(q, r_max=rmax, nthreads=8)).mean())
        out["configs"][cfg] = res
        print(cfg, res, f"{time.time() - t0:.1f}s", flush=True)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "must_visit.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
