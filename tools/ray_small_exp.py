"""C1-sized ray batches (64K rays, 20K-triangle sphere; also 256K / 1M rays on the 1M-triangle torus): the inline-leaf kernel against
the reference-order (parked) kernel at several flush / refill settings (GPU)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import snch_lbvh_b200 as pkg  # noqa: E402
from snch_lbvh_b200 import meshes as m  # noqa: E402


def run(sc, q, d, reps=30):
    for _ in range(5):
        sc.intersect(q, d)
    sc.set_option("query.time_kernels", 1)
    sc.counter("query.traversal_ms", reset=True)
    for _ in range(reps):
        sc.intersect(q, d)
    torch.cuda.synchronize()
    ms = sc.counter("query.traversal_ms", reset=True) / reps
    sc.set_option("query.time_kernels", 0)
    return ms


for name, (v, f), n in (("C1 sphere 20K tris", m.icosphere(5), 65536), ("torus 1M tris", m.bumpy_torus(708, 708), 262144), ("torus 1M tris", m.bumpy_torus(708, 708), 900000)):
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    lo, hi = m.mesh_bounds(v)
    q = torch.from_numpy(m.points_in_box(n, lo, hi, 1.1, seed=2025)).cuda()
    d = torch.from_numpy(m.unit_directions(n, seed=77)).cuda()
    for knobs in ({"query.ray_kernel": 0}, {"query.ray_kernel": 2, "query.ray_flush": 8, "query.ray_refill": 8}, {"query.ray_kernel": 2, "query.ray_flush": 1, "query.ray_refill": 1},
                  {"query.ray_kernel": 2, "query.ray_flush": 2, "query.ray_refill": 2}, {"query.ray_kernel": 2, "query.ray_flush": 4, "query.ray_refill": 4},
                  {"query.ray_kernel": 2, "query.ray_flush": 4, "query.ray_refill": 1}, {"query.ray_kernel": 2, "query.ray_flush": 1, "query.ray_refill": 8}):
        for k, val in knobs.items():
            sc.set_option(k, val)
        print(name, n, knobs, sc.last_kernel() if False else "", "kernel ms", round(run(sc, q, d), 4), flush=True)
