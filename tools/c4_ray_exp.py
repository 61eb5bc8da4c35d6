"""C4 (16.7M rays, 4M-triangle torus) under the ray scheduling knobs (GPU)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import snch_lbvh_b200 as pkg  # noqa: E402
from snch_lbvh_b200 import meshes as m  # noqa: E402

nu = int(os.environ.get("NU", 1416))
n = int(os.environ.get("N", 16777216))
v, f = m.bumpy_torus(nu, nu)
sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
lo, hi = m.mesh_bounds(v)
q = torch.from_numpy(m.points_in_box(n, lo, hi, 1.1, seed=2025)).cuda()
d = torch.from_numpy(m.unit_directions(n, seed=77)).cuda()
base = None
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for knobs in ({}, {"query.sort_rays": 1}, {"query.ray_flush": 12}, {"query.sort_rays": 1, "query.ray_flush": 12}, {"query.sort_rays": 1, "query.ray_flush": 10}, {"query.sort_rays": 1, "query.ray_flush": 12, "query.ray_refill": 10},
              {"query.sort_rays": 1, "query.ray_flush": 12, "query.ray_refill": 6}, {"query.sort_rays": 1, "query.ray_flush": 14}, {"query.sort_rays": 1, "query.sort_bits": 18}, {"query.sort_rays": 1, "query.sort_bits": 30}):
    for k, val in {"query.sort_rays": 0, "query.ray_flush": 8, "query.ray_refill": 8, "query.blocks_per_sm": 0, "query.ray_kernel": 1, **knobs}.items():
        sc.set_option(k, val)
    for _ in range(3):
        found, hits = sc.intersect(q, d)
    e0.record()
    for _ in range(5):
        found, hits = sc.intersect(q, d)
    e1.record()
    torch.cuda.synchronize()
    t = hits[:, 0].clone()
    if base is None:
        base = t
    print(nu, n, knobs, "ms", round(e0.elapsed_time(e1) / 5, 3), "M rays/s", round(n / (e0.elapsed_time(e1) / 5) / 1e3, 1), "same t", bool(torch.equal(t.view(torch.int32), base.view(torch.int32))), flush=True)
