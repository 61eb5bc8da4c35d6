"""e2e of the C3 batch (host buffers through snch_closest_silhouette_batch) under different chunk schedules (GPU; writes gpurun_out/e2e_exp.json)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import snch_lbvh_b200 as pkg  # noqa: E402
from snch_lbvh_b200 import meshes as m  # noqa: E402

n = int(os.environ.get("N", 16777216))
v, f = m.bumpy_torus(708, 708)
lo, hi = m.mesh_bounds(v)
q = m.points_in_box(n, lo, hi, 1.1, seed=2025)
sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
L = pkg.lib()
qd = torch.from_numpy(q).cuda()
_, d = sc.closest_point(qd)
rmax = (d * torch.from_numpy(m.star_radius_scale(n, seed=4242)).cuda()).contiguous()
out_d = sc.closest_silhouette(qd, r_max=rmax)
torch.cuda.synchronize()
qh = torch.from_numpy(q).pin_memory()
rh = rmax.cpu().pin_memory()
oh = torch.empty(n, dtype=torch.float32).pin_memory()
st = torch.cuda.Stream()


def run(steps=6):
    def step():
        rc = L.snch_closest_silhouette_batch(sc._h, qh.data_ptr(), None, rh.data_ptr(), n, oh.data_ptr(), None, None, st.cuda_stream)
        assert rc == 0
    step()
    step()
    t = []
    for _ in range(steps):
        w = time.perf_counter()
        step()
        t.append((time.perf_counter() - w) * 1e3)
    return float(np.median(t)), float(min(t))


e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    sc.closest_silhouette(qd, r_max=rmax)
e0.record()
for _ in range(5):
    sc.closest_silhouette(qd, r_max=rmax)
e1.record()
torch.cuda.synchronize()
res = {"device_ms": e0.elapsed_time(e1) / 5, "n": n, "variants": []}
print("device", res["device_ms"], flush=True)
variants = [(1 << 23, 0, 0), (1 << 23, 1 << 21, 0), (1 << 23, 1 << 21, 1), (1 << 23, 1 << 20, 1), (1 << 23, 1 << 19, 1), (1 << 23, 3 << 19, 1)]
for chunk, first, split in variants:
    sc.set_option("query.host_chunk", chunk).set_option("query.host_first", first).set_option("query.host_split_min", split)
    med, mn = run()
    assert torch.equal(oh.view(torch.int32), out_d.cpu().view(torch.int32))
    res["variants"].append({"host_chunk": chunk, "host_first": first, "host_split_min": split, "median_ms": med, "min_ms": mn})
    print(res["variants"][-1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"e2e_exp_{n}.json"), "w"), indent=1)
