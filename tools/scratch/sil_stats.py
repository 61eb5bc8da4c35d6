"""Counts cone tests / undecided tests of the compact silhouette kernel on a C3 sample."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import snch_lbvh_b200 as pkg
m = pkg.meshes
n = 1 << 21
v, f = m.bumpy_torus(708, 708)
lo, hi = m.mesh_bounds(v)
q = torch.from_numpy(m.points_in_box(n, lo, hi, 1.1, seed=2025)).cuda()
s = torch.from_numpy(m.star_radius_scale(n, seed=4242)).cuda()
sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
_, dcp = sc.closest_point(q)
rmax = (dcp * s).contiguous()
sc.set_option("query.sil_stats", 1)
for name, kw in (("bounded", dict(r_max=rmax)), ("unbounded", {})):
    sc.counter("query.sil_stats.7", reset=True)
    sc.closest_silhouette(q, **kw)
    torch.cuda.synchronize()
    st = [sc.counter(f"query.sil_stats.{i}") for i in range(8)]
    sc.counter("query.sil_stats.7", reset=True)
    print(name, dict(tests=st[0], undecided=st[1], exact_codes=st[2], warp_steps=st[3], warp_steps_undecided=st[4], visits=st[5]),
          "undecided/test %.4f  warp-steps with undecided %.3f  visits/query %.1f" % (st[1] / max(st[0], 1), st[4] / max(st[3], 1), st[5] / n), flush=True)
