"""e2e (host pinned buffers through the C-ABI) for several "query.host_chunk" settings on the C3 workload."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import snch_lbvh_b200 as pkg
m = pkg.meshes
n = 1 << 24
v, f = m.bumpy_torus(708, 708)
lo, hi = m.mesh_bounds(v)
q_h = m.points_in_box(n, lo, hi, 1.1, seed=2025)
s_h = m.star_radius_scale(n, seed=4242)
sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
q_d = torch.from_numpy(q_h).cuda()
_, dcp = sc.closest_point(q_d)
r_p = (dcp.cpu() * torch.from_numpy(s_h)).contiguous().pin_memory()
q_p = torch.from_numpy(q_h).pin_memory()
o_p = torch.empty(n, dtype=torch.float32).pin_memory()
L = pkg.lib()
stream = torch.cuda.Stream()
ref = None
for chunk in (0, 1 << 23, 6 << 20, 1 << 22, 3 << 20, 1 << 21, 1 << 20):
    sc.set_option("query.host_chunk", chunk)
    for _ in range(2):
        assert L.snch_closest_silhouette_batch(sc._h, q_p.data_ptr(), None, r_p.data_ptr(), n, o_p.data_ptr(), stream.cuda_stream) == 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        assert L.snch_closest_silhouette_batch(sc._h, q_p.data_ptr(), None, r_p.data_ptr(), n, o_p.data_ptr(), stream.cuda_stream) == 0
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    if ref is None:
        ref = o_p.clone()
    print(f"host_chunk {chunk:>9d}: {ms:7.2f} ms  {n / ms / 1e3:7.1f} Mq/s  identical={torch.equal(ref.view(torch.int32), o_p.view(torch.int32))}", flush=True)
