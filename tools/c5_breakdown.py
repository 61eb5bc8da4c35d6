"""Stage breakdown of config C5 (one wavefront walk-on-stars step, 10M-triangle mesh) on one GPU: the four stages called
separately (closest, silhouette with r_max = d, ray with t_max = star radius, sample) next to the fused call, for several
walker counts (the per-GPU shard at N = 1, 2, 4, 8).  Diagnostic; CUDA events, median of 3.
Usage: python tools/c5_breakdown.py [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import snch_lbvh_b200 as pkg

m = pkg.meshes


def timed(fn, reps=3):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), r


def main():
    v, f = m.bumpy_torus(2240, 2240)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    st = sc.stats()
    lo, hi = np.array(st["scene_lower"], np.float32), np.array(st["scene_upper"], np.float32)
    out = {"triangles": int(st["num_objects"])}
    for n in (1 << 26, 1 << 25, 1 << 24, 1 << 23):
        q = torch.from_numpy(m.points_in_box(n, lo, hi, 1.0, seed=31)).cuda()
        d = torch.from_numpy(m.unit_directions(n, seed=131)).cuda()
        rnd = torch.from_numpy(m.uniforms(n, 3, seed=99)).cuda()
        sc.wost_step(q[:1024], d[:1024], rnd[:1024])
        t_fused, r = timed(lambda: sc.wost_step(q, d, rnd))
        t_cp, (_, dist) = timed(lambda: sc.closest_point(q))
        t_sil, sd = timed(lambda: sc.closest_silhouette(q, r_max=dist))
        star = torch.minimum(dist, sd)
        t_ray, _ = timed(lambda: sc.intersect(q, d, t_max=star))
        sph = torch.cat([q, star[:, None]], dim=1).contiguous()
        t_smp, _ = timed(lambda: sc.sample_in_sphere(sph, rnd))
        out[str(n)] = {"fused_ms": t_fused, "closest_ms": t_cp, "silhouette_ms": t_sil, "ray_ms": t_ray, "sample_ms": t_smp,
                       "ns_per_walker_fused": t_fused * 1e6 / n}
        print(n, out[str(n)], flush=True)
        del q, d, rnd, r, dist, sd, star, sph
        torch.cuda.empty_cache()
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
