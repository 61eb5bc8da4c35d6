#!/usr/bin/env python
"""One launch of every kernel family the bench does not already cover, inside a cudaProfilerStart/Stop range, so that

    ncu --set full --profile-from-start off --clock-control none --import-source on -o gpurun_out/<tag>/prof_all \
        python tools/profile_all.py [--groups ray,ray4m,sample,wide,build,adjacency,2d]

captures them in ONE report (diagnostic; no number printed here is a bench value).  Groups:
  ray        k_intersect, 16.7M random rays on the 1 002 528-triangle torus (C3 mesh)
  ray4m      k_intersect, 2M rays on the 4 010 112-triangle torus (one rank's C4 shard)
  sample     k_sample, 16.7M star spheres
  wide       k_closest_wide (1M queries) and k_silhouette_wide (128K queries)
  build      every kernel of snch_scene_build on the 1M-triangle mesh
  adjacency  every kernel of the GPU silhouette adjacency
  2d         the k2_* traversal kernels, 4M queries on an 8 192-segment polyline
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import snch_lbvh_b200 as pkg
    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", default="ray,ray4m,sample,wide,build,adjacency,2d")
    ap.add_argument("--queries", type=int, default=1 << 24)
    args = ap.parse_args()
    groups = args.groups.split(",")
    m = pkg.meshes
    n = args.queries
    v, f = m.bumpy_torus(708, 708)
    lo, hi = m.mesh_bounds(v)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    for kv in filter(None, os.environ.get("SNCH_OPTIONS", "").replace(",", " ").split()):
        k, val = kv.split("=")
        sc.set_option(k, int(val))
    q = torch.from_numpy(m.points_in_box(n, lo, hi, 1.1, seed=2025)).cuda()
    d = torch.from_numpy(m.unit_directions(n, seed=77)).cuda()
    s = torch.from_numpy(m.star_radius_scale(n, seed=4242)).cuda()
    _, dcp = sc.closest_point(q)
    rmax = (dcp * s).contiguous()
    sph = torch.cat([q, rmax[:, None]], dim=1).contiguous()
    rnd = torch.from_numpy(m.uniforms(n, 3, seed=99)).cuda()
    sc4 = q4 = d4 = None
    if "ray4m" in groups:
        v4, f4 = m.bumpy_torus(1416, 1416)
        sc4 = pkg.Scene3(v4, f4).compute_silhouettes().build_bvh()
        lo4, hi4 = m.mesh_bounds(v4)
        q4 = torch.from_numpy(m.points_in_box(1 << 21, lo4, hi4, 1.1, seed=31)).cuda()
        d4 = torch.from_numpy(m.unit_directions(1 << 21, seed=32)).cuda()
    s2 = q2 = d2 = r2 = c2 = u2 = None
    if "2d" in groups:
        v2, g2 = m.wavy_circle(1 << 13, 37, 0.2)
        s2 = pkg.Scene2(v2, g2).compute_silhouettes().build_bvh()
        n2 = 1 << 22
        q2 = torch.from_numpy(m.points_in_box2(n2, v2.min(0), v2.max(0), 1.1, seed=41)).cuda()
        d2 = torch.from_numpy(m.unit_directions2(n2, seed=42)).cuda()
        _, dist2 = s2.closest_point(q2)
        r2 = (dist2 * 2.0).contiguous()
        c2 = torch.cat([q2, r2[:, None]], dim=1).contiguous()
        u2 = torch.from_numpy(m.uniforms(n2, 2, seed=43)).cuda()
    sc_adj = pkg.Scene3(v, f) if "adjacency" in groups else None
    torch.cuda.synchronize()

    torch.cuda.profiler.start()
    if "ray" in groups:
        sc.intersect(q, d)
    if "ray4m" in groups:
        sc4.intersect(q4, d4)
    if "sample" in groups:
        sc.sample_in_sphere(sph, rnd)
    if "wide" in groups:
        sc.closest_point(q[: 1 << 20].contiguous())
        sc.closest_silhouette(q[: 1 << 17].contiguous())
    if "silhouette" in groups:
        sc.closest_silhouette(q, r_max=rmax)
    if "closest" in groups:
        sc.closest_point(q)
    if "build" in groups:
        sc.build_bvh()
    if "adjacency" in groups:
        sc_adj.compute_silhouettes()
    if "2d" in groups:
        s2.closest_point(q2)
        s2.closest_silhouette(q2, r_max=r2)
        s2.intersect(q2, d2)
        s2.sample_in_sphere(c2, u2)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profile_all done:", groups)


if __name__ == "__main__":
    main()
