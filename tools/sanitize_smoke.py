#!/usr/bin/env python
"""Small workload that launches every kernel family once or twice, for `compute-sanitizer --tool memcheck|racecheck|initcheck`
(the warp-shared leaf queue, the tail list, the parked-leaf ray walk, the packet / one-query-per-warp kernels, the build, the
adjacency, the 2-D kernels, the wavefront step).  Sizes are chosen so that the per-lane AND the small-batch kernels both run:
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import snch_lbvh_b200 as pkg
    m = pkg.meshes
    v, f = m.bumpy_torus(48, 36)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    lo, hi = m.mesh_bounds(v)
    n = 20000
    q = m.points_in_box(n, lo, hi, 1.3, seed=1)
    d = m.unit_directions(n, seed=2)
    u = m.uniforms(n, 3, seed=3)
    flip = (np.arange(n) % 2).astype(np.uint8)
    for wide in (0, 1 << 30):  # per-lane / packet kernels, then the one-query-per-warp kernels
        sc.set_option("query.wide_max_n", wide).set_option("query.wide_max_n_sil", wide)
        for rk in (0, 2):
            sc.set_option("query.ray_kernel", rk)
            _, dist = sc.closest_point(q)
            sc.closest_silhouette(q, flip=flip)
            sc.set_option("query.sil_tail", 31)
            sc.closest_silhouette(q, r_max=dist * 2, with_edge=True)
            sc.set_option("query.sil_tail", 8)
            sc.intersect(q, d)
            sc.intersect(q, d, t_max=np.full(n, 0.5, np.float32), any_hit=True)
            sc.sample_in_sphere(np.concatenate([q, (dist + 0.05)[:, None]], 1).astype(np.float32), u)
            sc.wost_step(q, d, u, flip=flip, with_edge=True)
    sc.update_vertices((v * 1.1).astype(np.float32)).build_bvh(refit_only=True)
    sc.set_option("build.refit_kernel", 0).build_bvh()
    v2, s2 = m.wavy_circle(513, 5, 0.2)
    p2 = pkg.Scene2(v2, s2).compute_silhouettes().build_bvh()
    q2 = m.points_in_box2(8000, v2.min(0), v2.max(0), 1.2, seed=4)
    d2 = m.unit_directions2(8000, seed=5)
    for wide in (0, 1 << 30):
        p2.set_option("query.wide_max_n", wide)
        _, dd = p2.closest_point(q2)
        p2.closest_silhouette(q2, r_max=dd * 2, with_vertex=True)
        p2.intersect(q2, d2)
        p2.sample_in_sphere(np.concatenate([q2, (dd + 0.05)[:, None]], 1).astype(np.float32), m.uniforms(8000, 2, seed=6))
    print("SANITIZE_WORKLOAD_OK")


if __name__ == "__main__":
    main()
