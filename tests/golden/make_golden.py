"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libsnch_ref_cpu.so = the unmodified reference
headers executed on the CPU).  Run in the build container (needs /root/reference to have been compiled by
`make -C oracle ref`); the fixtures are committed so the GPU box, which has no /root/reference, can check against them.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import snch_lbvh_b200 as pkg  # noqa: E402
from conftest import small_cases, small_cases2  # noqa: E402
from oracle import RefScene, RefScene2  # noqa: E402

NQ = 256


def main():
    m = pkg.meshes
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, (v, f) in small_cases(m).items():
        r = RefScene(v, f, "cpu")
        nodes, aabbs, cones = r.tree()
        morton, sorted_idx = r.morton()
        edges, tri_edges, tri_owned = r.adjacency()
        lo, hi = m.mesh_bounds(v)
        q = m.points_in_box(NQ, lo, hi, 1.5, seed=11)
        d = m.unit_directions(NQ, seed=12)
        ci, cd = r.closest(q)
        s0 = r.silhouette(q, False)
        s1 = r.silhouette(q, True)
        rf, rt, ruv, rp = r.ray(q, d)
        tm = np.full(NQ, 0.75, np.float32)
        rf2, rt2, _, rp2 = r.ray(q, d, tm)
        sph = np.concatenate([q, (cd * 1.5 + 0.05)[:, None]], axis=1).astype(np.float32)
        u = m.uniforms(NQ, seed=13)
        si, sp = r.sample(sph, u)
        np.savez_compressed(
            os.path.join(out_dir, f"{name}.npz"), verts=v, tris=f, nodes=nodes, aabbs=aabbs, cones=cones, morton=morton,
            sorted_idx=sorted_idx, edges=edges, tri_edges=tri_edges, tri_owned=tri_owned, q=q, d=d, closest_idx=ci,
            closest_dist=cd, sil_noflip=s0, sil_flip=s1, ray_found=rf, ray_t=rt, ray_uv=ruv, ray_prim=rp, tmax=tm,
            ray_found_tmax=rf2, ray_t_tmax=rt2, ray_prim_tmax=rp2, sph=sph, u=u, sample_idx=si, sample_pdf=sp)
        print(name, len(f), "tris ->", f"{name}.npz")
    for name, (v, sg) in small_cases2(m).items():  # 2-D: lbvh::scene<2> of the reference
        r = RefScene2(v, sg, "cpu")
        nodes, aabbs, cones = r.tree()
        vert4, owned = r.adjacency()
        lo, hi = v.min(0).astype(np.float64), v.max(0).astype(np.float64)
        q = m.points_in_box2(NQ, lo, hi, 1.5, seed=21)
        d = m.unit_directions2(NQ, seed=22)
        ci, cd = r.closest(q)
        s0 = r.silhouette(q, False)
        s1 = r.silhouette(q, True)
        rf, rt, rs, rp = r.ray(q, d)
        tm = np.full(NQ, 0.4, np.float32)
        rf2, rt2, _, rp2 = r.ray(q, d, tm)
        sph = np.concatenate([q, (cd * 1.5 + 0.05)[:, None]], axis=1).astype(np.float32)
        u = m.uniforms(NQ, seed=23)
        si, sp = r.sample(sph, u)
        np.savez_compressed(
            os.path.join(out_dir, f"{name}.npz"), verts=v, segs=sg, nodes=nodes, aabbs=aabbs, cones=cones, vert4=vert4, owned=owned,
            q=q, d=d, closest_idx=ci, closest_dist=cd, sil_noflip=s0, sil_flip=s1, ray_found=rf, ray_t=rt, ray_s=rs, ray_prim=rp,
            tmax=tm, ray_found_tmax=rf2, ray_t_tmax=rt2, ray_prim_tmax=rp2, sph=sph, u=u, sample_idx=si, sample_pdf=sp)
        print(name, len(sg), "segments ->", f"{name}.npz")


if __name__ == "__main__":
    main()
