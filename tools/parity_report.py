"""Prints how closely the CUDA path matches the CPU oracle (and, when present, the reference's CUDA build) on seeded
inputs: exact-bit fractions and worst relative errors.  Diagnostic only; the asserted bars live in tests/."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import snch_lbvh_b200 as pkg  # noqa: E402
from oracle import OracleScene, RefScene, ref_available  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def rel(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
    r[np.isinf(a) & np.isinf(b)] = 0
    r[np.isnan(r)] = np.inf
    return r


def report(name, v, f, n, out):
    m = pkg.meshes
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    orc = OracleScene(v, f)
    lo, hi = m.mesh_bounds(v)
    q = m.points_in_box(n, lo, hi, 1.2, seed=71)
    d = m.unit_directions(n, seed=72)
    row = {"tris": len(f), "queries": n}
    idx, dist = sc.closest_point(q)
    oi, od = orc.closest(q, nthreads=16)
    row["closest_dist_biteq"] = float(np.mean(bits(dist) == bits(od)))
    row["closest_dist_maxrel"] = float(rel(dist, od).max())
    row["closest_idx_eq"] = float(np.mean(idx == oi))
    for flip in (False, True):
        s = sc.closest_silhouette(q, flip=flip)
        os_ = orc.silhouette(q, flip, nthreads=16)
        row[f"sil_flip{int(flip)}_biteq"] = float(np.mean(bits(s) == bits(os_)))
        row[f"sil_flip{int(flip)}_frac_rel_gt_1e-5"] = float(np.mean(rel(s, os_) > 1e-5))
    # star radii + the optional edge / point outputs against the oracle's silhouette_ex
    rmax = (od * m.star_radius_scale(n, seed=74)).astype(np.float32)
    sd, se, sp_ = sc.closest_silhouette(q, r_max=rmax, with_edge=True)
    d_o, e_o, p_o = orc.silhouette_ex(q, False, r_max=rmax, nthreads=16)
    fin = np.isfinite(d_o)
    row["sil_star_radius_biteq"] = float(np.mean(bits(sd) == bits(d_o)))
    row["sil_finite_fraction"] = float(fin.mean())
    d_at, p_at = orc.point_edge_distance(q[fin], se[fin].astype(np.int32))
    row["sil_edge_attains_distance"] = float(np.mean(bits(d_at) == bits(d_o[fin])))
    row["sil_point_is_closest_point_on_edge_biteq"] = float(np.mean(np.all(bits(p_at) == bits(sp_[fin]), axis=1)))
    same = se[fin] == e_o[fin].astype(np.uint32)
    row["sil_edge_same_as_oracle"] = float(same.mean())
    row["sil_point_biteq_when_same_edge"] = float(np.mean(np.all(bits(sp_[fin][same]) == bits(p_o[fin][same]), axis=1)))
    dev = np.abs(sp_[fin] - p_o[fin]).max(axis=1) / np.maximum(d_o[fin], 1e-30)
    row["sil_point_frac_beyond_1e-5_of_distance"] = float(np.mean(dev > 1e-5))   # exact float ties between different edges
    row["sil_point_max_dev_over_distance"] = float(dev.max())
    found, hits = sc.intersect(q, d)
    of, ot, ouv, op = orc.ray(q, d, nthreads=16)
    row["ray_found_eq"] = float(np.mean(found.astype(bool) == of.astype(bool)))
    row["ray_t_biteq"] = float(np.mean(bits(hits["t"]) == bits(ot)))
    row["ray_uv_biteq"] = float(np.mean((bits(hits["u"]) == bits(ouv[:, 0])) & (bits(hits["v"]) == bits(ouv[:, 1]))))
    row["ray_prim_eq"] = float(np.mean(hits["prim"] == op))
    sph = np.concatenate([q, (od * 1.5 + 0.05)[:, None]], axis=1).astype(np.float32)
    rnd = m.uniforms(n, 3, seed=73)
    si, sp, _ = sc.sample_in_sphere(sph, rnd)
    osi, osp = orc.sample(sph, rnd[:, 0].copy())
    row["sample_idx_eq"] = float(np.mean(si == osi))
    row["sample_pdf_biteq"] = float(np.mean(bits(sp) == bits(osp)))
    K = pkg.ExportKind
    _, _, oc = orc.tree()
    c = sc.export(K.CONES)
    ok = (~orc.q1_taint()) & (oc[:, 3] >= 0)
    row["cone_half_frac_rel_gt_1e-5"] = float(np.mean(rel(c[ok, 3], oc[ok, 3]) > 1e-5))
    row["cone_half_max_abs"] = float(np.abs(c[ok, 3] - oc[ok, 3]).max())
    row["cone_half_biteq"] = float(np.mean(bits(c[ok, 3]) == bits(oc[ok, 3])))
    row["cone_half_frac_abs_gt_1e-4"] = float(np.mean(np.abs(c[ok, 3] - oc[ok, 3]) > 1e-4))
    if ref_available("cuda"):
        ref = RefScene(v, f, "cuda")
        _, rd = ref.closest(q)
        row["refcuda_closest_frac_rel_gt_1e-5"] = float(np.mean(rel(dist, rd) > 1e-5))
        rs = ref.silhouette(q)
        row["refcuda_sil_frac_rel_gt_1e-5"] = float(np.mean(rel(sc.closest_silhouette(q), rs) > 1e-5))
        # the reference against ITSELF: same headers on Thrust's CPP backend (no FMA contraction) vs its CUDA build.  Above 200K
        # triangles the C oracle stands in for the CPP-backend build (tests/test_oracle_pinning.py pins them bit-identical;
        # the reference's host adjacency passes take minutes there).
        ns = min(n, 20000)
        if ref_available("cpu") and len(f) <= 200000:
            rcpu = RefScene(v, f, "cpu")
            rcs, rcd, who = rcpu.silhouette(q[:ns]), rcpu.closest(q[:ns])[1], "reference CPU build (oracle/_ref/libsnch_ref_cpu.so)"
        else:
            rcs, rcd, who = orc.silhouette(q[:ns], nthreads=16), od[:ns], "C oracle (pinned bit-identical to the reference CPU build)"
        row["refcpu_vs_refcuda_sil_frac"] = float(np.mean(rel(rcs, rs[:ns]) > 1e-5))
        row["refcpu_vs_refcuda_closest_frac"] = float(np.mean(rel(rcd, rd[:ns]) > 1e-5))
        row["refcpu_side"] = who
        rf, rt, _, _ = ref.ray(q, d)
        row["refcuda_ray_found_eq"] = float(np.mean(found.astype(bool) == rf.astype(bool)))
        both = found.astype(bool) & rf.astype(bool)
        row["refcuda_ray_t_frac_rel_gt_1e-5"] = float(np.mean(rel(hits["t"][both], rt[both]) > 1e-5))
        _, _, rc = ref.tree()
        row["refcuda_tainted_half_angles"] = [float(x) for x in rc[orc.q1_taint(), 3][:12]]
        row["refcuda_cone_half_frac_rel_gt_1e-5"] = float(np.mean(rel(c[ok, 3], rc[ok, 3]) > 1e-5))
        row["refcuda_cone_half_max_abs"] = float(np.abs(c[ok, 3] - rc[ok, 3]).max())
    out[name] = row
    print(name, json.dumps(row), flush=True)


def main():
    m = pkg.meshes
    out = {}
    report("ico5", *m.icosphere(5), 65536, out)
    report("grid40", *m.open_grid(40), 20000, out)
    report("torus300", *m.bumpy_torus(300, 300), 40000, out)
    report("torus708", *m.bumpy_torus(708, 708), 40000, out)
    if "--big" in sys.argv:
        report("torus1416", *m.bumpy_torus(1416, 1416), 20000, out)
        report("torus2240", *m.bumpy_torus(2240, 2240), 20000, out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
