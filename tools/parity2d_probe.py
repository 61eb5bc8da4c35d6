"""How close is the 2-D GPU path to the 2-D oracle, bit for bit?  (GPU; writes gpurun_out/parity2d_probe.json)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import snch_lbvh_b200 as pkg  # noqa: E402
from oracle import OracleScene2  # noqa: E402
from snch_lbvh_b200 import meshes  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def soup2(seed, nv, ns, offset, scale):
    rng = np.random.default_rng(seed)
    v = (rng.random((nv, 2)) * scale + offset).astype(np.float32)
    s = rng.integers(0, nv, (ns, 2))
    s = s[s[:, 0] != s[:, 1]]
    if ns >= 40:
        s = np.concatenate([s, s[:5], s[:3, ::-1]])
    return v, s.astype(np.int32)


cases = {}
g = os.path.join(ROOT, "tests", "golden")
for nm in ("poly_circle", "poly_open", "poly_soup"):
    z = np.load(os.path.join(g, nm + ".npz"))
    cases[nm] = (z["verts"], z["segs"])
v, s = meshes.wavy_circle(60000, 11, 0.3)
cases["wavy60000"] = (v, s[np.random.default_rng(5).permutation(len(s))].copy())
for c in [(1, 4, 5, 0.0, 1.0), (2, 8, 40, 0.0, 1.0), (3, 100, 300, 0.0, 1.0), (4, 300, 500, 1000.0, 1.0), (5, 300, 500, 0.0, 1e-3), (6, 3000, 20000, 0.0, 1.0)]:
    cases[f"soup{c[0]}"] = soup2(*c)

out = {}
for nm, (v, s) in cases.items():
    sc = pkg.Scene2(v, s).compute_silhouettes().build_bvh()
    orc = OracleScene2(v, s)
    n = 20000
    q = meshes.points_in_box2(n, v.min(0), v.max(0), 1.3, seed=6)
    d = meshes.unit_directions2(n, seed=7)
    r = {}
    nodes, aabbs, cones, q1 = orc.tree()
    r["nodes_equal"] = bool(np.array_equal(sc.export(pkg.ExportKind.NODES), nodes))
    r["aabbs_equal"] = bool(np.array_equal(bits(sc.export(pkg.ExportKind.AABBS)), bits(aabbs)))
    idx, dist = sc.closest_point(q)
    oidx, odist = orc.closest(q)
    r["closest_bits"] = float(np.mean(bits(dist) == bits(odist)))
    r["closest_idx"] = float(np.mean(idx == oidx))
    rmax = (odist * meshes.star_radius_scale(n)).astype(np.float32)
    for key, kw in (("sil", {}), ("sil_flip", {"flip": True}), ("sil_rmax", {"r_max": rmax})):
        a = sc.closest_silhouette(q, **kw)
        b = orc.silhouette(q, kw.get("flip", False), kw.get("r_max"))
        r[key + "_bits"] = float(np.mean(bits(a) == bits(b)))
        r[key + "_finite_agree"] = float(np.mean(np.isfinite(a) == np.isfinite(b)))
    found, hits = sc.intersect(q, d)
    of, ot, _, op = orc.ray(q, d)
    r["ray_found"] = float(np.mean(found.astype(bool) == of.astype(bool)))
    r["ray_t_bits"] = float(np.mean(bits(hits["t"]) == bits(ot)))
    r["ray_prim"] = float(np.mean(hits["prim"] == op.astype(np.uint32)))
    sph = np.concatenate([q, (odist * 1.5 + 0.02 * float(odist.max()))[:, None]], 1).astype(np.float32)
    u = meshes.uniforms(n, 2, seed=8)
    si, pdf, _ = sc.sample_in_sphere(sph, u)
    oi, opdf = orc.sample(sph, u[:, 0].copy())
    r["sample_idx"] = float(np.mean(si == oi))
    same = (si == oi) & (si >= 0)
    r["sample_pdf_bits"] = float(np.mean(bits(pdf[same]) == bits(opdf[same]))) if same.any() else 1.0
    out[nm] = r
    print(nm, r, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity2d_probe.json"), "w"), indent=1)
