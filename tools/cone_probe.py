"""Are the cones of the GPU build the oracle's, bit for bit?  (GPU; writes gpurun_out/cone_probe.json)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import snch_lbvh_b200 as pkg  # noqa: E402
from oracle import OracleScene, OracleScene2  # noqa: E402
from snch_lbvh_b200 import meshes  # noqa: E402
from test_gpu_fuzz import CASES, CASES2, soup, soup2  # noqa: E402


def bits(a):
    """bit patterns, every NaN mapped to one value (x86 and the GPU produce different default NaNs)"""
    a = np.ascontiguousarray(a, np.float32)
    return np.where(np.isnan(a), np.uint32(0x7FC00000), a.view(np.uint32))


out = {}
cases = {"sphere5": meshes.icosphere(5), "torus97x61": meshes.bumpy_torus(97, 61), "torus128": meshes.bumpy_torus(128, 128), "grid16": meshes.open_grid(16),
         "torus708": meshes.bumpy_torus(708, 708)}
for c in CASES:
    cases[f"soup{c[0]}"] = soup(*c)
for nm, (v, f) in cases.items():
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    orc = OracleScene(v, f)
    _, _, co = orc.tree()
    c = sc.export(pkg.ExportKind.CONES)
    taint = orc.q1_taint().astype(bool)
    valid = co[:, 3] >= 0
    ok = valid & ~taint
    eq = (bits(c)[ok] == bits(co)[ok]).all(axis=1)
    # tainted nodes: axis and radius are defined
    eqt = (bits(c)[taint & valid][:, [0, 1, 2, 4]] == bits(co)[taint & valid][:, [0, 1, 2, 4]]).all(axis=1) if (taint & valid).any() else np.ones(0, bool)
    bad = np.nonzero(ok)[0][~eq][:4]
    for b in bad:
        print("   node", int(b), "gpu", c[b].tolist(), "oracle", co[b].tolist(), "internal" if b < len(c) // 2 else "leaf", flush=True)
    ms = []
    for _ in range(5):
        sc.build_bvh()
        ms.append(sc.stats()["build_ms"])
    out[nm] = {"nodes": int(len(c)), "compared": int(ok.sum()), "bit_equal": int(eq.sum()), "tainted": int((taint & valid).sum()), "tainted_axis_radius_equal": int(eqt.sum()),
               "invalid_equal": bool(np.array_equal(bits(c)[~valid][:, 3], bits(co)[~valid][:, 3])), "build_ms": float(np.median(ms))}
    print(nm, out[nm], flush=True)
cases2 = {"wavy60000": meshes.wavy_circle(60000, 11, 0.3), "open449": meshes.open_polyline(449), "polysoup": meshes.polyline_soup(3, 5, 200)}
for c in CASES2:
    cases2[f"soup2_{c[0]}"] = soup2(*c)
for nm, (v, s) in cases2.items():
    sc = pkg.Scene2(v, s).compute_silhouettes().build_bvh()
    orc = OracleScene2(v, s)
    _, _, co, q1 = orc.tree()
    c = sc.export(pkg.ExportKind.CONES)
    valid = co[:, 2] >= 0
    ok = valid & ~q1.astype(bool)
    eq = (bits(c)[ok] == bits(co)[ok]).all(axis=1)
    out[nm] = {"nodes": int(len(c)), "compared": int(ok.sum()), "bit_equal": int(eq.sum())}
    print(nm, out[nm], flush=True)
    for b in np.nonzero(ok)[0][~eq][:4]:
        print("   node", int(b), "gpu", c[b].tolist(), "oracle", co[b].tolist(), "internal" if b < len(c) // 2 else "leaf", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cone_probe.json"), "w"), indent=1)
