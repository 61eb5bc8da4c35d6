"""Detail of the fuzz-sweep failures: which silhouette queries differ from the oracle, under which cone filter, and what the two edges look like (GPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
os.environ["SNCH_FUZZ_EXTRA"] = "150"
import snch_lbvh_b200 as pkg  # noqa: E402
from oracle import OracleScene  # noqa: E402
from snch_lbvh_b200 import meshes  # noqa: E402
from test_gpu_fuzz import CASES, soup  # noqa: E402

bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)  # noqa: E731
want = [int(x) for x in sys.argv[1:]] or [102, 113, 114, 154, 236]

# ---- an edge shared by a proper triangle and a ZERO-AREA one (collinear vertices): its second face normal is normalize(0) = NaN
vd = np.array([[0, 0, 0], [1, 0, 0], [0.5, 1, 0], [2, 0, 0], [0.5, -1, 0.3], [0.5, 0.4, 1.0]], np.float32)
fd = np.array([[0, 1, 2], [1, 0, 3], [0, 1, 4], [0, 2, 5], [2, 1, 5]], np.int32)  # (1, 0, 3) is degenerate; edge 0-1 has three faces
for name, ff in (("degenerate_face", fd[[0, 1, 3, 4]]), ("degenerate_face_fan", fd)):
    scd = pkg.Scene3(vd, ff).compute_silhouettes().build_bvh()
    od = OracleScene(vd, ff)
    qd = meshes.points_in_box(4000, vd.min(0), vd.max(0), 1.5, seed=5)
    for flip in (False, True):
        a, b = scd.closest_silhouette(qd, flip=flip), od.silhouette(qd, flip, nthreads=4)
        print(name, "flip", flip, "silhouette differs on", int(np.count_nonzero(bits(a) != bits(b))), "of", len(qd), "gpu<oracle", int(np.count_nonzero(a < b)), flush=True)
    _, dc = scd.closest_point(qd)
    print(name, "closest differs on", int(np.count_nonzero(bits(dc) != bits(od.closest(qd, nthreads=4)[1]))), flush=True)
for c in CASES:
    if c[0] not in want:
        continue
    v, f = soup(*c)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    orc = OracleScene(v, f)
    lo, hi = meshes.mesh_bounds(v)
    q = meshes.points_in_box(6000, lo, hi, 1.5, seed=1000 + c[0])
    e4, _, _ = orc.adjacency()
    _, dcp = orc.closest(q, nthreads=8)
    rmax = (dcp * meshes.star_radius_scale(len(q))).astype(np.float32)
    for flip in (False, True):
        d_o, e_o, p_o = orc.silhouette_ex(q, flip, r_max=rmax, nthreads=8)
        du_o = orc.silhouette(q, flip, nthreads=8)
        for cf in (1, 0):
            sc.set_option("query.cone_filter", cf)
            d, e, p = sc.closest_silhouette(q, flip=flip, r_max=rmax, with_edge=True)
            bad = np.nonzero(bits(d) != bits(d_o))[0]
            filt = np.where(du_o <= rmax, du_o, np.inf).astype(np.float32)
            print(f"seed {c[0]} {c[1:]} BOUNDED flip {flip} cone_filter {cf}: {len(bad)} differ; gpu<oracle {int(np.count_nonzero(d[bad] < d_o[bad]))} gpu>oracle {int(np.count_nonzero(d[bad] > d_o[bad]))};"
                  f" oracle bounded != filtered oracle unbounded on {int(np.count_nonzero(bits(d_o) != bits(filt)))}; gpu != filtered oracle unbounded on {int(np.count_nonzero(bits(d) != bits(filt)))}", flush=True)
            if cf == 1:
                for i in bad[:4]:
                    print(f"   q{i} rmax={rmax[i]!r} dcp={dcp[i]!r} gpu d={d[i]!r} edge={int(e[i])}  oracle d={d_o[i]!r} edge={int(e_o[i])} oracle unbounded={du_o[i]!r}", flush=True)
        sc.set_option("query.cone_filter", 1)
    dd = meshes.unit_directions(6000, seed=2000 + c[0])
    f_o, t_o, _, p_o = orc.ray(q, dd, nthreads=8)
    for rk in (1, 2, 0):
        sc.set_option("query.ray_kernel", rk)
        found, hits = sc.intersect(q, dd)
        bad = np.nonzero((bits(hits["t"]) != bits(t_o)) | (found.astype(bool) != f_o.astype(bool)))[0]
        print(f"seed {c[0]} RAYS ray_kernel {rk} ({sc.last_kernel()}): {len(bad)} differ; gpu<oracle {int(np.count_nonzero(hits['t'][bad] < t_o[bad]))} gpu>oracle {int(np.count_nonzero(hits['t'][bad] > t_o[bad]))}", flush=True)
        for i in bad[:3]:
            print(f"   ray{i} gpu t={hits['t'][i]!r} prim={int(hits['prim'][i])} oracle t={t_o[i]!r} prim={int(p_o[i])}", flush=True)
    sc.set_option("query.ray_kernel", 1)
