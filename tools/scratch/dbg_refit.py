import sys, numpy as np
sys.path.insert(0, ".")
import snch_lbvh_b200 as pkg
meshes = pkg.meshes
mode = sys.argv[1] if len(sys.argv) > 1 else "dev"
v, f = meshes.bumpy_torus(90, 60)
rng = np.random.default_rng(3)
v2 = (v * np.float32(1.07) + rng.normal(0, 0.01, v.shape)).astype(np.float32)
lo, hi = meshes.mesh_bounds(v2)
q = meshes.points_in_box(20000, lo, hi, 1.3, seed=9)
fresh = pkg.Scene3(v2, f).compute_silhouettes().build_bvh()
moved = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
if mode == "dev":
    import torch
    t = torch.from_numpy(v2).cuda()
    moved.update_vertices(t)
else:
    moved.update_vertices(v2)
moved.build_bvh(refit_only=True)
a = moved.closest_silhouette(q); b = fresh.closest_silhouette(q)
fin = np.isfinite(a) & np.isfinite(b)
rel = np.abs(a[fin]-b[fin])/np.maximum(b[fin],1e-9)
print("finite agree", np.mean(np.isfinite(a)==np.isfinite(b)), "within 1e-5", np.mean(rel<=1e-5), "max rel", rel.max(), "n bad", int((rel>1e-5).sum()))
print("refit smaller:", int((a[fin] < b[fin]*(1-1e-5)).sum()), "fresh smaller:", int((b[fin] < a[fin]*(1-1e-5)).sum()))
