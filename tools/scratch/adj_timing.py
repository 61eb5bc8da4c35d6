"""Times compute_silhouettes() (device adjacency) repeatedly in one process: is the first call's cost (context, lazy module
load) leaking into later calls?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import snch_lbvh_b200 as pkg

v, f = pkg.meshes.bumpy_torus(708, 708)
for mode in (1, 0, 1):
    sc = pkg.Scene3(v, f).set_option("adjacency.device", mode)
    for k in range(4):
        t0 = time.perf_counter()
        sc.compute_silhouettes()
        print(f"mode {mode} call {k}: adjacency_ms={sc.stats()['adjacency_ms']:.2f} wall={1e3 * (time.perf_counter() - t0):.2f}", flush=True)
