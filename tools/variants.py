#!/usr/bin/env python
"""A/B table of the scheduling knobs on the bench workload (GPU box only; diagnostic, not a bench value).

    python tools/variants.py [--queries N] [--torus 708] > gpurun_out/<tag>/variants.json

For every knob setting: device time (CUDA events, best of 3 after one warm-up) of the closest-point, bounded and
unbounded silhouette and ray kernels on the C3 query set, and whether the results are bit-identical to the default
setting's (they must be: knobs only change scheduling).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import snch_lbvh_b200 as pkg
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=1 << 24)
    ap.add_argument("--torus", type=int, default=708)
    ap.add_argument("--sets", default="all")
    args = ap.parse_args()
    m = pkg.meshes
    n = args.queries
    v, f = m.bumpy_torus(args.torus, args.torus)
    lo, hi = m.mesh_bounds(v)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    q = torch.from_numpy(m.points_in_box(n, lo, hi, 1.1, seed=2025)).cuda()
    d = torch.from_numpy(m.unit_directions(n, seed=77)).cuda()
    s = torch.from_numpy(m.star_radius_scale(n, seed=4242)).cuda()
    _, dcp = sc.closest_point(q)
    rmax = (dcp * s).contiguous()
    stream = torch.cuda.current_stream()

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        out = None
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            out = fn()
            b.record(stream)
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best, out

    defaults = {"query.sort_min_n": 16384, "query.sort_bits": 24, "query.sort_rays": 0, "query.cone_filter": 1, "query.seed": 1, "query.blocks_per_sm": 0,
                "query.sort_radius": 2, "query.sil_tail": 8, "query.sil_flush": 24, "query.sil_chunk": 0, "query.wide_max_n": 2097152, "query.wide_max_n_sil": 262144,
                "query.ray_kernel": 1, "query.ray_flush": 8, "query.ray_refill": 8}
    settings = [("default", {}), ("no_lower_bound", {"query.seed": 3}), ("no_seed", {"query.seed": 0}),
                ("sil_flush8", {"query.sil_flush": 8}), ("sil_flush16", {"query.sil_flush": 16}), ("sil_flush32", {"query.sil_flush": 32}), ("sil_tail0", {"query.sil_tail": 0}), ("sil_tail2", {"query.sil_tail": 2}), ("sil_tail4", {"query.sil_tail": 4}), ("sil_tail8", {"query.sil_tail": 8}), ("sil_tail16", {"query.sil_tail": 16}),
                ("sil_tail31", {"query.sil_tail": 31}), ("radius_none", {"query.sort_radius": 0}), ("radius_asc", {"query.sort_radius": 1}),
                ("sil_bps7", {"query.blocks_per_sm": 7}), ("no_sort", {"query.sort_min_n": 0}), ("no_cone_filter", {"query.cone_filter": 0}),
                ("ray_v1", {"query.ray_kernel": 0}), ("ray_v1_sorted", {"query.ray_kernel": 0, "query.sort_rays": 1}), ("ray_sorted", {"query.sort_rays": 1}),
                ("ray_octant", {"query.sort_rays": 2}), ("ray_flush1", {"query.ray_flush": 1}), ("ray_flush4", {"query.ray_flush": 4}),
                ("ray_flush12", {"query.ray_flush": 12}), ("ray_flush16", {"query.ray_flush": 16}), ("ray_flush24", {"query.ray_flush": 24}),
                ("ray_refill1", {"query.ray_refill": 1}), ("ray_refill2", {"query.ray_refill": 2}), 
                ("ray_refill4", {"query.ray_refill": 4}), ("ray_refill16", {"query.ray_refill": 16}), ("ray_f4_r2", {"query.ray_flush": 4, "query.ray_refill": 2}),
                ("ray_f16_r8", {"query.ray_flush": 16, "query.ray_refill": 8}), ("ray_bps8", {"query.blocks_per_sm": 8}), ("ray_bps6", {"query.blocks_per_sm": 6}),
                ("sort_bits_30", {"query.sort_bits": 30}), ("sort_bits_18", {"query.sort_bits": 18})]
    if args.sets != "all":
        settings = [x for x in settings if x[0] in args.sets.split(",")]
    ref = {}
    table = {}
    for name, kv in settings:
        for k, val in {**defaults, **kv}.items():
            sc.set_option(k, val)
        row = {}
        # a knob family only re-times the kernels it can affect
        fam = "ray" if name.startswith("ray") else ("sil" if name.startswith(("sil", "radius", "no_cone")) else ("closest" if name in ("no_lower_bound", "no_seed") else "all"))
        res = {}
        if fam in ("all", "closest"):
            t, (idx, dist) = timed(lambda: sc.closest_point(q))
            row["closest_ms"], row["closest_mqps"] = t, n / t / 1e3
            res["dist"] = dist
        if fam in ("all", "sil"):
            t, sb = timed(lambda: sc.closest_silhouette(q, r_max=rmax))
            row["sil_bounded_ms"], row["sil_bounded_mqps"] = t, n / t / 1e3
            t, su = timed(lambda: sc.closest_silhouette(q))
            row["sil_unbounded_ms"], row["sil_unbounded_mqps"] = t, n / t / 1e3
            res.update(sb=sb, su=su)
            if name in ("default", "sil_tail0"):
                t, (se, _, _) = timed(lambda: sc.closest_silhouette(q, r_max=rmax, with_edge=True))
                row["sil_bounded_with_edge_ms"] = t
                res["se"] = se
        if fam in ("all", "ray"):
            t, (fd, hits) = timed(lambda: sc.intersect(q, d))
            row["ray_ms"], row["ray_mqps"] = t, n / t / 1e3
            t, (fa, _) = timed(lambda: sc.intersect(q, d, any_hit=True))
            row["ray_any_ms"] = t
            res.update(t=hits[:, 0].contiguous(), found=fd)
        if not ref:
            ref = {k: x.clone() for k, x in res.items()}
            ref["se"] = ref["sb"].clone()
        row["identical_to_default"] = {k: bool(torch.equal(ref[k].view(torch.int32) if ref[k].dtype == torch.float32 else ref[k],
                                                           x.view(torch.int32) if x.dtype == torch.float32 else x)) for k, x in res.items()}
        table[name] = row
        print(name, json.dumps(row), file=sys.stderr, flush=True)
    # build: the refit kernel variants (identical arenas are asserted by tests/test_gpu_build.py; here only the time)
    build = {}
    if args.sets == "all" or "build" in args.sets.split(","):
        for name, rk in (("refit_coop", 1), ("refit_per_thread", 0)):
            sc.set_option("build.refit_kernel", rk)
            ts = []
            for _ in range(6):
                sc.build_bvh()
                ts.append(sc.stats()["build_ms"])
            build[name] = {"build_ms_min": min(ts), "build_ms_median": sorted(ts)[len(ts) // 2], "launches": int(sc.counter("build.launches"))}
            print(name, json.dumps(build[name]), file=sys.stderr, flush=True)
        sc.set_option("build.refit_kernel", 1)
    print(json.dumps({"queries": n, "triangles": len(f), "table": table, "build": build}, indent=1))


if __name__ == "__main__":
    main()
