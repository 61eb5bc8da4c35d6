#!/usr/bin/env python
"""Small-shard behaviour of the silhouette step (what one GPU sees at N = 2 / 4 / 8 of the strong-scaling bench): for shard
sizes 8M / 4M / 2M of the C3 batch and a few knob settings, the whole step and the traversal part alone (library CUDA
events), so that the fixed part of a step is visible.  Diagnostic, not a bench value.   python tools/shard_exp.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import snch_lbvh_b200 as pkg
    m = pkg.meshes
    v, f = m.bumpy_torus(708, 708)
    lo, hi = m.mesh_bounds(v)
    sc = pkg.Scene3(v, f).compute_silhouettes().build_bvh()
    n_all = 1 << 24
    q_all = torch.from_numpy(m.points_in_box(n_all, lo, hi, 1.1, seed=2025)).cuda()
    s_all = torch.from_numpy(m.star_radius_scale(n_all, seed=4242)).cuda()
    _, dcp = sc.closest_point(q_all)
    r_all = (dcp * s_all).contiguous()
    stream = torch.cuda.current_stream()
    settings = [("default", {}), ("chunk64", {"query.sil_chunk": 64}), ("chunk8", {"query.sil_chunk": 8}), ("chunk16_tail4", {"query.sil_chunk": 16, "query.sil_tail": 4}), ("chunk16_tail16", {"query.sil_chunk": 16, "query.sil_tail": 16}),
                ("chunk32", {"query.sil_chunk": 32}), ("tail0", {"query.sil_tail": 0})]
    defaults = {"query.sil_tail": 8, "query.sil_chunk": 0, "query.blocks_per_sm": 0, "query.sort_radius": 2, "query.sort_bits": 24}
    out = {}
    for n in (1 << 24, 1 << 23, 1 << 22, 1 << 21):
        q, r = q_all[:n].contiguous(), r_all[:n].contiguous()
        row = {}
        for name, kv in settings:
            for k, val in {**defaults, **kv}.items():
                sc.set_option(k, val)
            sc.closest_silhouette(q, r_max=r)
            torch.cuda.synchronize()
            sc.set_option("query.time_kernels", 1)
            sc.counter("query.traversal_ms", reset=True)
            best = 1e30
            for _ in range(4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                sc.closest_silhouette(q, r_max=r)
                b.record(stream)
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            kern = sc.counter("query.traversal_ms", reset=True) / 4
            sc.set_option("query.time_kernels", 0)
            row[name] = {"step_ms": best, "traversal_ms_mean": kern, "mqps": n / best / 1e3, "ideal_ms_from_16M": None}
            print(n, name, json.dumps(row[name]), file=sys.stderr, flush=True)
        dd = torch.from_numpy(m.unit_directions(n, seed=77)).cuda()
        for k, val in defaults.items():
            sc.set_option(k, val)
        for kind, fn in (("closest", lambda: sc.closest_point(q)), ("ray", lambda: sc.intersect(q, dd))):
            fn()
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                fn()
                b.record(stream)
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            row[kind] = {"step_ms": best, "mqps": n / best / 1e3, "ideal_ms_from_16M": None}
            print(n, kind, json.dumps(row[kind]), file=sys.stderr, flush=True)
        out[str(n)] = row
    for n, row in out.items():
        for name in row:
            base = out[str(1 << 24)][name if name in ("closest", "ray") else "default"]["step_ms"]
            row[name]["ideal_ms_from_16M"] = base * int(n) / (1 << 24)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
