#!/usr/bin/env python
"""bench.py — the headline measurement of the SNCH-LBVH hot path on B200.

Workload (BASELINE.json config C3, the one the north_star target is quoted on): 16 777 216 nearest-silhouette queries
with WoSt star radii (r_max = s * closest-point distance, s ~ U[0.5,4)) against the LBVH+SNCH of the 1 002 528-triangle
synthetic "bumpy torus".  A step = one pass of that batch through the traversal kernel.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

* value  : M queries/s with queries and results resident in HBM (CUDA events on the launching stream, max over ranks)
* e2e    : the same batch through the C-ABI with HOST buffers (H2D + kernel + D2H inside the timed region)
* roofline / cpu_baseline / extra: see DESIGN.md "Measurement"
* --impl reference: the CPU baseline arm (fcpw's CPU backend bundled with the reference, BASELINE.json north_star; the
  reference's own host query path is broken), on the box's host cores.
N > 1: launched by torchrun, one rank per GPU; rank 0 builds, the arena is broadcast over NCCL, every rank traverses its
own 16M-query shard (weak scaling), results stay per rank (e2e gathers them to the host).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_QUERIES = 1 << 24
TORUS = 708
WORKLOAD = "C3: 16777216 nearest-silhouette queries, WoSt star radii (r_max = s*d_closest, s~U[0.5,4)), 1002528-triangle bumpy torus"


class quiet_stdout:
    """The reference prints "Morton code collision detected." on stdout from C++ (bvh.cuh:466); the bench line must be the
    only thing on stdout, so fd 1 is pointed at stderr while reference code runs."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(m, n, seed_shift=0):
    v, f = m.bumpy_torus(TORUS, TORUS)
    lo, hi = m.mesh_bounds(v)
    q = m.points_in_box(n, lo, hi, 1.1, seed=2025 + seed_shift)
    s = m.star_radius_scale(n, seed=4242 + seed_shift)
    return v, f, q, s


# ----------------------------------------------------------------------------------------------------------------
# CPU baseline (fcpw CPU backend; falls back to the oracle port if the prebuilt fcpw library did not travel)
# ----------------------------------------------------------------------------------------------------------------
def cpu_silhouette_baseline(v, f, q, rmax, target_s=12.0, max_n=1 << 21):
    from oracle import FcpwScene, OracleScene, ref_available
    ncores = os.cpu_count() or 1
    if ref_available("fcpw"):
        sc = FcpwScene(v, f)
        kind, cores = "reference", sc.threads

        def run(a, b):
            sc.silhouette(q[a:b], r_max=rmax[a:b])
            return sc.last_ms / 1e3
        what = "fcpw CPU backend (ext/fcpw, Bvh_SurfaceArea, Enoki 8-wide), findClosestSilhouettePoints"
    else:
        sc = OracleScene(v, f)
        kind, cores = "port", ncores

        def run(a, b):
            t0 = time.perf_counter()
            sc.silhouette(q[a:b], r_max=rmax[a:b], nthreads=ncores)
            return time.perf_counter() - t0
        what = "oracle/snch_oracle.c (reference algorithm restated in C), pthreads"
    n0 = min(20000, len(q))
    t = run(0, n0)
    n1 = int(min(max_n, len(q) - n0, max(n0, n0 * target_s / max(t, 1e-6))))
    t1 = run(n0, n0 + n1)
    return {"value": n1 / t1 / 1e6, "unit": "M queries/s", "cores": int(cores), "kind": kind,
            "sample": f"{n1} of the {len(q)} C3 queries (same mesh, same star radii), {what}"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    import snch_lbvh_b200 as pkg
    from oracle import OracleScene
    m = pkg.meshes
    n = 1 << 21
    v, f, q, s = make_inputs(m, n)
    # star radii need closest-point distances: computed with the CPU oracle for this arm (no GPU code on this path)
    orc = OracleScene(v, f)
    _, dcp = orc.closest(q, nthreads=os.cpu_count() or 1)
    rmax = (dcp * s).astype(np.float32)
    vals = []
    info = None
    for i in range(args.warmup + args.steps):
        info = cpu_silhouette_baseline(v, f, q, rmax, target_s=6.0, max_n=1 << 20)
        if i >= args.warmup:
            vals.append(info["value"])
    val = statistics.median(vals)
    info["value"] = val
    line = {"impl": "reference", "metric": "M queries/sec (nearest-silhouette, star radii) @1M tris", "value": val, "unit": "M queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU arm: each step is a bounded sample of the workload"},
            "cpu_baseline": info, "e2e": {"value": val, "unit": "M queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    return line


def other_config(cfg, pkg, m, dev, stream):
    """BASELINE.json configs C4 / C5 at their per-GPU share of the 8-GPU batch (one rank's shard: the tree is replicated,
    so a rank's throughput does not depend on the other ranks).  Device-resident, CUDA events on `stream`."""
    import torch
    if cfg == "2d":
        return scene2_config(pkg, m, dev, stream)
    if cfg == "c1":
        return c1_config(pkg, m, dev, stream)
    nu, per_gpu = {"c4": (1416, (1 << 24) // 8), "c5": (2240, (1 << 26) // 8)}[cfg]
    v, f = m.bumpy_torus(nu, nu)
    t0 = time.perf_counter()
    sc = pkg.Scene3(v, f, device=dev).compute_silhouettes().build_bvh(stream=stream)
    setup_ms = (time.perf_counter() - t0) * 1e3
    st = sc.stats()
    lo, hi = m.mesh_bounds(v)
    out = {"triangles": int(st["num_objects"]), "queries_per_gpu": per_gpu, "build_ms": st["build_ms"], "adjacency_ms": st["adjacency_ms"],
           "arena_bytes": st["arena_bytes"], "scene_setup_wall_ms": setup_ms}
    q = torch.from_numpy(m.points_in_box(per_gpu, lo, hi, 1.1 if cfg == "c4" else 1.0, seed=31)).to(f"cuda:{dev}")
    d = torch.from_numpy(m.unit_directions(per_gpu, seed=32)).to(f"cuda:{dev}")

    def timed(fn, reps=5):
        fn()
        stream.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        stream.synchronize()
        return a.elapsed_time(b) / reps

    with torch.cuda.stream(stream):
        if cfg == "c4":
            ms = timed(lambda: sc.intersect(q, d, stream=stream))
            out.update(workload="C4 shard: closest-hit rays, t_max = inf", ms_per_step=ms, ray_mqps=per_gpu / ms / 1e3)
        else:
            u = torch.from_numpy(m.uniforms(per_gpu, 3, seed=33)).to(f"cuda:{dev}")
            ms = timed(lambda: sc.wost_step(q, d, u, stream=stream), reps=3)
            out.update(workload="C5 shard: one wavefront WoSt step per walker = closest point + silhouette (r_max = d) + ray (t_max = star "
                                "radius) + sample-in-sphere, fused in snch_wost_step_batch",
                       ms_per_step=ms, walker_steps_mqps=per_gpu / ms / 1e3, queries_mqps=4 * per_gpu / ms / 1e3)
    del sc
    torch.cuda.empty_cache()
    return out


def c1_config(pkg, m, dev, stream):
    """BASELINE.json config C1: 65 536 closest-point / silhouette / ray queries uniform in [-1.5, 1.5]^3 on the 20 480-triangle
    icosphere, device-resident (CUDA events on `stream`, mean of 20), with the reference's CUDA path on the same queries.  A
    batch this small does not fill the machine: its run time is the critical path of its most expensive query (the points
    near the centre of the sphere see every triangle at the same distance), which is what the one-query-per-warp kernels cut."""
    import torch
    v, f = m.icosphere(5)
    sc = pkg.Scene3(v, f, device=dev).compute_silhouettes().build_bvh(stream=stream)
    n = 65536
    q_h = np.random.default_rng(1234).uniform(-1.5, 1.5, (n, 3)).astype(np.float32)
    d_h = m.unit_directions(n, seed=3)
    q, d = torch.from_numpy(q_h).to(f"cuda:{dev}"), torch.from_numpy(d_h).to(f"cuda:{dev}")

    def timed(fn, reps=20):
        fn()
        stream.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        stream.synchronize()
        return a.elapsed_time(b) / reps

    out = {"triangles": len(f), "queries": n}
    with torch.cuda.stream(stream):
        out["closest_ms"] = timed(lambda: sc.closest_point(q, stream=stream))
        out["silhouette_ms"] = timed(lambda: sc.closest_silhouette(q, stream=stream))
        out["ray_ms"] = timed(lambda: sc.intersect(q, d, stream=stream))
    for k in ("closest", "silhouette", "ray"):
        out[f"{k}_mqps"] = n / out[f"{k}_ms"] / 1e3
    # the traversal kernels alone (the library's own CUDA events around them): what the reference's numbers below measure
    sc.set_option("query.time_kernels", 1)
    with torch.cuda.stream(stream):
        for k, fn in (("closest", lambda: sc.closest_point(q, stream=stream)), ("silhouette", lambda: sc.closest_silhouette(q, stream=stream)),
                      ("ray", lambda: sc.intersect(q, d, stream=stream))):
            fn()
            sc.counter("query.traversal_ms", reset=True)
            for _ in range(10):
                fn()
            out[f"{k}_kernel_ms"] = sc.counter("query.traversal_ms", reset=True) / 10
    sc.set_option("query.time_kernels", 0)
    out["note"] = "x_ms = whole call (ordering + kernels + per-call host overhead, back to back); x_kernel_ms = traversal kernel alone"
    try:
        from oracle import RefScene, ref_available
        if ref_available("cuda"):
            ref = RefScene(v, f, "cuda")
            rc = {}
            for name, fn in (("closest_ms", lambda: ref.closest(q_h)), ("silhouette_ms", lambda: ref.silhouette(q_h)), ("ray_ms", lambda: ref.ray(q_h, d_h))):
                fn()
                fn()
                rc[name] = ref.last_ms  # the traversal kernel alone (CUDA events inside the wrapper)
            out["reference_cuda"] = rc
    except Exception as ex:
        out["reference_cuda_error"] = repr(ex)
    del sc
    return out


def scene2_config(pkg, m, dev, stream):
    """2-D path (lbvh::scene<2>) at two polyline resolutions: 8 192 segments (segment length ~ the
    reference's 1e-3 absolute leaf padding: the regime the 2-D scenes of a walk-on-stars solver live in) and 1 048 576 segments
    (segments 150x shorter than the padding: ~300 leaf boxes overlap everywhere, for either implementation)."""
    return {"8192": scene2_case(pkg, m, dev, stream, 1 << 13), "1048576": scene2_case(pkg, m, dev, stream, 1 << 20)}


def scene2_case(pkg, m, dev, stream, n_segments):
    """A closed wavy polyline, 4 194 304 queries, device-resident, CUDA events on `stream`; the reference's own CUDA 2-D path
    (unmodified headers, one thread per query) beside it."""
    import torch
    v, s = m.wavy_circle(n_segments, 37, 0.2)
    n = 1 << 22
    t0 = time.perf_counter()
    sc = pkg.Scene2(v, s, device=dev).compute_silhouettes().build_bvh(stream=stream)
    out = {"segments": len(s), "queries": n, "scene_setup_wall_ms": (time.perf_counter() - t0) * 1e3, "build_ms": sc.stats()["build_ms"]}
    q_h = m.points_in_box2(n, v.min(0), v.max(0), 1.1, seed=41)
    d_h = m.unit_directions2(n, seed=42)
    q, d = torch.from_numpy(q_h).to(f"cuda:{dev}"), torch.from_numpy(d_h).to(f"cuda:{dev}")

    def timed(fn, reps=5):
        fn()
        stream.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        stream.synchronize()
        return a.elapsed_time(b) / reps

    with torch.cuda.stream(stream):
        _, dist = sc.closest_point(q, stream=stream)
        rmax = (dist * 2.0).contiguous()
        out["closest_mqps"] = n / timed(lambda: sc.closest_point(q, stream=stream)) / 1e3
        out["silhouette_unbounded_mqps"] = n / timed(lambda: sc.closest_silhouette(q, stream=stream)) / 1e3
        out["silhouette_star_radius_mqps"] = n / timed(lambda: sc.closest_silhouette(q, r_max=rmax, stream=stream)) / 1e3
        out["ray_mqps"] = n / timed(lambda: sc.intersect(q, d, stream=stream)) / 1e3
        circ = torch.cat([q, rmax[:, None]], dim=1).contiguous()
        rnd = torch.from_numpy(m.uniforms(n, 2, seed=43)).to(f"cuda:{dev}")
        out["sample_in_circle_mqps"] = n / timed(lambda: sc.sample_in_sphere(circ, rnd, stream=stream)) / 1e3
        builds = []
        for _ in range(3):
            sc.build_bvh(stream=stream)
            builds.append(sc.stats()["build_ms"])
        out["build_ms"] = min(builds)
    try:
        from oracle import RefScene2, ref_available
        if ref_available("cuda"):
            ns = 1 << 20
            t0 = time.perf_counter()
            ref = RefScene2(v, s, "cuda")
            rc = {"scene_setup_wall_ms": (time.perf_counter() - t0) * 1e3, "sample": f"first {ns} of the queries"}
            for name, fn in (("closest_mqps", lambda: ref.L.ref2_closest(ref.h, q_h[:ns], ns, np.zeros(ns, np.uint32), np.zeros(ns, np.float32), 1)),
                             ("silhouette_unbounded_mqps", lambda: ref.L.ref2_silhouette(ref.h, q_h[:ns], ns, 0, np.zeros(ns, np.float32), 1)),
                             ("ray_mqps", lambda: ref.L.ref2_ray(ref.h, q_h[:ns], d_h[:ns], np.full(ns, np.inf, np.float32), ns, np.zeros(ns, np.int32),
                                                                 np.zeros(ns, np.float32), np.zeros(ns, np.float32), np.zeros(ns, np.uint32), 1))):
                fn()
                rc[name] = ns / fn() / 1e3  # the wrappers return the traversal kernel's milliseconds (CUDA events)
            out["reference_cuda"] = rc
    except Exception as ex:
        out["reference_cuda_error"] = repr(ex)
    del sc
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------------------------
def main():
    """Everything but the result line is kept off stdout (NCCL prints its version banner there, the reference prints its
    collision notice): fd 1 points at stderr until the line is ready."""
    with quiet_stdout():
        line = _run()
    if line is not None:
        print(json.dumps(line), flush=True)


def _run():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--queries", type=int, default=N_QUERIES, help="queries per GPU per step (default: the C3 batch)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary measurements (closest/ray/build/reference CUDA)")
    ap.add_argument("--also", default="", help="comma list of further configs to time into `extra` on rank 0 (c1, c4, c5 of BASELINE.json; 2d = the scene<2> path): "
                    "c4 (16M rays / 8 GPUs on a 4M-triangle mesh), c5 (wavefront WoSt step, 64M walkers / 8 GPUs on a 10M-triangle mesh)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import snch_lbvh_b200 as pkg
    from snch_lbvh_b200 import distributed as sd  # noqa: F401  (import through the shim)
    m = pkg.meshes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: snch-lbvh_b200 has no CPU fallback")
    dev = local_rank
    torch.cuda.set_device(dev)
    n = args.queries

    # ---- scene: built on rank 0, replicated by broadcasting the arena -----------------------------------------------
    v, f, q_h, s_h = make_inputs(m, n, seed_shift=rank)
    t0 = time.perf_counter()
    scene = None
    if rank == 0:
        scene = pkg.Scene3(v, f, device=dev).compute_silhouettes().build_bvh()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    scene = sd.replicate_scene(scene, rank, world, dev, dist)
    torch.cuda.synchronize()
    replicate_ms = (time.perf_counter() - t1) * 1e3
    stats = scene.stats()
    for kv in filter(None, os.environ.get("SNCH_OPTIONS", "").replace(",", " ").split()):  # diagnostic runs only (tools/gpu_exp.sh)
        k, val = kv.split("=")
        scene.set_option(k, int(val))

    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        q_d = torch.from_numpy(q_h).to(f"cuda:{dev}")
        _, dcp = scene.closest_point(q_d, stream=stream)  # WoSt: the star radius derives from the closest-point distance
        rmax_d = (dcp * torch.from_numpy(s_h).to(f"cuda:{dev}")).contiguous()
        out_d = torch.empty(n, dtype=torch.float32, device=f"cuda:{dev}")
        L = pkg.lib()

        def step_device():
            st = L.snch_closest_silhouette_batch(scene._h, q_d.data_ptr(), None, rmax_d.data_ptr(), n, out_d.data_ptr(), None, None, stream.cuda_stream)
            assert st == 0, L.snch_last_error()

        def barrier():
            stream.synchronize()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        for _ in range(args.warmup):
            step_device()
        barrier()
        scene.set_option("query.time_kernels", 1)  # CUDA events around the traversal kernel itself, on `stream` (roofline.achieved)
        scene.counter("query.launches", reset=True)
        sampler = ClockSampler(dev)
        if rank == 0:
            sampler.start()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        evs[0].record(stream)
        for k in range(args.steps):
            step_device()
            evs[k + 1].record(stream)
        barrier()
        launches = int(scene.counter("query.launches"))
        trav_launches = int(scene.counter("query.traversal_launches"))
        trav_ms = scene.counter("query.traversal_ms", reset=True)
        scene.set_option("query.time_kernels", 0)
        clocks = sampler.stop() if rank == 0 else None
        step_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
        total_ms = evs[0].elapsed_time(evs[-1])
        tt = torch.tensor([total_ms], dtype=torch.float64, device=f"cuda:{dev}")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms_max = float(tt.item())
        finite_frac = float(torch.isfinite(out_d).float().mean().item())

        # ---- e2e: HOST buffers through the C-ABI (H2D + kernel + D2H every step) ----------------------------------------
        q_p = torch.from_numpy(q_h).pin_memory()
        r_p = rmax_d.cpu().pin_memory()
        o_p = torch.empty(n, dtype=torch.float32).pin_memory()

        def step_host():
            st = L.snch_closest_silhouette_batch(scene._h, q_p.data_ptr(), None, r_p.data_ptr(), n, o_p.data_ptr(), None, None, stream.cuda_stream)
            assert st == 0, L.snch_last_error()

        for _ in range(2):
            step_host()
        barrier()
        e2e_steps = max(3, min(args.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host()  # synchronises the stream before returning (results are on the host)
        e1.record(stream)
        barrier()
        e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3 * 0.0)  # device clock; host wall only as a floor guard
        et = torch.tensor([e2e_ms], dtype=torch.float64, device=f"cuda:{dev}")
        if dist is not None:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        e2e_ms_max = float(et.item())
        assert np.array_equal(o_p.numpy().view(np.uint32), out_d.cpu().numpy().view(np.uint32)), "host path != device path"

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return None

    # ---- rank 0: roofline, secondary numbers, CPU baseline ------------------------------------------------------------
    ms_per_step = total_ms_max / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6
    peak, peak_src = load_peaks()
    mv = None
    mvp = os.path.join(ROOT, "profiles", "must_visit.json")
    if os.path.exists(mvp):
        mv = json.load(open(mvp))["configs"].get("torus708")
    roofline = None
    if mv is not None and n == N_QUERIES:
        V, Lv = mv["silhouette_star_radius"]["V"], mv["silhouette_star_radius"]["L"]
        io_q = 12 + 4 + 4                       # point + r_max in, distance out (flip omitted: NULL)
        b_q = io_q + 64.0 * V + 192.0 * Lv      # B_q = IO_q + 64*V* + P*L*, P = 3 edges x 64 B
        kernel_ms = trav_ms / max(trav_launches, 1)  # the traversal kernel alone (CUDA events on its stream, inside the timed region)
        achieved = b_q * n / (kernel_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "kernel": "snch::k_silhouette_coop<3, 0, 0, 1>", "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / statistics.mean(step_ms),
                    "algorithmic_bytes_per_query": b_q, "must_visit_internal": V, "must_visit_leaves": Lv,
                    "peak_source": peak_src,
                    "note": "divergent gather over SNode 96 MB + LEdge 96 MB; the kernel is issue/L1-bound, not DRAM-bound: traffic << algorithmic "
                            "bytes because records are re-read from L1/L2 (hit rates in profiles/)"}
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            roofline["traffic"] = tj.get("k_silhouette_coop_bytes_per_launch", tj.get("k_silhouette_bytes_per_launch"))

    extra = {"build_ms": stats["build_ms"], "adjacency_ms": stats["adjacency_ms"], "arena_bytes": stats["arena_bytes"],
             "replicate_ms": replicate_ms if world > 1 else 0.0, "scene_setup_wall_ms": (t1 - t0) * 1e3,
             "finite_fraction": finite_frac, "step_ms_min": min(step_ms), "step_ms_max": max(step_ms)}
    cpu_base = None
    if not args.no_extra:
        with torch.cuda.stream(stream):
            def timed(fn, reps=5):
                fn()
                stream.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(reps):
                    fn()
                b.record(stream)
                stream.synchronize()
                return a.elapsed_time(b) / reps
            d_d = torch.from_numpy(m.unit_directions(n, seed=77)).to(f"cuda:{dev}")
            extra["closest_mqps"] = n / timed(lambda: scene.closest_point(q_d, stream=stream)) / 1e3
            extra["silhouette_unbounded_mqps"] = n / timed(lambda: scene.closest_silhouette(q_d, stream=stream)) / 1e3
            extra["ray_mqps"] = n / timed(lambda: scene.intersect(q_d, d_d, stream=stream)) / 1e3
            sph = torch.cat([q_d, rmax_d[:, None]], dim=1).contiguous()
            rnd = torch.from_numpy(m.uniforms(n, 3, seed=99)).to(f"cuda:{dev}")
            extra["sample_in_sphere_mqps"] = n / timed(lambda: scene.sample_in_sphere(sph, rnd, stream=stream)) / 1e3
            builds = []
            for _ in range(5):
                scene0 = scene if world == 1 else None
                if scene0 is None:
                    break
                scene0.build_bvh(stream=stream)
                builds.append(scene0.stats()["build_ms"])
            if builds:
                extra["build_ms"] = min(builds)
                extra["build_roofline_frac"] = (334.0 * stats["num_objects"] / (min(builds) * 1e-3) / 1e9) / peak
        for cfg in [c for c in args.also.split(",") if c]:
            try:
                extra[cfg] = other_config(cfg, pkg, m, dev, stream)
            except Exception as ex:
                extra[cfg] = {"error": repr(ex)}
        # the reference's own CUDA path on this B200 (prebuilt from the unmodified headers; absent -> skipped)
        try:
            from oracle import RefScene, ref_available
            if ref_available("cuda"):
                with quiet_stdout():
                    ns = 1 << 22
                    ref = RefScene(v, f, "cuda")
                    ref.time_construct(3)
                    rc = {"construct_ms": ref.timings()["construct_ms"], "build_bvh_ms_incl_host": ref.timings()["build_bvh_ms"],
                          "compute_silhouettes_host_ms": ref.timings()["silhouettes_ms"], "sample": f"first {ns} of the C3 queries"}
                    d_h = m.unit_directions(ns, seed=77)
                    ref.silhouette(q_h[:ns])
                    ref.silhouette(q_h[:ns])
                    rc["silhouette_unbounded_mqps"] = ns / ref.last_ms / 1e3
                    ref.closest(q_h[:ns])
                    ref.closest(q_h[:ns])
                    rc["closest_mqps"] = ns / ref.last_ms / 1e3
                    ref.ray(q_h[:ns], d_h)
                    ref.ray(q_h[:ns], d_h)
                    rc["ray_mqps"] = ns / ref.last_ms / 1e3
                    rc["note"] = ("the reference has no r_max input (SURVEY Q5): its answer to the C3 workload is the unbounded "
                                  "query followed by a filter, i.e. silhouette_unbounded_mqps is its C3 throughput")
                    extra["reference_cuda"] = rc
                    extra["speedup_vs_reference_cuda_c3"] = value / world / rc["silhouette_unbounded_mqps"]
        except Exception as ex:  # baseline legs never break the bench line
            extra["reference_cuda_error"] = repr(ex)
        try:
            cpu_base = cpu_silhouette_baseline(v, f, q_h, r_p.numpy())
        except Exception as ex:
            cpu_base = {"value": None, "unit": "M queries/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex!r}"}

    line = {"metric": "M queries/sec (nearest-silhouette, star radii) @1M tris", "value": value, "unit": "M queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "queries_per_gpu_per_step": n, "triangles": stats["num_objects"],
                       "parallelism": f"replicated tree, query batch sharded x{world}",
                       "l2": "inputs larger than L2 (268 MB of queries + 192 MB of tree records + 201 MB of ordering buffers per step vs 126 MB L2); no explicit flush",
                       "step": "Morton ordering of the batch (bounds, keys, 3-pass onesweep radix sort) + persistent traversal kernel",
                       "e2e": "snch_closest_silhouette_batch on pinned HOST buffers: H2D, kernels and D2H inside the call, pipelined in chunks "
                              "of query.host_chunk = 8388608 queries over a copy-in stream, two compute streams and a copy-out stream"},
            "e2e": {"value": world * n * e2e_steps / (e2e_ms_max * 1e-3) / 1e6, "unit": "M queries/s", "h2d_bytes_per_step": n * 16,
                    "d2h_bytes_per_step": n * 4, "ms_per_step": e2e_ms_max / e2e_steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_base, "extra": extra}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return line


if __name__ == "__main__":
    main()
