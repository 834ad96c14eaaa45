#!/usr/bin/env python3
"""bench.py -- LM edges/s of the B200-native bundle-adjustment engine (metric of BASELINE.json).

One "step" = one pass of the hot path over one synthetic graph: `optimize(10)` (10 LM iterations) on a
KITTI-00-shaped graph (1 322 poses / 133 383 landmarks / 131 233 mono + 429 883 stereo edges, fp64).

  value  : LM edge-iterations/s with every input already resident in HBM (state reset by a device copy,
           L2 flushed before each step), CUDA-event timed on the engine's stream, max over ranks.
  e2e    : the same metric through the reference-facing C ABI with HOST buffers: set_problem (H2D copies +
           structure build) + optimize(10) + get_state (D2H) per step -- the reference's own timed window
           `initialize(); optimize(10)` (samples/sample_ba_from_file.cpp:52-57).
  roofline: the Jacobian+Hessian landmark-pass kernel (the HBM-dominant kernel), algorithmic bytes of
           SURVEY.md 8(d) / CUDA-event time per launch, against MEASURED_PEAKS.json.
  cpu_baseline: the CPU oracle (g2o-equivalent restatement, oracle/ba_oracle.c) on the same graph.

`--impl reference` times the UNMODIFIED reference (compiled for sm_100 into oracle/_ref/libcuba_ref.so by
oracle/build_ref.sh) through its own public API on the same graph -- the reference is a GPU library, so
its arm runs on the GPU; when the library is missing the arm falls back to the CPU oracle port.

Launch: python bench.py [--gpus N --steps K --warmup W]; for N>1 under torch.distributed.run.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

KERNELS = {"none": ((0, 0), (0.0, 0.0)), "huber": ((1, 1), (5.991 ** 0.5, 7.815 ** 0.5)), "tukey": ((2, 2), (4.0, 5.0))}
LM_ITERS = 10


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def jh_bytes(sz, s=8):
    """algorithmic bytes of the Jacobian+Hessian stage.  Returns (landmark-pass kernel, whole stage)."""
    E2, E3, nhpl, Pall, Lall, P, L = sz["E2"], sz["E3"], sz["nhpl"], sz["Pall"], sz["Lall"], sz["numP"], sz["numL"]
    common = E2 * (3 * s + 8) + E3 * (4 * s + 8) + nhpl * 18 * s + Pall * 12 * s + Lall * 3 * s + L * 12 * s
    return common, common + P * 42 * s


def build_problem(pkg, workload):
    g = pkg.synth.make_config(workload)
    return pkg.graphio.flatten(g)


def run_reference(args, prob, rk, rank):
    """--impl reference: the compiled reference (GPU build) through its own API; CPU oracle port if absent."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reference
    E = prob.nedges
    use_ref = reference.available(args.fp32)
    times, iters = [], LM_ITERS
    kind, sample = "reference", ""
    import torch
    if use_ref and torch.cuda.is_available():
        for i in range(args.warmup + args.steps):
            r = reference.run(prob, LM_ITERS, rk[0], rk[1], warmup=0, fp32=args.fp32)
            if r is None:
                use_ref = False
                break
            if i >= args.warmup:
                times.append(r["seconds"]); iters = len(r["chi2"])
        sample = "unmodified reference compiled for sm_100 (oracle/_ref/libcuba_ref.so), initialize()+optimize(10) on the full graph, " \
                 "host buffers; runs on the GPU because the reference has no CPU path (its CPU comparator g2o is not in the image)"
    if not use_ref or not times:
        kind = "port"
        oracle = ge.load_oracle()
        for i in range(min(args.warmup, 1) + min(args.steps, 2)):
            o = oracle.Oracle(prob, rk[0], rk[1])
            t0 = time.perf_counter(); chi, lam, tr = o.optimize(LM_ITERS); dt = time.perf_counter() - t0
            if i >= min(args.warmup, 1):
                times.append(dt); iters = len(chi)
        sample = "CPU oracle port (oracle/ba_oracle.c), optimize(10) on the full graph"
    sec = float(np.mean(times))
    val = E * iters / sec
    cores = 1 if kind == "port" else 0
    line = {"impl": "reference", "metric": "LM edges/sec (10 iters)", "value": val, "unit": "edge-iterations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if args.fp32 else "f64", "data": "synthetic",
            "config": {"workload": args.workload, "robust_kernel": args.robust, "lm_iterations": LM_ITERS, "edges": E},
            "cpu_baseline": {"value": val, "unit": "edge-iterations/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "edge-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="kitti00_shaped")
    ap.add_argument("--robust", default="none", choices=list(KERNELS))
    ap.add_argument("--fp32", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    rk = KERNELS[args.robust]
    pkg = ge.load_package()

    if args.impl == "reference":
        if rank != 0:
            return 0      # rank 0 alone runs and prints the reference arm
        run_reference(args, build_problem(pkg, args.workload), rk, rank)
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    prob = build_problem(pkg, args.workload)
    E = prob.nedges
    # the e2e leg copies its inputs from PINNED host memory (bench contract): page-lock the flat problem once
    import dataclasses
    def _pin(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() if isinstance(a, np.ndarray) and a.size else a
    prob = dataclasses.replace(prob, **{f.name: _pin(getattr(prob, f.name)) for f in dataclasses.fields(prob)
                                        if f.name in ("q", "t", "cam", "Xw", "idx2", "meas2", "omega2", "idx3", "meas3", "omega3")})
    eng = pkg.Engine(device=local, use_fp32=args.fp32)
    for et in (0, 1):
        eng.set_robust_kernels(rk[0][et], rk[1][et], et)
    if world > 1:
        eng.set_comm(rank, world, pkg.sharding.broadcast_unique_id(pkg.Engine, rank, world))
    stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- e2e: host buffers -> set_problem (H2D + structure) -> optimize -> get_state (D2H)
    e2e_ms, e2e_iters = [], LM_ITERS
    h2d0 = d2h0 = 0
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            h2d0, d2h0 = pkg.transfer_bytes()
        barrier()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(stream)
        eng.initialize(prob)
        stats = eng.optimize(LM_ITERS)
        q, t, Xw = eng.state()
        b.record(stream)
        barrier()
        if i >= args.warmup:
            e2e_ms.append(max_over_ranks(a.elapsed_time(b))); e2e_iters = len(stats)
    h2d1, d2h1 = pkg.transfer_bytes()
    prof = eng.time_profile()     # buckets of exactly one e2e step (set_problem resets them)
    e2e_t = float(np.mean(e2e_ms)) * 1e-3
    sizes = eng.sizes

    # ---------------- value: inputs resident in HBM, L2 flushed, device-timed optimize(10)
    clocks = ClockSampler(local); clocks.start()
    dev_ms, iters_done, launches0, pcg_total = [], LM_ITERS, 0, 0
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            launches0 = eng.launch_count()
        eng.reset_state(); eng.flush_l2()
        barrier()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(stream)
        stats = eng.optimize(LM_ITERS)
        b.record(stream)
        barrier()
        if i >= args.warmup:
            dev_ms.append(max_over_ranks(a.elapsed_time(b))); iters_done = len(stats)
            pcg_total = sum(s["pcg_iters"] for s in stats)
    launches = eng.launch_count() - launches0
    clocks.stop_flag.set(); clocks.join(timeout=3)
    ms_per_step = float(np.mean(dev_ms))
    value = E * iters_done / (ms_per_step * 1e-3)
    final_chi2 = stats[-1]["chi2"]

    # ---------------- roofline of the J+H landmark-pass kernel + per-stage device times (L2 flushed between reps)
    eng.reset_state(); eng.linearize()
    lam = 1e-5 * eng.max_diagonal()
    stage_ms = {}
    for label, st in (("jh_landmark_pass", 1), ("jh_pose_pass", 2), ("schur", 3), ("pcg_solve", 4), ("backsub_update_chi2", 5), ("chi2_only", 6)):
        stage_ms[label] = eng.bench_stage(st, reps=20, flush_l2=True, lam=lam)
    s = 4 if args.fp32 else 8
    # per-rank bytes: each rank streams its shard of the edges
    b_kernel, b_stage = jh_bytes(sizes, s)
    b_kernel /= world; b_stage /= world
    peak, peak_src = measured_peak()
    ach = b_kernel / (stage_ms["jh_landmark_pass"] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "jh_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": ("k_linearize_landmark<float>" if args.fp32 else "k_linearize_landmark4"), "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes": b_kernel, "ms_per_launch": stage_ms["jh_landmark_pass"],
                "stage_frac_jh_both_kernels": b_stage / ((stage_ms["jh_landmark_pass"] + stage_ms["jh_pose_pass"]) * 1e-3) / 1e9 / peak}

    # ---------------- CPU baseline: the oracle port on this box's host cores (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        oracle = ge.load_oracle()
        o = oracle.Oracle(prob, rk[0], rk[1])
        t0 = time.perf_counter(); chi, lam_o, tr = o.optimize(LM_ITERS); dt = time.perf_counter() - t0
        cpu = {"value": E * len(chi) / dt, "unit": "edge-iterations/s", "cores": 1, "kind": "port",
               "sample": "oracle/ba_oracle.c (g2o-equivalent: Schur + sparse block Cholesky + LM), one optimize(10) on the full %s graph, %.1f s; "
                         "host has %d cores, the port is single-threaded like g2o's default" % (args.workload, dt, os.cpu_count()),
               "final_chi2": float(chi[-1]), "chi2_rel_diff_vs_gpu": float(abs(chi[-1] - final_chi2) / chi[-1])}

    line = {"metric": "LM edges/sec (10 iters)", "value": value, "unit": "edge-iterations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if args.fp32 else "f64", "data": "synthetic",
            "config": {"workload": args.workload, "poses": sizes["Pall"], "landmarks": sizes["Lall"], "edges": E, "robust_kernel": args.robust,
                       "lm_iterations": iters_done, "l2": "flushed between timed steps (320 MB fill)", "parallelism": "landmark-sharded x%d" % world,
                       "pcg_iterations_per_step": pcg_total, "final_chi2": final_chi2},
            "e2e": {"value": E * e2e_iters / e2e_t, "unit": "edge-iterations/s", "ms_per_step": 1e3 * e2e_t,
                    "h2d_bytes_per_step": (h2d1 - h2d0) // max(args.steps, 1),
                    "d2h_bytes_per_step": (d2h1 - d2h0) // max(args.steps, 1),
                    "window": "set_problem(H2D from pinned host buffers + structure build) + optimize(10) + get_state(D2H) == reference's initialize()+optimize(10)"},
            "gpu_launches": launches, "clocks": clocks.summary(), "roofline": roofline, "stage_ms": stage_ms,
            "profile_ms_e2e_step": {k: round(1e3 * v, 4) for k, v in prof.items()}}
    if cpu:
        line["cpu_baseline"] = cpu
    if rank == 0:
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
