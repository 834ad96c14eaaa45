#!/usr/bin/env python3
"""Side-by-side comparison report, the equivalent of the reference's samples/sample_comparison_with_g2o.cpp:62-136 with the
roles "CPU (g2o)" -> the CPU oracle (oracle/ba_oracle.c, the g2o-equivalent restatement), "GPU" -> this engine driven through
the drop-in C++ class (samples/sample_ba_from_file), plus a third column: the UNMODIFIED reference GPU build
(oracle/_ref/libcuba_ref.so).  Lives under tests/ because it executes oracle/ (test infrastructure).

  python tests/compare_with_reference.py [ba_kitti_07|ba_kitti_00|small|kitti07_shaped|...] [--kernel none|huber] [--no-cpu]

Protocol = the reference samples': all three start from the same file estimate, initialize() + optimize(10); table of chi2 per
iteration, processing times, RMSE of q / t / Xw between the estimates (README.md:177-191 reports <= ~1e-12 vs g2o)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

KERNELS = {"none": ((0, 0), (0.0, 0.0)), "huber": ((1, 1), (5.991 ** 0.5, 7.815 ** 0.5))}


def build_sample(pkg, outdir):
    out = os.path.join(outdir, "sample_ba_from_file")
    libdir = os.path.dirname(pkg.library_path())
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-DCUBA_FORCE_EIGEN_COMPAT", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "samples", "sample_ba_from_file.cpp"), "-L", libdir, "-lcuba_b200",
                           "-Wl,-rpath," + libdir, "-o", out])
    return out


def rmse(a, b):
    d = np.asarray(a) - np.asarray(b)
    return float(np.sqrt((d * d).sum() / max(len(d), 1)))


def compare(workload, kernel="none", with_cpu=True, iters=10, out=sys.stdout):
    """returns dict(chi2_ours, chi2_ref, chi2_cpu, rmse_ref, rmse_cpu, seconds...) and prints the report"""
    pkg = ge.load_package()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reference
    path = os.path.join(ROOT, "oracle", "_ref", "fixtures", workload + ".cubagraph")
    tmp = tempfile.mkdtemp(prefix="cuba_cmp_")
    if workload.startswith("ba_"):
        g = pkg.graphio.read_graph(path)
    else:
        g = pkg.synth.make_config(workload)
        path = os.path.join(tmp, workload + ".cubagraph")
        pkg.graphio.write_graph(path, g)
    prob = pkg.graphio.flatten(g)
    rk = KERNELS[kernel]
    exe = build_sample(pkg, tmp)
    dump = os.path.join(tmp, "state.bin")
    res = subprocess.run([exe, path, "--json", "--no-warmup", "--iters", str(iters), "--dump", dump] + (["--huber"] if kernel == "huber" else []),
                         capture_output=True, text=True, timeout=1200)
    if res.returncode != 0:
        raise RuntimeError("sample failed: " + res.stderr)
    ours = json.loads(res.stdout)
    nP, nL = len(g["pose_id"]), len(g["lm_id"])
    raw = np.fromfile(dump, dtype=np.float64)
    q = raw[:4 * nP].reshape(nP, 4)[prob.pose_rows]; t = raw[4 * nP:7 * nP].reshape(nP, 3)[prob.pose_rows]; Xw = raw[7 * nP:].reshape(nL, 3)[prob.lm_rows]
    r = reference.run(prob, iters, rk[0], rk[1]) if reference.available() else None
    cpu = None
    if with_cpu:
        oracle = ge.load_oracle()
        o = oracle.Oracle(prob, *rk)
        t0 = time.perf_counter(); chi, lam, tr = o.optimize(iters); dt = time.perf_counter() - t0
        oq, ot, oX = o.state()
        cpu = dict(chi2=np.array(chi), q=oq, t=ot, Xw=oX, seconds=dt)
    w = out.write
    w("=== Graph size : \nnum poses      : %d\nnum landmarks  : %d\nnum edges      : %d\n\n" % (ours["nposes"], ours["nlandmarks"], ours["nedges"]))
    w("=== Processing time (initialize() + optimize(%d)) : \n" % iters)
    if cpu:
        w("CPU oracle (1 thread)        : %9.4f [sec]\n" % cpu["seconds"])
    if r:
        w("reference GPU build          : %9.4f [sec]\n" % r["seconds"])
    w("this engine (drop-in class)  : %9.4f [sec]\n\n" % ours["seconds"])
    w("=== Objective function value : \n%10s|%16s|%16s|%16s\n" % ("iteration", "chi2 CPU", "chi2 reference", "chi2 this engine"))
    n = max(len(ours["chi2"]), len(r["chi2"]) if r else 0, len(cpu["chi2"]) if cpu else 0)
    cell = lambda a, i: ("%16.1f" % a[i]) if a is not None and i < len(a) else "%16s" % "N/A"
    for i in range(n):
        w("%10d|%s|%s|%s\n" % (i + 1, cell(cpu["chi2"] if cpu else None, i), cell(r["chi2"] if r else None, i), cell(ours["chi2"], i)))
    result = dict(chi2_ours=np.array(ours["chi2"]), seconds_ours=ours["seconds"])
    for name, other in (("reference GPU build", r), ("CPU oracle", cpu)):
        if other is None:
            continue
        e = dict(Rotation=rmse(q, other["q"]), Translation=rmse(t, other["t"]), Landmark=rmse(Xw, other["Xw"]))
        w("\n=== RMSE between the %s's estimates and this engine's : \n" % name)
        for k, v in e.items():
            w("%-12s: %.2e\n" % (k, v))
        m = min(len(other["chi2"]), len(ours["chi2"]))
        rel = float(np.abs(np.array(ours["chi2"][:m]) - other["chi2"][:m]).max() / np.abs(other["chi2"][:m]).max()) if m else 0.0
        w("max relative chi2 difference : %.2e\n" % rel)
        key = "ref" if other is r else "cpu"
        result["rmse_" + key] = e; result["chi2_" + key] = np.array(other["chi2"]); result["chi2_rel_" + key] = rel
        result["seconds_" + key] = other["seconds"]
    return result


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default="ba_kitti_07")
    ap.add_argument("--kernel", default="none", choices=list(KERNELS))
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    ge.build()
    compare(a.workload, a.kernel, not a.no_cpu)
