"""Stage-by-stage parity of the CUDA engine against the CPU oracle (and the compiled reference when
present).  Run on a GPU box: python tools/gpu_diag.py [config ...]"""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
oracle = ge.load_oracle()
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reference  # noqa: E402


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    d = np.abs(a - b).max()
    return float(d / max(np.abs(b).max(), 1e-300))


def load(name):
    if name.startswith("ba_"):
        return pkg.graphio.read_graph(os.path.join(ROOT, "oracle", "_ref", "fixtures", name + ".cubagraph"))
    return pkg.synth.make_config(name)


def stages(name, rk, delta, fp32=False):
    print("=== %s  kernels=%s fp32=%s" % (name, rk, fp32), flush=True)
    prob = pkg.graphio.flatten(load(name))
    eng = pkg.Engine(device=0, use_fp32=fp32)
    for et in (0, 1):
        eng.set_robust_kernels(rk[et], delta[et], et)
    t0 = time.time(); sz = eng.initialize(prob); print("initialize %.1f ms" % (1e3 * (time.time() - t0)), sz, flush=True)
    o = oracle.Oracle(prob, rk, delta)
    cp, ri, e2h = eng.hpl_structure(); ocp, ori, oe2h = o.hpl_structure()
    rp, ci = eng.hsc_structure(); orp, oci = o.hsc_structure()
    print("structure bit-exact:", np.array_equal(cp, ocp), np.array_equal(ri, ori), np.array_equal(e2h, oe2h), np.array_equal(rp, orp), np.array_equal(ci, oci))
    chi = eng.linearize(); ochi = o.compute_errors(); o.build_system()
    print("chi2 %.12g oracle %.12g rel %.2e" % (chi, ochi, abs(chi - ochi) / ochi))
    names = ("Hpp", "bp", "Hll", "bl", "Hpl")
    for n, a, b in zip(names, eng.system(), o.system()):
        print("  %-4s rel %.2e" % (n, rel(a, b)))
    md, omd = eng.max_diagonal(), o.max_diagonal()
    print("maxdiag %.12g oracle %.12g" % (md, omd))
    lam = 1e-5 * omd
    it, ok = eng.solve(lam); ook = o.solve(lam)
    print("solve: pcg iters %d ok %s (oracle ok %s)" % (it, ok, ook))
    for n, a, b in zip(("Hsc", "bsc", "invHll"), eng.schur(), o.schur()):
        print("  %-6s rel %.2e" % (n, rel(a, b)))
    for n, a, b in zip(("xp", "xl"), eng.delta(), o.delta()):
        print("  %-6s rel %.2e" % (n, rel(a, b)))
    fh, sc = eng.update(lam); o.update(); ofh = o.compute_errors(); osc = o.compute_scale(lam)
    print("trial chi2 %.12g oracle %.12g rel %.2e | scale %.12g oracle %.12g rel %.2e" % (fh, ofh, abs(fh - ofh) / ofh, sc, osc, abs(sc - osc) / abs(osc)))
    eng.commit(True)
    for n, a, b in zip(("q", "t", "Xw"), eng.state(), o.state()):
        print("  state %-3s rel %.2e" % (n, rel(a, b)))
    print("  per-edge chi2 rel %.2e" % rel(eng.chi_squared(), o.chi_sqs()))
    eng.close()


def full(name, rk, delta, niter=10, fp32=False, use_ref=True):
    print("=== full optimize %s kernels=%s fp32=%s" % (name, rk, fp32), flush=True)
    prob = pkg.graphio.flatten(load(name))
    eng = pkg.Engine(device=0, use_fp32=fp32)
    for et in (0, 1):
        eng.set_robust_kernels(rk[et], delta[et], et)
    eng.initialize(prob); eng.optimize(1)     # warm-up on the same engine
    t0 = time.time()
    eng.initialize(prob)
    stats = eng.optimize(niter)
    dt = time.time() - t0
    prof = eng.time_profile()
    o = oracle.Oracle(prob, rk, delta)
    t1 = time.time(); chi, lam, tr = o.optimize(niter); odt = time.time() - t1
    got = np.array([s["chi2"] for s in stats])
    print("engine %.1f ms (%.3g edge-iters/s), oracle CPU %.1f ms" % (1e3 * dt, prob.nedges * len(got) / dt, 1e3 * odt))
    for i, s in enumerate(stats):
        oc = chi[i] if i < len(chi) else float("nan")
        print("  it %2d chi2 %.10f oracle %.10f rel %.1e trials %d/%d lambda %.6g/%.6g pcg %d fail %d" % (
            i, s["chi2"], oc, abs(s["chi2"] - oc) / oc, s["trials"], tr[i] if i < len(tr) else -1, s["lambda_"], lam[i] if i < len(lam) else -1,
            s["pcg_iters"], s["pcg_failed"]))
    for n, a, b in zip(("q", "t", "Xw"), eng.state(), o.state()):
        print("  final %-3s rel %.2e  maxabs %.2e" % (n, rel(a, b), np.abs(np.asarray(a) - np.asarray(b)).max()))
    print("  profile ms:", {k: round(1e3 * v, 3) for k, v in prof.items()})
    if use_ref and reference.available(fp32):
        r = reference.run(prob, niter, rk, delta, warmup=1, fp32=fp32)
        if r is not None:
            n = min(len(r["chi2"]), len(got))
            print("  reference GPU build: %.1f ms (%.3g edge-iters/s); chi2 rel vs engine %.2e, vs oracle %.2e" % (
                1e3 * r["seconds"], prob.nedges * len(r["chi2"]) / r["seconds"], rel(got[:n], r["chi2"][:n]), rel(chi[:n], r["chi2"][:n])))
            print("  reference profile ms:", [round(1e3 * v, 2) for v in r["profile"]])
            for nme, a, b in zip(("q", "t", "Xw"), eng.state(), (r["q"], r["t"], r["Xw"])):
                print("  final %-3s engine vs reference rel %.2e" % (nme, rel(a, b)))
    eng.close()


if __name__ == "__main__":
    cfgs = sys.argv[1:] or ["tiny", "small"]
    NONE = ((0, 0), (0.0, 0.0)); HUBER = ((1, 1), (5.991 ** 0.5, 7.815 ** 0.5)); TUKEY = ((2, 2), (4.0, 5.0))
    for c in cfgs:
        for rk, d in (NONE, HUBER):
            try:
                stages(c, rk, d)
            except Exception:
                traceback.print_exc()
        for rk, d in (NONE, HUBER):
            try:
                full(c, rk, d)
            except Exception:
                traceback.print_exc()
    try:
        stages(cfgs[0], *TUKEY)
        stages(cfgs[0], *HUBER, fp32=True)
        full(cfgs[-1], *HUBER, fp32=True)
    except Exception:
        traceback.print_exc()
