"""Parity of the 10-iteration trajectory vs the CPU oracle as a function of the PCG tolerance."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package(); oracle = ge.load_oracle()
HUBER = ((1, 1), (5.991 ** 0.5, 7.815 ** 0.5))
for workload in sys.argv[1:] or ["kitti07_shaped", "ba_kitti_07"]:
    path = os.path.join(ROOT, "oracle", "_ref", "fixtures", workload + ".cubagraph")
    if workload.startswith("ba_") and not os.path.exists(path):
        continue
    g = pkg.graphio.read_graph(path) if workload.startswith("ba_") else pkg.synth.make_config(workload)
    prob = pkg.graphio.flatten(g)
    o = oracle.Oracle(prob, *HUBER); chi, lam, tr = o.optimize(10); oq, ot, oX = o.state()
    for tol in (1e-13, 1e-12, 1e-11, 1e-10, 1e-9):
        eng = pkg.Engine(device=0, pcg_tol=tol)
        for et in (0, 1):
            eng.set_robust_kernels(HUBER[0][et], HUBER[1][et], et)
        eng.initialize(prob)
        st = eng.optimize(10)
        got = np.array([s["chi2"] for s in st]); q, t, X = eng.state()
        print("%s tol %.0e: pcg iters %6d  chi2 rel %.1e  q %.1e t %.1e Xw %.1e  trials ok %s" % (workload, tol, sum(s["pcg_iters"] for s in st),
              np.abs(got - chi).max() / chi.max(), np.abs(q - oq).max(), np.abs(t - ot).max() / np.abs(ot).max(), np.abs(X - oX).max() / np.abs(oX).max(),
              [s["trials"] for s in st] == list(tr)), flush=True)
        eng.close()
