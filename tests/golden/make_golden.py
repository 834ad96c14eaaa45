#!/usr/bin/env python3
"""Generates tests/golden/*.json.  Run in the authoring container (needs /root/reference for the KITTI
fixtures; the synthetic vectors need nothing).  The vectors are outputs of the CPU oracle
(oracle/ba_oracle.c), which is itself pinned to the reference's published chi2 table
(README.md:141-150) -- that table is stored here verbatim as `readme_chi2_kitti00_none`."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
oracle = ge.load_oracle()
HUBER = ((1, 1), (5.991 ** 0.5, 7.815 ** 0.5))
NONE = ((0, 0), (0.0, 0.0))
TUKEY = ((2, 2), (4.0, 5.0))


def trajectory(prob, rk, warmup):
    out = {}
    o = oracle.Oracle(prob, *rk)
    out["initial_chi2"] = o.compute_errors()
    if warmup:   # reference protocol: initialize();optimize(1) writes back, then initialize();optimize(10)
        chi, lam, tr = o.optimize(1)
        out["warmup_chi2"] = float(chi[0])
        q, t, Xw = o.state()
        prob = prob.copy(); prob.q, prob.t, prob.Xw = q, t, Xw
        o = oracle.Oracle(prob, *rk)
    chi, lam, tr = o.optimize(10)
    out["chi2"] = [float(v) for v in chi]; out["lambda"] = [float(v) for v in lam]; out["trials"] = [int(v) for v in tr]
    q, t, Xw = o.state()
    out["state_checksum"] = [float(np.abs(q).sum()), float(np.abs(t).sum()), float(np.abs(Xw).sum())]
    out["sizes"] = dict(nhpl=o.nhpl, nblk=o.nblk, nmul=o.nmul, numP=o.numP, numL=o.numL, E=o.E)
    return out


def main():
    gold = {"readme_chi2_kitti00_none": [334210.0, 331822.8, 329700.4, 327743.4, 326123.2, 324876.6, 323698.5, 322572.7, 321410.3, 320086.4]}
    fx = os.path.join(ROOT, "oracle", "_ref", "fixtures")
    for name in ("ba_kitti_07", "ba_kitti_00"):
        path = os.path.join(fx, name + ".cubagraph")
        if not os.path.exists(path):
            continue
        prob = pkg.graphio.flatten(pkg.graphio.read_graph(path))
        for label, rk in (("none", NONE), ("huber", HUBER)):
            gold["%s_%s" % (name, label)] = trajectory(prob, rk, warmup=True)
            print(name, label, gold["%s_%s" % (name, label)]["chi2"][-1], flush=True)
    for name in ("tiny", "small"):
        prob = pkg.graphio.flatten(pkg.synth.make_config(name))
        for label, rk in (("none", NONE), ("huber", HUBER), ("tukey", TUKEY)):
            gold["synth_%s_%s" % (name, label)] = trajectory(prob, rk, warmup=False)
    with open(os.path.join(ROOT, "tests", "golden", "oracle_trajectories.json"), "w") as f:
        json.dump(gold, f, indent=1)


if __name__ == "__main__":
    main()
