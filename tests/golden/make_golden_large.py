#!/usr/bin/env python3
"""Generates tests/golden/oracle_large.json: the CPU oracle's 10-iteration trajectories (no warm-up) on the BASELINE configs
that are too big for a live oracle run inside bench.py / the GPU tests -- synth_mono_5m (C3), synth_stereo_10m (C4), both Huber,
and kitti00_shaped NONE/Huber.  Minutes of CPU time; run once in the authoring container: python tests/golden/make_golden_large.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
oracle = ge.load_oracle()
HUBER = ((1, 1), (5.991 ** 0.5, 7.815 ** 0.5))
NONE = ((0, 0), (0.0, 0.0))
out_path = os.path.join(ROOT, "tests", "golden", "oracle_large.json")
gold = json.load(open(out_path)) if os.path.exists(out_path) else {}
for name, label, rk in (("kitti00_shaped", "none", NONE), ("kitti00_shaped", "huber", HUBER), ("synth_mono_5m", "huber", HUBER), ("synth_stereo_10m", "huber", HUBER)):
    key = "%s_%s" % (name, label)
    if key in gold and "--force" not in sys.argv:
        continue
    prob = pkg.graphio.flatten(pkg.synth.make_config(name))
    o = oracle.Oracle(prob, *rk)
    t0 = time.time()
    chi, lam, tr = o.optimize(10)
    q, t, Xw = o.state()
    gold[key] = {"chi2": [float(v) for v in chi], "lambda": [float(v) for v in lam], "trials": [int(v) for v in tr],
                 "state_checksum": [float(np.abs(q).sum()), float(np.abs(t).sum()), float(np.abs(Xw).sum())], "oracle_seconds": time.time() - t0}
    print(key, gold[key]["chi2"][-1], "%.1f s" % (time.time() - t0), flush=True)
    with open(out_path, "w") as f:
        json.dump(gold, f, indent=1)
