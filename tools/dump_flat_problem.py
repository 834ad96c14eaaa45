"""python tools/dump_flat_problem.py <workload> <out.bin>: sizes + (iP, iL) index pairs of a flattened graph for the host-only
C++ harnesses (tools/host_plan_bench.cpp).  workload: a fixture name (ba_kitti_00) or a synthetic config (kitti00_shaped, ...)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
workload, out = sys.argv[1], sys.argv[2]
path = os.path.join(ROOT, "oracle", "_ref", "fixtures", workload + ".cubagraph")
prob = pkg.graphio.flatten(pkg.graphio.read_graph(path) if workload.startswith("ba_") else pkg.synth.make_config(workload))
with open(out, "wb") as f:
    np.array([prob.Pall, prob.numP, prob.Lall, prob.numL, prob.E2, prob.E3], dtype=np.int64).tofile(f)
    np.ascontiguousarray(prob.idx2, dtype=np.int32).tofile(f)
    np.ascontiguousarray(prob.idx3, dtype=np.int32).tofile(f)
