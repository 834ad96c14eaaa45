"""Short run for ncu: python tools/profile_run.py <workload> [iters] [robust]  (see profiles/README.md)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
workload = sys.argv[1] if len(sys.argv) > 1 else "kitti00_shaped"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
robust = sys.argv[3] if len(sys.argv) > 3 else "none"
KERNELS = {"none": ((0, 0), (0.0, 0.0)), "huber": ((1, 1), (5.991 ** 0.5, 7.815 ** 0.5))}
path = os.path.join(ROOT, "oracle", "_ref", "fixtures", workload + ".cubagraph")
prob = pkg.graphio.flatten(pkg.graphio.read_graph(path) if workload.startswith("ba_") else pkg.synth.make_config(workload))
eng = pkg.Engine(device=0)
for et in (0, 1):
    eng.set_robust_kernels(KERNELS[robust][0][et], KERNELS[robust][1][et], et)
eng.initialize(prob)
stats = eng.optimize(iters)
print(workload, eng.sizes, [round(s["chi2"], 3) for s in stats], [s["pcg_iters"] for s in stats])
