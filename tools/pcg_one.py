"""One linearisation + solves at a few dampings with a given PCG variant (for ncu): python tools/pcg_one.py <workload> <variant>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
workload = sys.argv[1] if len(sys.argv) > 1 else "kitti00_shaped"
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 3
path = os.path.join(ROOT, "oracle", "_ref", "fixtures", workload + ".cubagraph")
g = pkg.graphio.read_graph(path) if workload.startswith("ba_") else pkg.synth.make_config(workload)
prob = pkg.graphio.flatten(g)
eng = pkg.Engine(device=0, pcg_variant=variant)
eng.initialize(prob)
eng.linearize()
md = eng.max_diagonal()
for lam in (1e-5 * md, 1e-8 * md, 1e-10 * md):
    print(workload, variant, lam, eng.solve(lam), flush=True)
