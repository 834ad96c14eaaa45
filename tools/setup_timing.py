"""Wall-clock marks of set_problem (CUBA_SETUP_TIMING=1): python tools/setup_timing.py [workload]"""
import os
import sys
os.environ["CUBA_SETUP_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
pkg = ge.load_package()
prob = pkg.graphio.flatten(pkg.synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "kitti00_shaped"))
eng = pkg.Engine(device=0)
for i in range(4):
    print("---- initialize", i, file=sys.stderr, flush=True)
    eng.initialize(prob)
