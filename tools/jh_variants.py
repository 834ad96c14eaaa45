"""J+H landmark-pass variants: device time per launch (L2 flushed) on one workload."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
for workload in sys.argv[1:] or ["kitti00_shaped"]:
    prob = pkg.graphio.flatten(pkg.synth.make_config(workload))
    for v in (0, 5, 4):
        eng = pkg.Engine(device=0, jh_variant=v)
        eng.initialize(prob)
        chi = eng.linearize()
        ms1 = eng.bench_stage(1, reps=30, flush_l2=True)
        ms1w = eng.bench_stage(1, reps=30, flush_l2=False)
        ms5 = eng.bench_stage(5, reps=10, flush_l2=True)
        print("%s jh_variant %d: landmark pass %.1f us (L2 flushed) %.1f us (warm); backsub+update+chi2 %.1f us; chi2 %.9g" % (workload, v, 1e3 * ms1, 1e3 * ms1w, 1e3 * ms5, chi), flush=True)
        eng.close()
