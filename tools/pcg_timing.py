"""Per-phase clock64 breakdown of k_pcg3 (default) or k_pcg4 (PCG_VARIANT=3; MAX_AGG, LAM_SCALE optional).
Needs cuda-bundle-adjustment_b200/libcuba_b200_timing.so, the library built with -DCUBA_PCG_TIMING:
  cd cuda-bundle-adjustment_b200 && nvcc -DCUBA_PCG_TIMING -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
      -Xcompiler -fPIC -shared -I ../include -o libcuba_b200_timing.so csrc/cuba_engine.cu csrc/cuba_structure.cpp csrc/cuba_api.cpp -ldl"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
from cuda_bundle_adjustment_b200 import binding  # noqa: E402
binding.library_path = lambda: os.path.join(ROOT, "cuda-bundle-adjustment_b200", "libcuba_b200_timing.so")
L = pkg.load_library()
L.cuba_debug_get_pcg_timing.restype = C.c_int
L.cuba_debug_get_pcg_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
VARIANT = int(os.environ.get("PCG_VARIANT", "0"))
names = ["poll", "sync_after_poll", "scalars", "update+sync", "spmv", "reduce+sync", "publish"] if VARIANT not in (3, 5, 6) else \
    ["loads+scalars", "rc+owners+gather", "coarse slices", "u", "spmv+restrict", "reduce+publish", "grid barrier"]
if VARIANT in (5, 6):   # k_pcg5
    # one-GPU on-chip solves run the tuned shape (cuba_pcg5t.cuh), CUBA_PCG5_LEGACY=1 / large systems the legacy one (cuba_pcg5.cuh)
    names = ["poll w + partials", "local sums (+rank hop) + scalars", "advance rc (legacy: + r,s,p,y)", "coarse rows (tuned: + advance r,s,p,y), poll c, u", "spmv + row sums", "publish w + partial products", "publish partials"]
for workload in sys.argv[1:] or ["kitti00_shaped"]:
    path = os.path.join(ROOT, "oracle", "_ref", "fixtures", workload + ".cubagraph")
    g = pkg.graphio.read_graph(path) if workload.startswith("ba_") else pkg.synth.make_config(workload)
    prob = pkg.graphio.flatten(g)
    eng = pkg.Engine(device=0, pcg_variant=VARIANT, max_aggregates=int(os.environ.get("MAX_AGG", "0")))
    eng.initialize(prob)
    eng.linearize()
    lam = float(os.environ.get("LAM_SCALE", "1e-8")) * eng.max_diagonal()
    it, ok = eng.solve(lam)
    ms = eng.bench_stage(4, reps=3, flush_l2=False, lam=lam)
    buf = np.zeros((256, 8), dtype=np.int64)
    n = L.cuba_debug_get_pcg_timing(eng.h, buf.ctypes.data_as(C.c_void_p), 256)
    t = buf[:n].astype(np.float64)
    iters = t[:, 7].max()
    print("%s: %d iters, %.2f us/iter, %d CTAs; cycles per iteration (thread 0 of each CTA): mean / min / max over CTAs" % (workload, it, 1e3 * ms / it, n))
    for i, nm in enumerate(names):
        c = t[:, i] / iters
        print("   %-42s %8.0f %8.0f %8.0f" % (nm, c.mean(), c.min(), c.max()))
    print("   total            %8.0f" % (t[:, :7].sum(1) / iters).mean())
    eng.close()
