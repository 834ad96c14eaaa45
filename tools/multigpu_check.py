"""Landmark-sharded run under torchrun: parity of the N-GPU trajectory with the CPU oracle (rank 0 prints).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py [workload]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

pkg = ge.load_package()
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
HUBER = ((1, 1), (5.991 ** 0.5, 7.815 ** 0.5))
for workload in sys.argv[1:] or ["small", "kitti07_shaped"]:
    prob = pkg.graphio.flatten(pkg.synth.make_config(workload))
    eng = pkg.Engine(device=local)
    for et in (0, 1):
        eng.set_robust_kernels(HUBER[0][et], HUBER[1][et], et)
    eng.set_comm(rank, world, pkg.sharding.broadcast_unique_id(pkg.Engine, rank, world))
    eng.initialize(prob)
    torch.cuda.synchronize(); dist.barrier()
    import time
    t0 = time.perf_counter()
    stats = eng.optimize(10)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    q, t, Xw = eng.state()
    if rank == 0:
        oracle = ge.load_oracle()
        o = oracle.Oracle(prob, *HUBER)
        chi, lam, tr = o.optimize(10)
        got = np.array([s["chi2"] for s in stats])
        oq, ot, oX = o.state()
        print("%s world %d: %.1f ms; chi2 rel diff vs oracle %.2e; trials %s vs %s; state diff q %.1e t %.1e Xw %.1e; pcg iters %s" % (
            workload, world, 1e3 * dt, np.abs(got - chi).max() / chi.max(), [s["trials"] for s in stats], list(tr),
            np.abs(q - oq).max(), np.abs(t - ot).max() / np.abs(ot).max(), np.abs(Xw - oX).max() / np.abs(oX).max(), [s["pcg_iters"] for s in stats]), flush=True)
    eng.close()
dist.destroy_process_group()
