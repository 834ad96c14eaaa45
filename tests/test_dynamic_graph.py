"""Dynamic graph semantics of the drop-in C++ API (reference src/cuda_bundle_adjustment.cpp:677-781,
README.md:46): remove*/re-add/re-initialize(), repeated optimize() without initialize(), chiSquared()-driven outlier
removal, vertices fixed after the fact, per-pose cameras.  The reference has no test for any of this.

tests/cpp/dynamic_graph_driver.cpp executes an op list through cuba::CudaBundleAdjustment; this file mirrors the same ops
on the graph arrays and checks every optimize() against the CPU oracle started from the same (written-back) estimate."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import KERNELS, ROOT


def _build_driver(tmp_path_factory, pkg):
    out = str(tmp_path_factory.mktemp("cppdyn") / "dynamic_graph_driver")
    libdir = os.path.dirname(pkg.library_path())
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-DCUBA_FORCE_EIGEN_COMPAT", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "dynamic_graph_driver.cpp"), "-L", libdir, "-lcuba_b200",
                           "-Wl,-rpath," + libdir, "-o", out])
    return out


def test_driver_compiles(tmp_path_factory, pkg):
    assert os.path.exists(_build_driver(tmp_path_factory, pkg))


class Mirror:
    """the same graph edits on the arrays of a graph dict; optimize() = the CPU oracle on the flattened live graph"""

    def __init__(self, pkg, oracle, g, rk):
        self.pkg, self.oracle, self.rk = pkg, oracle, rk
        self.g = {k: np.array(v, copy=True) for k, v in g.items()}
        self.live_p = np.ones(len(g["pose_id"]), bool); self.live_l = np.ones(len(g["lm_id"]), bool)
        self.live_m = np.ones(len(g["mono_vP"]), bool); self.live_s = np.ones(len(g["stereo_vP"]), bool)
        self.o = None; self.prob = None; self.stats = []; self.chisq = {}

    def live_graph(self):
        g = dict(self.g)
        for k in ("pose_id", "pose_fixed", "q", "t", "cam"):
            g[k] = self.g[k][self.live_p]
        for k in ("lm_id", "lm_fixed", "Xw"):
            g[k] = self.g[k][self.live_l]
        for k in ("mono_vP", "mono_vL", "mono_meas", "mono_info"):
            g[k] = self.g[k][self.live_m]
        for k in ("stereo_vP", "stereo_vL", "stereo_meas", "stereo_info"):
            g[k] = self.g[k][self.live_s]
        return g

    def init(self):
        self.lg = self.live_graph()
        self.prob = self.pkg.graphio.flatten(self.lg)
        self.o = self.oracle.Oracle(self.prob, *self.rk)
        self.stats = []

    def opt(self, n):
        if self.o is None:
            self.init()
        chi, lam, tr = self.o.optimize(n)
        self.stats += list(chi)
        q, t, Xw = self.o.state()
        # finalize(): write back into the live rows of the full arrays
        prow = np.nonzero(self.live_p)[0][self.prob.pose_rows]; lrow = np.nonzero(self.live_l)[0][self.prob.lm_rows]
        self.g["q"][prow] = q; self.g["t"][prow] = t; self.g["Xw"][lrow] = Xw
        cs = self.o.chi_sqs()
        self.chisq = {}
        mrow = np.nonzero(self.live_m)[0][self.prob.mono_rows]; srow = np.nonzero(self.live_s)[0][self.prob.stereo_rows]
        for k, r in enumerate(mrow):
            self.chisq[("m", int(r))] = cs[k]
        for k, r in enumerate(srow):
            self.chisq[("s", int(r))] = cs[len(mrow) + k]
        return list(self.stats), float(cs.sum())

    def rmpose(self, pid):
        self.live_m &= self.g["mono_vP"] != pid; self.live_s &= self.g["stereo_vP"] != pid
        self.live_p &= self.g["pose_id"] != pid

    def rmlm(self, lid):
        self.live_m &= self.g["mono_vL"] != lid; self.live_s &= self.g["stereo_vL"] != lid
        self.live_l &= self.g["lm_id"] != lid

    def outliers(self, T):
        n = 0
        for (kind, r), v in self.chisq.items():
            live = self.live_m if kind == "m" else self.live_s
            if v > T and live[r]:
                live[r] = False; n += 1
        return n

    def nedges(self):
        return int(self.live_m.sum() + self.live_s.sum())


def _run(pkg, oracle, tmp_path_factory, g, ops, kernel="huber", tol=1e-10):
    exe = _build_driver(tmp_path_factory, pkg)
    d = tmp_path_factory.mktemp("dyn")
    path = str(d / "g.cubagraph"); dump = str(d / "state.bin")
    pkg.graphio.write_graph(path, g)
    cmd = [exe, path, ";".join(ops), dump] + (["--huber"] if kernel == "huber" else [])
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    steps = json.loads(out.stdout)["steps"]
    m = Mirror(pkg, oracle, g, KERNELS[kernel])
    for st, op in zip(steps, ops):
        f = op.split(":")
        if f[0] == "init":
            m.init()
        elif f[0] == "opt":
            chi, total = m.opt(int(f[1]))
            assert len(st["chi2"]) == len(chi), (op, st["chi2"], chi)
            assert np.allclose(st["chi2"], chi, rtol=tol), (op, st["chi2"], chi)
            assert st["sum_edge_chi2"] == pytest.approx(total, rel=1e-8), op
        elif f[0] == "rmpose":
            m.rmpose(int(f[1]))
        elif f[0] == "rmlm":
            m.rmlm(int(f[1]))
        elif f[0] == "rmedge":
            (m.live_m if f[1] == "m" else m.live_s)[int(f[2])] = False
        elif f[0] == "addedge":
            (m.live_m if f[1] == "m" else m.live_s)[int(f[2])] = True
        elif f[0] == "fixp":
            m.g["pose_fixed"][m.g["pose_id"] == int(f[1])] = 1
        elif f[0] == "fixl":
            m.g["lm_fixed"][m.g["lm_id"] == int(f[1])] = 1
        elif f[0] == "outliers":
            assert st["removed"] == m.outliers(float(f[1])), op
        assert st["nedges"] == m.nedges(), (op, st["nedges"], m.nedges())
        assert st["nposes"] == int(m.live_p.sum()) and st["nlandmarks"] == int(m.live_l.sum()), op
    # final estimate of every vertex object (removed ones keep their last value), file order
    nP, nL = len(g["pose_id"]), len(g["lm_id"])
    raw = np.fromfile(dump, dtype=np.float64)
    q = raw[:4 * nP].reshape(nP, 4); t = raw[4 * nP:7 * nP].reshape(nP, 3); Xw = raw[7 * nP:].reshape(nL, 3)
    assert np.abs(q - m.g["q"]).max() < 1e-9 and np.abs(t - m.g["t"]).max() < 1e-8 * max(1.0, np.abs(m.g["t"]).max())
    assert np.abs(Xw - m.g["Xw"]).max() < 1e-8 * max(1.0, np.abs(m.g["Xw"]).max())
    return steps, m


@pytest.mark.gpu
def test_remove_readd_reinitialize(tmp_path_factory, pkg, oracle):
    g = pkg.synth.make_config("small")
    pid = int(g["pose_id"][len(g["pose_id"]) // 2]); lid = int(g["stereo_vL"][100]); lid2 = int(g["mono_vL"][7])
    ops = ["init", "opt:2", "rmpose:%d" % pid, "rmlm:%d" % lid, "rmedge:m:5", "rmedge:s:9", "rmedge:s:10", "addedge:s:9", "addedge:s:9",
           "init", "opt:3", "rmlm:%d" % lid2, "rmedge:m:5", "init", "opt:2"]
    _run(pkg, oracle, tmp_path_factory, g, ops)


@pytest.mark.gpu
def test_repeated_optimize_without_initialize(tmp_path_factory, pkg, oracle):
    """optimize() twice in a row continues from the current estimate and appends to batchStatistics (cpp:848)"""
    g = pkg.synth.make_config("small")
    steps, m = _run(pkg, oracle, tmp_path_factory, g, ["init", "opt:2", "opt:3", "opt:1"], kernel="none")
    assert [s.get("stats_before") for s in steps if s["op"].startswith("opt")] == [0, 2, 5]


@pytest.mark.gpu
def test_chi_squared_outlier_loop(tmp_path_factory, pkg, oracle):
    g = pkg.synth.make_config("small")
    steps, m = _run(pkg, oracle, tmp_path_factory, g, ["init", "opt:3", "outliers:7.815", "init", "opt:3", "outliers:5.0", "init", "opt:2"])
    assert steps[2]["removed"] > 0


@pytest.mark.gpu
def test_fixing_vertices_between_runs(tmp_path_factory, pkg, oracle):
    """vertices fixed after the first run move to the end of the index order; an edge whose two ends are now fixed is dropped"""
    g = pkg.synth.make_config("tiny")
    p = int(g["stereo_vP"][3]); l = int(g["stereo_vL"][3])
    _run(pkg, oracle, tmp_path_factory, g, ["init", "opt:2", "fixp:%d" % p, "fixl:%d" % l, "fixl:%d" % int(g["lm_id"][5]), "init", "opt:3"])


@pytest.mark.gpu
def test_per_pose_cameras(tmp_path_factory, pkg, oracle):
    """every pose carries its own intrinsics (README.md:46 of the reference): perturb them per pose, regenerate nothing else"""
    g = pkg.synth.make_config("small")
    rng = np.random.default_rng(5)
    g["cam"] = g["cam"] * (1.0 + 0.002 * rng.standard_normal(g["cam"].shape))
    _run(pkg, oracle, tmp_path_factory, g, ["init", "opt:4"])
