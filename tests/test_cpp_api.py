"""The drop-in C++ API (include/cuda_bundle_adjustment.h) exercised exactly like the reference's
sample program: graph objects -> warm-up -> initialize()+optimize(10) -> batchStatistics/timeProfile/chiSquared."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import KERNELS, ROOT


def _build_sample(tmp_path_factory, pkg):
    out = str(tmp_path_factory.mktemp("cpp") / "sample_ba_from_file")
    libdir = os.path.dirname(pkg.library_path())
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-DCUBA_FORCE_EIGEN_COMPAT", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "samples", "sample_ba_from_file.cpp"), "-L", libdir, "-lcuba_b200",
                           "-Wl,-rpath," + libdir, "-o", out])
    return out


def test_sample_compiles_against_dropin_headers(tmp_path_factory, pkg):
    """compile + link only: user code written for the reference API builds against our headers"""
    exe = _build_sample(tmp_path_factory, pkg)
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_cpp_api_matches_oracle(tmp_path_factory, pkg, oracle):
    exe = _build_sample(tmp_path_factory, pkg)
    g = pkg.synth.make_config("small")
    path = str(tmp_path_factory.mktemp("graph") / "small.cubagraph")
    pkg.graphio.write_graph(path, g)
    out = subprocess.run([exe, path, "--json", "--huber"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    res = json.loads(out.stdout)
    prob = pkg.graphio.flatten(g)
    assert res["nedges"] == prob.nedges
    rk = KERNELS["huber"]
    o = oracle.Oracle(prob, *rk)
    o.optimize(1)                      # the sample's warm-up, written back into the graph
    q, t, Xw = o.state()
    p2 = prob.copy(); p2.q, p2.t, p2.Xw = q, t, Xw
    o2 = oracle.Oracle(p2, *rk)
    chi, lam, tr = o2.optimize(10)
    assert np.allclose(res["chi2"], chi, rtol=1e-10)
    assert res["sum_edge_chi2"] == pytest.approx(o2.chi_sqs().sum(), rel=1e-9)
    oq, ot, oX = o2.state()
    last_row = int(np.nonzero(prob.pose_rows == len(g["pose_id"]) - 1)[0][0])
    assert np.allclose(res["t_last"], ot[last_row], rtol=1e-9, atol=1e-12)
    assert set(res["profile"]) == set(pkg.PROFILE_ITEMS)


@pytest.mark.gpu
@pytest.mark.parametrize("name,kernel", [("small", "huber"), ("ba_kitti_07", "none")])
def test_comparison_report(pkg, name, kernel, capsys):
    """tests/compare_with_reference.py = the reference's sample_comparison_with_g2o.cpp:62-136 printout (chi2 columns side by
    side, RMSE of q/t/Xw) with the CPU oracle and the compiled reference GPU build beside this engine's drop-in class"""
    import io
    from conftest import have_fixture
    import compare_with_reference as cmp
    if name.startswith("ba_") and not have_fixture(name):
        pytest.skip("fixture not extracted")
    buf = io.StringIO()
    r = cmp.compare(name, kernel, with_cpu=True, out=buf)
    text = buf.getvalue()
    assert "=== Objective function value" in text and "=== RMSE between" in text
    assert r["chi2_rel_cpu"] < 1e-10 and max(r["rmse_cpu"].values()) < 1e-8
    if "chi2_rel_ref" in r:
        assert r["chi2_rel_ref"] < 1e-10 and max(r["rmse_ref"].values()) < 1e-8
    print(text)
