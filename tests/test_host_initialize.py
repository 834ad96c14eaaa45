"""Host side of the drop-in class without a GPU: initialize() numbers the vertices exactly like the reference
(src/cuda_bundle_adjustment.cpp:142-200) -- checked against graphio.flatten on a graph with fixed vertices on both sides,
vertices without edges and sparse ids, large enough that the landmark passes run on several host threads."""
import json
import os
import subprocess

import numpy as np

from conftest import ROOT


def _build(tmp_path_factory, pkg):
    out = str(tmp_path_factory.mktemp("cpp") / "index_assignment_driver")
    libdir = os.path.dirname(pkg.library_path())
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-DCUBA_FORCE_EIGEN_COMPAT", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "index_assignment_driver.cpp"), "-L", libdir, "-lcuba_b200",
                           "-Wl,-rpath," + libdir, "-o", out])
    return out


def test_initialize_numbering_matches_flatten(tmp_path_factory, pkg):
    exe = _build(tmp_path_factory, pkg)
    g = pkg.synth.make_config("kitti07_shaped")
    rng = np.random.default_rng(5)
    nP, nL = len(g["pose_id"]), len(g["lm_id"])
    assert nL > 16384                                       # two slices of the threaded landmark passes
    g["pose_fixed"] = g["pose_fixed"].copy(); g["lm_fixed"] = g["lm_fixed"].copy()
    g["pose_fixed"][rng.choice(nP, 7, replace=False)] = 1
    g["lm_fixed"][rng.choice(nL, nL // 30, replace=False)] = 1
    # sparse, shuffled ids (the class orders by id, not by insertion)
    g["pose_id"] = (rng.permutation(nP) * 3 + 1).astype(g["pose_id"].dtype)
    new_lid = (rng.permutation(nL) * 2 + 5).astype(g["lm_id"].dtype)
    old_to_row = np.full(int(g["lm_id"].max()) + 1, -1, np.int64); old_to_row[g["lm_id"]] = np.arange(nL)
    # edges refer to ids: remap them with the vertices
    pid_old = pkg.synth.make_config("kitti07_shaped")["pose_id"]
    prow = np.full(int(pid_old.max()) + 1, -1, np.int64); prow[pid_old] = np.arange(nP)
    for k in ("mono", "stereo"):
        g[k + "_vP"] = g["pose_id"][prow[g[k + "_vP"]]]
        g[k + "_vL"] = new_lid[old_to_row[g[k + "_vL"]]]
    g["lm_id"] = new_lid
    # landmarks without edges: drop every edge of 2 % of them
    lonely = set(int(v) for v in g["lm_id"][rng.choice(nL, nL // 50, replace=False)])
    for k in ("mono", "stereo"):
        keep = np.array([int(v) not in lonely for v in g[k + "_vL"]], dtype=bool)
        g[k + "_vP"], g[k + "_vL"] = g[k + "_vP"][keep], g[k + "_vL"][keep]
        g[k + "_meas"] = g[k + "_meas"][keep]
        g[k + "_info"] = g[k + "_info"][keep]
    path = str(tmp_path_factory.mktemp("graph") / "g.cubagraph")
    pkg.graphio.write_graph(path, g)
    dump = path + ".idx"
    out = subprocess.run([exe, path, dump, "3"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    info = json.loads(out.stdout)
    prob = pkg.graphio.flatten(pkg.graphio.read_graph(path))
    idx = np.fromfile(dump, dtype=np.int32)
    iP, iL = idx[:nP], idx[nP:]
    expP = np.full(nP, -1, np.int64); expP[prob.pose_rows] = np.arange(len(prob.pose_rows))
    expL = np.full(nL, -1, np.int64); expL[prob.lm_rows] = np.arange(len(prob.lm_rows))
    assert (expL < 0).sum() >= nL // 50 and prob.numL < prob.Lall and prob.numP < prob.Pall
    np.testing.assert_array_equal(iP, expP)
    np.testing.assert_array_equal(iL, expL)
    # both-fixed edges are dropped by initialize() only from the flat arrays; nedges() counts the graph's edges
    assert info["nedges"] == len(g["mono_vP"]) + len(g["stereo_vP"])
