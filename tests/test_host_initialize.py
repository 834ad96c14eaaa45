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


def _build_flatten_driver(tmp_path_factory, pkg):
    out = str(tmp_path_factory.mktemp("cppflat") / "flatten_driver")
    libdir = os.path.dirname(pkg.library_path())
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-DCUBA_FORCE_EIGEN_COMPAT", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "flatten_driver.cpp"), "-L", libdir, "-lcuba_b200",
                           "-Wl,-rpath," + libdir, "-o", out])
    return out


def _read_dumps(path, n):
    raw = open(path, "rb").read()
    off = 0
    out = []

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=off)
        off += a.nbytes
        return a
    for _ in range(n):
        Pall, numP, Lall, numL, E2, E3 = (int(v) for v in take(np.int64, 6))
        d = dict(Pall=Pall, numP=numP, Lall=Lall, numL=numL, q=take(np.float64, 4 * Pall).reshape(-1, 4), t=take(np.float64, 3 * Pall).reshape(-1, 3),
                 cam=take(np.float64, 5 * Pall).reshape(-1, 5), Xw=take(np.float64, 3 * Lall).reshape(-1, 3))
        d["idx2"] = take(np.int32, 2 * E2).reshape(-1, 2); d["meas2"] = take(np.float64, 2 * E2).reshape(-1, 2); d["omega2"] = take(np.float64, E2)
        d["idx3"] = take(np.int32, 2 * E3).reshape(-1, 2); d["meas3"] = take(np.float64, 3 * E3).reshape(-1, 3); d["omega3"] = take(np.float64, E3)
        out.append(d)
    assert off == len(raw)
    return out


def _check_flatten_ops(pkg, tmp_path_factory, g, ops):
    """runs `ops` through the C++ class and through the array mirror of tests/test_dynamic_graph.py; every initialize() must produce
    exactly graphio.flatten of the live graph"""
    from test_dynamic_graph import Mirror
    exe = _build_flatten_driver(tmp_path_factory, pkg)
    d = tmp_path_factory.mktemp("flat")
    path = str(d / "g.cubagraph"); dump = str(d / "flat.bin")
    pkg.graphio.write_graph(path, g)
    out = subprocess.run([exe, path, ";".join(ops), dump], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    info = json.loads(out.stdout)
    dumps = _read_dumps(dump, info["inits"])
    m = Mirror(pkg, None, g, None)
    # the class keeps its edges in insertion order: a removed edge that is added again goes to the end of its list
    order = {"m": list(range(len(g["mono_vP"]))), "s": list(range(len(g["stereo_vP"])))}

    def live_graph():
        lg = m.live_graph()
        for kind, pre, live in (("m", "mono", m.live_m), ("s", "stereo", m.live_s)):
            rows = np.array([r for r in order[kind] if live[r]], dtype=np.int64)
            for key in ("_vP", "_vL", "_meas", "_info"):
                lg[pre + key] = m.g[pre + key][rows]
        return lg

    def drop_dead():
        for kind, live in (("m", m.live_m), ("s", m.live_s)):
            order[kind] = [r for r in order[kind] if live[r]]
    k = 0
    for op in ops:
        f = op.split(":")
        if f[0] == "init":
            prob = pkg.graphio.flatten(live_graph())
            got = dumps[k]; k += 1
            for name in ("Pall", "numP", "Lall", "numL"):
                assert got[name] == getattr(prob, name), (op, k, name)
            for name in ("q", "t", "cam", "Xw", "idx2", "meas2", "omega2", "idx3", "meas3", "omega3"):
                a, b = got[name], np.asarray(getattr(prob, name))
                assert a.shape == b.reshape(a.shape).shape and np.array_equal(a, b.reshape(a.shape)), (op, k, name)
        elif f[0] == "rmpose":
            m.rmpose(int(f[1])); drop_dead()
        elif f[0] == "rmlm":
            m.rmlm(int(f[1])); drop_dead()
        elif f[0] == "rmedge":
            (m.live_m if f[1] == "m" else m.live_s)[int(f[2])] = False; drop_dead()
        elif f[0] == "addedge":
            live = m.live_m if f[1] == "m" else m.live_s
            if not live[int(f[2])]:
                live[int(f[2])] = True; order[f[1]].append(int(f[2]))
        elif f[0] == "fixp":
            m.g["pose_fixed"][m.g["pose_id"] == int(f[1])] = 1
        elif f[0] == "fixl":
            m.g["lm_fixed"][m.g["lm_id"] == int(f[1])] = 1
        elif f[0] == "unfixl":
            m.g["lm_fixed"][m.g["lm_id"] == int(f[1])] = 0
        elif f[0] == "meas":
            if f[1] == "m":
                m.g["mono_meas"][int(f[2]), 0] += float(f[3])
            else:
                m.g["stereo_meas"][int(f[2]), 2] += float(f[3])
        elif f[0] == "movel":
            m.g["Xw"][m.g["lm_id"] == int(f[1]), 1] += float(f[2])
    assert k == info["inits"]
    assert info["nedges"] == m.nedges()


def test_flat_arrays_follow_graph_edits(tmp_path_factory, pkg):
    """tombstoned and re-added edges (a re-added edge moves to the end of its list), removed vertices, vertices fixed and released
    between two initialize() calls (both-fixed edges drop out of the flat arrays and come back), measurements and estimates edited
    in place, repeated initialize() without any change"""
    g = pkg.synth.make_config("small")
    g["pose_fixed"] = g["pose_fixed"].copy(); g["lm_fixed"] = g["lm_fixed"].copy()
    p0, p1 = int(g["pose_id"][0]), int(g["pose_id"][7])
    # a landmark seen by pose p0 through a mono edge and one through a stereo edge
    lm_m = int(g["mono_vL"][np.nonzero(g["mono_vP"] == p0)[0][0]]); lm_s = int(g["stereo_vL"][np.nonzero(g["stereo_vP"] == p0)[0][0]])
    ops = ["init", "init",
           "meas:m:5:0.25", "meas:s:9:-0.5", "movel:%d:0.125" % lm_m, "init",
           "rmedge:m:3", "rmedge:s:11", "init",
           "addedge:m:3", "init",
           "fixp:%d" % p0, "fixl:%d" % lm_m, "fixl:%d" % lm_s, "init", "init",
           "meas:m:7:1.5", "init",
           "unfixl:%d" % lm_m, "init",
           "rmlm:%d" % lm_s, "rmpose:%d" % p1, "init",
           "rmedge:s:100", "addedge:s:100", "rmedge:s:100", "init", "addedge:s:100", "addedge:s:11", "init", "init"]
    _check_flatten_ops(pkg, tmp_path_factory, g, ops)
