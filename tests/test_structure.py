"""Host-side structure builder (csrc/cuba_structure.cpp) against the oracle's restatement of the
reference's index structures: bit-exact."""
import numpy as np
import pytest

from conftest import have_fixture


def _check(pkg, oracle, prob):
    s = pkg.build_structure_host(prob)
    o = oracle.Oracle(prob)
    cp, ri, e2h = o.hpl_structure(); rp, ci = o.hsc_structure()
    assert np.array_equal(cp, s["hplColPtr"]) and np.array_equal(ri, s["hplRowInd"]) and np.array_equal(e2h, s["edge2Hpl"])
    assert np.array_equal(rp, s["hscRowPtr"]) and np.array_equal(ci, s["hscColInd"])
    assert s["nhpl"] == o.nhpl and s["nblk"] == o.nblk and s["nmul"] == o.nmul
    # symmetric-full BSR: columns ascending per row, contains (a,b) iff it contains (b,a), diagonal present
    fr, fc = s["fullRowPtr"], s["fullColInd"]
    pairs = set()
    for r in range(prob.numP):
        cols = fc[fr[r]:fr[r + 1]]
        assert np.all(np.diff(cols) > 0)
        assert r in cols
        pairs.update((r, int(c)) for c in cols)
    assert all((b, a) in pairs for a, b in pairs)
    assert s["nblk_full"] == 2 * s["nblk"] - prob.numP
    return s


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_structure_synthetic(pkg, oracle, problems, name):
    _check(pkg, oracle, problems(name))


@pytest.mark.skipif(not have_fixture("ba_kitti_07"), reason="reference fixture absent")
def test_structure_kitti07(pkg, oracle, problems):
    s = _check(pkg, oracle, problems("ba_kitti_07"))
    # SURVEY.md section 8 sizes, verified against the reference's own structures
    assert (s["nhpl"], s["nblk"], s["nmul"]) == (94605, 4776, 308963)


def _variant(pkg, prob, fixed_poses=(), fixed_lms=()):
    """re-flatten `prob` with extra fixed vertices (exercises flags, the appended-fixed ordering and Hpl gaps)"""
    g = {"pose_id": np.arange(prob.Pall, dtype=np.int32), "pose_fixed": np.zeros(prob.Pall, np.int32),
         "q": prob.q.copy(), "t": prob.t.copy(), "cam": prob.cam.copy(),
         "lm_id": (prob.Pall + np.arange(prob.Lall)).astype(np.int32), "lm_fixed": np.zeros(prob.Lall, np.int32), "Xw": prob.Xw.copy(),
         "mono_vP": prob.idx2[:, 0].copy(), "mono_vL": (prob.Pall + prob.idx2[:, 1]).astype(np.int32), "mono_meas": prob.meas2.copy(),
         "mono_info": prob.omega2.copy(), "stereo_vP": prob.idx3[:, 0].copy(), "stereo_vL": (prob.Pall + prob.idx3[:, 1]).astype(np.int32),
         "stereo_meas": prob.meas3.copy(), "stereo_info": prob.omega3.copy()}
    g["pose_fixed"][prob.numP:] = 1
    g["lm_fixed"][prob.numL:] = 1
    g["pose_fixed"][list(fixed_poses)] = 1
    g["lm_fixed"][list(fixed_lms)] = 1
    return pkg.graphio.flatten(g)


def test_structure_fixed_vertices(pkg, oracle, problems):
    base = problems("tiny")
    p = _variant(pkg, base, fixed_poses=(0, 3, 7), fixed_lms=range(0, base.Lall, 5))
    assert p.numP == base.numP - 3 and p.numL < base.numL
    # edges with both ends fixed were dropped by flatten (reference cpp:212,233)
    assert not np.any((p.idx2[:, 0] >= p.numP) & (p.idx2[:, 1] >= p.numL))
    _check(pkg, oracle, p)


def test_structure_pose_only_and_landmark_only(pkg, oracle, problems):
    base = problems("tiny")
    pose_only = _variant(pkg, base, fixed_lms=range(base.Lall))
    assert pose_only.numL == 0
    s = pkg.build_structure_host(pose_only)
    assert s["nhpl"] == 0 and s["nblk"] == pose_only.numP   # only the diagonal blocks
    lm_only = _variant(pkg, base, fixed_poses=range(base.Pall))
    assert lm_only.numP == 0
    s = pkg.build_structure_host(lm_only)
    assert s["nhpl"] == 0 and s["nblk"] == 0


def test_structure_rejects_bad_input(pkg, problems):
    p = problems("tiny").copy()
    p.idx2 = p.idx2.copy(); p.idx2[0, 0] = p.Pall + 5
    with pytest.raises(pkg.CubaError):
        pkg.build_structure_host(p)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_structure_shards_partition_the_graph(pkg, problems, world):
    prob = problems("small")
    iL = np.concatenate([prob.idx2[:, 1], prob.idx3[:, 1]])
    bounds = pkg.sharding.shard_bounds(iL, prob.Lall, world)
    edges = prods = 0
    full = pkg.build_structure_host(prob)
    for r in range(world):
        s = pkg.build_structure_host(prob, r, world)
        assert (s["shard"][0], s["shard"][1]) == (bounds[r], bounds[r + 1])
        # the global structures do not depend on the shard
        assert np.array_equal(s["hscColInd"], full["hscColInd"]) and np.array_equal(s["hplRowInd"], full["hplRowInd"])
        edges += s["shard"][2]; prods += s["shard"][3]
    assert edges == prob.nedges and prods == full["nmul"]
    # balanced by edge count to within one landmark's degree
    per = [pkg.build_structure_host(prob, r, world)["shard"][2] for r in range(world)]
    assert max(per) - min(per) <= 2 * 64


@pytest.mark.parametrize("name,n_ctas,max_agg", [("tiny", 148, 74), ("small", 148, 74), ("small", 7, 3), ("kitti07_shaped", 148, 74),
                                                  ("kitti07_shaped", 31, 37), ("kitti07_shaped", 148, 37)])
def test_pcg_partition_host_logic(pkg, problems, name, n_ctas, max_agg):
    """host side of the PCG setup (csrc/cuba_structure.cpp, shared with the engine): row partition over the persistent CTAs, need
    lists, block-local column positions, pose aggregates of the two-level PCG and the coarse-block lists -- the library builds them
    and verifies every invariant (rows cover [0,P), own rows in the need list, diagonal encoding, aggregates aligned with CTA
    groups, every lower-triangle block in exactly one ascending coarse list)"""
    prob = problems(name)
    info = pkg.pcg_partition_host(prob, n_ctas, max_agg)
    assert info["G"] == min(n_ctas, prob.numP)
    assert 1 <= info["A"] <= max_agg and info["A"] == -(-info["G"] // info["gs"])
    assert info["maxRows"] >= -(-prob.numP // info["G"]) and info["needMax"] >= info["maxRows"]
    assert 1 <= info["maxNeedAgg"] <= info["A"]
    s = pkg.build_structure_host(prob)
    # lower block triangle of the coarse matrix: at least the diagonal blocks of the fine matrix, at most all of them
    assert prob.numP <= info["coarse_list_size"] <= s["nblk_full"]


@pytest.mark.parametrize("name", ["small", "kitti07_shaped", "kitti00_shaped"])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_row_distributed_pcg_plan(pkg, problems, name, world):
    """plan of k_pcg5 for `world` GPUs (csrc/cuba_structure.cpp::build_pcg5_plan, the engine's own code): rows over world x G
    virtual CTAs, aggregates that never straddle two ranks, and the halo masks -- for every row exactly the ranks, other than
    its owner, whose rows couple to it.  The library checks the invariants; here: sizes and that the halo grows with the cut count"""
    prob = problems(name)
    info = pkg.pcg5_plan_host(prob, world)
    if not info["ok"]:
        assert world * 8 > prob.numP or name == "small"
        return
    assert info["G"] % info["gs"] == 0 and info["A"] == world * info["G"] // info["gs"] <= 74
    assert info["maxRows"] * 6 <= 256 and world * info["G"] <= prob.numP
    assert (info["halo_rows"] == 0) == (world == 1)
    if world > 1:
        assert info["halo_rows"] >= world - 1


def test_pcg_partition_rejects_bad_arguments(pkg, problems):
    with pytest.raises(pkg.CubaError):
        pkg.pcg_partition_host(problems("tiny"), 0, 74)
