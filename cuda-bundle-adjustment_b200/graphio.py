"""Graph files and the flat problem layout.

`.cubagraph` (little endian): magic b"CUBAGRF1", int64[4] = nposes, nlandmarks, nmono, nstereo, then
  pose_id i32[nP], pose_fixed i32[nP], q f64[nP,4] (x,y,z,w), t f64[nP,3], cam f64[nP,5] (fx,fy,cx,cy,bf),
  lm_id i32[nL], lm_fixed i32[nL], Xw f64[nL,3],
  mono_vP i32[nM], mono_vL i32[nM], mono_meas f64[nM,2], mono_info f64[nM],
  stereo_vP i32[nS], stereo_vL i32[nS], stereo_meas f64[nS,3], stereo_info f64[nS]
(vertex references in edges are vertex *ids*, like the reference's JSON: samples/sample_ba_from_file.cpp:93-157).

`flatten()` mirrors CudaBlockSolver::initialize (reference src/cuda_bundle_adjustment.cpp:115-261):
index assignment iP / iL in ascending id order, free vertices first, fixed appended; vertices without
edges skipped; edges with both ends fixed dropped; monocular edges get ids 0..E2-1, stereo E2..E2+E3-1.
The reference walks an unordered_set for the edges (order = heap addresses); we use file order.
"""
from __future__ import annotations

import dataclasses

import numpy as np

MAGIC = b"CUBAGRF1"

_FIELDS = (("pose_id", np.int32, 1), ("pose_fixed", np.int32, 1), ("q", np.float64, 4), ("t", np.float64, 3),
           ("cam", np.float64, 5), ("lm_id", np.int32, 1), ("lm_fixed", np.int32, 1), ("Xw", np.float64, 3),
           ("mono_vP", np.int32, 1), ("mono_vL", np.int32, 1), ("mono_meas", np.float64, 2), ("mono_info", np.float64, 1),
           ("stereo_vP", np.int32, 1), ("stereo_vL", np.int32, 1), ("stereo_meas", np.float64, 3),
           ("stereo_info", np.float64, 1))


def _count(name, n):
    if name.startswith("pose") or name in ("q", "t", "cam"):
        return n[0]
    if name.startswith("lm") or name == "Xw":
        return n[1]
    return n[2] if name.startswith("mono") else n[3]


def read_graph(path):
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError("%s: not a cubagraph file" % path)
        n = np.fromfile(f, dtype=np.int64, count=4)
        g = {}
        for name, dt, width in _FIELDS:
            cnt = int(_count(name, n))
            a = np.fromfile(f, dtype=dt, count=cnt * width)
            g[name] = a.reshape(cnt, width) if width > 1 else a
    return g


def write_graph(path, g):
    with open(path, "wb") as f:
        f.write(MAGIC)
        np.array([len(g["pose_id"]), len(g["lm_id"]), len(g["mono_vP"]), len(g["stereo_vP"])], dtype=np.int64).tofile(f)
        for name, dt, _ in _FIELDS:
            np.ascontiguousarray(g[name], dtype=dt).tofile(f)


@dataclasses.dataclass
class FlatProblem:
    """Host-side flat problem = argument of cuba_engine_set_problem (include/cuba_b200.h)."""
    Pall: int
    numP: int
    Lall: int
    numL: int
    q: np.ndarray       # [Pall,4]
    t: np.ndarray       # [Pall,3]
    cam: np.ndarray     # [Pall,5]
    Xw: np.ndarray      # [Lall,3]
    idx2: np.ndarray    # [E2,2] int32 (iP,iL)
    meas2: np.ndarray   # [E2,2]
    omega2: np.ndarray  # [E2]
    idx3: np.ndarray    # [E3,2]
    meas3: np.ndarray   # [E3,3]
    omega3: np.ndarray  # [E3]
    # bookkeeping for writing results back into a graph dict
    pose_rows: np.ndarray = None   # row in graph["pose_id"] of each iP
    lm_rows: np.ndarray = None     # row in graph["lm_id"] of each iL
    mono_rows: np.ndarray = None   # row in graph mono arrays of each kept mono edge
    stereo_rows: np.ndarray = None

    @property
    def E2(self):
        return int(self.idx2.shape[0])

    @property
    def E3(self):
        return int(self.idx3.shape[0])

    @property
    def nedges(self):
        return self.E2 + self.E3

    def copy(self):
        return dataclasses.replace(self, **{f.name: (getattr(self, f.name).copy()
                                                     if isinstance(getattr(self, f.name), np.ndarray) else getattr(self, f.name))
                                            for f in dataclasses.fields(self)})


def _assign(ids, fixed, used):
    """free (ascending id) first, then fixed (ascending id); unused skipped. Returns rows in iX order, nfree."""
    order = np.argsort(ids, kind="stable")
    order = order[used[order]]
    free = order[fixed[order] == 0]
    fix = order[fixed[order] != 0]
    return np.concatenate([free, fix]).astype(np.int64), int(len(free))


def flatten(g) -> FlatProblem:
    pid, lid = g["pose_id"], g["lm_id"]
    pfix, lfix = g["pose_fixed"], g["lm_fixed"]
    # id -> row lookups (ids may be sparse)
    prow = {int(v): i for i, v in enumerate(pid)} if len(pid) and (pid.max() > 4 * len(pid) + 16) else None
    def rows_of(ids, table_ids, lut):
        if lut is not None:
            return np.array([lut[int(v)] for v in ids], dtype=np.int64)
        m = np.full(int(table_ids.max()) + 1 if len(table_ids) else 1, -1, dtype=np.int64)
        m[table_ids] = np.arange(len(table_ids))
        return m[ids]
    m_p = rows_of(g["mono_vP"], pid, prow); s_p = rows_of(g["stereo_vP"], pid, prow)
    m_l = rows_of(g["mono_vL"], lid, None); s_l = rows_of(g["stereo_vL"], lid, None)
    pused = np.zeros(len(pid), dtype=bool); lused = np.zeros(len(lid), dtype=bool)
    pused[m_p] = True; pused[s_p] = True; lused[m_l] = True; lused[s_l] = True
    prow_order, numP = _assign(pid, pfix, pused)
    lrow_order, numL = _assign(lid, lfix, lused)
    iP_of_row = np.full(len(pid), -1, dtype=np.int64); iP_of_row[prow_order] = np.arange(len(prow_order))
    iL_of_row = np.full(len(lid), -1, dtype=np.int64); iL_of_row[lrow_order] = np.arange(len(lrow_order))

    def edges(rp, rl, meas, info):
        keep = ~((pfix[rp] != 0) & (lfix[rl] != 0))
        rows = np.nonzero(keep)[0]
        idx = np.stack([iP_of_row[rp[rows]], iL_of_row[rl[rows]]], axis=1).astype(np.int32)
        return rows, np.ascontiguousarray(idx), np.ascontiguousarray(meas[rows], dtype=np.float64), \
            np.ascontiguousarray(info[rows], dtype=np.float64)

    mrows, idx2, meas2, om2 = edges(m_p, m_l, g["mono_meas"].reshape(-1, 2), g["mono_info"])
    srows, idx3, meas3, om3 = edges(s_p, s_l, g["stereo_meas"].reshape(-1, 3), g["stereo_info"])
    return FlatProblem(
        Pall=len(prow_order), numP=numP, Lall=len(lrow_order), numL=numL,
        q=np.ascontiguousarray(g["q"][prow_order], dtype=np.float64), t=np.ascontiguousarray(g["t"][prow_order], dtype=np.float64),
        cam=np.ascontiguousarray(g["cam"][prow_order], dtype=np.float64), Xw=np.ascontiguousarray(g["Xw"][lrow_order], dtype=np.float64),
        idx2=idx2.reshape(-1, 2), meas2=meas2.reshape(-1, 2), omega2=om2, idx3=idx3.reshape(-1, 2), meas3=meas3.reshape(-1, 3), omega3=om3,
        pose_rows=prow_order, lm_rows=lrow_order, mono_rows=mrows, stereo_rows=srows)


def write_back(g, prob: FlatProblem, q, t, Xw):
    """finalize(): reference src/cuda_bundle_adjustment.cpp:512-526 (fixed vertices are written back too)."""
    g["q"][prob.pose_rows] = q
    g["t"][prob.pose_rows] = t
    g["Xw"][prob.lm_rows] = Xw
