"""The premise of the two-level PCG (csrc/cuba_pcg4.cuh), checked on the CPU with scipy on the reduced system the oracle
assembles: adding the coarse correction over rigid-motion aggregates, M^-1 = D^-1 + Z (Z^T S Z)^-1 Z^T with Z_i = Ad(T_i), cuts the
iteration count of block-Jacobi PCG several times at a low damping and converges to the same solution; a coarse inverse taken
at a 9x larger damping preconditions as well as the current one (the engine rebuilds it only now and then)."""
import numpy as np
import pytest

sp = pytest.importorskip("scipy.sparse")


def _system(o, P, lam):
    assert o.solve(lam)
    Hsc, bsc, _ = o.schur()
    rp, ci = o.hsc_structure()
    B = Hsc.reshape(-1, 6, 6).transpose(0, 2, 1)            # column-major blocks -> [k][r][c]
    rows = np.repeat(np.arange(P), np.diff(rp))
    rr, cc = np.meshgrid(np.arange(6), np.arange(6), indexing="ij")
    I = (6 * rows[:, None, None] + rr).ravel(); J = (6 * ci[:, None, None] + cc).ravel(); V = B.ravel()
    off = np.repeat(rows != ci, 36)
    A = sp.csr_matrix((np.concatenate([V, V[off]]), (np.concatenate([I, J[off]]), np.concatenate([J, I[off]]))), shape=(6 * P, 6 * P))
    return A, bsc.reshape(-1).copy()


def _pcg(A, b, Minv, tol=1e-11, maxit=5000):
    x = np.zeros_like(b); r = b.copy(); z = Minv(r); p = z.copy(); rz = r @ z; rz0 = rz; it = 0
    while it < maxit:
        Ap = A @ p; al = rz / (p @ Ap); x += al * p; r -= al * Ap; z = Minv(r); rzn = r @ z; it += 1
        if rzn <= tol * tol * rz0:
            break
        p = z + (rzn / rz) * p; rz = rzn
    return x, it


def _adjoints(prob, P):
    out = []
    for i in range(P):
        x, y, z, w = prob.q[i]; t = prob.t[i]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        K = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        Ad = np.zeros((6, 6)); Ad[:3, :3] = R; Ad[3:, 3:] = R; Ad[3:, :3] = K @ R      # delta = [omega; upsilon]
        out.append(Ad)
    return out


def test_coarse_correction_cuts_the_iteration_count(pkg, oracle, problems):
    prob = problems("kitti07_shaped")
    P = prob.numP
    o = oracle.Oracle(prob, (0, 0), (0.0, 0.0))
    o.compute_errors(); o.build_system()
    md = o.max_diagonal()
    m = 8                                                   # poses per aggregate
    na = -(-P // m)
    adj = _adjoints(prob, P)
    rr, cc = np.meshgrid(np.arange(6), np.arange(6), indexing="ij")
    I = np.concatenate([(6 * i + rr).ravel() for i in range(P)]); J = np.concatenate([(6 * (i // m) + cc).ravel() for i in range(P)])
    Z = sp.csr_matrix((np.concatenate([a.ravel() for a in adj]), (I, J)), shape=(6 * P, 6 * na))

    def block_jacobi(A):
        D = sp.block_diag([sp.csr_matrix(np.linalg.inv(A[6 * i:6 * i + 6, 6 * i:6 * i + 6].toarray())) for i in range(P)], format="csr")
        return lambda r: D @ r

    lam = 1e-9 * md
    A, b = _system(o, P, lam)
    bj = block_jacobi(A)
    x0, it_bj = _pcg(A, b, bj)
    Aci = np.linalg.inv((Z.T @ A @ Z).toarray())
    x1, it_tl = _pcg(A, b, lambda r: bj(r) + Z @ (Aci @ (Z.T @ r)))
    assert np.abs(x1 - x0).max() <= 1e-6 * np.abs(x0).max()
    assert it_tl * 3 < it_bj, (it_tl, it_bj)
    # a coarse inverse from a 9x larger damping is as good
    A9, _ = _system(o, P, 9 * lam)
    Aci9 = np.linalg.inv((Z.T @ A9 @ Z).toarray())
    x2, it_stale = _pcg(A, b, lambda r: bj(r) + Z @ (Aci9 @ (Z.T @ r)))
    assert it_stale <= 1.25 * it_tl + 5, (it_stale, it_tl)
    assert np.abs(x2 - x0).max() <= 1e-6 * np.abs(x0).max()
