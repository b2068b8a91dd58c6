"""Host-side checks of the pose-graph oracle (no GPU): the dense restatement satisfies the normal equations it is
defined by, honours `freen`, and the block assembly equals the explicit J^T J."""
import numpy as np

from oracle import pgo_oracle


def test_oracle_solves_its_normal_equations():
    rng = np.random.default_rng(0)
    n, r = 7, 10
    ii = rng.integers(1, n, r); jj = np.array([rng.integers(0, i) for i in ii])
    J_i = rng.standard_normal((r, 7, 7)).astype(np.float32); J_j = rng.standard_normal((r, 7, 7)).astype(np.float32)
    res = rng.standard_normal((r, 7)).astype(np.float32)
    delta, A, b = pgo_oracle.solve_system(J_i, J_j, ii, jj, res, 1e-3, 1e-4, -1)
    n = int(max(ii.max(), jj.max())) + 1
    assert delta.shape == (n, 7)
    assert np.abs(A @ delta.reshape(-1).astype(np.float64) - b).max() < 1e-4 * np.abs(b).max()
    d2, _, _ = pgo_oracle.solve_system(J_i, J_j, ii, jj, res, 1e-3, 1e-4, 3)
    assert (d2.reshape(-1)[21:] == 0).all() and np.abs(d2).max() > 0
    assert np.abs(A[:21, :21] @ d2.reshape(-1)[:21].astype(np.float64) - b[:21]).max() < 1e-4 * np.abs(b).max()


def test_dense_oracle_equals_independent_sparse_solve():
    """Pins pgo_oracle.solve_system: the dense restatement against an implementation of the same algorithm built the way
    the reference builds it (sparse triplets -> J^T J -> sparse direct solve in double), chain + loop-closure graphs, all
    `freen` modes."""
    for n, loops, freen, seed in ((30, 5, -1, 0), (120, 15, -1, 1), (120, 15, 90, 2), (25, 3, 0, 3)):
        rng = np.random.default_rng(seed)
        kk = np.arange(1, n); ll = kk - 1
        li = rng.integers(8, n, loops); lj = np.array([rng.integers(0, i - 5) for i in li])
        ii = np.concatenate([kk, li]); jj = np.concatenate([ll, lj])
        r = len(ii)
        J_i = (np.eye(7)[None] + 0.1 * rng.standard_normal((r, 7, 7))).astype(np.float32)
        J_j = (-np.eye(7)[None] + 0.1 * rng.standard_normal((r, 7, 7))).astype(np.float32)
        res = (0.05 * rng.standard_normal((r, 7))).astype(np.float32)
        dense, A, b = pgo_oracle.solve_system(J_i, J_j, ii, jj, res, 1e-3, 1e-6, freen)
        sparse = pgo_oracle.solve_system_sparse(J_i, J_j, ii, jj, res, 1e-3, 1e-6, freen)
        assert dense.shape == sparse.shape == (n, 7)
        if freen == 0:
            assert (dense == 0).all() and (sparse == 0).all()
            continue
        assert np.abs(dense - sparse).max() <= 2e-6 * np.abs(sparse).max()
