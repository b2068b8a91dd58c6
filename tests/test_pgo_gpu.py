"""cuda_ba.solve_system on the device (pgo.cu) against the numpy float64 restatement of ba.cpp:99-180.
Both sides work in double on the same float32 inputs; the bound is the forward error of a double solve,
cond(A) * 2^-52 with a safety factor, plus the float32 rounding of the returned delta."""
import numpy as np
import pytest
import torch

import cuda_ba
from oracle import pgo_oracle

pytestmark = pytest.mark.gpu


def _graph(n, n_loops, seed, noise=0.1):
    """Chain edges (k, k-1) followed by loop edges (i, j), j < i - 5, like optim_utils.residual builds them; Jacobian
    blocks near +-identity (the Sim3 residual Jacobians are adjoint-like)."""
    rng = np.random.default_rng(seed)
    kk = np.arange(1, n); ll = kk - 1
    li = rng.integers(8, n, n_loops); lj = np.array([rng.integers(0, i - 5) for i in li])
    ii = np.concatenate([kk, li]).astype(np.int64); jj = np.concatenate([ll, lj]).astype(np.int64)
    r = len(ii)
    J_i = (np.eye(7)[None] + noise * rng.standard_normal((r, 7, 7))).astype(np.float32)
    J_j = (-np.eye(7)[None] + noise * rng.standard_normal((r, 7, 7))).astype(np.float32)
    res = (0.05 * rng.standard_normal((r, 7))).astype(np.float32)
    return J_i, J_j, ii, jj, res


@pytest.mark.parametrize("n,n_loops,ep,lm,freen", [(12, 3, 1e-3, 1e-4, -1), (40, 6, 1e-2, 1e-6, -1), (40, 6, 1e-3, 1e-4, 25),
                                                   (150, 20, 1e-3, 1e-6, -1), (150, 20, 0.0, 1e-6, 120), (9, 2, 1e-3, 0.0, 0)])
def test_solve_system_matches_oracle(n, n_loops, ep, lm, freen):
    J_i, J_j, ii, jj, res = _graph(n, n_loops, seed=n + n_loops)
    want, A, b = pgo_oracle.solve_system(J_i, J_j, ii, jj, res, ep, lm, freen)
    t = lambda a: torch.as_tensor(a, device="cuda")
    got, = cuda_ba.solve_system(t(J_i), t(J_j), t(ii), t(jj), t(res), ep, lm, freen)
    assert got.shape == (n, 7) and got.dtype == torch.float32 and got.is_cuda
    got = got.cpu().numpy()
    m = n * 7 if freen < 0 else min(freen * 7, n * 7)
    assert (got.reshape(-1)[m:] == 0).all()
    if m == 0:
        return
    cond = np.linalg.cond(A[:m, :m])
    tol = max(1e-6, 50 * cond * 2.0 ** -52) + 2.0 ** -23
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= tol * scale, (np.abs(got - want).max() / scale, cond)
    assert scale > 1e-4


def test_solve_system_rejects_self_edges():
    J_i, J_j, ii, jj, res = _graph(10, 2, seed=1)
    t = lambda a: torch.as_tensor(a, device="cuda")
    jj2 = jj.copy(); jj2[3] = ii[3]
    with pytest.raises(RuntimeError):
        cuda_ba.solve_system(t(J_i), t(J_j), t(ii), t(jj2), t(res), 1e-3, 1e-4, -1)
