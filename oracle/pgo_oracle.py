"""CPU restatement (numpy, float64) of cuda_ba.solve_system (reference: cdvslam/fastba/ba.cpp:99-180).
TEST INFRASTRUCTURE ONLY: imported by tests/ (the product path never touches it).

Parity pinning: the reference ships no golden vector for this function and its own implementation needs Eigen (absent
here).  The dense restatement below is pinned (tests/test_pgo_oracle_cpu.py) against solve_system_sparse(), an independent
implementation of the same published algorithm the reference delegates to Eigen (sparse triplet assembly of J,
A = J^T J as a sparse product, sparse direct factorisation in double: scipy's SuperLU standing in for Eigen's
SimplicialCholesky, which for an SPD system computes the same unique solution).  Dense form of the reference's algebra:
J [7r x 7n] from the two 7x7 blocks per residual (ba.cpp:140-156), b = -J^T res, A = J^T J (:160-162),
A.diag += A.diag * lm; A.diag += ep (:164-165), solve the leading 7*freen block, or all of A when 7*freen < 0
(:101-118, :166).  lm and ep are C floats promoted to double, as in the reference."""
import numpy as np


def solve_system(J_Ginv_i, J_Ginv_j, ii, jj, res, ep, lm, freen):
    J_i = np.asarray(J_Ginv_i, np.float32).astype(np.float64)
    J_j = np.asarray(J_Ginv_j, np.float32).astype(np.float64)
    ii = np.asarray(ii, np.int64); jj = np.asarray(jj, np.int64)
    v = np.asarray(res, np.float32).astype(np.float64).reshape(-1)
    r = J_i.shape[0]
    n = int(max(ii.max(), jj.max())) + 1
    if (ii == jj).any():
        raise ValueError("self edge")                    # the reference calls exit(1) (ba.cpp:151-152)
    J = np.zeros((r * 7, n * 7))
    for x in range(r):
        J[x * 7:(x + 1) * 7, ii[x] * 7:(ii[x] + 1) * 7] += J_i[x]
        J[x * 7:(x + 1) * 7, jj[x] * 7:(jj[x] + 1) * 7] += J_j[x]
    b = -(J.T @ v)
    A = J.T @ J
    d = np.arange(n * 7)
    A[d, d] += A[d, d] * np.float64(np.float32(lm))
    A[d, d] += np.float64(np.float32(ep))
    m = freen * 7
    delta = np.zeros(n * 7)
    if m < 0:
        delta = np.linalg.solve(A, b)
    elif m > 0:
        m = min(m, n * 7)
        delta[:m] = np.linalg.solve(A[:m, :m], b[:m])
    return delta.astype(np.float32).reshape(n, 7), A, b


def solve_system_sparse(J_Ginv_i, J_Ginv_j, ii, jj, res, ep, lm, freen):
    """Second opinion, sharing no code with solve_system(): the reference's own data flow (ba.cpp:136-166) -- triplets
    (x*7+k, i*7+l, J_i[x][k][l]) / (x*7+k, j*7+l, J_j[x][k][l]) -> sparse J, b = -J^T v, A = J^T J sparse, diagonal
    damping, sparse direct solve of the leading block in double."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    J_i = np.asarray(J_Ginv_i, np.float32); J_j = np.asarray(J_Ginv_j, np.float32)
    ii = np.asarray(ii, np.int64); jj = np.asarray(jj, np.int64)
    r, n = len(ii), int(max(ii.max(), jj.max())) + 1
    x, k, l = np.meshgrid(np.arange(r), np.arange(7), np.arange(7), indexing="ij")
    rows = (x * 7 + k).reshape(-1)
    J = sp.coo_matrix((np.concatenate([J_i.reshape(-1), J_j.reshape(-1)]).astype(np.float64),
                       (np.concatenate([rows, rows]),
                        np.concatenate([(ii[x] * 7 + l).reshape(-1), (jj[x] * 7 + l).reshape(-1)]))),
                      shape=(r * 7, n * 7)).tocsr()
    b = -(J.T @ np.asarray(res, np.float32).reshape(-1).astype(np.float64))
    A = (J.T @ J).tolil()
    d = A.diagonal()
    A.setdiag(d + d * np.float64(np.float32(lm)) + np.float64(np.float32(ep)))
    m = n * 7 if freen < 0 else min(freen * 7, n * 7)
    delta = np.zeros(n * 7)
    if m > 0:
        delta[:m] = spla.spsolve(A.tocsc()[:m, :m], b[:m])
    return delta.astype(np.float32).reshape(n, 7)
