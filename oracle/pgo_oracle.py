"""CPU restatement (numpy, float64) of cuda_ba.solve_system (reference: cdvslam/fastba/ba.cpp:99-180).
TEST INFRASTRUCTURE ONLY: imported by tests/ (the product path never touches it).

Parity pinning: the reference ships no golden vector for this function and its own implementation needs Eigen (absent
here), so this oracle is pinned by construction only -- it is the literal dense form of the reference's sparse algebra:
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
