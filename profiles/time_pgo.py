"""cuda_ba.solve_system (pgo.cu) on a 1000-pose graph (chain + 200 loop edges, 7000 unknowns): device time per call next
to the same normal equations solved on the host with scipy's sparse LU (the reference uses Eigen's sparse Cholesky on
the CPU, ba.cpp:99-118), and the agreement of the two solutions.  Usage: python profiles/time_pgo.py"""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "cdv-slam_b200")]
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch
import cuda_ba


def graph(n, n_loops, seed, noise=0.1):
    rng = np.random.default_rng(seed)
    kk = np.arange(1, n); ll = kk - 1
    li = rng.integers(40, n, n_loops); lj = np.array([rng.integers(0, i - 30) for i in li])
    ii = np.concatenate([kk, li]).astype(np.int64); jj = np.concatenate([ll, lj]).astype(np.int64)
    r = len(ii)
    J_i = (np.eye(7)[None] + noise * rng.standard_normal((r, 7, 7))).astype(np.float32)
    J_j = (-np.eye(7)[None] + noise * rng.standard_normal((r, 7, 7))).astype(np.float32)
    res = (0.05 * rng.standard_normal((r, 7))).astype(np.float32)
    return J_i, J_j, ii, jj, res


def host_sparse(J_i, J_j, ii, jj, res, ep, lm):
    r, n = len(ii), int(max(ii.max(), jj.max())) + 1
    rows = np.repeat(np.arange(r * 7), 7)
    cols_i = (ii[:, None, None] * 7 + np.arange(7)[None, None, :]).repeat(7, 1).reshape(-1)
    cols_j = (jj[:, None, None] * 7 + np.arange(7)[None, None, :]).repeat(7, 1).reshape(-1)
    J = sp.csr_matrix((np.concatenate([J_i.reshape(-1), J_j.reshape(-1)]).astype(np.float64),
                       (np.concatenate([rows, rows]), np.concatenate([cols_i, cols_j]))), shape=(r * 7, n * 7))
    A = (J.T @ J).tocsc()
    d = A.diagonal()
    A.setdiag(d + d * np.float64(np.float32(lm)) + np.float64(np.float32(ep)))
    b = -(J.T @ res.reshape(-1).astype(np.float64))
    return spla.spsolve(A, b).reshape(n, 7)


def main():
    n, loops, ep, lm = 1000, 200, 1e-3, 1e-6
    J_i, J_j, ii, jj, res = graph(n, loops, 7)
    t = lambda a: torch.as_tensor(a, device="cuda")
    args = (t(J_i), t(J_j), t(ii), t(jj), t(res), ep, lm, -1)
    for _ in range(2):
        got, = cuda_ba.solve_system(*args)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); got, = cuda_ba.solve_system(*args); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t0 = time.perf_counter(); want = host_sparse(J_i, J_j, ii, jj, res, ep, lm); t_host = time.perf_counter() - t0
    err = float(np.abs(got.cpu().numpy() - want).max() / np.abs(want).max())
    print(json.dumps({"poses": n, "loop_edges": loops, "unknowns": 7 * n, "gpu_ms": float(np.median(ts)),
                      "host_scipy_sparse_ms": 1e3 * t_host, "rel_diff": err}))


if __name__ == "__main__":
    main()
