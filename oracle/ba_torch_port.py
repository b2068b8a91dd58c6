"""torch/CPU restatement of the reference's *torch* bundle adjustment (oracle; test infrastructure only).

This is the "port" that bench.py times as `cpu_baseline` / `--impl reference`: the reference's CPU-runnable BA is
cdvslam/ba.py:86-185 driving cdvslam/projective_ops.py:53-113 (`transform(..., jacobian=True)`), which cannot be
imported on the GPU box (it needs torch_scatter and the lietorch C++ backend).  The
restatement keeps the reference's op structure -- batched per-edge Jacobians, 6x6 block products by matmul, block
scatter-adds, dense Schur complement, Cholesky -- so that its CPU time is representative:

  transform_jac()  <- projective_ops.py:19-29 (iproj), :32-50 (proj, Z.clamp(min=0.1)), :53-108 (Jacobians)
  ba_torch()       <- ba.py:86-185 (residual gate 250 px :98, bounds :100-104, scatter :140-153, Q :161,
                      Schur :170-171, block_solve with (ep + lm*A)*I :66-76, retractions :49-56, clamp :179)

SE3 helpers follow lietorch/include/se3.h:36-95 and so3.h:30-170 (quaternion normalised on construction).
tests/test_oracle_golden.py checks this port against outputs of the verbatim reference files.
"""
import torch


def _cross(a, b):
    return torch.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                        a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                        a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], -1)


def _rot(q, X):
    uv = 2.0 * _cross(q[..., :3], X)
    return X + q[..., 3:4] * uv + _cross(q[..., :3], uv)


def _qmul(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                        aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz], -1)


def _normq(q):
    return q / q.norm(dim=-1, keepdim=True)


def se3_rel(Gi, Gj):
    """Gj * Gi^-1 for [.,7] pose tensors."""
    ti, qi = Gi[..., :3], _normq(Gi[..., 3:])
    tj, qj = Gj[..., :3], _normq(Gj[..., 3:])
    qi_inv = torch.cat([-qi[..., :3], qi[..., 3:]], -1)
    qij = _normq(_qmul(qj, qi_inv))
    tij = tj + _rot(qj, -_rot(qi_inv, ti))
    return tij, qij


def se3_adjT(t, q, a):
    """Ad(G)^T a with Ad = [[R, [t]x R], [0, R]] (se3.h:57-66, 84-86)."""
    qinv = torch.cat([-q[..., :3], q[..., 3:]], -1)
    top = _rot(qinv, a[..., :3])
    bot = _rot(qinv, a[..., 3:]) + _rot(qinv, _cross(a[..., :3], t))
    return torch.cat([top, bot], -1)


def se3_exp(xi):
    tau, phi = xi[..., :3], xi[..., 3:]
    th2 = (phi * phi).sum(-1, keepdim=True)
    th = th2.sqrt()
    small = th < 1e-6
    s = torch.where(small, torch.ones_like(th), th)
    imag = torch.where(small, 0.5 - th2 / 48.0 + th2 * th2 / 3840.0, torch.sin(0.5 * s) / s)
    real = torch.where(small, 1.0 - th2 / 8.0 + th2 * th2 / 384.0, torch.cos(0.5 * s))
    q = _normq(torch.cat([imag * phi, real], -1))
    c1 = torch.where(small, 0.5 - th2 / 24.0, (1 - torch.cos(s)) / (s * s))
    c2 = torch.where(small, 1.0 / 6.0 - th2 / 120.0, (s - torch.sin(s)) / (s * s * s))
    px = _cross(phi, tau)
    t = tau + c1 * px + c2 * _cross(phi, px)
    return t, q


def se3_retr(poses, dx):
    """Exp(dx) * poses (groups.py `retr`)."""
    dt, dq = se3_exp(dx)
    t, q = poses[..., :3], _normq(poses[..., 3:])
    return torch.cat([_rot(dq, t) + dt, _normq(_qmul(dq, q))], -1)


def transform_jac(poses, patches, intrinsics, ii, jj, kk):
    """coords [1,E,P,P,2], valid [1,E], (Ji, Jj [1,E,2,6], Jz [1,E,2,1]) as projective_ops.transform(jacobian=True)."""
    pk = patches[:, kk]
    x, y, d = pk.unbind(dim=2)
    fx, fy, cx, cy = intrinsics[:, ii][..., None, None].unbind(dim=2)
    X0 = torch.stack([(x - cx) / fx, (y - cy) / fy, torch.ones_like(d), d], dim=-1)      # [1,E,P,P,4]

    tij, qij = se3_rel(poses[:, ii], poses[:, jj])
    R3 = _rot(qij[:, :, None, None], X0[..., :3]) + tij[:, :, None, None] * X0[..., 3:4]
    X1 = torch.cat([R3, X0[..., 3:4]], -1)

    fxj, fyj, cxj, cyj = intrinsics[:, jj][..., None, None].unbind(dim=2)
    dd = 1.0 / X1[..., 2].clamp(min=0.1)
    coords = torch.stack([fxj * (dd * X1[..., 0]) + cxj, fyj * (dd * X1[..., 1]) + cyj], dim=-1)

    p = X1.shape[2]
    X, Y, Z, H = X1[..., p // 2, p // 2, :].unbind(dim=-1)
    o = torch.zeros_like(H)
    fx, fy, cx, cy = intrinsics[:, jj].unbind(dim=-1)
    d = torch.zeros_like(Z)
    big = Z.abs() > 0.2
    d[big] = 1.0 / Z[big]
    Ja = torch.stack([H, o, o, o, Z, -Y,
                      o, H, o, -Z, o, X,
                      o, o, H, Y, -X, o,
                      o, o, o, o, o, o], dim=-1).view(1, len(ii), 4, 6)
    Jp = torch.stack([fx * d, o, -fx * X * d * d, o,
                      o, fy * d, -fy * Y * d * d, o], dim=-1).view(1, len(ii), 2, 4)
    Jj = torch.matmul(Jp, Ja)
    Ji = -se3_adjT(tij[:, :, None], qij[:, :, None], Jj)
    T3 = torch.cat([tij, torch.ones_like(tij[..., :1])], -1)[..., None]                 # last column of Gij.matrix()
    Jz = torch.matmul(Jp, T3)
    return coords, (Z > 0.2).float(), (Ji, Jj, Jz)


def _scatter_mat(A, ii, jj, n, m):
    v = (ii >= 0) & (jj >= 0) & (ii < n) & (jj < m)
    out = torch.zeros(A.shape[0], n * m, *A.shape[2:], dtype=A.dtype)
    out.index_add_(1, (ii[v] * m + jj[v]), A[:, v])
    return out


def _scatter_vec(b, ii, n):
    v = (ii >= 0) & (ii < n)
    out = torch.zeros(b.shape[0], n, *b.shape[2:], dtype=b.dtype)
    out.index_add_(1, ii[v], b[:, v])
    return out


def _block_matmul(A, B):
    b, n1, m1, p1, q1 = A.shape
    b, n2, m2, p2, q2 = B.shape
    A = A.permute(0, 1, 3, 2, 4).reshape(b, n1 * p1, m1 * q1)
    B = B.permute(0, 1, 3, 2, 4).reshape(b, n2 * p2, m2 * q2)
    return torch.matmul(A, B).reshape(b, n1, p1, m2, q2).permute(0, 1, 3, 2, 4)


def _block_solve(A, B, ep, lm):
    b, n1, m1, p1, q1 = A.shape
    b, n2, m2, p2, q2 = B.shape
    A = A.permute(0, 1, 3, 2, 4).reshape(b, n1 * p1, m1 * q1)
    B = B.permute(0, 1, 3, 2, 4).reshape(b, n2 * p2, m2 * q2)
    A = A + (ep + lm * A) * torch.eye(n1 * p1, dtype=A.dtype)
    U, info = torch.linalg.cholesky_ex(A)
    X = torch.zeros_like(B) if bool(info.any()) else torch.cholesky_solve(B, U)
    return X.reshape(b, n1, p1, m2, q2).permute(0, 1, 3, 2, 4)


def ba_torch(poses, patches, intrinsics, targets, weights, lmbda, ii, jj, kk, bounds, ep=100.0, fixedp=1,
             structure_only=False):
    """One Gauss-Newton step; returns new (poses [1,F,7], patches [1,K,3,P,P])  (ba.py:86-185)."""
    b = 1
    n = int(max(ii.max().item(), jj.max().item())) + 1
    coords, v, (Ji, Jj, Jz) = transform_jac(poses, patches, intrinsics, ii, jj, kk)
    p = coords.shape[3]
    c = coords[..., p // 2, p // 2, :]
    r = targets - c
    v = v * (r.norm(dim=-1) < 250).float()
    v = v * ((c[..., 0] > bounds[0]) & (c[..., 1] > bounds[1]) & (c[..., 0] < bounds[2]) & (c[..., 1] < bounds[3])).float()
    r = (v[..., None] * r).unsqueeze(-1)
    weights = (v[..., None] * weights).unsqueeze(-1)

    wJiT = (weights * Ji).transpose(2, 3)
    wJjT = (weights * Jj).transpose(2, 3)
    wJzT = (weights * Jz).transpose(2, 3)
    Bii, Bij = torch.matmul(wJiT, Ji), torch.matmul(wJiT, Jj)
    Bji, Bjj = torch.matmul(wJjT, Ji), torch.matmul(wJjT, Jj)
    Eik, Ejk = torch.matmul(wJiT, Jz), torch.matmul(wJjT, Jz)
    vi, vj = torch.matmul(wJiT, r), torch.matmul(wJjT, r)

    n = n - fixedp
    ii = ii - fixedp
    jj = jj - fixedp
    kx, kk = torch.unique(kk, return_inverse=True, sorted=True)
    m = len(kx)

    B = (_scatter_mat(Bii, ii, ii, n, n) + _scatter_mat(Bij, ii, jj, n, n) +
         _scatter_mat(Bji, jj, ii, n, n) + _scatter_mat(Bjj, jj, jj, n, n)).view(b, n, n, 6, 6)
    E = (_scatter_mat(Eik, ii, kk, n, m) + _scatter_mat(Ejk, jj, kk, n, m)).view(b, n, m, 6, 1)
    C = _scatter_vec(torch.matmul(wJzT, Jz), kk, m)
    vv = (_scatter_vec(vi, ii, n) + _scatter_vec(vj, jj, n)).view(b, n, 1, 6, 1)
    w = _scatter_vec(torch.matmul(wJzT, r), kk, m)

    if isinstance(lmbda, torch.Tensor):
        lmbda = lmbda.item() if lmbda.numel() == 1 else lmbda.reshape(*C.shape)
    Q = 1.0 / (C + lmbda)
    EQ = E * Q[:, None]

    if structure_only or n == 0:
        dZ = (Q * w).view(b, -1, 1, 1)
        dX = None
    else:
        Et = E.permute(0, 2, 1, 4, 3)
        S = B - _block_matmul(EQ, Et)
        y = vv - _block_matmul(EQ, w.unsqueeze(2))
        dX = _block_solve(S, y, ep, 1e-4)
        dZ = (Q * (w - _block_matmul(Et, dX).squeeze(-1))).view(b, -1, 1, 1)
        dX = dX.reshape(b, -1, 6)

    x, y_, disps = patches.unbind(dim=2)
    upd = torch.zeros_like(disps)
    upd.index_add_(1, kx, dZ.expand(-1, -1, *disps.shape[2:]).contiguous())
    disps = (disps + upd).clamp(min=1e-3, max=10.0)
    patches = torch.stack([x, y_, disps], dim=2)
    if dX is not None:
        # functional (no in-place write into a tensor whose slice fed the retraction: keeps the port differentiable)
        poses = torch.cat([poses[:, :fixedp], se3_retr(poses[:, fixedp:fixedp + n], dX), poses[:, fixedp + n:]], 1)
    return poses, patches


def run(problem_tensors, iterations=2, ep=1.0):
    """Convenience wrapper used by the bench CPU baseline: `iterations` GN steps on one problem."""
    poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, bounds = problem_tensors
    for _ in range(iterations):
        poses, patches = ba_torch(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, bounds,
                                  ep=ep, fixedp=t0)
    return poses, patches
