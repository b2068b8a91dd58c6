"""Differentiable bundle adjustment for the training path.

Counterpart of the reference's torch BA -- `BA()` cdvslam/ba.py:86-185 with `CholeskySolver` (:11-37) and the Jacobians of
`pops.transform(..., jacobian=True)` (cdvslam/projective_ops.py:53-108) -- which `CDVNet.forward` calls once per update
iteration with gradients flowing into the network's `targets` / `weights` (net_cdv.py:550).  Same semantics as that path
(NOT the CUDA inference path: damping `ep` + 1e-4 S on the diagonal, 250 px residual gate, explicit `bounds`, per-frame
intrinsics, `Z.clamp(min=0.1)` projection, depth clamp to [1e-3, 10], functional update instead of in place), so it is
checked against vectors produced by the verbatim reference files (tests/golden/make_golden.py), gradients included.

Like the reference's, this operator is built from batched tensor ops and relies on autograd -- one step of Gauss-Newton is
differentiated *through* (including the Jacobians' dependence on poses and depths), which a hand-derived adjoint of the
inference kernels would not give; the only custom backward is the SPD solve (`SolveSPD`, the implicit-function form of
ba.py:26-37).  It runs on whatever device the tensors live on (CUDA in training) and never synchronises with the host
unless `compact_patches=True` asks for the reference's `torch.unique(kk)` compaction (the Cholesky `info` check of the
failure convention is the one remaining synchronisation, as in the reference).

Differences in form (not in result): poses are plain `[B, F, 7]` tensors (tx ty tz qx qy qz qw, world -> camera) or any
object with a `.data` tensor of that layout (lietorch `SE3`; the same type is returned); the normal equations are
assembled as dense `[6n, 6n]` / `[6n, m]` matrices with `index_add_` instead of 6x6 block tensors and `scatter_sum`; the
number of optimised poses is taken from the pose tensor, not from `ii.max().item()` (frames without edges get a zero
step either way).
"""
import torch

MIN_DEPTH = 0.2           # projective_ops.py:6


# ---- SE3 on [., 7] tensors; quaternions are normalised on use, as lietorch's SO3 constructor does (so3.h:31-37) --------
def _cross(a, b):
    return torch.stack((a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                        a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                        a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]), -1)


def _unit(q):
    return q / q.norm(dim=-1, keepdim=True)


def _conj(q):
    return torch.cat((-q[..., :3], q[..., 3:]), -1)


def _rotate(q, p):
    u = q[..., :3]
    c = 2.0 * _cross(u, p)
    return p + q[..., 3:] * c + _cross(u, c)


def _quat_mul(a, b):
    av, aw = a[..., :3], a[..., 3:]
    bv, bw = b[..., :3], b[..., 3:]
    return torch.cat((aw * bv + bw * av + _cross(av, bv), aw * bw - (av * bv).sum(-1, keepdim=True)), -1)


def relative_pose(Gi, Gj):
    """(t, q) of Gj * Gi^-1."""
    qi, qj = _unit(Gi[..., 3:]), _unit(Gj[..., 3:])
    qi_inv = _conj(qi)
    t = Gj[..., :3] - _rotate(qj, _rotate(qi_inv, Gi[..., :3]))
    return t, _unit(_quat_mul(qj, qi_inv))


def adjoint_transpose(t, q, a):
    """Ad(G)^T a for G = (t, q), Ad = [[R, [t]x R], [0, R]], a = (tau, phi) rows."""
    qi = _conj(q)
    tau, phi = a[..., :3], a[..., 3:]
    return torch.cat((_rotate(qi, tau), _rotate(qi, phi) + _rotate(qi, _cross(tau, t))), -1)


def se3_exp(xi):
    """Exp of (tau, phi) -> (t, q); series below |phi| = 1e-6 as lietorch (so3.h:150-190, se3.h:137-145)."""
    tau, phi = xi[..., :3], xi[..., 3:]
    th2 = (phi * phi).sum(-1, keepdim=True)
    th = th2.clamp_min(1e-30).sqrt()
    small = th2 < 1e-12
    ths = torch.where(small, torch.ones_like(th), th)
    imag = torch.where(small, 0.5 - th2 / 48.0 + th2 * th2 / 3840.0, torch.sin(0.5 * ths) / ths)
    real = torch.where(small, 1.0 - th2 / 8.0 + th2 * th2 / 384.0, torch.cos(0.5 * ths))
    a = torch.where(small, 0.5 - th2 / 24.0, (1.0 - torch.cos(ths)) / (ths * ths))
    b = torch.where(small, 1.0 / 6.0 - th2 / 120.0, (ths - torch.sin(ths)) / (ths * ths * ths))
    w = _cross(phi, tau)
    return tau + a * w + b * _cross(phi, w), _unit(torch.cat((imag * phi, real), -1))


def se3_retract(G, xi):
    """Exp(xi) * G (left perturbation, lietorch `retr`)."""
    dt, dq = se3_exp(xi)
    q = _unit(G[..., 3:])
    return torch.cat((dt + _rotate(dq, G[..., :3]), _unit(_quat_mul(dq, q))), -1)


# ---- SPD solve with the reference's failure convention ----------------------------------------------------------------
class SolveSPD(torch.autograd.Function):
    """x = A^-1 b by Cholesky.  A factorisation that fails returns x = 0 and blocks the gradient, so that one degenerate
    window does not crash a training run (ba.py:13-20, 27-29).  Backward: with z = A^-1 g, dL/db = z and dL/dA = -x z^T as
    ba.py:31-35 has it (the transpose of the textbook -z x^T: the same thing for every A that is a symmetric function of
    the inputs, which S is)."""

    @staticmethod
    def forward(ctx, A, b):
        L, info = torch.linalg.cholesky_ex(A)
        ctx.failed = bool(torch.any(info))
        if ctx.failed:
            return torch.zeros_like(b)
        x = torch.cholesky_solve(b, L)
        ctx.save_for_backward(L, x)
        return x

    @staticmethod
    def backward(ctx, g):
        if ctx.failed:
            return None, None
        L, x = ctx.saved_tensors
        z = torch.cholesky_solve(g, L)
        return -torch.matmul(x, z.transpose(-1, -2)), z


# ---- projection + Jacobians (projective_ops.py:19-108) ----------------------------------------------------------------
def transform_with_jacobians(poses, patches, intrinsics, ii, jj, kk):
    """Reprojection of all P x P pixels of every edge's patch and the Jacobians of the centre pixel.
    Returns coords [B,E,P,P,2], valid [B,E], Ji [B,E,2,6], Jj [B,E,2,6], Jz [B,E,2]."""
    pk = patches[:, kk]                                            # [B,E,3,P,P]
    Ki, Kj = intrinsics[:, ii], intrinsics[:, jj]                  # [B,E,4]
    fxi, fyi, cxi, cyi = (Ki[..., c, None, None] for c in range(4))
    X0 = torch.stack(((pk[:, :, 0] - cxi) / fxi, (pk[:, :, 1] - cyi) / fyi, torch.ones_like(pk[:, :, 2]), pk[:, :, 2]), -1)
    t, q = relative_pose(poses[:, ii], poses[:, jj])               # [B,E,3], [B,E,4]
    tb, qb = t[:, :, None, None], q[:, :, None, None]
    X1 = torch.cat((_rotate(qb, X0[..., :3]) + tb * X0[..., 3:], X0[..., 3:]), -1)      # homogeneous action (se3.h:53-56)
    fxj, fyj, cxj, cyj = (Kj[..., c] for c in range(4))
    d_proj = 1.0 / X1[..., 2].clamp(min=0.1)
    coords = torch.stack((fxj[..., None, None] * (d_proj * X1[..., 0]) + cxj[..., None, None],
                          fyj[..., None, None] * (d_proj * X1[..., 1]) + cyj[..., None, None]), -1)
    c = patches.shape[-1] // 2
    X, Y, Z, H = X1[:, :, c, c].unbind(-1)
    d = torch.where(Z.abs() > MIN_DEPTH, 1.0 / torch.where(Z.abs() > MIN_DEPTH, Z, torch.ones_like(Z)), torch.zeros_like(Z))
    o = torch.zeros_like(Z)
    # Jj = Jp Ja with Jp = d proj / d X1 (rows x, y) and Ja = d (G X) / d xi at the identity
    px = torch.stack((fxj * d, o, -fxj * X * d * d), -1)           # d x / d (X, Y, Z)
    py = torch.stack((o, fyj * d, -fyj * Y * d * d), -1)
    point = torch.stack((X, Y, Z), -1)

    def pose_row(pr):                                              # pr^T [H I | -[point]x]
        return torch.cat((H[..., None] * pr, _cross(point, pr)), -1)
    Jj = torch.stack((pose_row(px), pose_row(py)), -2)
    Ji = -adjoint_transpose(t[:, :, None], q[:, :, None], Jj)
    Jz = torch.stack(((px * t).sum(-1), (py * t).sum(-1)), -1)     # Jp (t, 1): the last component of Jp is zero
    return coords, (Z > MIN_DEPTH).to(Z.dtype), Ji, Jj, Jz


def BA(poses, patches, intrinsics, targets, weights, lmbda, ii, jj, kk, bounds, ep=100.0, PRINT=False, fixedp=1,
       structure_only=False, compact_patches=False):
    """One differentiable Gauss-Newton step; returns (poses, patches) like cdvslam/ba.py:86-185.

    poses [B,F,7] (or SE3-like with `.data`), patches [B,K,3,P,P], intrinsics [B,F,4], targets / weights [B,E,2],
    lmbda float or tensor, ii / jj / kk int64 [E], bounds (x0, y0, x1, y1); the first `fixedp` poses stay fixed."""
    wrap = None
    if not isinstance(poses, torch.Tensor):
        wrap, poses = poses.__class__, poses.data
    Bn, F = poses.shape[0], poses.shape[1]
    E = ii.numel()
    coords, valid, Ji, Jj, Jz = transform_with_jacobians(poses, patches, intrinsics, ii, jj, kk)
    c = patches.shape[-1] // 2
    centre = coords[:, :, c, c]
    r = targets - centre
    inside = (centre[..., 0] > bounds[0]) & (centre[..., 1] > bounds[1]) & (centre[..., 0] < bounds[2]) & \
             (centre[..., 1] < bounds[3])
    valid = valid * (r.norm(dim=-1) < 250).to(r.dtype) * inside.to(r.dtype)
    if PRINT:
        print((r * valid[..., None]).norm(dim=-1).mean().item())
    r = valid[..., None] * r                                       # [B,E,2]
    w = valid[..., None] * weights

    n = max(F - fixedp, 0)
    if compact_patches:                                            # the reference's compaction (host sync: dynamic size)
        kx, kc = torch.unique(kk, return_inverse=True, sorted=True)
    else:
        kx, kc = torch.arange(patches.shape[1], device=kk.device), kk
    m = kx.numel()
    wJz = w * Jz
    C = torch.zeros((Bn, m), dtype=r.dtype, device=r.device).index_add_(1, kc, (wJz * Jz).sum(-1))
    g_z = torch.zeros((Bn, m), dtype=r.dtype, device=r.device).index_add_(1, kc, (wJz * r).sum(-1))
    lam = lmbda
    if isinstance(lam, torch.Tensor) and lam.numel() > 1:
        lam = lam.reshape(C.shape)
    Q = 1.0 / (C + lam)

    if structure_only or n == 0:
        dZ = Q * g_z
        dX = None
    else:
        i0, j0 = ii - fixedp, jj - fixedp
        ok_i, ok_j = i0 >= 0, j0 >= 0                              # indices never exceed n - 1 (n is taken from `poses`)
        wJi, wJj = w[..., None] * Ji, w[..., None] * Jj            # [B,E,2,6]
        n6 = 6 * n

        six = torch.arange(6, device=ii.device)

        def add_blocks(dst, rows, cols, blocks, mask):             # dst [B, n6*n6] += 6x6 blocks at (rows, cols)
            # masked-out edges (a fixed pose) add zeros at block (0, 0): no filtering, hence no host synchronisation
            rr, cc = torch.where(mask, rows, 0), torch.where(mask, cols, 0)
            base = (6 * rr[:, None, None] + six[None, :, None]) * n6 + 6 * cc[:, None, None] + six[None, None, :]
            return dst.index_add_(1, base.reshape(-1), (blocks * mask[None, :, None, None].to(blocks.dtype)).reshape(Bn, -1))
        Hp = torch.zeros((Bn, n6 * n6), dtype=r.dtype, device=r.device)
        Hp = add_blocks(Hp, i0, i0, torch.matmul(wJi.transpose(2, 3), Ji), ok_i)
        Hp = add_blocks(Hp, i0, j0, torch.matmul(wJi.transpose(2, 3), Jj), ok_i & ok_j)
        Hp = add_blocks(Hp, j0, i0, torch.matmul(wJj.transpose(2, 3), Ji), ok_i & ok_j)
        Hp = add_blocks(Hp, j0, j0, torch.matmul(wJj.transpose(2, 3), Jj), ok_j)
        Hp = Hp.view(Bn, n6, n6)

        def add_rows(dst, rows, cols, vals, mask):                 # dst [B, n6*ncol] += 6-vectors at (rows, cols)
            ncol = dst.shape[1] // n6
            rr = torch.where(mask, rows, 0)
            base = (6 * rr[:, None] + six[None, :]) * ncol + cols[:, None]
            return dst.index_add_(1, base.reshape(-1), (vals * mask[None, :, None].to(vals.dtype)).reshape(Bn, -1))
        Ei = (wJi * Jz[..., None]).sum(2)                          # [B,E,6]  = (w Ji)^T Jz
        Ej = (wJj * Jz[..., None]).sum(2)
        Em = torch.zeros((Bn, n6 * m), dtype=r.dtype, device=r.device)
        Em = add_rows(Em, i0, kc, Ei, ok_i)
        Em = add_rows(Em, j0, kc, Ej, ok_j).view(Bn, n6, m)
        zero_col = torch.zeros_like(ii)
        g_x = torch.zeros((Bn, n6), dtype=r.dtype, device=r.device)
        g_x = add_rows(g_x, i0, zero_col, (wJi * r[..., None]).sum(2), ok_i)
        g_x = add_rows(g_x, j0, zero_col, (wJj * r[..., None]).sum(2), ok_j)

        EQ = Em * Q[:, None, :]
        S = Hp - torch.matmul(EQ, Em.transpose(1, 2))
        y = g_x - torch.matmul(EQ, g_z[..., None])[..., 0]
        S = S + torch.diag_embed(ep + 1e-4 * torch.diagonal(S, dim1=1, dim2=2))         # block_solve, ba.py:66-76
        dX = SolveSPD.apply(S, y[..., None])[..., 0]               # [B, 6n]
        dZ = Q * (g_z - torch.matmul(Em.transpose(1, 2), dX[..., None])[..., 0])

    disps = patches[:, :, 2]
    step = torch.zeros((Bn, patches.shape[1]), dtype=r.dtype, device=r.device).index_add_(1, kx, dZ)
    disps = (disps + step[:, :, None, None]).clamp(min=1e-3, max=10.0)
    patches = torch.stack((patches[:, :, 0], patches[:, :, 1], disps), 2)
    if dX is not None:
        moved = se3_retract(poses[:, fixedp:], dX.view(Bn, n, 6))
        poses = torch.cat((poses[:, :fixedp], moved), 1)
    return (wrap(poses) if wrap is not None else poses), patches
