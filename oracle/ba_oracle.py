"""float64 numpy restatement of the reference's CUDA bundle adjustment (oracle; test infrastructure only).

Follows cuda_ba() and its kernels in the reference, cdvslam/fastba/ba_cuda.cu:
  * per-edge residual / Jacobians / masks ........ ba_cuda.cu:265-343  (`linearize_edges`)
  * normal-equation assembly B, E, C, v, u ....... ba_cuda.cu:346-403  (`assemble`)
  * Q, Schur complement, damping, solve .......... ba_cuda.cu:548-592  (`schur_solve`)
  * structure-only branch (t1 == t0) ............. ba_cuda.cu:550-560
  * pose / inverse-depth retraction .............. ba_cuda.cu:178-229
The block-sparse `eff_impl=True` storage of the reference (block_e.cu:43-300) is only a storage change
(SURVEY.md appendix C item 11); the oracle keeps E as a scipy CSR matrix so that the same code covers the
window case (config c2) and the 1000-frame global case (c4) and therefore *is* the parity target for both.

Everything is computed in `dtype` (float64 by default).  Passing float32 gives a CPU estimate of the fp32
noise floor of the reference arithmetic, which tests use to justify tolerances.
"""
import numpy as np
import scipy.linalg
import scipy.sparse as sp

from . import se3_np as se3


def linearize_edges(poses, patches, intrinsics, target, weight, ii, jj, kk, dtype=np.float64):
    """Per-edge quantities exactly as the reference kernel forms them before accumulation.

    poses [F,7], patches [K,3,P,P], intrinsics [F,4] (row 0 only is used: ba_cuda.cu:253-259),
    target/weight [E,2], ii/jj/kk int [E].
    Returns dict with r [E,2], w [E,2] (mask*weight), Ji/Jj [E,2,6], Jz [E,2], coords [E,2], mask [E].
    """
    poses = np.asarray(poses, dtype)
    patches = np.asarray(patches, dtype)
    fx, fy, cx, cy = (dtype(x) for x in np.asarray(intrinsics, dtype).reshape(-1, 4)[0])
    target = np.asarray(target, dtype).reshape(-1, 2)
    weight = np.asarray(weight, dtype).reshape(-1, 2)

    ti, qi = poses[ii, :3], poses[ii, 3:]
    tj, qj = poses[jj, :3], poses[jj, 3:]
    tij, qij = se3.rel_se3(ti, qi, tj, qj)

    # patch centre [1][1] (ba_cuda.cu:282-285)
    px = patches[kk, 0, 1, 1]
    py = patches[kk, 1, 1, 1]
    pd = patches[kk, 2, 1, 1]
    Xi = np.stack([(px - cx) / fx, (py - cy) / fy, np.ones_like(px), pd], axis=-1)
    Xj = se3.act_se3(tij, qij, Xi)
    X, Y, Z, W = (Xj[:, c] for c in range(4))

    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.where(Z >= 0.2, 1.0 / Z, 0.0).astype(dtype)   # ba_cuda.cu:296
        d2 = d * d
        x1 = fx * (X / Z) + cx                                   # unguarded division, ba_cuda.cu:299-300
        y1 = fy * (Y / Z) + cy
        rx = target[:, 0] - x1
        ry = target[:, 1] - y1
        in_bounds = (np.sqrt(rx * rx + ry * ry) < 128) & (Z > 0.2) & \
            (x1 > -64) & (y1 > -64) & (x1 < 2 * cx + 64) & (y1 < 2 * cy + 64)
    mask = in_bounds.astype(dtype)

    o = np.zeros_like(X)
    Jj = np.empty((len(ii), 2, 6), dtype)
    Jj[:, 0] = np.stack([fx * W * d, o, -fx * X * W * d2, -fx * X * Y * d2, fx * (1 + X * X * d2), -fx * Y * d], -1)
    Jj[:, 1] = np.stack([o, fy * W * d, -fy * Y * W * d2, -fy * (1 + Y * Y * d2), fy * X * Y * d2, fy * X * d], -1)
    Jz = np.stack([fx * (tij[:, 0] * d - tij[:, 2] * X * d2),
                   fy * (tij[:, 1] * d - tij[:, 2] * Y * d2)], -1)
    Ji = np.stack([se3.adj_se3(tij, qij, Jj[:, 0]), se3.adj_se3(tij, qij, Jj[:, 1])], axis=1)

    r = np.stack([rx, ry], -1)
    w = mask[:, None] * weight
    return dict(r=r, w=w, Ji=Ji, Jj=Jj, Jz=Jz, coords=np.stack([x1, y1], -1), mask=mask, tij=tij, qij=qij)


def assemble(lin, ii, jj, ku, n_patches, t0, N):
    """Normal equations B [6N,6N], E (CSR [6N,M]), C [M], v [6N], u [M], r_total  (ba_cuda.cu:346-403).

    Frames < t0 are fixed: their rows are skipped, but their edges still add to C and u.
    Sign conventions: B_ij -= w Ji Jj^T (and transpose), E_i -= w Jz Ji, E_j += w Jz Jj, v_i -= w r Ji, v_j += w r Jj.
    """
    r, w, Ji, Jj, Jz = lin["r"], lin["w"], lin["Ji"], lin["Jj"], lin["Jz"]
    dtype = r.dtype
    with np.errstate(invalid="ignore"):
        wr = w * r                   # 0 * inf = nan is reproduced on purpose
        wz = w * Jz
    ix = ii - t0
    jx = jj - t0
    fi = ix >= 0
    fj = jx >= 0
    fij = fi & fj

    B = np.zeros((N, N, 6, 6), dtype)
    Bii = np.einsum("er,era,erb->eab", w, Ji, Ji)
    Bjj = np.einsum("er,era,erb->eab", w, Jj, Jj)
    Bij = -np.einsum("er,era,erb->eab", w, Ji, Jj)
    np.add.at(B, (ix[fi], ix[fi]), Bii[fi])
    np.add.at(B, (jx[fj], jx[fj]), Bjj[fj])
    np.add.at(B, (ix[fij], jx[fij]), Bij[fij])
    np.add.at(B, (jx[fij], ix[fij]), Bij[fij].transpose(0, 2, 1))
    B = B.transpose(0, 2, 1, 3).reshape(6 * N, 6 * N)

    Ei = -np.einsum("er,era->ea", wz, Ji)      # [E,6]
    Ej = np.einsum("er,era->ea", wz, Jj)
    six = np.arange(6)
    rows = np.concatenate([(6 * ix[fi, None] + six).ravel(), (6 * jx[fj, None] + six).ravel()])
    cols = np.concatenate([np.repeat(ku[fi], 6), np.repeat(ku[fj], 6)])
    vals = np.concatenate([Ei[fi].ravel(), Ej[fj].ravel()])
    E = sp.coo_matrix((vals, (rows, cols)), shape=(6 * N, n_patches), dtype=dtype).tocsr()  # sums duplicates

    v = np.zeros((N, 6), dtype)
    np.add.at(v, ix[fi], -np.einsum("er,era->ea", wr, Ji)[fi])
    np.add.at(v, jx[fj], np.einsum("er,era->ea", wr, Jj)[fj])
    v = v.reshape(6 * N)

    C = np.zeros(n_patches, dtype)
    u = np.zeros(n_patches, dtype)
    np.add.at(C, ku, np.sum(wz * Jz, -1))
    np.add.at(u, ku, np.sum(wr * Jz, -1))
    r_total = float(np.sum(wr * r))
    return dict(B=B, E=E, C=C, v=v, u=u, r_total=r_total)


def schur_solve(B, E, C, v, u, lmbda):
    """Q = 1/(C+lmbda); S = B - E Q E^T; y = v - E Q u; S += I*(1e-4*S + 1); Cholesky; dZ = Q (u - E^T dX)
    (ba_cuda.cu:548, 583-592 and 569-579)."""
    dtype = B.dtype
    Q = (1.0 / (C + dtype.type(lmbda))).astype(dtype)
    EQ = E.multiply(Q[None, :]).tocsr()
    S = B - np.asarray((EQ @ E.T).todense(), dtype)
    y = v - EQ @ u
    S_damped = S.copy()
    idx = np.arange(S.shape[0])
    S_damped[idx, idx] += 1e-4 * S[idx, idx] + 1.0
    try:
        L = scipy.linalg.cholesky(S_damped, lower=True, check_finite=False)
        dX = scipy.linalg.cho_solve((L, True), y, check_finite=False).astype(dtype)
    except Exception:  # reference ignores `info`; a failed factorisation yields NaNs
        dX = np.full_like(y, np.nan)
    dZ = Q * (u - E.T @ dX)
    return dict(Q=Q, S=S, S_damped=S_damped, y=y, dX=dX, dZ=dZ.astype(dtype))


def retract(poses, patches, kx, dX, dZ, t0, t1):
    """In-place pose and inverse-depth retraction (ba_cuda.cu:178-229)."""
    if dX is not None and t1 > t0:
        xi = dX.reshape(t1 - t0, 6)
        tn, qn = se3.retr_se3(xi, poses[t0:t1, :3], poses[t0:t1, 3:])
        poses[t0:t1, :3] = tn
        poses[t0:t1, 3:] = qn
    d = patches[kx, 2, 0, 0] + dZ                 # read [2][0][0] (ba_cuda.cu:218)
    d = np.where(d > 20, 1.0, d)                   # ba_cuda.cu:220
    d = np.maximum(d, 1e-4)                        # ba_cuda.cu:221
    patches[kx, 2, :, :] = d[:, None, None]        # broadcast to all P x P cells (ba_cuda.cu:223-227)


def ba(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, iterations=2,
       dtype=np.float64, debug=False):
    """Restatement of cuda_ba(): returns (poses, patches[, per-iteration debug list]) as new arrays of `dtype`.

    Shapes as in the reference API with the leading batch dimension of 1 removed or kept:
    poses [.,F,7], patches [.,K,3,P,P], intrinsics [.,F,4], target/weight [.,E,2], lmbda scalar / [1].
    """
    P = np.asarray(patches).shape[-1]
    poses = np.array(np.asarray(poses).reshape(-1, 7), dtype)
    patches = np.array(np.asarray(patches).reshape(-1, 3, P, P), dtype)
    ii = np.asarray(ii, np.int64).reshape(-1)
    jj = np.asarray(jj, np.int64).reshape(-1)
    kk = np.asarray(kk, np.int64).reshape(-1)
    lm = float(np.asarray(lmbda).reshape(-1)[0])
    kx, ku = np.unique(kk, return_inverse=True)          # ba_cuda.cu:476-478
    N = t1 - t0
    M = len(kx)
    dbg = []
    for _ in range(iterations):
        lin = linearize_edges(poses, patches, intrinsics, target, weight, ii, jj, kk, dtype)
        asm = assemble(lin, ii, jj, ku, M, t0, max(N, 0))
        if N == 0:                                       # structure-only branch, ba_cuda.cu:550-560
            Q = 1.0 / (asm["C"] + dtype(lm))
            sol = dict(Q=Q, dZ=(Q * asm["u"]).astype(dtype), dX=None, S=None, y=None)
        else:
            sol = schur_solve(asm["B"], asm["E"], asm["C"], asm["v"], asm["u"], lm)
        retract(poses, patches, kx, sol["dX"], sol["dZ"], t0, t1)
        if debug:
            dbg.append(dict(lin=lin, kx=kx, ku=ku, **asm, **sol))
    if debug:
        return poses, patches, dbg
    return poses, patches
