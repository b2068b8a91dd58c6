"""numpy restatement of altcorr corr / patchify and of the reproject kernel (oracle; test infrastructure only).

  corr()       <- cdvslam/altcorr/correlation_kernel.cu:83-136 (window dot products, zero outside the map) and
                  :221-232 (4-tap bilinear blend with dx on the dim-3 / x shift, then permute(0,1,3,2,4,5))
  patchify()   <- correlation_kernel.cu:17-47 (integer gather of a (2R+2)^2 window, output pre-zeroed :297) and
                  cdvslam/altcorr/correlation.py:51-71 ('bilinear' blend / 'upperleft' crop done in python)
  reproject()  <- cdvslam/fastba/ba_cuda.cu:408-458 (all PxP pixels, no depth clamp)
Accumulation is done in `acc_dtype` (float64 by default) on the inputs as given, so for float16 inputs the
oracle is the exact result of the half-precision data (the reference accumulates halves in half,
correlation_kernel.cu:121-131; tolerances for that path are stated in the tests).
"""
import numpy as np

from . import se3_np as se3


def corr_volume(fmap1, fmap2, coords, ii, jj, radius, acc_dtype=np.float64):
    """Raw (2R+2)^2 volume: out[b,m,i_off,j_off,i0,j0] (correlation_kernel.cu:105-134).
    fmap1 [B,K,C,P,P], fmap2 [B,F,C,H2,W2], coords [B,E,2,P,P] (ch0 = x, ch1 = y), ii/jj [E]."""
    B, E = coords.shape[:2]
    Pp = coords.shape[3]
    D = 2 * radius + 2
    C, H2, W2 = fmap2.shape[2:]
    f1 = np.asarray(fmap1, acc_dtype)[:, ii]                      # [B,E,C,P,P]
    f2 = np.asarray(fmap2, acc_dtype)
    x = np.asarray(coords[:, :, 0], np.float32)
    y = np.asarray(coords[:, :, 1], np.float32)
    fx = np.floor(x).astype(np.int64)
    fy = np.floor(y).astype(np.int64)
    out = np.zeros((B, E, D, D, Pp, Pp), acc_dtype)
    b_idx = np.arange(B)[:, None, None, None]
    j_idx = np.asarray(jj)[None, :, None, None]
    for io in range(D):
        for jo in range(D):
            i1 = fy + (io - radius)
            j1 = fx + (jo - radius)
            ok = (i1 >= 0) & (i1 < H2) & (j1 >= 0) & (j1 < W2)
            i1c = np.clip(i1, 0, H2 - 1)
            j1c = np.clip(j1, 0, W2 - 1)
            g = f2[b_idx, j_idx, :, i1c, j1c]                     # [B,E,P,P,C]
            s = np.einsum("bepqc,becpq->bepq", g, f1)
            out[:, :, io, jo] = np.where(ok, s, 0.0)
    return out


def corr(fmap1, fmap2, coords, ii, jj, radius, acc_dtype=np.float64):
    """Final correlation [B,E,2R+1 (x-off),2R+1 (y-off),P,P] as returned by cuda_corr.forward."""
    D = 2 * radius + 2
    vol = corr_volume(fmap1, fmap2, coords, ii, jj, radius, acc_dtype)
    x = np.asarray(coords[:, :, 0], np.float32)
    y = np.asarray(coords[:, :, 1], np.float32)
    dx = (x - np.floor(x)).astype(acc_dtype)[:, :, None, None]
    dy = (y - np.floor(y)).astype(acc_dtype)[:, :, None, None]
    out = (1 - dx) * (1 - dy) * vol[:, :, 0:D - 1, 0:D - 1]
    out = out + dx * (1 - dy) * vol[:, :, 0:D - 1, 1:D]
    out = out + (1 - dx) * dy * vol[:, :, 1:D, 0:D - 1]
    out = out + dx * dy * vol[:, :, 1:D, 1:D]
    return out.transpose(0, 1, 3, 2, 4, 5)


def patchify_raw(net, coords, radius):
    """patches[b,m,c,ii,jj] = net[b,c,floor(y)+ii-R,floor(x)+jj-R], zero outside (correlation_kernel.cu:31-46)."""
    B, C, H, W = net.shape
    M = coords.shape[1]
    D = 2 * radius + 2
    x = np.floor(np.asarray(coords[..., 0], np.float32)).astype(np.int64)
    y = np.floor(np.asarray(coords[..., 1], np.float32)).astype(np.int64)
    out = np.zeros((B, M, C, D, D), net.dtype)
    b_idx = np.arange(B)[:, None]
    for io in range(D):
        for jo in range(D):
            i = y + (io - radius)
            j = x + (jo - radius)
            ok = (i >= 0) & (i < H) & (j >= 0) & (j < W)
            g = net[b_idx, :, np.clip(i, 0, H - 1), np.clip(j, 0, W - 1)]       # [B,M,C]
            out[:, :, :, io, jo] = np.where(ok[..., None], g, 0)
    return out


def patchify(net, coords, radius, mode="bilinear"):
    """correlation.py:51-71."""
    patches = patchify_raw(net, coords, radius)
    if mode == "bilinear":
        # torch evaluates the reference's expression in float32 (float32 weights; a half window is promoted), one rounding
        # per operation, left to right -- numpy float32 arithmetic does exactly the same, so this is bit-exact
        c32 = np.asarray(coords, np.float32)
        off = c32 - np.floor(c32)
        dx = off[..., 0][:, :, None, None, None]
        dy = off[..., 1][:, :, None, None, None]
        one = np.float32(1)
        w = patches.astype(np.float32)
        d = 2 * radius + 1
        return (((one - dy) * (one - dx)) * w[..., :d, :d] + ((one - dy) * dx) * w[..., :d, 1:] +
                (dy * (one - dx)) * w[..., 1:, :d] + (dy * dx) * w[..., 1:, 1:])
    if mode == "upperleft":
        return patches[..., :1, :1]
    return patches


def reproject(poses, patches, intrinsics, ii, jj, kk, dtype=np.float64):
    """coords [1,E,2,P,P] of all patch pixels in frame j (ba_cuda.cu:427-457)."""
    P = np.asarray(patches).shape[-1]
    poses = np.asarray(poses, dtype).reshape(-1, 7)
    patches = np.asarray(patches, dtype).reshape(-1, 3, P, P)
    fx, fy, cx, cy = np.asarray(intrinsics, dtype).reshape(-1, 4)[0]
    tij, qij = se3.rel_se3(poses[ii, :3], poses[ii, 3:], poses[jj, :3], poses[jj, 3:])
    pk = patches[kk]                                                   # [E,3,P,P]
    Xi = np.stack([(pk[:, 0] - cx) / fx, (pk[:, 1] - cy) / fy, np.ones_like(pk[:, 0]), pk[:, 2]], -1)  # [E,P,P,4]
    Xj = se3.act_se3(tij[:, None, None], qij[:, None, None], Xi)
    with np.errstate(divide="ignore", invalid="ignore"):
        u = fx * (Xj[..., 0] / Xj[..., 2]) + cx
        v = fy * (Xj[..., 1] / Xj[..., 2]) + cy
    return np.stack([u, v], 1)[None]


def transform_pops(poses, patches, intrinsics, ii, jj, kk, dtype=np.float64):
    """pops.transform(SE3(poses), patches, intrinsics, ii, jj, kk) without jacobian / valid / depth
    (cdvslam/projective_ops.py:53-68 with iproj :19-30 and proj :32-50), laid out as slam.py:329 hands it to corr:
    coords [1,E,2,P,P].  Differences from reproject(): intrinsics of the source frame for the back-projection (:57), of the
    target frame for the projection (:68), and d = 1 / Z.clamp(min=0.1) (:43) instead of the unguarded division."""
    P = np.asarray(patches).shape[-1]
    poses = np.asarray(poses, dtype).reshape(-1, 7)
    patches = np.asarray(patches, dtype).reshape(-1, 3, P, P)
    K = np.asarray(intrinsics, dtype).reshape(-1, 4)
    Ki, Kj = K[ii][:, :, None, None], K[jj][:, :, None, None]               # [E,4,1,1]
    tij, qij = se3.rel_se3(poses[ii, :3], poses[ii, 3:], poses[jj, :3], poses[jj, 3:])
    pk = patches[kk]
    Xi = np.stack([(pk[:, 0] - Ki[:, 2]) / Ki[:, 0], (pk[:, 1] - Ki[:, 3]) / Ki[:, 1], np.ones_like(pk[:, 0]), pk[:, 2]], -1)
    Xj = se3.act_se3(tij[:, None, None], qij[:, None, None], Xi)
    d = 1.0 / np.maximum(Xj[..., 2], 0.1)
    u = Kj[:, 0] * (d * Xj[..., 0]) + Kj[:, 2]
    v = Kj[:, 1] * (d * Xj[..., 1]) + Kj[:, 3]
    return np.stack([u, v], 1)[None]
