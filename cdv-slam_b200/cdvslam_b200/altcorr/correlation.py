"""Mirror of the reference's cdvslam/altcorr/correlation.py: same classes, functions, signatures and modes."""
import torch

import cuda_corr


class CorrLayer(torch.autograd.Function):
    """correlation.py:4-30"""

    @staticmethod
    def forward(ctx, fmap1, fmap2, coords, ii, jj, radius, dropout):
        ctx.save_for_backward(fmap1, fmap2, coords, ii, jj)
        ctx.radius = radius
        ctx.dropout = dropout
        corr, = cuda_corr.forward(fmap1, fmap2, coords, ii, jj, radius)
        return corr

    @staticmethod
    def backward(ctx, grad):
        fmap1, fmap2, coords, ii, jj = ctx.saved_tensors
        if ctx.dropout < 1:
            perm = torch.rand(len(ii), device=ii.device) < ctx.dropout
            coords = coords[:, perm]
            grad = grad[:, perm]
            ii = ii[perm]
            jj = jj[perm]
        fmap1_grad, fmap2_grad = cuda_corr.backward(fmap1, fmap2, coords, ii, jj, grad, ctx.radius)
        return fmap1_grad, fmap2_grad, None, None, None, None, None


class PatchLayer(torch.autograd.Function):
    """correlation.py:33-49"""

    @staticmethod
    def forward(ctx, net, coords, radius):
        ctx.radius = radius
        ctx.save_for_backward(net, coords)
        patches, = cuda_corr.patchify_forward(net, coords, radius)
        return patches

    @staticmethod
    def backward(ctx, grad):
        net, coords = ctx.saved_tensors
        grad, = cuda_corr.patchify_backward(net, coords, grad, ctx.radius)
        return grad, None, None


def patchify(net, coords, radius, mode='bilinear'):
    """extract patches (correlation.py:51-71)"""
    patches = PatchLayer.apply(net, coords, radius)

    if mode == 'bilinear':
        offset = (coords - coords.floor()).to(net.device)
        dx, dy = offset[:, :, None, None, None].unbind(dim=-1)
        d = 2 * radius + 1
        x00 = (1 - dy) * (1 - dx) * patches[..., :d, :d]
        x01 = (1 - dy) * (dx) * patches[..., :d, 1:]
        x10 = (dy) * (1 - dx) * patches[..., 1:, :d]
        x11 = (dy) * (dx) * patches[..., 1:, 1:]
        return x00 + x01 + x10 + x11

    elif mode == 'upperleft':
        return patches[..., :1, :1]

    return patches


def corr(fmap1, fmap2, coords, ii, jj, radius=1, dropout=1):
    """correlation.py:74-75"""
    return CorrLayer.apply(fmap1, fmap2, coords, ii, jj, radius, dropout)


def corr_pyramid2(fmap1, pyramid, coords, ii, jj, radius=3):
    """Extension: the two-level lookup of cdvslam/slam.py:316-323 in one kernel -> [B, E, (2R+1)^2 * P*P * 2]."""
    out = cuda_corr.forward_pyramid2(fmap1, pyramid[0], pyramid[1], coords, ii, jj, radius)
    return out.view(out.shape[0], out.shape[1], -1)
