"""Host-side mirror of the reference's cdvslam/altcorr/correlation.py: the names a caller imports (CorrLayer, PatchLayer,
patchify, corr) with the reference's signatures; all arithmetic is in libpgba.so (cuda_corr -> include/pcorr.h)."""
import torch

import cuda_corr


class CorrLayer(torch.autograd.Function):
    """Correlation lookup with gradients w.r.t. both feature tensors (reference: correlation.py:4-30).  `dropout` < 1
    back-propagates through a random subset of the edges only, as the reference does."""

    @staticmethod
    def forward(ctx, fmap1, fmap2, coords, ii, jj, radius, dropout):
        ctx.save_for_backward(fmap1, fmap2, coords, ii, jj)
        ctx.cfg = (radius, dropout)
        return cuda_corr.forward(fmap1, fmap2, coords, ii, jj, radius)[0]

    @staticmethod
    def backward(ctx, grad):
        fmap1, fmap2, coords, ii, jj = ctx.saved_tensors
        radius, dropout = ctx.cfg
        if dropout < 1:
            keep = torch.rand(len(ii), device=ii.device) < dropout
            coords, grad, ii, jj = coords[:, keep], grad[:, keep], ii[keep], jj[keep]
        g1, g2 = cuda_corr.backward(fmap1, fmap2, coords, ii, jj, grad, radius)
        return g1, g2, None, None, None, None, None


class PatchLayer(torch.autograd.Function):
    """Patch gather with a gradient w.r.t. the map (reference: correlation.py:33-49).  `mode` (extension, default raw
    window) selects the fused blend / crop of cuda_corr.patchify_mode_forward."""

    @staticmethod
    def forward(ctx, net, coords, radius, mode=0):
        ctx.save_for_backward(net, coords)
        ctx.cfg = (radius, mode)
        return cuda_corr.patchify_mode_forward(net, coords, radius, mode)

    @staticmethod
    def backward(ctx, grad):
        net, coords = ctx.saved_tensors
        radius, mode = ctx.cfg
        return cuda_corr.patchify_mode_backward(net, coords, grad, radius, mode), None, None, None


def patchify(net, coords, radius, mode='bilinear'):
    """Extract (2R+1)^2 bilinear / 1x1 upper-left / raw (2R+2)^2 patches around coords (reference: correlation.py:51-71).
    One kernel per call; any other mode string returns the raw window like the reference."""
    return PatchLayer.apply(net, coords, radius, cuda_corr.PATCH_MODES.get(mode, 0))


def corr(fmap1, fmap2, coords, ii, jj, radius=1, dropout=1):
    """reference: correlation.py:74-75"""
    return CorrLayer.apply(fmap1, fmap2, coords, ii, jj, radius, dropout)


def corr_pyramid2(fmap1, pyramid, coords, ii, jj, radius=3):
    """Extension: the two-level lookup of cdvslam/slam.py:316-323 in one kernel -> [B, E, (2R+1)^2 * P*P * 2]."""
    out = cuda_corr.forward_pyramid2(fmap1, pyramid[0], pyramid[1], coords, ii, jj, radius)
    return out.view(out.shape[0], out.shape[1], -1)


PyramidRing = cuda_corr.PyramidRing     # extension: persistent channel-last pyramid mirror (one slot re-copied per new frame)
