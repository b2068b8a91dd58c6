"""Drop-in replacement for the reference's compiled extension module `cuda_corr`
(pybind table: cdvslam/altcorr/correlation.cpp:57-63; imported by cdvslam/altcorr/correlation.py:2).
Same function names, argument order and return conventions (lists of tensors); the work is done by libpgba.so
through the C ABI of include/pcorr.h."""
import ctypes
import os

import torch

from cdvslam_b200 import native

_DT = {torch.float32: 0, torch.float16: 1}


def _dt(t, what):
    if t.dtype not in _DT:
        raise RuntimeError("cuda_corr.%s: feature maps must be float32 or float16 (got %s)" % (what, t.dtype))
    return _DT[t.dtype]


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _tma(fmap1, maps, coords, ii, jj, radius, out):
    """Default path of the production shapes (P = 3, R = 3; fp16 with C in {24, 32, 128}, fp32 with C = 128): channel-last
    maps + TMA region tiles + tensor cores (corr_tma.cu; fp32: 3xTF32); returns False when the shape does not qualify.
    PCORR_TMA=0 disables it (A/B runs)."""
    if os.environ.get("PCORR_TMA", "1") == "0":
        return False
    L = native.lib()
    B, E, _, P, _ = coords.shape
    K, C = fmap1.shape[1], fmap1.shape[2]
    F = maps[0].shape[1]
    if fmap1.dtype not in _DT or not L.pcorr_tma_supported(C, P, int(radius), _DT[fmap1.dtype]) or E == 0:
        return False
    nlev = len(maps)
    H0, W0 = maps[0].shape[3], maps[0].shape[4]
    H1, W1 = (maps[1].shape[3], maps[1].shape[4]) if nlev == 2 else (0, 0)
    if min(H0, W0) < 12 or (nlev == 2 and min(H1, W1) < 12):        # the 12 x 12 TMA box must fit the map
        return False
    nbytes = ctypes.c_size_t(0)
    native.check(L.pcorr_tma_workspace_bytes_dt(nlev, B, F, C, _DT[fmap1.dtype], H0, W0, H1, W1, ctypes.byref(nbytes)),
                 "pcorr_tma_workspace_bytes_dt")
    with torch.cuda.device(fmap1.device):
        ws = native.workspace(nbytes.value, fmap1.device, pool="corr_tma")
        rc = L.pcorr_forward_tma(fmap1.data_ptr(), maps[0].data_ptr(), maps[1].data_ptr() if nlev == 2 else None,
                                 coords.data_ptr(), ii.data_ptr(), jj.data_ptr(), nlev, B, E, K, F, C, H0, W0, H1, W1,
                                 P, int(radius), _DT[fmap1.dtype], out.data_ptr(), ws.data_ptr(), ws.numel(),
                                 native.stream_ptr(fmap1.device))
    native.check(rc, "pcorr_forward_tma")
    return True


def forward(fmap1, fmap2, coords, ii, jj, radius):
    """cuda_corr.forward (correlation.cpp:28-35 -> corr_cuda_forward, correlation_kernel.cu:193-233).
    Returns [corr] with corr [B, E, 2R+1 (x-off), 2R+1 (y-off), P, P] in fmap1's dtype."""
    native.require_cuda(fmap1, fmap2, coords, ii, jj)
    dt = _dt(fmap1, "forward")
    if fmap2.dtype != fmap1.dtype:
        raise RuntimeError("cuda_corr.forward: fmap1 and fmap2 dtypes differ")
    fmap1, fmap2, ii, jj = _c(fmap1), _c(fmap2), _c(ii.long()), _c(jj.long())
    coords = _c(coords.float())
    B, E, _, P, _ = coords.shape
    K, C = fmap1.shape[1], fmap1.shape[2]
    F, H2, W2 = fmap2.shape[1], fmap2.shape[3], fmap2.shape[4]
    D = 2 * radius + 1
    out = torch.empty((B, E, D, D, P, P), dtype=fmap1.dtype, device=fmap1.device)
    if _tma(fmap1, [fmap2], coords, ii, jj, radius, out):
        return [out]
    with torch.cuda.device(fmap1.device):
        rc = native.lib().pcorr_forward(fmap1.data_ptr(), fmap2.data_ptr(), coords.data_ptr(), ii.data_ptr(),
                                        jj.data_ptr(), B, E, K, F, C, H2, W2, P, int(radius), dt, out.data_ptr(),
                                        native.stream_ptr(fmap1.device))
    native.check(rc, "pcorr_forward")
    return [out]


def forward_pyramid2(fmap1, fmap2_l0, fmap2_l1, coords, ii, jj, radius):
    """Fused two-level lookup (extension): equals torch.stack([corr(l0, coords), corr(l1, coords / 4)], -1) of
    cdvslam/slam.py:321-323.  Returns [B, E, 2R+1, 2R+1, P, P, 2]."""
    native.require_cuda(fmap1, fmap2_l0, fmap2_l1, coords, ii, jj)
    dt = _dt(fmap1, "forward_pyramid2")
    fmap1, fmap2_l0, fmap2_l1, ii, jj = _c(fmap1), _c(fmap2_l0), _c(fmap2_l1), _c(ii.long()), _c(jj.long())
    coords = _c(coords.float())
    B, E, _, P, _ = coords.shape
    K, C = fmap1.shape[1], fmap1.shape[2]
    F = fmap2_l0.shape[1]
    D = 2 * radius + 1
    out = torch.empty((B, E, D, D, P, P, 2), dtype=fmap1.dtype, device=fmap1.device)
    if fmap2_l1.dtype == fmap1.dtype == fmap2_l0.dtype and _tma(fmap1, [fmap2_l0, fmap2_l1], coords, ii, jj, radius, out):
        return out
    with torch.cuda.device(fmap1.device):
        rc = native.lib().pcorr_forward_pyramid2(fmap1.data_ptr(), fmap2_l0.data_ptr(), fmap2_l1.data_ptr(),
                                                 coords.data_ptr(), ii.data_ptr(), jj.data_ptr(), B, E, K, F, C,
                                                 fmap2_l0.shape[3], fmap2_l0.shape[4], fmap2_l1.shape[3],
                                                 fmap2_l1.shape[4], P, int(radius), dt, out.data_ptr(),
                                                 native.stream_ptr(fmap1.device))
    native.check(rc, "pcorr_forward_pyramid2")
    return out


def backward(fmap1, fmap2, coords, ii, jj, corr_grad, radius):
    """cuda_corr.backward (correlation.cpp:37-45 -> corr_cuda_backward, correlation_kernel.cu:236-286)."""
    native.require_cuda(fmap1, fmap2, coords, ii, jj, corr_grad)
    dt = _dt(fmap1, "backward")
    fmap1, fmap2, ii, jj = _c(fmap1), _c(fmap2), _c(ii.long()), _c(jj.long())
    coords = _c(coords.float())
    grad = _c(corr_grad.float())
    B, E, _, P, _ = coords.shape
    K, C = fmap1.shape[1], fmap1.shape[2]
    F, H2, W2 = fmap2.shape[1], fmap2.shape[3], fmap2.shape[4]
    g1 = torch.zeros_like(fmap1)
    g2 = torch.zeros_like(fmap2)
    with torch.cuda.device(fmap1.device):
        rc = native.lib().pcorr_backward(fmap1.data_ptr(), fmap2.data_ptr(), coords.data_ptr(), ii.data_ptr(),
                                         jj.data_ptr(), grad.data_ptr(), B, E, K, F, C, H2, W2, P, int(radius), dt,
                                         g1.data_ptr(), g2.data_ptr(), native.stream_ptr(fmap1.device))
    native.check(rc, "pcorr_backward")
    return [g1, g2]


def patchify_forward(net, coords, radius):
    """cuda_corr.patchify_forward (correlation.cpp:47-50 -> patchify_cuda_forward, correlation_kernel.cu:288-307)."""
    native.require_cuda(net, coords)
    dt = _dt(net, "patchify_forward")
    net = _c(net)
    coords = _c(coords.float())
    B, C, H, W = net.shape
    M = coords.shape[1]
    D = 2 * radius + 2
    patches = torch.empty((B, M, C, D, D), dtype=net.dtype, device=net.device)
    with torch.cuda.device(net.device):
        rc = native.lib().pcorr_patchify_forward(net.data_ptr(), coords.data_ptr(), B, M, C, H, W, int(radius), dt,
                                                 patches.data_ptr(), native.stream_ptr(net.device))
    native.check(rc, "pcorr_patchify_forward")
    return [patches]


def patchify_backward(net, coords, gradient, radius):
    """cuda_corr.patchify_backward (correlation.cpp:52-55 -> patchify_cuda_backward, correlation_kernel.cu:310-333)."""
    native.require_cuda(net, coords, gradient)
    dt = _dt(net, "patchify_backward")
    coords = _c(coords.float())
    gradient = _c(gradient.to(net.dtype))
    B, C, H, W = net.shape
    M = coords.shape[1]
    net_grad = torch.zeros_like(net, memory_format=torch.contiguous_format)
    with torch.cuda.device(net.device):
        rc = native.lib().pcorr_patchify_backward(gradient.data_ptr(), coords.data_ptr(), B, M, C, H, W, int(radius),
                                                  dt, net_grad.data_ptr(), native.stream_ptr(net.device))
    native.check(rc, "pcorr_patchify_backward")
    return [net_grad]


PATCH_MODES = {"none": 0, "raw": 0, "bilinear": 1, "upperleft": 2}


def patchify_mode_forward(net, coords, radius, mode):
    """Extension: altcorr.patchify(net, coords, radius, mode) in ONE kernel -- the gather of patchify_forward with the
    'bilinear' blend / 'upperleft' crop of cdvslam/altcorr/correlation.py:56-69 fused in (pcorr_patchify_mode_forward).
    'bilinear' returns float32 (torch's promotion of the half window against the float32 weights)."""
    native.require_cuda(net, coords)
    dt = _dt(net, "patchify_mode_forward")
    net = _c(net)
    coords = _c(coords.float())
    B, C, H, W = net.shape
    M = coords.shape[1]
    if mode == 1:
        d = 2 * radius + 1
        out = torch.empty((B, M, C, d, d), dtype=torch.float32, device=net.device)
    elif mode == 2:
        out = torch.empty((B, M, C, 1, 1), dtype=net.dtype, device=net.device)
    else:
        return patchify_forward(net, coords, radius)[0]
    with torch.cuda.device(net.device):
        rc = native.lib().pcorr_patchify_mode_forward(net.data_ptr(), coords.data_ptr(), B, M, C, H, W, int(radius), mode,
                                                      dt, out.data_ptr(), native.stream_ptr(net.device))
    native.check(rc, "pcorr_patchify_mode_forward")
    return out


def patchify_mode_backward(net, coords, gradient, radius, mode):
    """Adjoint of patchify_mode_forward w.r.t. net (pcorr_patchify_mode_backward)."""
    if mode == 0:
        return patchify_backward(net, coords, gradient, radius)[0]
    native.require_cuda(net, coords, gradient)
    dt = _dt(net, "patchify_mode_backward")
    coords = _c(coords.float())
    gradient = _c(gradient.float() if mode == 1 else gradient.to(net.dtype))
    B, C, H, W = net.shape
    M = coords.shape[1]
    net_grad = torch.zeros_like(net, memory_format=torch.contiguous_format)
    with torch.cuda.device(net.device):
        rc = native.lib().pcorr_patchify_mode_backward(gradient.data_ptr(), coords.data_ptr(), B, M, C, H, W, int(radius),
                                                       mode, dt, net_grad.data_ptr(), native.stream_ptr(net.device))
    native.check(rc, "pcorr_patchify_mode_backward")
    return net_grad


class PyramidRing:
    """Extension: persistent channel-last mirror of the frame-map pyramid ring for the production lookup (fp16, C in
    {24, 32, 128}, P = 3, radius 3).  cuda_corr.forward / forward_pyramid2 re-copy ALL frame maps to the channel-last layout
    on every call; slam.py rewrites one ring slot per new frame (slam.py:681-682), so an integrator keeps a PyramidRing
    next to `self.pyramid`, calls `ring.update(n % mem)` after those two lines and `ring.lookup(gmap, coords, ii, jj)` in
    SLAM.corr().  Results are bit-identical to forward_pyramid2 / forward on the same maps."""

    def __init__(self, pyramid):
        self.maps = [_c(m) for m in pyramid]
        m0 = self.maps[0]
        native.require_cuda(*self.maps)
        if len(self.maps) not in (1, 2) or any(m.dtype != torch.float16 for m in self.maps):
            raise RuntimeError("PyramidRing: one or two float16 maps [B, F, C, H, W] expected")
        for m, src in zip(self.maps, pyramid):
            if m.data_ptr() != src.data_ptr():
                raise RuntimeError("PyramidRing: the pyramid tensors must be contiguous (they are mirrored in place)")
        self.B, self.F, self.C, self.H0, self.W0 = m0.shape
        self.H1, self.W1 = (self.maps[1].shape[3], self.maps[1].shape[4]) if len(self.maps) == 2 else (0, 0)
        L = native.lib()
        if not L.pcorr_tma_supported(self.C, 3, 3, 1) or min(self.H0, self.W0) < 12 or (len(self.maps) == 2 and min(self.H1, self.W1) < 12):
            raise RuntimeError("PyramidRing: shape outside the TMA lookup's range (C in {24, 32, 128}, maps >= 12 x 12)")
        n = ctypes.c_size_t(0)
        native.check(L.pcorr_tma_workspace_bytes(len(self.maps), self.B, self.F, self.C, self.H0, self.W0, self.H1, self.W1,
                                                 ctypes.byref(n)), "pcorr_tma_workspace_bytes")
        self.ring = torch.empty(n.value, dtype=torch.uint8, device=m0.device)
        self.update(0, self.F)

    def update(self, first, count=1):
        """Re-mirror ring slots [first, first + count) from the pyramid tensors (call after writing them)."""
        m = self.maps
        with torch.cuda.device(m[0].device):
            rc = native.lib().pcorr_ring_update(m[0].data_ptr(), m[1].data_ptr() if len(m) == 2 else None, len(m), self.B,
                                                self.F, self.C, self.H0, self.W0, self.H1, self.W1, int(first), int(count),
                                                self.ring.data_ptr(), self.ring.numel(), native.stream_ptr(m[0].device))
        native.check(rc, "pcorr_ring_update")

    def lookup(self, fmap1, coords, ii, jj, radius=3):
        """== forward_pyramid2(fmap1, *pyramid, coords, ii, jj, radius).view(B, E, -1) (two levels) or forward(...)[0]
        (one level), without touching the NCHW maps."""
        native.require_cuda(fmap1, coords, ii, jj)
        if fmap1.dtype != torch.float16:
            raise RuntimeError("PyramidRing.lookup: fmap1 must be float16")
        fmap1, ii, jj = _c(fmap1), _c(ii.long()), _c(jj.long())
        coords = _c(coords.float())
        B, E, _, P, _ = coords.shape
        D = 2 * radius + 1
        nlev = len(self.maps)
        shape = (B, E, D, D, P, P, 2) if nlev == 2 else (B, E, D, D, P, P)
        out = torch.empty(shape, dtype=torch.float16, device=fmap1.device)
        with torch.cuda.device(fmap1.device):
            rc = native.lib().pcorr_forward_ring(fmap1.data_ptr(), coords.data_ptr(), ii.data_ptr(), jj.data_ptr(), nlev, B, E,
                                                 fmap1.shape[1], self.F, self.C, self.H0, self.W0, self.H1, self.W1, P,
                                                 int(radius), 1, out.data_ptr(), self.ring.data_ptr(), self.ring.numel(),
                                                 native.stream_ptr(fmap1.device))
        native.check(rc, "pcorr_forward_ring")
        return out.view(B, E, -1) if nlev == 2 else out
