"""Extensions beyond the reference API: batched independent windows (BASELINE config c5) and the debug export of
the normal equations that the parity tolerance is stated on."""
import ctypes

import torch

from cdvslam_b200 import native


def BA_batched(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, M, iterations, eff_impl=False,
               n_edges=None):
    """`fastba.BA` over a batch of independent windows with identical shapes, in one launch sequence.

    poses [B,F,7], patches [B,K,3,P,P], intrinsics [B,F,4] or [1,F,4], target/weight [B,E,2], lmbda [1] or [B],
    ii/jj/kk i64 [B,E] (or [E] to share the graph), n_edges optional i32 [B] for ragged batches.  In place."""
    B = poses.shape[0]
    native.require_cuda(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk)
    for t in (poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk):
        if not t.is_contiguous():
            raise RuntimeError("BA_batched: all tensors must be contiguous")
    P = patches.shape[-1]
    F = poses.shape[1]
    K = patches.shape[1]
    E = ii.shape[-1]
    st = native.Strides()
    st.poses = F * 7
    st.patches = K * 3 * P * P
    st.intrinsics = intrinsics.shape[1] * 4 if intrinsics.shape[0] == B and B > 1 else 0
    st.target = E * 2
    st.weight = E * 2
    st.lmbda = 1 if lmbda.numel() == B and B > 1 else 0
    st.ii = E if ii.dim() == 2 and ii.shape[0] == B and B > 1 else 0
    st.jj = E if jj.dim() == 2 and jj.shape[0] == B and B > 1 else 0
    st.kk = E if kk.dim() == 2 and kk.shape[0] == B and B > 1 else 0
    L = native.lib()
    nbytes = ctypes.c_size_t(0)
    native.check(L.pgba_ba_workspace_bytes(E, F, K, int(t0), int(t1), B, ctypes.byref(nbytes)),
                 "pgba_ba_workspace_bytes")
    with torch.cuda.device(poses.device):
        ws = native.workspace(nbytes.value, poses.device)
        rc = L.pgba_ba_solve_batched(poses.data_ptr(), patches.data_ptr(), intrinsics.data_ptr(), target.data_ptr(),
                                     weight.data_ptr(), lmbda.data_ptr(), ii.data_ptr(), jj.data_ptr(), kk.data_ptr(),
                                     n_edges.data_ptr() if n_edges is not None else None, ctypes.byref(st), B, E, F, K,
                                     P, int(M), int(t0), int(t1), int(iterations), int(bool(eff_impl)), ws.data_ptr(),
                                     ws.numel(), native.stream_ptr(poses.device))
    native.check(rc, "pgba_ba_solve_batched")
    native.note_ba_call(ws, poses.device, E, F, K, t0, t1, B)
    return []


def linearize_debug(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, with_schur=True):
    """One linearisation without retraction; returns a dict of torch tensors: S [6N,6N], y, dX [6N], and per unique
    patch (sorted by patch id, as at::_unique does in the reference: ba_cuda.cu:476) kx, C, u, Q, dZ; status int."""
    native.require_cuda(poses, patches)
    dev = poses.device
    f32 = lambda t: t.contiguous().float()
    poses, patches, intrinsics, target, weight, lmbda = map(f32, (poses, patches, intrinsics, target, weight, lmbda))
    ii, jj, kk = (t.contiguous().long() for t in (ii, jj, kk))
    P = patches.shape[-1]
    F = poses.numel() // 7
    K = patches.numel() // (3 * P * P)
    E = ii.numel()
    N = t1 - t0
    S = torch.zeros(6 * N, 6 * N, device=dev)
    y = torch.zeros(6 * N, device=dev)
    dX = torch.zeros(6 * N, device=dev)
    cap = max(min(E, K), 1)
    ids = torch.zeros(cap, dtype=torch.int64, device=dev)
    C, u, Q, dZ = (torch.zeros(cap, device=dev) for _ in range(4))
    nu = torch.zeros(1, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    L = native.lib()
    nbytes = ctypes.c_size_t(0)
    native.check(L.pgba_ba_workspace_bytes(E, F, K, int(t0), int(t1), 1, ctypes.byref(nbytes)),
                 "pgba_ba_workspace_bytes")
    with torch.cuda.device(dev):
        ws = native.workspace(nbytes.value, dev)
        rc = L.pgba_ba_linearize_debug(poses.data_ptr(), patches.data_ptr(), intrinsics.data_ptr(), target.data_ptr(),
                                       weight.data_ptr(), lmbda.data_ptr(), ii.data_ptr(), jj.data_ptr(), kk.data_ptr(),
                                       E, F, K, P, int(t0), int(t1), int(bool(with_schur)), S.data_ptr(), y.data_ptr(),
                                       dX.data_ptr(), ids.data_ptr(), C.data_ptr(), u.data_ptr(), Q.data_ptr(),
                                       dZ.data_ptr(), nu.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(),
                                       native.stream_ptr(dev))
    native.check(rc, "pgba_ba_linearize_debug")
    native.note_ba_call(ws, dev, E, F, K, int(t0), int(t1), 1)
    n = int(nu.item())
    order = torch.argsort(ids[:n])
    return dict(S=S, y=y, dX=dX, kx=ids[:n][order], C=C[:n][order], u=u[:n][order], Q=Q[:n][order], dZ=dZ[:n][order],
                status=int(status.item()))
