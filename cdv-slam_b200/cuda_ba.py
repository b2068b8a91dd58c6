"""Drop-in replacement for the reference's compiled extension module `cuda_ba`
(pybind table: cdvslam/fastba/ba.cpp:183-188; imported by cdvslam/fastba/ba.py:2 and loop_closure/optim_utils.py:1).

Same function names, argument order and in-place semantics; the work is done by libpgba.so through the C ABI of
include/pgba.h.  Put the directory containing this file on sys.path ahead of the reference's build products and
`import cuda_ba` resolves here.
"""
import ctypes

import torch

from cdvslam_b200 import native


def _prep(t, dtype, name):
    if t.dtype != dtype:
        raise RuntimeError("cuda_ba: %s must be %s (got %s)" % (name, dtype, t.dtype))
    native.require_cuda(t)
    return t if t.is_contiguous() else t.contiguous()


def forward(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, PPF, t0, t1, iterations, eff_impl=False):
    """cuda_ba.forward (ba.cpp:32-45 -> cuda_ba(), ba_cuda.cu:462-611).  Mutates `poses` rows t0..t1-1 and the
    inverse-depth channel of `patches` in place; returns [] like the reference."""
    for name, t in (("poses", poses), ("patches", patches)):
        if not t.is_contiguous():
            raise RuntimeError("cuda_ba.forward: %s must be contiguous (it is updated in place)" % name)
    poses_ = _prep(poses, torch.float32, "poses")
    patches_ = _prep(patches, torch.float32, "patches")
    intrinsics = _prep(intrinsics, torch.float32, "intrinsics")
    target = _prep(target, torch.float32, "target")
    weight = _prep(weight, torch.float32, "weight")
    lmbda = _prep(lmbda, torch.float32, "lmbda")
    ii = _prep(ii, torch.int64, "ii")
    jj = _prep(jj, torch.int64, "jj")
    kk = _prep(kk, torch.int64, "kk")
    P = patches_.shape[-1]
    F = poses_.numel() // 7
    K = patches_.numel() // (3 * P * P)
    E = ii.numel()
    if target.numel() != 2 * E or weight.numel() != 2 * E or jj.numel() != E or kk.numel() != E:
        raise RuntimeError("cuda_ba.forward: target/weight/ii/jj/kk sizes disagree")
    L = native.lib()
    nbytes = ctypes.c_size_t(0)
    native.check(L.pgba_ba_workspace_bytes(E, F, K, int(t0), int(t1), 1, ctypes.byref(nbytes)),
                 "pgba_ba_workspace_bytes")
    with torch.cuda.device(poses_.device):
        ws = native.workspace(nbytes.value, poses_.device)
        rc = L.pgba_ba_solve(poses_.data_ptr(), patches_.data_ptr(), intrinsics.data_ptr(), target.data_ptr(),
                             weight.data_ptr(), lmbda.data_ptr(), ii.data_ptr(), jj.data_ptr(), kk.data_ptr(),
                             E, F, K, P, int(PPF), int(t0), int(t1), int(iterations), int(bool(eff_impl)),
                             ws.data_ptr(), ws.numel(), native.stream_ptr(poses_.device))
    native.check(rc, "pgba_ba_solve")
    native.note_ba_call(ws, poses_.device, E, F, K, t0, t1, 1)
    return []


_aux_streams = {}


def forward_host(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, PPF, t0, t1, iterations,
                 eff_impl=False, device=None):
    """Extension: cuda_ba.forward for HOST tensors (pinned memory for asynchronous copies).  Uploads the inputs, runs
    the same kernels and writes the updated `poses` / `patches` back into the host tensors, all enqueued on the
    current stream of `device` (plus an auxiliary stream: the upload of everything but ii/jj/kk overlaps the graph
    analysis).  Asynchronous like every CUDA op: synchronise the stream before reading the host tensors."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    tens = dict(poses=poses, patches=patches, intrinsics=intrinsics, target=target, weight=weight, lmbda=lmbda)
    for name, t in tens.items():
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("cuda_ba.forward_host: %s must be a contiguous float32 CPU tensor" % name)
    for name, t in dict(ii=ii, jj=jj, kk=kk).items():
        if t.is_cuda or t.dtype not in (torch.int64, torch.int32) or t.dtype != ii.dtype or not t.is_contiguous():
            raise RuntimeError("cuda_ba.forward_host: %s must be a contiguous int64 (or, all three, int32) CPU tensor" % name)
    P = patches.shape[-1]
    F = poses.numel() // 7
    K = patches.numel() // (3 * P * P)
    E = ii.numel()
    if target.numel() != 2 * E or weight.numel() != 2 * E or jj.numel() != E or kk.numel() != E:
        raise RuntimeError("cuda_ba.forward_host: target/weight/ii/jj/kk sizes disagree")
    L = native.lib()
    nws, nst = ctypes.c_size_t(0), ctypes.c_size_t(0)
    native.check(L.pgba_ba_workspace_bytes(E, F, K, int(t0), int(t1), 1, ctypes.byref(nws)), "pgba_ba_workspace_bytes")
    native.check(L.pgba_ba_host_staging_bytes(E, F, K, P, ctypes.byref(nst)), "pgba_ba_host_staging_bytes")
    with torch.cuda.device(dev):
        ws = native.workspace(nws.value, dev)
        stg = native.workspace(nst.value, dev, pool="ba_host_staging")
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        aux = _aux_streams.get(key)
        if aux is None:
            aux = _aux_streams[key] = torch.cuda.Stream(device=dev)
        entry = L.pgba_ba_solve_host_i32 if ii.dtype == torch.int32 else L.pgba_ba_solve_host
        rc = entry(poses.data_ptr(), patches.data_ptr(), intrinsics.data_ptr(), target.data_ptr(),
                                  weight.data_ptr(), lmbda.data_ptr(), ii.data_ptr(), jj.data_ptr(), kk.data_ptr(),
                                  E, F, K, P, int(PPF), int(t0), int(t1), int(iterations), int(bool(eff_impl)),
                                  stg.data_ptr(), stg.numel(), ws.data_ptr(), ws.numel(), native.stream_ptr(dev),
                                  aux.cuda_stream)
    native.check(rc, "pgba_ba_solve_host")
    native.note_ba_call(ws, dev, E, F, K, t0, t1, 1)
    return []


def reproject(poses, patches, intrinsics, ii, jj, kk, clamp_depth=False):
    """cuda_ba.reproject (ba.cpp:48-56 -> cuda_reproject(), ba_cuda.cu:614-645): coords f32 [1, E, 2, P, P]."""
    poses = _prep(poses, torch.float32, "poses")
    patches = _prep(patches, torch.float32, "patches")
    intrinsics = _prep(intrinsics, torch.float32, "intrinsics")
    ii = _prep(ii, torch.int64, "ii")
    jj = _prep(jj, torch.int64, "jj")
    kk = _prep(kk, torch.int64, "kk")
    P = patches.shape[-1]
    E = ii.numel()
    coords = torch.empty((1, E, 2, P, P), dtype=torch.float32, device=poses.device)
    with torch.cuda.device(poses.device):
        rc = native.lib().pgba_reproject(poses.data_ptr(), patches.data_ptr(), intrinsics.data_ptr(), ii.data_ptr(),
                                         jj.data_ptr(), kk.data_ptr(), E, poses.numel() // 7,
                                         patches.numel() // (3 * P * P), P, int(bool(clamp_depth)), coords.data_ptr(),
                                         native.stream_ptr(poses.device))
    native.check(rc, "pgba_reproject")
    return coords


def neighbors(ii, jj):
    """cuda_ba.neighbors (ba.cpp:59-97): for edges grouped by `ii` and stably ordered by `jj`, the index of the
    previous / next edge of the same group (-1 at the ends).  One cluster kernel on the device the tensors live on
    (pgba_neighbors), without the reference's D2H -> CPU stable_sort -> H2D round trip; returns [ix, jx] (i64, CUDA)."""
    ii = _prep(ii, torch.int64, "ii")
    jj = _prep(jj, torch.int64, "jj")
    n = ii.numel()
    if jj.numel() != n:
        raise RuntimeError("cuda_ba.neighbors: ii and jj sizes disagree")
    dev = ii.device
    ix = torch.empty(n, dtype=torch.int64, device=dev)
    jx = torch.empty(n, dtype=torch.int64, device=dev)
    if n == 0:
        return [ix, jx]
    L = native.lib()
    nbytes = ctypes.c_size_t(0)
    native.check(L.pgba_neighbors_workspace_bytes(n, ctypes.byref(nbytes)), "pgba_neighbors_workspace_bytes")
    with torch.cuda.device(dev):
        ws = native.workspace(nbytes.value, dev, pool="neighbors")
        rc = L.pgba_neighbors(ii.data_ptr(), jj.data_ptr(), n, ix.data_ptr(), jx.data_ptr(), ws.data_ptr(), ws.numel(),
                              native.stream_ptr(dev))
    native.check(rc, "pgba_neighbors")
    return [ix, jx]


def solve_system(J_Ginv_i, J_Ginv_j, ii, jj, res, ep, lm, freen):
    """cuda_ba.solve_system (ba.cpp:120-180): pose-graph normal equations A = J^T J (+ lm * diag + ep), b = -J^T res and
    the double-precision Cholesky solve of the leading 7*freen block (all when freen < 0), on the device (the reference
    moves everything to the CPU and uses Eigen).  Returns [delta] with delta f32 [n, 7] on res.device.

    Like the reference (ba.cpp:123-127, 171) the inputs may live on any device: its only caller passes CPU tensors from a
    worker process (loop_closure/optim_utils.py:230, fed by long_term.py:259-266).  They are moved to the current CUDA
    device, solved there, and delta goes back to res.device.  (A forked worker cannot initialise CUDA: start the pool
    with the 'spawn' context, see INTEGRATION.md.)"""
    out_device = res.device
    if not torch.cuda.is_available():
        raise RuntimeError("cuda_ba.solve_system: no CUDA device (there is no CPU path)")
    dev = res.device if res.is_cuda else torch.device("cuda", torch.cuda.current_device())
    J_Ginv_i, J_Ginv_j, res, ii, jj = (t.to(dev) for t in (J_Ginv_i, J_Ginv_j, res, ii, jj))
    J_Ginv_i = _prep(J_Ginv_i, torch.float32, "J_Ginv_i")
    J_Ginv_j = _prep(J_Ginv_j, torch.float32, "J_Ginv_j")
    res = _prep(res, torch.float32, "res")
    ii = _prep(ii, torch.int64, "ii")
    jj = _prep(jj, torch.int64, "jj")
    r = res.shape[0]
    if J_Ginv_i.shape != (r, 7, 7) or J_Ginv_j.shape != (r, 7, 7) or res.shape != (r, 7) or ii.numel() != r or jj.numel() != r:
        raise RuntimeError("cuda_ba.solve_system: expected J_Ginv_i/j [r,7,7], res [r,7], ii/jj [r]")
    if r == 0:
        raise RuntimeError("cuda_ba.solve_system: no residuals")
    if bool((ii == jj).any()):                    # the reference calls exit(1) here (ba.cpp:151-152)
        raise RuntimeError("cuda_ba.solve_system: self edge (ii[x] == jj[x])")
    n = int(torch.maximum(ii.max(), jj.max()).item()) + 1          # like the reference (ba.cpp:131), a host sync
    L = native.lib()
    nbytes = ctypes.c_size_t(0)
    native.check(L.pgba_pgo_workspace_bytes(n, ctypes.byref(nbytes)), "pgba_pgo_workspace_bytes")
    delta = torch.empty((n, 7), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = native.workspace(nbytes.value, dev, pool="pgo")
        rc = L.pgba_pgo_solve(J_Ginv_i.data_ptr(), J_Ginv_j.data_ptr(), ii.data_ptr(), jj.data_ptr(), res.data_ptr(), r, n,
                              float(ep), float(lm), int(freen), delta.data_ptr(), None, ws.data_ptr(), ws.numel(),
                              native.stream_ptr(dev))
    native.check(rc, "pgba_pgo_solve")
    return [delta.to(out_device)]
