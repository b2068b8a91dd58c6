"""ctypes binding of libpgba.so (the C ABI declared in include/pgba.h and include/pcorr.h).

PyTorch is only the owner of device memory and streams here: tensors are passed as raw device pointers
(`tensor.data_ptr()`) together with `torch.cuda.current_stream().cuda_stream`.  There is no CPU or eager fallback:
importing this module without the built library, or calling an op with CPU tensors, raises.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("PGBA_LIB") or os.path.join(PKG_ROOT, "lib", "libpgba.so")   # PGBA_LIB: A/B builds
BUILD_SCRIPT = os.path.join(PKG_ROOT, "csrc", "build.sh")

c_i64, c_int, c_vp, c_sz = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t


class Strides(ctypes.Structure):
    _fields_ = [(n, c_i64) for n in ("poses", "patches", "intrinsics", "target", "weight", "lmbda", "ii", "jj", "kk")]


# name -> (restype, argtypes); mirrors include/pgba.h and include/pcorr.h one to one
SIGNATURES = {
    "pgba_error_string": (ctypes.c_char_p, [c_int]),
    "pgba_version": (c_int, []),
    "pgba_ba_workspace_bytes": (c_int, [c_i64, c_i64, c_i64, c_int, c_int, c_i64, ctypes.POINTER(c_sz)]),
    "pgba_ba_solve": (c_int, [c_vp] * 9 + [c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_sz,
                                           c_vp]),
    "pgba_ba_solve_batched": (c_int, [c_vp] * 10 + [ctypes.POINTER(Strides), c_i64, c_i64, c_i64, c_i64, c_int, c_int,
                                                     c_int, c_int, c_int, c_int, c_vp, c_sz, c_vp]),
    "pgba_ba_host_staging_bytes": (c_int, [c_i64, c_i64, c_i64, c_int, ctypes.POINTER(c_sz)]),
    "pgba_ba_host_arena_offsets": (c_int, [c_i64, c_i64, c_i64, c_int, ctypes.POINTER(c_sz), ctypes.POINTER(c_sz)]),
    "pgba_ba_host_arena_offsets_i32": (c_int, [c_i64, c_i64, c_i64, c_int, ctypes.POINTER(c_sz), ctypes.POINTER(c_sz)]),
    "pgba_ba_solve_host": (c_int, [c_vp] * 9 + [c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_sz,
                                                c_vp, c_sz, c_vp, c_vp]),
    "pgba_ba_solve_host_i32": (c_int, [c_vp] * 9 + [c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_sz,
                                                c_vp, c_sz, c_vp, c_vp]),
    "pgba_ba_linearize_debug": (c_int, [c_vp] * 9 + [c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int] + [c_vp] * 10 +
                                [c_vp, c_sz, c_vp]),
    "pgba_ba_solve_profiled": (c_int, [c_vp] * 9 + [ctypes.POINTER(Strides), c_i64, c_i64, c_i64, c_i64, c_int, c_int,
                                                     c_int, c_int, c_vp, c_sz, c_vp, ctypes.POINTER(ctypes.c_float)]),
    "pgba_launch_count": (ctypes.c_longlong, []),
    "pgba_ba_status_ptr": (c_vp, [c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_i64, c_i64]),
    "pgba_ba_plan_hit_ptr": (c_vp, [c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_i64, c_i64]),
    "pgba_ba_order_ptr": (c_vp, [c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_i64, c_i64, ctypes.POINTER(c_int),
                                 ctypes.POINTER(c_int), ctypes.POINTER(c_vp)]),
    "pgba_pgo_workspace_bytes": (c_int, [c_i64, ctypes.POINTER(c_sz)]),
    "pgba_pgo_solve": (c_int, [c_vp] * 5 + [c_i64, c_i64, ctypes.c_float, ctypes.c_float, c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "pgba_neighbors_workspace_bytes": (c_int, [c_i64, ctypes.POINTER(c_sz)]),
    "pgba_neighbors": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "pgba_reproject": (c_int, [c_vp] * 6 + [c_i64, c_i64, c_i64, c_int, c_int, c_vp, c_vp]),
    "pcorr_forward": (c_int, [c_vp] * 5 + [c_int, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_vp,
                                           c_vp]),
    "pcorr_forward_pyramid2": (c_int, [c_vp] * 6 + [c_int, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int,
                                                    c_int, c_int, c_int, c_vp, c_vp]),
    "pcorr_tma_supported": (c_int, [c_int, c_int, c_int, c_int]),
    "pcorr_tma_workspace_bytes": (c_int, [c_int, c_int, c_i64, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_sz)]),
    "pcorr_tma_workspace_bytes_dt": (c_int, [c_int, c_int, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_sz)]),
    "pcorr_forward_tma": (c_int, [c_vp] * 6 + [c_int, c_int, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int,
                                               c_int, c_int, c_int, c_vp, c_vp, c_sz, c_vp]),
    "pcorr_ring_update": (c_int, [c_vp, c_vp, c_int, c_int, c_i64, c_int, c_int, c_int, c_int, c_int, c_i64, c_i64, c_vp, c_sz,
                                  c_vp]),
    "pcorr_forward_ring": (c_int, [c_vp] * 4 + [c_int, c_int, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_int,
                                                c_int, c_int, c_vp, c_vp, c_sz, c_vp]),
    "pcorr_backward": (c_int, [c_vp] * 6 + [c_int, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_vp,
                                            c_vp, c_vp]),
    "pcorr_patchify_forward": (c_int, [c_vp, c_vp, c_int, c_i64, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "pcorr_patchify_backward": (c_int, [c_vp, c_vp, c_int, c_i64, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "pcorr_patchify_mode_forward": (c_int, [c_vp, c_vp, c_int, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "pcorr_patchify_mode_backward": (c_int, [c_vp, c_vp, c_int, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
}

_lib = None


def build(verbose=False):
    """Compile libpgba.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["bash", BUILD_SCRIPT], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libpgba.so failed")
    return LIB_PATH


def lib():
    """The loaded library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libpgba.so is missing at %s -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(or bash cdv-slam_b200/csrc/build.sh); there is no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s (code %d)" % (what, lib().pgba_error_string(rc).decode(), rc))


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("cdvslam_b200 ops run on CUDA tensors only (got a %s tensor); there is no CPU path"
                               % t.device)


_workspaces = {}
_retired = []          # outgrown buffers stay alive: a CUDA graph captured earlier may still point into them


def workspace(nbytes, device, pool="ba"):
    """Grow-only scratch buffer per (pool, device, current stream) -- the C ABI never allocates.  Keyed by the stream as
    well: two streams calling BA concurrently (or a captured graph replayed on a side stream while eager calls run on
    another) must not share S, y and the plan tables, as the reference's temporaries do not (stream-aware caching
    allocator).  When a larger buffer is needed the old one is kept alive, so pointers baked into previously captured
    CUDA graphs remain valid."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (pool, device.type, idx, torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _retired.append(buf)
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


_last_ba = {}          # device index -> (workspace tensor, E, F, K, t0, t1, batch) of the most recent BA call


def note_ba_call(ws, device, E, F, K, t0, t1, batch):
    """Remember where the status words of the call just enqueued live; with PGBA_CHECK_STATUS=1 synchronise and raise when
    a window could not be processed completely (the reference has no such limits, so a partial update must not pass
    silently).  Default: asynchronous; poll with last_ba_status()."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    _last_ba[key] = (ws, int(E), int(F), int(K), int(t0), int(t1), int(batch))
    if os.environ.get("PGBA_CHECK_STATUS", "0") == "1":
        st = last_ba_status(device)
        if st:
            raise RuntimeError("fastba.BA: status 0x%x -- %s; the update is partial or missing (limits: INTEGRATION.md)"
                               % (st, ", ".join(n for b, n in STATUS_BITS if st & b)))


STATUS_BITS = ((1, "an ii/jj/kk index is outside the pose / patch buffers (edge skipped)"),
               (2, "a chunk sees more than 128 distinct target frames (its edges were dropped)"),
               (4, "internal table capacity exceeded (chunk dropped or whole solve skipped)"),
               (8, "a patch id appears with two source frames"))


def last_ba_status(device=None):
    """OR of the device-side status words (PGBA_ST_* of include/pgba.h) of every window of the most recent BA call on
    `device`; synchronises that device's current stream.  0 = every edge was processed."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    rec = _last_ba.get(key)
    if rec is None:
        return 0
    ws, E, F, K, t0, t1, batch = rec
    torch.cuda.current_stream(dev).synchronize()
    L = lib()
    st = 0
    for b in range(batch):
        ptr = L.pgba_ba_status_ptr(ws.data_ptr(), E, F, K, t0, t1, batch, b)
        if not ptr:
            continue
        off = ptr - ws.data_ptr()
        st |= int(ws[off:off + 4].view(torch.int32).item())
    return st


def last_ba_plan_hits(device=None):
    """Number of windows of the most recent BA call on `device` that reused their plan tables (fingerprint match);
    synchronises.  See include/pgba.h, "Plan cache"."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    rec = _last_ba.get(key)
    if rec is None:
        return 0
    ws, E, F, K, t0, t1, batch = rec
    torch.cuda.current_stream(dev).synchronize()
    L = lib()
    hits = 0
    for b in range(batch):
        ptr = L.pgba_ba_plan_hit_ptr(ws.data_ptr(), E, F, K, t0, t1, batch, b)
        if ptr:
            off = ptr - ws.data_ptr()
            hits += int(ws[off:off + 4].view(torch.int32).item())
    return hits


def last_ba_order(device=None, b=0):
    """Frame ordering the large solve of the most recent BA call on `device` used for window b (include/pgba.h,
    pgba_ba_order_ptr): dict(pos=i32 [F] numpy, segments, tile_capacity, tiles, border_base, border_tiles, border_frames,
    seg_tiles, seg_base), or None when the call did not take the reordered solve.  Synchronises."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    rec = _last_ba.get(key)
    if rec is None:
        return None
    ws, E, F, K, t0, t1, batch = rec
    torch.cuda.current_stream(dev).synchronize()
    P, cap, hdr = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_void_p(0)
    ptr = lib().pgba_ba_order_ptr(ws.data_ptr(), E, F, K, t0, t1, batch, b, ctypes.byref(P), ctypes.byref(cap), ctypes.byref(hdr))
    if not ptr:
        return None
    off, hoff = ptr - ws.data_ptr(), hdr.value - ws.data_ptr()
    pos = ws[off:off + 4 * F].view(torch.int32).cpu().numpy()
    h = ws[hoff:hoff + 4 * 68].view(torch.int32).cpu().numpy()
    return dict(pos=pos, segments=P.value, tile_capacity=cap.value, tiles=int(h[0]), border_base=int(h[1]),
                border_tiles=int(h[2]), border_frames=int(h[3]), seg_tiles=h[4:4 + P.value].copy(),
                seg_base=h[36:36 + P.value].copy())


def invalidate_plan_cache():
    """Forget every cached plan (zero the descriptor at the start of every BA workspace): the next call on any of them
    rebuilds its tables.  Used by bench.py to time the cold path."""
    for key, buf in list(_workspaces.items()):
        if key[0] == "ba":
            buf[:256].zero_()
    for buf in _retired:
        buf[:256].zero_()


def host_arena(n_edges, n_pose_rows, n_patch_rows, P=3, index_dtype=torch.int64):
    """One pinned host allocation holding the nine input tensors of fastba.BA_host at the offsets of
    pgba_ba_host_arena_offsets (views with the reference's shapes, leading batch dim of 1).  BA_host then moves the
    problem with two uploads and one download instead of nine + two.  index_dtype=torch.int32: ii / jj / kk views are
    32-bit (pgba_ba_host_arena_offsets_i32): half the index upload."""
    offs = (c_sz * 9)()
    total = c_sz(0)
    fn = lib().pgba_ba_host_arena_offsets_i32 if index_dtype == torch.int32 else lib().pgba_ba_host_arena_offsets
    check(fn(n_edges, n_pose_rows, n_patch_rows, P, offs, ctypes.byref(total)), "pgba_ba_host_arena_offsets")
    buf = torch.zeros(total.value, dtype=torch.uint8).pin_memory()
    E, F, K = n_edges, n_pose_rows, n_patch_rows

    def view(i, dtype, shape):
        n = int(torch.tensor(shape).prod().item()) * torch.empty((), dtype=dtype).element_size()
        return buf[offs[i]:offs[i] + n].view(dtype).view(shape)
    return dict(_arena=buf, poses=view(0, torch.float32, (1, F, 7)), patches=view(1, torch.float32, (1, K, 3, P, P)),
                intrinsics=view(2, torch.float32, (1, F, 4)), target=view(3, torch.float32, (1, E, 2)),
                weight=view(4, torch.float32, (1, E, 2)), lmbda=view(5, torch.float32, (1,)),
                ii=view(6, index_dtype, (E,)), jj=view(7, index_dtype, (E,)), kk=view(8, index_dtype, (E,)))
