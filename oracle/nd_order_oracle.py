"""TEST INFRASTRUCTURE ONLY (imported by tests/ alone; the product path never touches oracle/).

CPU restatement of the frame ordering of the large (global BA) solve -- cdv-slam_b200/csrc/ba_bignd.cu: nd_stats_kernel +
nd_order_kernel -- and the property that makes it a valid elimination order for the reference's dense system
(cdvslam/fastba/ba_cuda.cu:575-578, 589-591 factor S in natural order; any symmetric permutation gives the same solution).

The reference has no counterpart (it factors the dense matrix as is), so there is nothing of the reference to pin this
against; what is checked instead is (a) the invariant the parallel elimination relies on -- no patch couples free frames of two
different chain segments -- straight from the edge list, and (b) that the device computes exactly this ordering."""
import numpy as np

ND_MIN_N, ND_MAXP, NB_FRAMES = 27, 32, 8      # ba_common.cuh: ND_MIN_N, ND_MAXP; 48 unknowns = 8 frames per tile


def parameters(N):
    """make_layout (ba_common.cuh): segments P, longest segment, far-edge distance R; None below ND_MIN_N."""
    if N < ND_MIN_N:
        return None
    P = max(1, min(N // 60, ND_MAXP))
    lseg = (N + P - 1) // P
    return dict(P=P, tmax=(lseg + 7) // 8, R=max(lseg // 4, 4), nt=(N + 7) // 8 + P + 1)


def order(ii, jj, kk, F, K, t0, t1):
    ii, jj, kk = (np.asarray(a, np.int64) for a in (ii, jj, kk))
    N = t1 - t0
    par = parameters(N)
    P, R = par["P"], par["R"]
    ok = (ii >= 0) & (ii < F) & (jj >= 0) & (jj < F) & (kk >= 0) & (kk < K)          # as the plan: other edges are ignored
    i, j = ii[ok], jj[ok]
    far = np.abs(j - i) > R
    border = np.zeros(F, bool)
    border[j[far]] = True                                                          # loop-closure targets
    lo, hi = np.arange(F), np.arange(F)                                            # extent of a source frame's near edges
    np.minimum.at(lo, i[~far], j[~far])
    np.maximum.at(hi, i[~far], j[~far])
    cut = [t0 + (p * N) // P for p in range(P + 1)]
    for p in range(1, P):                                                          # separators
        c = cut[p]
        m = (lo < c) & (c <= hi)
        r = hi[m].max() if m.any() else c - 1
        border[c:min(r, t1 - 1) + 1] = True
    free = np.arange(t0, t1)
    seg = np.searchsorted(np.asarray(cut[1:P]), free, side="right")
    pos = np.zeros(F, np.int64)
    T, base, run = [], [], 0
    for p in range(P):
        m = (seg == p) & ~border[free]
        cnt = int(m.sum())
        T.append((cnt + 7) // 8); base.append(run)
        pos[free[m]] = run * 8 + np.arange(cnt)
        run += T[-1]
    mb = border[free]
    nb = int(mb.sum())
    pos[free[mb]] = run * 8 + np.arange(nb)
    Bt = (nb + 7) // 8
    return dict(pos=pos, border=border, seg=seg, segments=P, tiles=run + Bt, border_base=run, border_tiles=Bt,
                border_frames=nb, seg_tiles=np.asarray(T), seg_base=np.asarray(base), tile_capacity=par["nt"],
                steps=max(T) + Bt)


def coupled_pairs_cross_segments(ii, jj, kk, o, t0, t1):
    """Number of (patch, frame a, frame b) couplings of S -- a, b free, both among the patch's source frame and targets -- with
    a and b in different segments and neither in the border.  Must be 0."""
    ii, jj, kk = (np.asarray(a, np.int64) for a in (ii, jj, kk))
    bad = 0
    order_k = np.argsort(kk, kind="stable")
    ks, starts = np.unique(kk[order_k], return_index=True)
    ends = np.append(starts[1:], len(kk))
    seg_of = np.full(len(o["border"]), -1)
    seg_of[t0:t1] = o["seg"]
    for s, e in zip(starts, ends):
        idx = order_k[s:e]
        frames = np.unique(np.concatenate([ii[idx], jj[idx]]))
        frames = frames[(frames >= t0) & (frames < t1)]
        frames = frames[~o["border"][frames]]
        if len(frames) and len(np.unique(seg_of[frames])) > 1:
            bad += 1
    return bad
