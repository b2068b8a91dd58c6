"""Restatement of cuda_ba.neighbors (cdvslam/fastba/ba.cpp:59-97) (oracle; test infrastructure only).

Edges are grouped by `ii`; inside a group they are stably sorted by `jj`; ix[e] / jx[e] are the edge indices of
the previous / next edge of the same group in that order, -1 at the ends."""
import numpy as np


def neighbors(ii, jj):
    ii = np.asarray(ii, np.int64)
    jj = np.asarray(jj, np.int64)
    n = len(ii)
    ix = np.full(n, -1, np.int64)
    jx = np.full(n, -1, np.int64)
    order = np.lexsort((np.arange(n), jj, ii))          # by ii, then jj, ties by original index (stable_sort)
    same = ii[order][1:] == ii[order][:-1]
    ix[order[1:][same]] = order[:-1][same]
    jx[order[:-1][same]] = order[1:][same]
    return ix, jx
