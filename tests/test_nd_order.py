"""Frame ordering of the large solve (csrc/ba_bignd.cu): the CPU restatement keeps the invariant the parallel elimination
relies on; the device computes exactly that ordering (GPU)."""
import numpy as np
import pytest
import torch

from cdvslam_b200 import synth
from oracle import nd_order_oracle as ndo
from tests.helpers import to_dev


def _edges(kind, F, M, rng):
    if kind == "chain+loops":
        return synth.global_edges(F, M, 25, rng)
    if kind == "chain":
        return synth.global_edges(F, M, 0, rng)
    if kind == "many-loops":
        return synth.global_edges(F, M, 4 * F, rng)
    if kind == "wide-band":
        kk_l, jj_l = [], []
        for dlt in (-9, -4, -1, 1, 3, 9):
            f = np.arange(max(0, -dlt), min(F, F - dlt))
            kk_l.append((f[:, None] * M + np.arange(M)[None, :]).ravel())
            jj_l.append(np.repeat(f + dlt, M))
        kk, jj = np.concatenate(kk_l), np.concatenate(jj_l)
        return kk // M, jj, kk
    if kind == "shuffled":
        ii, jj, kk = synth.global_edges(F, M, 25, rng)
        o = rng.permutation(len(kk))
        return ii[o], jj[o], kk[o]
    raise ValueError(kind)


KINDS = ["chain+loops", "chain", "many-loops", "wide-band", "shuffled"]


@pytest.mark.parametrize("kind", KINDS)
def test_ordering_invariants(kind):
    F, M, t0 = 300, 6, 1
    ii, jj, kk = _edges(kind, F, M, np.random.default_rng(77))
    o = ndo.order(ii, jj, kk, F, F * M, t0, F)
    assert ndo.coupled_pairs_cross_segments(ii, jj, kk, o, t0, F) == 0
    pos = o["pos"][t0:F]
    assert len(np.unique(pos)) == F - t0                                  # a permutation into the padded slots
    assert pos.max() < o["tiles"] * 8 <= o["tile_capacity"] * 8
    for p in range(o["segments"]):                                        # segment p occupies its own tiles
        m = (o["seg"] == p) & ~o["border"][t0:F]
        if m.any():
            assert pos[m].min() >= o["seg_base"][p] * 8 and pos[m].max() < (o["seg_base"][p] + o["seg_tiles"][p]) * 8
    assert (pos[o["border"][t0:F]] >= o["border_base"] * 8).all()
    if kind == "many-loops":
        assert o["border_frames"] > 0.7 * (F - t0)                        # most frames are loop-closure targets: a large border
    if kind == "chain":
        assert o["border_frames"] == 4 * (o["segments"] - 1)              # only the separators: 4 frames per cut


def test_c4_takes_36_dependent_steps_instead_of_125():
    p = synth.config_c4()
    o = ndo.order(p.ii, p.jj, p.kk, 1000, 1000 * p.M, p.t0, p.t1)
    assert o["segments"] == 16 and o["steps"] == 36 and (p.N * 6 + 47) // 48 == 125
    assert ndo.coupled_pairs_cross_segments(p.ii, p.jj, p.kk, o, p.t0, p.t1) == 0


def test_segment_count_by_size():
    assert ndo.parameters(26) is None and ndo.parameters(27)["P"] == 1 and ndo.parameters(256)["P"] == 4


def test_single_segment_below_120_poses():
    F, M, t0 = 75, 6, 1
    ii, jj, kk = synth.global_edges(F, M, 10, np.random.default_rng(5))
    o = ndo.order(ii, jj, kk, F, F * M, t0, F)
    assert o["segments"] == 1 and o["border_frames"] == len(np.unique(jj[np.abs(jj - ii) > 18]))
    assert ndo.coupled_pairs_cross_segments(ii, jj, kk, o, t0, F) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("kind,F", [(k, 300) for k in KINDS] + [("chain+loops", 75), ("wide-band", 140)])
def test_device_ordering_equals_oracle(kind, F):
    from cdvslam_b200 import fastba, native
    M = 6
    p = synth.make_problem("nd-" + kind, F, _edges(kind, F, M, np.random.default_rng(77)), 1, F, 11, M, eff_impl=True)
    d = to_dev(p)
    fastba.BA(d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"], d["jj"], d["kk"],
              p.t0, p.t1, M=p.M, iterations=1, eff_impl=True)
    got = native.last_ba_order()
    assert got is not None
    o = ndo.order(p.ii, p.jj, p.kk, F, F * M, p.t0, p.t1)
    np.testing.assert_array_equal(got["pos"][p.t0:p.t1], o["pos"][p.t0:p.t1])
    for k in ("segments", "tile_capacity", "tiles", "border_base", "border_tiles", "border_frames"):
        assert got[k] == o[k], k
    np.testing.assert_array_equal(got["seg_tiles"], o["seg_tiles"])
    np.testing.assert_array_equal(got["seg_base"], o["seg_base"])


def test_ordering_invariant_on_random_graphs():
    """Random mixtures of near edges (random per-frame reach), far edges and edges from / to fixed frames, random edge
    order: no patch ever couples free frames of two different segments, and the positions are a permutation."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.integers(0, 2 ** 31 - 1))
    def run(seed):
        rng = np.random.default_rng(seed)
        F = int(rng.integers(40, 700))
        t0 = int(rng.integers(1, min(F - 28, 40)))
        M = int(rng.integers(1, 4))
        reach = rng.integers(1, int(rng.integers(2, 30)), F)                 # how far frame f's patches are seen
        kk_l, jj_l = [], []
        for f in range(F):
            for m in range(M):
                n = int(rng.integers(1, 6))
                js = np.clip(f + rng.integers(-reach[f], reach[f] + 1, n), 0, F - 1)
                kk_l.append(np.full(n, f * M + m)); jj_l.append(js)
        n_far = int(rng.integers(0, 3 * F))
        if n_far:
            kf = rng.integers(0, F * M, n_far)
            kk_l.append(kf); jj_l.append(rng.integers(0, F, n_far))
        kk, jj = np.concatenate(kk_l), np.concatenate(jj_l)
        perm = rng.permutation(len(kk))
        kk, jj = kk[perm], jj[perm]
        ii = kk // M
        o = ndo.order(ii, jj, kk, F, F * M, t0, F)
        assert ndo.coupled_pairs_cross_segments(ii, jj, kk, o, t0, F) == 0
        pos = o["pos"][t0:F]
        assert len(np.unique(pos)) == F - t0 and pos.max() < o["tiles"] * 8 <= o["tile_capacity"] * 8

    run()


def _random_graph(seed):
    rng = np.random.default_rng(seed)
    F = int(rng.integers(120, 420))
    M = 3
    reach = rng.integers(1, int(rng.integers(2, 16)), F)
    kk_l, jj_l = [], []
    for f in range(F):
        for m in range(M):
            n = int(rng.integers(2, 6))
            js = np.clip(f + rng.integers(-reach[f], reach[f] + 1, n), 0, F - 1)
            kk_l.append(np.full(n, f * M + m)); jj_l.append(np.unique(js))
            kk_l[-1] = kk_l[-1][:len(jj_l[-1])]
    n_far = int(rng.integers(0, F))
    if n_far:
        kf = rng.integers(0, F * M, n_far)
        kk_l.append(kf); jj_l.append(rng.integers(0, F, n_far))
    kk, jj = np.concatenate(kk_l), np.concatenate(jj_l)
    keep = np.unique(np.stack([kk, jj]), axis=1)                       # no duplicated (patch, frame) edge
    kk, jj = keep[0], keep[1]
    perm = rng.permutation(len(kk))
    kk, jj = kk[perm], jj[perm]
    return F, M, kk // M, jj, kk


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_graphs_device_ordering_and_solve(seed):
    """Random near / far edge mixtures in random order: the device ordering equals the restatement and the solve is the solve
    of the exported damped system (backward error, and against a float64 solve)."""
    from cdvslam_b200 import fastba, native
    from tests.helpers import rel_err
    F, M, ii, jj, kk = _random_graph(seed)
    p = synth.make_problem("nd-rand", F, (ii, jj, kk), 1, F, 30 + seed, M, eff_impl=True)
    d = to_dev(p)
    g = fastba.linearize_debug(d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"],
                               d["jj"], d["kk"], p.t0, p.t1, with_schur=True)
    assert g["status"] == 0
    got = native.last_ba_order()
    o = ndo.order(p.ii, p.jj, p.kk, F, F * M, p.t0, p.t1)
    np.testing.assert_array_equal(got["pos"][p.t0:p.t1], o["pos"][p.t0:p.t1])
    assert (got["tiles"], got["border_tiles"]) == (o["tiles"], o["border_tiles"])
    S, y = g["S"].cpu().numpy().astype(np.float64), g["y"].cpu().numpy().astype(np.float64).reshape(-1)
    A = S + np.diag(1e-4 * np.diag(S) + 1.0)
    dX = g["dX"].cpu().numpy().astype(np.float64).reshape(-1)
    r = A @ dX - y
    assert np.abs(r).max() / (np.abs(A).sum(1).max() * np.abs(dX).max() + np.abs(y).max()) < 1e-5
    assert rel_err(dX, np.linalg.solve(A, y)) < max(1e-3, 1.5 * np.linalg.cond(A) * 2.0 ** -23)
