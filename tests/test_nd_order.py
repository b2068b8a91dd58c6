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
