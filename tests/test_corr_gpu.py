"""Parity of the CUDA correlation lookup / patch gather against the numpy oracle.
fp32: values within 1e-5 (BASELINE.json); index gather (patchify) bit-exact.
fp16: the kernel accumulates in fp32 and rounds once, the oracle is the exact result of the fp16 inputs, so the
bound is the output rounding: |err| <= 2^-11 * |value| + 1e-4 (stated here; the reference accumulates in half)."""
import numpy as np
import pytest
import torch

from cdvslam_b200 import synth, altcorr
from oracle import corr_oracle
from tests.helpers import to_dev, f32_problem

pytestmark = pytest.mark.gpu


def _setup(C, dtype, seed=0, F=6, M=8, n_mem=6, edge_noise=True):
    p = synth.small_problem(seed=seed, F=F, M=M, t0=2, lifetime=4)
    gmap, pyr = synth.make_fmaps(p, C=C, n_mem=n_mem, dtype=dtype)
    q = f32_problem(p)
    coords = corr_oracle.reproject(q["poses"], q["patches"], q["intrinsics"], p.ii, p.jj, p.kk).astype(np.float32)
    if edge_noise:   # push some windows across the map border and onto exact integers
        rng = np.random.default_rng(seed)
        coords[0, :10] -= 70.0
        coords[0, 10:20] += 60.0
        coords[0, 20:30] = np.round(coords[0, 20:30])
        coords[0, 30:35] = -0.5
        coords += rng.uniform(-0.5, 0.5, coords.shape).astype(np.float32) * (np.arange(coords.shape[1]) % 3 == 0)[None, :, None, None, None]
    return p, gmap, pyr, coords


@pytest.mark.parametrize("C", [8, 24, 128])
def test_corr_fp32(C):
    p, gmap, pyr, coords = _setup(C, np.float32)
    dev = "cuda"
    ii = torch.as_tensor(p.kk % gmap.shape[0], device=dev)
    jj = torch.as_tensor(p.jj % pyr[0].shape[0], device=dev)
    for lvl, scale in ((0, 1.0), (1, 4.0)):
        c = (coords / np.float32(scale)).astype(np.float32)
        got = altcorr.corr(torch.as_tensor(gmap, device=dev)[None], torch.as_tensor(pyr[lvl], device=dev)[None],
                           torch.as_tensor(c, device=dev), ii, jj, 3)
        want = corr_oracle.corr(gmap[None], pyr[lvl][None], c, p.kk % gmap.shape[0], p.jj % pyr[0].shape[0], 3)
        assert got.shape == (1, p.E, 7, 7, 3, 3) and got.dtype == torch.float32
        assert np.abs(got.cpu().numpy() - want).max() < 1e-5
        assert np.abs(want).max() > 0.05


def test_corr_fp16():
    p, gmap, pyr, coords = _setup(24, np.float16)
    dev = "cuda"
    ii = torch.as_tensor(p.kk, device=dev)
    jj = torch.as_tensor(p.jj, device=dev)
    got = altcorr.corr(torch.as_tensor(gmap, device=dev)[None], torch.as_tensor(pyr[0], device=dev)[None],
                       torch.as_tensor(coords, device=dev), ii, jj, 3)
    want = corr_oracle.corr(gmap[None], pyr[0][None], coords, p.kk, p.jj, 3)
    assert got.dtype == torch.float16
    err = np.abs(got.float().cpu().numpy() - want)
    assert (err <= 2.0 ** -11 * np.abs(want) + 1e-4).all()


def test_corr_pyramid2_equals_two_calls():
    p, gmap, pyr, coords = _setup(24, np.float32, seed=2)
    dev = "cuda"
    g = torch.as_tensor(gmap, device=dev)[None]
    p0, p1 = (torch.as_tensor(x, device=dev)[None] for x in pyr)
    c = torch.as_tensor(coords, device=dev)
    ii = torch.as_tensor(p.kk, device=dev); jj = torch.as_tensor(p.jj, device=dev)
    c1 = altcorr.corr(g, p0, c / 1, ii, jj, 3)
    c2 = altcorr.corr(g, p1, c / 4, ii, jj, 3)
    ref = torch.stack([c1, c2], -1).view(1, len(ii), -1)            # slam.py:323
    fused = altcorr.corr_pyramid2(g, [p0, p1], c, ii, jj, 3)
    assert fused.shape == ref.shape == (1, p.E, 882)
    assert torch.equal(fused, ref)


@pytest.mark.parametrize("radius,mode", [(0, "bilinear"), (1, "bilinear"), (1, "upperleft"), (3, "none")])
@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_patchify(radius, mode, dtype):
    rng = np.random.default_rng(7)
    net = rng.standard_normal((2, 12, 30, 40)).astype(dtype)
    coords = rng.uniform(-3, 43, (2, 50, 2)).astype(np.float32)
    coords[:, :5] = np.round(coords[:, :5])
    got = altcorr.patchify(torch.as_tensor(net, device="cuda"), torch.as_tensor(coords, device="cuda"), radius, mode=mode)
    want = corr_oracle.patchify(net, coords, radius, mode=mode)
    assert got.shape == want.shape
    assert got.cpu().numpy().dtype == want.dtype                        # 'bilinear' is float32 also for half maps
    np.testing.assert_array_equal(got.cpu().numpy(), want)              # gather AND blend: bit exact (see corr_oracle.patchify)


def test_corr_backward_matches_autograd_of_oracle_formula():
    """Gradient of sum(out * G) w.r.t. fmap1 / fmap2, against a torch float64 restatement of the forward formula."""
    p, gmap, pyr, coords = _setup(8, np.float32, seed=5, edge_noise=False)
    dev = "cuda"
    E = p.E
    g = torch.as_tensor(gmap, device=dev)[None].requires_grad_(True)
    f2 = torch.as_tensor(pyr[0], device=dev)[None].requires_grad_(True)
    c = torch.as_tensor(coords, device=dev)
    ii = torch.as_tensor(p.kk, device=dev); jj = torch.as_tensor(p.jj, device=dev)
    G = torch.randn(1, E, 7, 7, 3, 3, device=dev)
    out = altcorr.corr(g, f2, c, ii, jj, 3)
    (out * G).sum().backward()
    # float64 torch restatement (dense gather) for the same scalar
    g64 = torch.as_tensor(gmap, dtype=torch.float64)[None].requires_grad_(True)
    f64 = torch.as_tensor(pyr[0], dtype=torch.float64)[None].requires_grad_(True)
    x = torch.as_tensor(coords[:, :, 0]); y = torch.as_tensor(coords[:, :, 1])
    fx, fy = torch.floor(x).long(), torch.floor(y).long()
    dx, dy = (x - torch.floor(x)).double(), (y - torch.floor(y)).double()
    H2, W2 = pyr[0].shape[-2:]
    f1e = g64[0, torch.as_tensor(p.kk)]                                 # [E,C,3,3]
    vol = torch.zeros(E, 8, 8, 3, 3, dtype=torch.float64)
    vols = []
    for io in range(8):
        row = []
        for jo in range(8):
            i1 = fy[0] + io - 3; j1 = fx[0] + jo - 3
            ok = (i1 >= 0) & (i1 < H2) & (j1 >= 0) & (j1 < W2)
            gat = f64[0, torch.as_tensor(p.jj)[:, None, None], :, i1.clamp(0, H2 - 1), j1.clamp(0, W2 - 1)]  # [E,3,3,C]
            row.append(torch.where(ok, (gat * f1e.permute(0, 2, 3, 1)).sum(-1), torch.zeros((), dtype=torch.float64)))
        vols.append(torch.stack(row, 1))
    vol = torch.stack(vols, 1)                                          # [E,8(y),8(x),3,3]
    o = ((1 - dx[0])[:, None, None] * (1 - dy[0])[:, None, None] * vol[:, :7, :7] + dx[0][:, None, None] * (1 - dy[0])[:, None, None] * vol[:, :7, 1:] +
         (1 - dx[0])[:, None, None] * dy[0][:, None, None] * vol[:, 1:, :7] + dx[0][:, None, None] * dy[0][:, None, None] * vol[:, 1:, 1:])
    o = o.permute(0, 2, 1, 3, 4)
    (o * G[0].double().cpu()).sum().backward()
    assert np.abs(g.grad.cpu().numpy() - g64.grad.numpy()).max() < 1e-3 * max(1.0, g64.grad.abs().max().item())
    assert np.abs(f2.grad.cpu().numpy() - f64.grad.numpy()).max() < 1e-3 * max(1.0, f64.grad.abs().max().item())


@pytest.mark.parametrize("C", [24, 32, 128])
def test_corr_fp16_tma_path(C, monkeypatch):
    """fp16, C in {24, 32}, P = 3, R = 3 takes the TMA + tensor-core path (corr_tma.cu) by default: both levels, windows
    across the map border, far outside the map, on exact integers, windows too far apart for one region (per-tap path);
    the fused two-level call must equal the two single-level calls bit for bit, and the staged kernel (PCORR_TMA=0)
    must agree to output rounding."""
    p, gmap, pyr, coords = _setup(C, np.float16, seed=3, F=8, M=24, n_mem=8)
    coords[0, 40:44] += 5000.0                       # far outside: all-zero windows
    coords[0, 44:46, :, 0, 0] += 9.0                 # windows too far apart for one region: per-tap path
    coords[0, 46, 0, 1, 1] = np.float32("nan")       # non-finite coordinate: that pixel's window is all zeros
    dev = "cuda"
    g = torch.as_tensor(gmap, device=dev)[None]
    maps = [torch.as_tensor(x, device=dev)[None] for x in pyr]
    ii = torch.as_tensor(p.kk, device=dev); jj = torch.as_tensor(p.jj, device=dev)
    c = torch.as_tensor(coords, device=dev)
    finite = np.isfinite(coords).all(axis=(2,))[0]                      # [E, 3, 3]
    outs = []
    for lvl, scale in ((0, 1.0), (1, 4.0)):
        cl = (coords / np.float32(scale)).astype(np.float32)
        got = altcorr.corr(g, maps[lvl], torch.as_tensor(cl, device=dev), ii, jj, 3)
        want = corr_oracle.corr(gmap[None], pyr[lvl][None], np.nan_to_num(cl, nan=-1e7), p.kk, p.jj, 3)
        assert got.dtype == torch.float16 and got.shape == (1, p.E, 7, 7, 3, 3)
        gotn = got.float().cpu().numpy()
        err = np.abs(gotn - want)
        ok = np.broadcast_to(finite[None, :, None, None], err.shape)
        assert (err[ok] <= 2.0 ** -11 * np.abs(want[ok]) + 1e-4).all()
        assert np.abs(want).max() > 0.05
        monkeypatch.setenv("PCORR_TMA", "0")
        staged = altcorr.corr(g, maps[lvl], torch.as_tensor(cl, device=dev), ii, jj, 3)
        monkeypatch.delenv("PCORR_TMA")
        d = (got.float() - staged.float()).abs().cpu().numpy()
        assert (d[ok] <= 2.0 ** -10 * np.abs(want[ok]) + 1e-4).all()
        outs.append(got)
    fused = altcorr.corr_pyramid2(g, maps, c, ii, jj, 3)
    both = torch.stack(outs, -1).view(1, p.E, -1)
    okf = torch.as_tensor(np.broadcast_to(finite[None, :, None, None, :, :, None], (1, p.E, 7, 7, 3, 3, 2)).reshape(1, p.E, -1).copy(), device=dev)
    assert torch.equal(fused[okf], both[okf])


def test_corr_fp32_tile_path(monkeypatch):
    """fp32, C = 128, P = 3, R = 3 takes the TMA tile kernel with the 3xTF32 contraction (corr_tma_wide32_kernel): both
    levels, windows across the map border, far outside the map, windows too far apart for one region (per-tap path), a
    non-finite coordinate; against the float64 oracle at north_star's 1e-5 absolute, against the staged FFMA kernel
    (PCORR_TMA=0), and the fused two-level call against the two single-level calls bit for bit."""
    p, gmap, pyr, coords = _setup(128, np.float32, seed=7, F=8, M=24, n_mem=8)
    coords[0, 40:44] += 5000.0
    coords[0, 44:46, :, 0, 0] += 9.0
    coords[0, 46, 0, 1, 1] = np.float32("nan")
    dev = "cuda"
    g = torch.as_tensor(gmap, device=dev)[None]
    maps = [torch.as_tensor(x, device=dev)[None] for x in pyr]
    ii = torch.as_tensor(p.kk, device=dev); jj = torch.as_tensor(p.jj, device=dev)
    c = torch.as_tensor(coords, device=dev)
    finite = np.isfinite(coords).all(axis=(2,))[0]
    outs = []
    for lvl, scale in ((0, 1.0), (1, 4.0)):
        cl = (coords / np.float32(scale)).astype(np.float32)
        got = altcorr.corr(g, maps[lvl], torch.as_tensor(cl, device=dev), ii, jj, 3)
        want = corr_oracle.corr(gmap[None], pyr[lvl][None], np.nan_to_num(cl, nan=-1e7), p.kk, p.jj, 3)
        assert got.dtype == torch.float32 and got.shape == (1, p.E, 7, 7, 3, 3)
        err = np.abs(got.cpu().numpy() - want)
        ok = np.broadcast_to(finite[None, :, None, None], err.shape)
        assert err[ok].max() < 1e-5, err[ok].max()
        assert np.abs(want).max() > 0.05
        monkeypatch.setenv("PCORR_TMA", "0")
        staged = altcorr.corr(g, maps[lvl], torch.as_tensor(cl, device=dev), ii, jj, 3)
        monkeypatch.delenv("PCORR_TMA")
        assert (got - staged).abs().cpu().numpy()[ok].max() < 1e-5
        outs.append(got)
    fused = altcorr.corr_pyramid2(g, maps, c, ii, jj, 3)
    both = torch.stack(outs, -1).view(1, p.E, -1)
    okf = torch.as_tensor(np.broadcast_to(finite[None, :, None, None, :, :, None], (1, p.E, 7, 7, 3, 3, 2)).reshape(1, p.E, -1).copy(), device=dev)
    assert torch.equal(fused[okf], both[okf])


def test_corr_fp16_tma_batch2():
    """B = 2: per-batch maps and coordinates, shared edge lists (the reference's [B, ...] layout)."""
    p, gmap, pyr, coords = _setup(24, np.float16, seed=4, F=6, M=16, n_mem=6)
    rng = np.random.default_rng(11)
    gm = np.stack([gmap, rng.permutation(gmap.reshape(-1)).reshape(gmap.shape)])
    m0 = np.stack([pyr[0], pyr[0][::-1].copy()])
    cc = np.concatenate([coords, coords + rng.uniform(-2, 2, coords.shape).astype(np.float32)], 0)
    dev = "cuda"
    got = altcorr.corr(torch.as_tensor(gm, device=dev), torch.as_tensor(m0, device=dev), torch.as_tensor(cc, device=dev),
                       torch.as_tensor(p.kk, device=dev), torch.as_tensor(p.jj, device=dev), 3)
    want = corr_oracle.corr(gm, m0, cc, p.kk, p.jj, 3)
    err = np.abs(got.float().cpu().numpy() - want)
    assert (err <= 2.0 ** -11 * np.abs(want) + 1e-4).all()


def test_corr_fp16_tma_c2_shape_matches_staged_fp32():
    """Production shape (c2 graph, 120x160 + 30x40 maps, C = 24): default (TMA) fp16 result vs the fp32 staged kernel run
    on the same (fp16-representable) inputs: differences are output rounding only."""
    p = synth.config_c2()
    gmap, pyr = synth.make_fmaps(p, C=24, dtype=np.float16)
    dev = "cuda"
    d = to_dev(p)
    from cdvslam_b200 import fastba
    coords = fastba.reproject(d["poses"], d["patches"], d["intrinsics"], d["ii"], d["jj"], d["kk"])
    g16 = torch.as_tensor(gmap, device=dev)[None]
    m16 = [torch.as_tensor(x, device=dev)[None] for x in pyr]
    got = altcorr.corr_pyramid2(g16, m16, coords, d["kk"], d["jj"], 3).float()
    ref = altcorr.corr_pyramid2(g16.float(), [m.float() for m in m16], coords, d["kk"], d["jj"], 3)
    err = (got - ref).abs()
    assert bool((err <= 2.0 ** -11 * ref.abs() + 1e-4).all())
    assert ref.abs().max() > 0.1


def test_corr_fp16_tiny_maps_fall_back():
    """Maps smaller than the 12 x 12 TMA box take the staged kernel; same oracle bound."""
    rng = np.random.default_rng(3)
    C, K, F, E = 24, 20, 3, 50
    gmap = (rng.standard_normal((K, C, 3, 3)) / 4).astype(np.float16)
    fmap = (rng.standard_normal((F, C, 9, 10)) / 4).astype(np.float16)
    coords = rng.uniform(-2, 11, (1, E, 2, 3, 3)).astype(np.float32)
    ii = rng.integers(0, K, E); jj = rng.integers(0, F, E)
    got = altcorr.corr(torch.as_tensor(gmap, device="cuda")[None], torch.as_tensor(fmap, device="cuda")[None],
                       torch.as_tensor(coords, device="cuda"), torch.as_tensor(ii, device="cuda"),
                       torch.as_tensor(jj, device="cuda"), 3)
    want = corr_oracle.corr(gmap[None], fmap[None], coords, ii, jj, 3)
    err = np.abs(got.float().cpu().numpy() - want)
    assert (err <= 2.0 ** -11 * np.abs(want) + 1e-4).all()


@pytest.mark.parametrize("C", [24, 128])
def test_pyramid_ring_equals_per_call_transposition(C):
    """altcorr.PyramidRing (persistent channel-last mirror, one slot re-copied per new frame) gives bit-identical lookups
    to corr_pyramid2 / corr, which re-copy every frame map on every call; overwriting a ring slot of the pyramid without
    update() leaves the mirror stale (old result), update(slot) brings it back in line."""
    p, gmap, pyr, coords = _setup(C, np.float16, seed=6, F=8, M=24, n_mem=8)
    dev = "cuda"
    g = torch.as_tensor(gmap, device=dev)[None]
    maps = [torch.as_tensor(x, device=dev)[None].contiguous() for x in pyr]
    ii = torch.as_tensor(p.kk, device=dev); jj = torch.as_tensor(p.jj, device=dev)
    c = torch.as_tensor(coords, device=dev)
    ring = altcorr.PyramidRing(maps)
    ref = altcorr.corr_pyramid2(g, maps, c, ii, jj, 3)
    assert torch.equal(ring.lookup(g, c, ii, jj, 3), ref)
    one = altcorr.PyramidRing(maps[:1])
    assert torch.equal(one.lookup(g, c, ii, jj, 3), altcorr.corr(g, maps[0], c, ii, jj, 3))
    # a new frame arrives in slot 3 (slam.py:681-682)
    torch.manual_seed(1)
    new0 = (torch.randn_like(maps[0][:, 3].float()) / 4).half()
    maps[0][:, 3] = new0
    maps[1][:, 3] = torch.nn.functional.avg_pool2d(new0.float(), 4, 4).half()
    ref2 = altcorr.corr_pyramid2(g, maps, c, ii, jj, 3)
    assert not torch.equal(ref2, ref)
    assert torch.equal(ring.lookup(g, c, ii, jj, 3), ref)          # stale until told
    ring.update(3)
    assert torch.equal(ring.lookup(g, c, ii, jj, 3), ref2)
