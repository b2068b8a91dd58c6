"""Round-2 parity cases: production-size correlation against the oracle AND the reference's own kernels, reproject with
pops.transform semantics, backward kernels against the reference's, fused patchify modes, solve_system against an
independent sparse host solve, reported capacity failures."""
import glob
import importlib.util
import os

import numpy as np
import pytest
import torch

import cuda_ba
import cuda_corr
from cdvslam_b200 import synth, fastba, altcorr
from oracle import corr_oracle, pgo_oracle
from tests.helpers import to_dev, f32_problem

pytestmark = pytest.mark.gpu
REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


def _load_ref(name):
    hits = glob.glob(os.path.join(REF_DIR, name + "*.so"))
    if not hits:
        pytest.skip("oracle/_ref/%s not built" % name)
    spec = importlib.util.spec_from_file_location(name, hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ---- c3-size correlation (c2 graph: 37 824 edges, 22 referenced frames, maps 120x160 + 30x40) -------------------------
@pytest.mark.parametrize("C,dtype", [(24, np.float16), (128, np.float16), (128, np.float32), (24, np.float32)])
def test_corr_c3_size_against_oracle_and_reference_kernel(C, dtype):
    """The production shape through the kernel the dispatcher picks for it (C=24 f16: corr_tma_kernel, C=128 f16:
    corr_tma_wide_kernel, f32: corr_tile32 / staged kernel), both pyramid levels:
      * against corr_oracle.corr (float64 accumulation) on every 9th edge plus the 600 edges whose windows reach furthest
        outside the map (the oracle is per-edge independent; the full edge list at C=128 would take minutes in numpy):
        f32 <= 1e-5 absolute (north_star), f16 <= 2^-11 |v| + 1e-4 (one output rounding);
      * against the reference's own kernel (oracle/_ref, compiled unmodified) on ALL edges: f32 <= 1e-5; f16: the
        reference accumulates C products in half (correlation_kernel.cu:121-131), so its own error is ~ sqrt(C) 2^-11
        per unit of |v|; bound stated as 2^-11 (2 + sqrt(C)) max|v|."""
    ref = _load_ref("ref_cuda_corr")
    p = synth.config_c2()
    gmap, pyr = synth.make_fmaps(p, C=C, dtype=dtype)
    d = to_dev(p)
    dev = "cuda"
    coords = fastba.reproject(d["poses"], d["patches"], d["intrinsics"], d["ii"], d["jj"], d["kk"])
    cn = coords.cpu().numpy()
    g = torch.as_tensor(gmap, device=dev)[None]
    maps = [torch.as_tensor(x, device=dev)[None] for x in pyr]
    ii1 = p.kk % gmap.shape[0]
    jj1 = p.jj % pyr[0].shape[0]
    ti, tj = torch.as_tensor(ii1, device=dev), torch.as_tensor(jj1, device=dev)
    fused = altcorr.corr_pyramid2(g, maps, coords, ti, tj, 3).view(1, p.E, 7, 7, 3, 3, 2)
    # edges for the oracle: a stride + the ones furthest out of the map
    out_of_map = np.maximum(np.maximum(-cn[0, :, 0].min((1, 2)), cn[0, :, 0].max((1, 2)) - p.wd),
                            np.maximum(-cn[0, :, 1].min((1, 2)), cn[0, :, 1].max((1, 2)) - p.ht))
    pick = np.unique(np.concatenate([np.arange(0, p.E, 9), np.argsort(out_of_map)[-600:]]))
    f16 = dtype == np.float16
    for lvl, scale in ((0, 1.0), (1, 4.0)):
        cl = (cn / np.float32(scale)).astype(np.float32)
        got = altcorr.corr(g, maps[lvl], torch.as_tensor(cl, device=dev), ti, tj, 3)
        assert got.shape == (1, p.E, 7, 7, 3, 3) and got.dtype == g.dtype
        assert torch.equal(got, fused[..., lvl])                         # fused two-level call == single-level calls
        gn = got.float().cpu().numpy()
        want = corr_oracle.corr(gmap[None], pyr[lvl][None], cl[:, pick], ii1[pick], jj1[pick], 3)
        err = np.abs(gn[:, pick] - want)
        if f16:
            assert (err <= 2.0 ** -11 * np.abs(want) + 1e-4).all(), err.max()
        else:
            assert err.max() < 1e-5, err.max()
        assert np.abs(want).max() > 0.1
        r, = ref.forward(g, maps[lvl], torch.as_tensor(cl, device=dev), ti, tj, 3)
        assert r.shape == got.shape
        dr = (r.float() - got.float()).abs().max().item()
        vmax = float(np.abs(gn).max())
        bound = 1e-5 if not f16 else 2.0 ** -11 * (2 + np.sqrt(C)) * vmax
        assert dr <= bound, (dr, bound)


# ---- reproject with pops.transform semantics -----------------------------------------------------
def test_reproject_clamp_depth_matches_pops_transform():
    """fastba.reproject(..., clamp_depth=True) == pops.transform(SE3(poses), patches, intrinsics, ii, jj, kk) as
    slam.py:325-329 lays it out (projective_ops.py:19-68): per-frame intrinsics (source frame for the back-projection,
    target frame for the projection) and d = 1 / Z.clamp(min=0.1), including points with 0 < Z < 0.1 and Z < 0."""
    p = synth.small_problem(seed=12, F=9, M=20, t0=3, lifetime=6)
    rng = np.random.default_rng(3)
    # the camera backs away by 0.03 per frame, so t_ij,z = -0.03 (j - i) for j > i; with inverse depths of 20..60 the
    # reprojected Z = 1 + d t_ij,z + ... ends up in (0, 0.1) for some edges and below 0 for others
    p.poses[:, 2] -= 0.03 * np.arange(p.poses.shape[0])
    p.patches[:40, 2] = rng.uniform(20.0, 60.0, (40, 1, 1))
    d = to_dev(p)
    F = d["intrinsics"].shape[1]
    K = np.asarray(p.intrinsics[:1], np.float64) * rng.uniform(0.9, 1.1, (F, 4))     # a different camera per frame
    d["intrinsics"] = torch.as_tensor(K, dtype=torch.float32, device="cuda")[None]
    q = f32_problem(p)
    K32 = np.asarray(K, np.float32).astype(np.float64)
    want = corr_oracle.transform_pops(q["poses"], q["patches"], K32, p.ii, p.jj, p.kk)
    got = fastba.reproject(d["poses"], d["patches"], d["intrinsics"], d["ii"], d["jj"], d["kk"], clamp_depth=True)
    assert got.shape == (1, p.E, 2, 3, 3)
    got = got.cpu().numpy()
    # the test must actually reach the clamp: recompute Z of the centre pixel in float64
    plain = corr_oracle.reproject(q["poses"], q["patches"], K32[:1], p.ii, p.jj, p.kk)
    n_diff = int((np.abs(plain - corr_oracle.transform_pops(q["poses"], q["patches"], np.tile(K32[:1], (F, 1)), p.ii,
                                                              p.jj, p.kk)) > 1e-3).any(axis=(0, 2, 3, 4)).sum())
    assert n_diff >= 10, n_diff                                  # edges where the clamp changes the result
    assert np.isfinite(got).all()
    scale = np.maximum(np.abs(want), 100.0)                      # pixels: fp32 evaluation of ~100..1e4 px coordinates
    assert (np.abs(got - want) / scale).max() < 2e-5
    # clamp_depth=False stays the reference kernel (row 0, unguarded division)
    got0 = fastba.reproject(d["poses"], d["patches"], d["intrinsics"], d["ii"], d["jj"], d["kk"]).cpu().numpy()
    fin = np.isfinite(plain) & (np.abs(plain) < 1e4)
    assert (np.abs(got0 - plain)[fin] / np.maximum(np.abs(plain[fin]), 100.0)).max() < 2e-5


# ---- backward kernels against the reference's own --------------------------------------------------
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.float16, 2e-2)])
def test_corr_backward_matches_reference_kernel(dtype, tol):
    """cuda_corr.backward vs ref_cuda_corr.backward (correlation_kernel.cu:140-190, 236-286) on the same inputs: both
    scatter with atomics (order-dependent rounding; in half for f16 maps), so agreement is to rounding, relative to the
    largest gradient entry."""
    ref = _load_ref("ref_cuda_corr")
    p = synth.small_problem(seed=21, F=8, M=24, t0=3, lifetime=5)
    gmap, pyr = synth.make_fmaps(p, C=24, n_mem=8)
    d = to_dev(p)
    coords = fastba.reproject(d["poses"], d["patches"], d["intrinsics"], d["ii"], d["jj"], d["kk"])
    coords[0, :10] -= 70.0                                        # windows across / outside the border
    coords[0, 10:20] += 60.0
    g = torch.as_tensor(gmap, device="cuda")[None].to(dtype)
    f2 = torch.as_tensor(pyr[0], device="cuda")[None].to(dtype)
    G = torch.randn(1, p.E, 7, 7, 3, 3, device="cuda") * 0.1
    a1, a2 = ref.backward(g, f2, coords, d["kk"], d["jj"], G, 3)
    b1, b2 = cuda_corr.backward(g, f2, coords, d["kk"], d["jj"], G, 3)
    for a, b in ((a1, b1), (a2, b2)):
        assert a.shape == b.shape and a.dtype == b.dtype
        scale = a.float().abs().max().item()
        assert scale > 0.01
        assert (a.float() - b.float()).abs().max().item() <= tol * scale


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_patchify_backward_matches_reference_kernel(dtype):
    ref = _load_ref("ref_cuda_corr")
    torch.manual_seed(3)
    net = torch.randn(2, 12, 30, 40, device="cuda").to(dtype)
    xy = torch.rand(2, 50, 2, device="cuda") * torch.tensor([46.0, 36.0], device="cuda") - 3.0
    grad = torch.randn(2, 50, 12, 4, 4, device="cuda").to(dtype)
    a, = ref.patchify_backward(net, xy, grad, 1)
    b, = cuda_corr.patchify_backward(net, xy, grad, 1)
    assert a.shape == b.shape and a.dtype == b.dtype
    tol = 1e-6 if dtype == torch.float32 else 2e-2              # half atomics: order-dependent rounding
    assert (a.float() - b.float()).abs().max().item() <= tol * a.float().abs().max().item()


# ---- fused patchify modes ---------------------------------------------------------------------------------------------
def _reference_patchify_blend(patches, coords, radius, mode):
    """What cdvslam/altcorr/correlation.py:56-69 does with torch ops on the raw window (restated for the test)."""
    if mode == "bilinear":
        off = coords - coords.floor()
        dx, dy = off[:, :, None, None, None].unbind(dim=-1)
        d = 2 * radius + 1
        return ((1 - dy) * (1 - dx) * patches[..., :d, :d] + (1 - dy) * dx * patches[..., :d, 1:] +
                dy * (1 - dx) * patches[..., 1:, :d] + dy * dx * patches[..., 1:, 1:])
    if mode == "upperleft":
        return patches[..., :1, :1]
    return patches


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("radius,mode", [(0, "bilinear"), (1, "bilinear"), (3, "bilinear"), (1, "upperleft"), (0, "upperleft")])
def test_patchify_fused_modes_equal_reference_blend_bitwise(radius, mode, dtype):
    """altcorr.patchify(..., mode) is ONE kernel here; the reference gathers the raw window with its kernel and blends with
    torch ops.  Forward: bit-identical (same float32 operations in the same order) to that pipeline run on the reference's
    own gather kernel.  Backward: the gradient w.r.t. the map equals autograd through the torch blend + reference scatter."""
    ref = _load_ref("ref_cuda_corr")
    torch.manual_seed(5)
    net = torch.randn(2, 12, 30, 40, device="cuda").to(dtype)
    xy = torch.rand(2, 50, 2, device="cuda") * torch.tensor([46.0, 36.0], device="cuda") - 3.0
    xy[:, :5] = xy[:, :5].round()
    raw, = ref.patchify_forward(net, xy, radius)
    want = _reference_patchify_blend(raw, xy, radius, mode)
    got = altcorr.patchify(net, xy, radius, mode=mode)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert torch.equal(got, want)
    # gradient
    netg = net.clone().requires_grad_(True)
    out = altcorr.patchify(netg, xy, radius, mode=mode)
    G = torch.randn_like(out)
    (out * G).sum().backward()
    rawg = raw.clone().requires_grad_(True)
    (_reference_patchify_blend(rawg, xy, radius, mode) * G).sum().backward()
    want_g, = ref.patchify_backward(net, xy, rawg.grad, radius)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert (netg.grad.float() - want_g.float()).abs().max().item() <= tol * max(1.0, want_g.float().abs().max().item())


# ---- solve_system: second opinion + CPU-tensor call shape ----------------------------------------
def _pose_graph(n, n_loops, seed, noise=0.1, min_gap=5):
    rng = np.random.default_rng(seed)
    kk = np.arange(1, n); ll = kk - 1
    li = rng.integers(min_gap + 3, n, n_loops); lj = np.array([rng.integers(0, i - min_gap) for i in li])
    ii = np.concatenate([kk, li]).astype(np.int64); jj = np.concatenate([ll, lj]).astype(np.int64)
    r = len(ii)
    J_i = (np.eye(7)[None] + noise * rng.standard_normal((r, 7, 7))).astype(np.float32)
    J_j = (-np.eye(7)[None] + noise * rng.standard_normal((r, 7, 7))).astype(np.float32)
    res = (0.05 * rng.standard_normal((r, 7))).astype(np.float32)
    return J_i, J_j, ii, jj, res


@pytest.mark.parametrize("n,n_loops,freen", [(60, 8, -1), (300, 40, -1), (300, 40, 250), (1000, 200, -1)])
def test_solve_system_matches_independent_sparse_solve(n, n_loops, freen):
    """pgba_pgo_solve against scipy's sparse LU of the same normal equations built as a sparse J^T J (an implementation
    that shares nothing with oracle/pgo_oracle.py or pgo.cu): the stand-in for Eigen's SimplicialCholesky (ba.cpp:99-118)."""
    J_i, J_j, ii, jj, res = _pose_graph(n, n_loops, seed=n + n_loops, min_gap=5 if n < 500 else 30)
    want = pgo_oracle.solve_system_sparse(J_i, J_j, ii, jj, res, 1e-3, 1e-6, freen)
    t = lambda a: torch.as_tensor(a, device="cuda")
    got, = cuda_ba.solve_system(t(J_i), t(J_j), t(ii), t(jj), t(res), 1e-3, 1e-6, freen)
    got = got.cpu().numpy()
    scale = np.abs(want).max()
    assert scale > 1e-4
    assert np.abs(got - want).max() <= 2e-6 * scale, np.abs(got - want).max() / scale


def test_solve_system_accepts_cpu_tensors_like_the_reference_caller():
    """loop_closure/optim_utils.py:230 calls cuda_ba.solve_system with CPU tensors (pred_poses.cpu(), long_term.py:259-266)
    from a worker process; the reference moves everything to the CPU and returns delta on res.device (ba.cpp:120-127,171)."""
    J_i, J_j, ii, jj, res = _pose_graph(40, 6, seed=9)
    c = lambda a: torch.as_tensor(a)
    got, = cuda_ba.solve_system(c(J_i), c(J_j), c(ii), c(jj), c(res), 1e-3, 1e-6, -1)
    assert got.device.type == "cpu" and got.shape == (40, 7) and got.dtype == torch.float32
    want, _, _ = pgo_oracle.solve_system(J_i, J_j, ii, jj, res, 1e-3, 1e-6, -1)
    assert np.abs(got.numpy() - want).max() <= 2e-6 * np.abs(want).max()


# ---- capacity failures are reported, not silent --------------------------------------------------------------
def test_too_many_target_frames_per_chunk_is_reported():
    """A source frame whose patches see more than PGBA_MAX_SLOTS (128) distinct target frames cannot be processed; with
    PGBA_CHECK_STATUS=1 the call raises instead of returning a partial update, and fastba.last_status() exposes the word."""
    F, M = 140, 4
    rng = np.random.default_rng(0)
    kk, jj = np.meshgrid(np.arange(M), np.arange(1, F), indexing="ij")       # patches of frame 0 -> 139 target frames
    edges = (np.zeros(kk.size, np.int64), jj.reshape(-1).astype(np.int64), kk.reshape(-1).astype(np.int64))
    p = synth.make_problem("wide", F, edges, 1, F, 3, M)
    d = to_dev(p)
    args = (d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"], d["jj"], d["kk"],
            p.t0, p.t1)
    before = d["poses"].clone()
    fastba.BA(*args, M=M, iterations=1)                                   # default: asynchronous, no host check
    torch.cuda.synchronize()
    st = fastba.last_status(d["poses"].device)
    assert st & 2, st                                                     # PGBA_ST_TOO_MANY_SLOTS
    d["poses"].copy_(before)
    os.environ["PGBA_CHECK_STATUS"] = "1"
    try:
        with pytest.raises(RuntimeError, match="status"):
            fastba.BA(*args, M=M, iterations=1)
    finally:
        del os.environ["PGBA_CHECK_STATUS"]
    # an index outside the buffers is reported the same way
    d["jj"][5] = 10 ** 6
    fastba.BA(*args, M=M, iterations=1)
    torch.cuda.synchronize()
    assert fastba.last_status(d["poses"].device) & 1                       # PGBA_ST_INDEX_RANGE
