"""Differentiable BA (cdvslam_b200.diffba, counterpart of cdvslam/ba.py:86-185 + CholeskySolver :11-37) against
  * the vectors of the verbatim reference files (tests/golden/*.npz: Jacobians, two Gauss-Newton steps, and the gradients of
    the new inverse depths w.r.t. targets / weights -- through the reference's own CholeskySolver.backward),
  * the oracle's torch port under native autograd (all inputs, pose outputs included: an independent derivative path --
    torch differentiates through the Cholesky factorisation itself instead of the implicit-function backward).
float64 on the CPU here; the same checks run on the GPU in float64 / float32 under -m gpu."""
import os

import numpy as np
import pytest
import torch

from cdvslam_b200 import synth, diffba
from oracle import ba_torch_port

CASES = {"small": lambda: synth.small_problem(seed=3, F=6, M=8, t0=2, lifetime=4), "c1": synth.config_c1}


def _inputs(p, device="cpu", dtype=torch.float64):
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype, device=device)[None]
    ix = lambda a: torch.as_tensor(np.asarray(a), device=device)
    fx, fy, cx, cy = p.intrinsics[0]
    return dict(poses=t(p.poses), patches=t(p.patches), intr=t(p.intrinsics), target=t(p.target), weight=t(p.weight),
                ii=ix(p.ii), jj=ix(p.jj), kk=ix(p.kk), bounds=[-64.0, -64.0, 2 * cx + 64.0, 2 * cy + 64.0])


def _step(d, p, ep, poses=None, patches=None, **kw):
    return diffba.BA(d["poses"] if poses is None else poses, d["patches"] if patches is None else patches, d["intr"],
                     d["target"], d["weight"], p.lmbda, d["ii"], d["jj"], d["kk"], d["bounds"], ep=ep, fixedp=p.t0, **kw)


@pytest.mark.parametrize("name", list(CASES))
def test_jacobians_match_reference_transform(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "ba_ref_%s_ep1.npz" % name))
    p = CASES[name]()
    d = _inputs(p)
    coords, valid, Ji, Jj, Jz = diffba.transform_with_jacobians(d["poses"], d["patches"], d["intr"], d["ii"], d["jj"], d["kk"])
    np.testing.assert_allclose(Jj[0].numpy(), g["Jj"], rtol=2e-5, atol=2e-5)          # golden stored as f32
    np.testing.assert_allclose(Ji[0].numpy(), g["Ji"], rtol=2e-5, atol=2e-4)
    np.testing.assert_allclose(Jz[0].numpy(), g["Jz"][..., 0], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(coords[0, :, 1, 1].numpy(), g["coords_centre"], rtol=1e-9, atol=1e-9)
    assert float(valid.min()) == 1.0


@pytest.mark.parametrize("compact", [False, True])
@pytest.mark.parametrize("ep", [1.0, 100.0])
@pytest.mark.parametrize("name", list(CASES))
def test_two_steps_match_reference_ba(golden_dir, name, ep, compact):
    g = np.load(os.path.join(golden_dir, "ba_ref_%s_ep%g.npz" % (name, ep)))
    p = CASES[name]()
    d = _inputs(p)
    poses, patches = d["poses"], d["patches"]
    for it in (1, 2):
        poses, patches = _step(d, p, ep, poses, patches, compact_patches=compact)
        np.testing.assert_allclose(poses[0].numpy(), g["poses_it%d" % it], rtol=0, atol=1e-9)
        np.testing.assert_allclose(patches[0, :, 2, 0, 0].numpy(), g["patches_it%d" % it], rtol=1e-9, atol=1e-10)
    assert torch.equal(patches[0, :, :2], d["patches"][0, :, :2])


@pytest.mark.parametrize("ep", [1.0, 100.0])
def test_gradients_match_reference_cholesky_backward(golden_dir, ep):
    """dL/dtargets, dL/dweights of L = sum_k c_k d_k(new depth): the verbatim reference BA with its own
    CholeskySolver.backward (make_golden.run_reference_grad) vs diffba with SolveSPD."""
    g = np.load(os.path.join(golden_dir, "ba_ref_small_ep%g.npz" % ep))
    p = CASES["small"]()
    d = _inputs(p)
    d["target"].requires_grad_(True)
    d["weight"].requires_grad_(True)
    _, patches = _step(d, p, ep)
    c = torch.as_tensor(np.random.default_rng(7).standard_normal(patches.shape[1]), dtype=torch.float64)
    loss = (c * patches[0, :, 2, 0, 0]).sum()
    loss.backward()
    assert abs(loss.item() - float(g["grad_loss"])) < 1e-9
    scale_t, scale_w = np.abs(g["grad_target"]).max(), np.abs(g["grad_weight"]).max()
    assert scale_t > 1e-3 and scale_w > 1e-3
    np.testing.assert_allclose(d["target"].grad[0].numpy(), g["grad_target"], rtol=0, atol=1e-9 * scale_t)
    np.testing.assert_allclose(d["weight"].grad[0].numpy(), g["grad_weight"], rtol=0, atol=1e-9 * scale_w)


def _all_grads(fn, d, p, ep):
    leaves = {k: d[k].clone().requires_grad_(True) for k in ("poses", "patches", "target", "weight")}
    poses, patches = fn(leaves)
    rng = np.random.default_rng(11)
    cp = torch.as_tensor(rng.standard_normal(tuple(poses.shape)), dtype=poses.dtype, device=poses.device)
    cd = torch.as_tensor(rng.standard_normal(tuple(patches.shape)), dtype=poses.dtype, device=poses.device)
    loss = (cp * poses).sum() + (cd * patches).sum()
    loss.backward()
    return loss.item(), {k: v.grad.detach().cpu().numpy() for k, v in leaves.items()}


@pytest.mark.parametrize("ep", [1.0, 100.0])
def test_all_gradients_match_the_port_under_native_autograd(ep):
    """Every input (poses, patches, targets, weights), loss on BOTH outputs: diffba (implicit-function solve backward) vs
    oracle/ba_torch_port.py differentiated by torch through cholesky_ex / cholesky_solve."""
    p = CASES["small"]()
    d = _inputs(p)
    mine = lambda L: diffba.BA(L["poses"], L["patches"], d["intr"], L["target"], L["weight"], p.lmbda, d["ii"], d["jj"],
                               d["kk"], d["bounds"], ep=ep, fixedp=p.t0)
    port = lambda L: ba_torch_port.ba_torch(L["poses"], L["patches"], d["intr"], L["target"], L["weight"], p.lmbda, d["ii"],
                                            d["jj"], d["kk"], d["bounds"], ep=ep, fixedp=p.t0)
    l1, g1 = _all_grads(mine, d, p, ep)
    l2, g2 = _all_grads(port, d, p, ep)
    assert abs(l1 - l2) < 1e-9 * max(1.0, abs(l2))
    for k in g1:
        scale = np.abs(g2[k]).max()
        assert scale > 1e-6, k
        np.testing.assert_allclose(g1[k], g2[k], rtol=0, atol=1e-8 * scale, err_msg=k)


def test_solve_spd_gradcheck_and_failure_convention():
    torch.manual_seed(0)
    M = torch.randn(2, 7, 7, dtype=torch.float64)
    b = torch.randn(2, 7, 1, dtype=torch.float64, requires_grad=True)
    M.requires_grad_(True)
    # A = M M^T + I is a symmetric function of M, the situation in which the -x z^T backward of ba.py:33 is exact
    fn = lambda M_, b_: diffba.SolveSPD.apply(torch.matmul(M_, M_.transpose(1, 2)) + torch.eye(7, dtype=torch.float64), b_)
    assert torch.autograd.gradcheck(fn, (M, b), eps=1e-6, atol=1e-7)
    A = (-torch.eye(4, dtype=torch.float64)[None]).requires_grad_(True)       # not positive definite: x = 0, no gradient
    rhs = torch.ones(1, 4, 1, dtype=torch.float64, requires_grad=True)
    x = diffba.SolveSPD.apply(A, rhs)
    assert torch.equal(x, torch.zeros_like(x))
    (x.sum() + (rhs * 2).sum()).backward()
    assert A.grad is None and torch.equal(rhs.grad, torch.full_like(rhs, 2.0))


def test_structure_only_and_pose_objects():
    """structure_only=True leaves the poses alone (ba.py:165-166); an SE3-like object (anything with `.data`) comes back as
    the same type."""
    p = CASES["small"]()
    d = _inputs(p)
    poses, patches = _step(d, p, 100.0, structure_only=True)
    assert torch.equal(poses, d["poses"]) and not torch.equal(patches, d["patches"])

    class PoseBox:                       # stands in for lietorch.SE3 (cdvslam/lietorch/groups.py): holds `.data`
        def __init__(self, data):
            self.data = data
    out, _ = diffba.BA(PoseBox(d["poses"]), d["patches"], d["intr"], d["target"], d["weight"], p.lmbda, d["ii"], d["jj"],
                       d["kk"], d["bounds"], ep=100.0, fixedp=p.t0)
    ref, _ = _step(d, p, 100.0)
    assert isinstance(out, PoseBox) and torch.equal(out.data, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 2e-3)])
def test_on_the_gpu(golden_dir, dtype, tol):
    """The same operator on CUDA tensors (no host synchronisation besides the Cholesky info check): forward against the
    golden, gradients against the CPU float64 run."""
    g = np.load(os.path.join(golden_dir, "ba_ref_small_ep100.npz"))
    p = CASES["small"]()
    d = _inputs(p, device="cuda", dtype=dtype)
    d["target"].requires_grad_(True)
    d["weight"].requires_grad_(True)
    poses, patches = _step(d, p, 100.0)
    assert poses.is_cuda and patches.is_cuda
    np.testing.assert_allclose(poses[0].detach().cpu().numpy(), g["poses_it1"], rtol=0, atol=max(tol, 1e-9))
    c = torch.as_tensor(np.random.default_rng(7).standard_normal(patches.shape[1]), dtype=dtype, device="cuda")
    (c * patches[0, :, 2, 0, 0]).sum().backward()
    for k, gk in (("target", "grad_target"), ("weight", "grad_weight")):
        scale = np.abs(g[gk]).max()
        np.testing.assert_allclose(d[k].grad[0].cpu().numpy(), g[gk], rtol=0, atol=tol * scale)
