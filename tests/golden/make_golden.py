"""Generate golden vectors by running the reference's OWN python BA (verbatim files) in the build container.

    python tests/golden/make_golden.py            # writes tests/golden/ba_ref_*.npz

What runs: /root/reference/cdvslam/ba.py (BA, ba.py:86-185), /root/reference/cdvslam/projective_ops.py
(transform with jacobian=True, :53-113) and the reference's python lietorch classes
(/root/reference/cdvslam/lietorch/groups.py) -- all imported unmodified from the read-only reference tree.
Three dependencies that cannot be installed/built here are replaced by minimal shims
defined below: `torch_scatter.scatter_sum` (an index_add), the compiled `lietorch_backends` extension (pure-torch
SE3 forward ops following lietorch/include/se3.h:36-95 and so3.h:30-60; needs Eigen to build) and an empty
`cuda_ba` stub (imported by cdvslam/fastba/__init__.py, never called by ba.py).
The reference cannot travel to the GPU box, so the vectors are committed; tests/test_oracle_golden.py checks
oracle/ba_oracle.py and oracle/ba_torch_port.py against them.
"""
import os
import sys
import types
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def _install_shims():
    ts = types.ModuleType("torch_scatter")

    def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
        shape = list(src.shape)
        shape[dim] = int(dim_size) if dim_size is not None else int(index.max()) + 1
        res = torch.zeros(shape, dtype=src.dtype, device=src.device)
        return res.index_add_(dim, index.reshape(-1), src)
    ts.scatter_sum = scatter_sum
    sys.modules["torch_scatter"] = ts

    lb = types.ModuleType("lietorch_backends")

    def cross(a, b):
        return torch.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                            a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                            a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], -1)

    def split(X):                                   # so3.h:31-37: quaternion normalised on construction
        q = X[..., 3:7]
        return X[..., :3], q / q.norm(dim=-1, keepdim=True)

    def rot(q, p):                                  # so3.h:56-61
        uv = 2 * cross(q[..., :3], p)
        return p + q[..., 3:4] * uv + cross(q[..., :3], uv)

    def qmul(a, b):                                 # Eigen quaternion product, (x,y,z,w) storage
        ax, ay, az, aw = a.unbind(-1)
        bx, by, bz, bw = b.unbind(-1)
        return torch.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                            aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz], -1)

    def only_se3(gid):
        assert gid == 3, "shim implements SE3 only"

    def inv(gid, X):                                # se3.h:36-38
        only_se3(gid)
        t, q = split(X)
        qi = torch.cat([-q[..., :3], q[..., 3:]], -1)
        return torch.cat([-rot(qi, t), qi], -1)

    def mul(gid, X, Y):                             # se3.h:45-47
        only_se3(gid)
        tx, qx = split(X)
        ty, qy = split(Y)
        q = qmul(qx, qy)
        return torch.cat([tx + rot(qx, ty), q / q.norm(dim=-1, keepdim=True)], -1)

    def act4(gid, X, p):                            # se3.h:53-56
        only_se3(gid)
        t, q = split(X)
        return torch.cat([rot(q, p[..., :3]) + t * p[..., 3:4], p[..., 3:4]], -1)

    def adjT(gid, X, a):                            # se3.h:57-66, 84-86:  Ad^T a, Ad = [[R, [t]x R],[0, R]]
        only_se3(gid)
        t, q = split(X)
        qi = torch.cat([-q[..., :3], q[..., 3:]], -1)
        return torch.cat([rot(qi, a[..., :3]), rot(qi, a[..., 3:]) + rot(qi, cross(a[..., :3], t))], -1)

    def expm(gid, a):                               # se3.h:137-145, so3.h:150-167, left_jacobian so3.h:169-190
        only_se3(gid)
        tau, phi = a[..., :3], a[..., 3:]
        th2 = (phi * phi).sum(-1, keepdim=True)
        th = th2.sqrt()
        small = th < 1e-6
        s = torch.where(small, torch.ones_like(th), th)
        imag = torch.where(small, 0.5 - th2 / 48 + th2 * th2 / 3840, torch.sin(0.5 * s) / s)
        real = torch.where(small, 1 - th2 / 8 + th2 * th2 / 384, torch.cos(0.5 * s))
        q = torch.cat([imag * phi, real], -1)
        q = q / q.norm(dim=-1, keepdim=True)
        c1 = torch.where(small, 0.5 - th2 / 24, (1 - torch.cos(s)) / (s * s))
        c2 = torch.where(small, 1.0 / 6 - th2 / 120, (s - torch.sin(s)) / (s * s * s))
        px = cross(phi, tau)
        return torch.cat([tau + c1 * px + c2 * cross(phi, px), q], -1)

    def missing(*a, **k):
        raise NotImplementedError("not needed for the forward BA path")
    for name in ["expm_backward", "logm", "logm_backward", "inv_backward", "mul_backward", "adj", "adj_backward",
                 "adjT_backward", "act", "act_backward", "act4_backward", "Jinv", "as_matrix", "projector"]:
        setattr(lb, name, missing)
    lb.expm, lb.inv, lb.mul, lb.adjT, lb.act4 = expm, inv, mul, adjT, act4
    sys.modules["lietorch_backends"] = lb

    # cdvslam/fastba/__init__.py imports the compiled `cuda_ba`; ba.py never calls it -> empty stub.
    cb = types.ModuleType("cuda_ba")
    cb.forward = cb.neighbors = cb.reproject = cb.solve_system = missing
    sys.modules["cuda_ba"] = cb


def _import_reference():
    sys.dont_write_bytecode = True                  # the reference tree is read-only
    _install_shims()
    sys.path.insert(0, REF)
    import cdvslam.ba as ref_ba                     # noqa: E402  (verbatim reference module)
    import cdvslam.projective_ops as ref_pops       # noqa: E402
    from cdvslam.lietorch import SE3                # noqa: E402
    return ref_ba, ref_pops, SE3


def run_reference(problem, iterations, ep, dtype=torch.float64):
    ref_ba, ref_pops, SE3 = _import_reference()
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype)[None]
    poses = SE3(t(problem.poses))
    patches = t(problem.patches)
    intr = t(problem.intrinsics)
    target, weight = t(problem.target), t(problem.weight)
    ii, jj, kk = (torch.as_tensor(x) for x in (problem.ii, problem.jj, problem.kk))
    fx, fy, cx, cy = problem.intrinsics[0]
    bounds = [-64.0, -64.0, 2 * cx + 64.0, 2 * cy + 64.0]       # the CUDA path's bounds (ba_cuda.cu:306)
    out = {}
    coords, valid, (Ji, Jj, Jz) = ref_pops.transform(poses, patches, intr, ii, jj, kk, jacobian=True)
    out.update(coords=coords[0].numpy(), valid=valid[0].numpy(), Ji=Ji[0].numpy(), Jj=Jj[0].numpy(), Jz=Jz[0].numpy())
    for it in range(iterations):
        poses, patches = ref_ba.BA(poses, patches, intr, target, weight, problem.lmbda, ii, jj, kk, bounds,
                                   ep=ep, fixedp=problem.t0)
        out["poses_it%d" % (it + 1)] = poses.data[0].numpy().copy()
        out["patches_it%d" % (it + 1)] = patches[0].numpy().copy()
    return out


def run_reference_grad(problem, ep, dtype=torch.float64):
    """Gradients through ONE step of the verbatim reference BA, including its own CholeskySolver.backward (ba.py:26-37):
    L = sum_k c_k * d_k(new inverse depths) with fixed pseudo-random c; dL/dtargets, dL/dweights.  (Gradients w.r.t. the
    poses would need the compiled lietorch backward ops, which cannot be built here; the depth output depends on targets /
    weights through the whole solve -- S, y, dX, dZ -- without passing through them.)"""
    ref_ba, ref_pops, SE3 = _import_reference()
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype)[None]
    poses = SE3(t(problem.poses))
    patches, intr = t(problem.patches), t(problem.intrinsics)
    target = t(problem.target).requires_grad_(True)
    weight = t(problem.weight).requires_grad_(True)
    ii, jj, kk = (torch.as_tensor(x) for x in (problem.ii, problem.jj, problem.kk))
    fx, fy, cx, cy = problem.intrinsics[0]
    bounds = [-64.0, -64.0, 2 * cx + 64.0, 2 * cy + 64.0]
    _, new_patches = ref_ba.BA(poses, patches, intr, target, weight, problem.lmbda, ii, jj, kk, bounds, ep=ep,
                               fixedp=problem.t0)
    c = torch.as_tensor(np.random.default_rng(7).standard_normal(new_patches.shape[1]), dtype=dtype)
    loss = (c * new_patches[0, :, 2, 0, 0]).sum()
    loss.backward()
    return dict(grad_target=target.grad[0].numpy().copy(), grad_weight=weight.grad[0].numpy().copy(), loss=float(loss))


def main():
    sys.path.insert(0, os.path.join(REPO, "cdv-slam_b200"))
    from cdvslam_b200 import synth
    cases = {
        "small": synth.small_problem(seed=3, F=6, M=8, t0=2, lifetime=4),
        "c1": synth.config_c1(),
    }
    for name, prob in cases.items():
        for ep in (1.0, 100.0):
            res = run_reference(prob, iterations=2, ep=ep)
            keep = {k: v for k, v in res.items() if k.startswith(("poses_it", "patches_it"))}
            keep = {k: (v[:, 2, 0, 0].copy() if k.startswith("patches") else v) for k, v in keep.items()}
            if ep == 1.0:
                keep.update(Ji=res["Ji"].astype(np.float32), Jj=res["Jj"].astype(np.float32),
                            Jz=res["Jz"].astype(np.float32), coords_centre=res["coords"][:, 1, 1, :])
            meta = dict(seed_note="inputs are regenerated from cdvslam_b200.synth with the recorded constructor",
                        ep=ep, t0=prob.t0, t1=prob.t1, E=prob.E)
            if name == "small":
                g = run_reference_grad(prob, ep)
                keep.update(grad_target=g["grad_target"], grad_weight=g["grad_weight"], grad_loss=np.asarray(g["loss"]))
            path = os.path.join(HERE, "ba_ref_%s_ep%g.npz" % (name, ep))
            np.savez_compressed(path, **keep, **{"meta_" + k: np.asarray(v) for k, v in meta.items()})
            print("wrote", path, {k: v.shape for k, v in keep.items()})


if __name__ == "__main__":
    main()
