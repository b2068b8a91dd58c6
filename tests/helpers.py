"""Shared helpers for the parity tests (torch <-> numpy plumbing for synthetic problems)."""
import numpy as np
import torch


def to_dev(p, device="cuda", pad_pose_rows=0, pad_patch_rows=0):
    """Problem -> dict of torch tensors in the reference's layouts (leading batch dim of 1)."""
    F = p.poses.shape[0] + pad_pose_rows
    K = p.patches.shape[0] + pad_patch_rows
    poses = np.zeros((F, 7)); poses[:, 6] = 1.0
    poses[:p.poses.shape[0]] = p.poses
    patches = np.zeros((K,) + p.patches.shape[1:])
    patches[:p.patches.shape[0]] = p.patches
    intr = np.tile(p.intrinsics[:1], (F, 1))
    f = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=device)
    l = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.int64, device=device)
    return dict(poses=f(poses)[None], patches=f(patches)[None], intrinsics=f(intr)[None], target=f(p.target)[None],
                weight=f(p.weight)[None], lmbda=f([p.lmbda]), ii=l(p.ii), jj=l(p.jj), kk=l(p.kk))


def f32_problem(p):
    """The float32-rounded inputs as float64 arrays: what the GPU actually sees, fed to the float64 oracle."""
    r = lambda a: np.asarray(a, np.float32).astype(np.float64)
    return dict(poses=r(p.poses), patches=r(p.patches), intrinsics=r(p.intrinsics), target=r(p.target),
                weight=r(p.weight), lmbda=float(np.float32(p.lmbda)))


def rel_err(a, b):
    """max |a-b| / max |b|  (norm-wise relative error)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)
