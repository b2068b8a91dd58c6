"""Seeded synthetic patch graphs shaped like the tensors cdvslam/slam.py hands to fastba.BA / altcorr.corr.

Layouts follow SURVEY.md appendix A (reference: cdvslam/patchgraph.py:28-30, cdvslam/slam.py:229-235, 331-337,
528-541): poses [F,7] = (tx,ty,tz,qx,qy,qz,qw) world->camera, patches [K,3,3,3] with channel 0/1 = pixel x/y at
1/4 resolution and channel 2 = inverse depth, intrinsics [F,4] = (fx,fy,cx,cy)/4, global patch id k = frame*M + m,
ii = source frame of k, jj = target frame.  Everything is generated in float64 numpy and cast by the caller.

Configurations (BASELINE.json `configs`, sized in SURVEY.md section 8(d)):
  c1  F=10, complete bipartite patch->frame graph, E=9600, t0=1
  c2  F=22, |frame(k)-j| <= 12 window graph in the reference's append order, E=37824, t0=12, t1=22
  c4  F=1000, temporal edges 1<=|d|<=2 plus 200 loop-closure frame pairs, E=402624, t0=1 (eff_impl)
  c5  c2-shaped windows with EuRoC intrinsics (calib/euroc.txt of the reference) and seeds 1234+s
"""
from dataclasses import dataclass, field
import numpy as np

M_DEFAULT = 96
P = 3


# ---------------------------------------------------------------- small SE3 helpers (generator only)
def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], -1)


def _rot(q, X):
    uv = 2.0 * _cross(q[..., :3], X)
    return X + q[..., 3:4] * uv + _cross(q[..., :3], uv)


def _qmul(a, b):
    ax, ay, az, aw = (a[..., i] for i in range(4))
    bx, by, bz, bw = (b[..., i] for i in range(4))
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz], -1)


def _exp(xi):
    tau, phi = xi[..., :3], xi[..., 3:]
    th2 = np.sum(phi * phi, -1, keepdims=True)
    th = np.sqrt(th2)
    small = th < 1e-6
    s = np.where(small, 1.0, th)
    q = np.concatenate([np.where(small, 0.5 - th2 / 48.0, np.sin(0.5 * s) / s) * phi,
                        np.where(small, 1.0 - th2 / 8.0, np.cos(0.5 * s))], -1)
    a = np.where(small, 0.5, (1 - np.cos(s)) / (s * s))
    b = np.where(small, 1.0 / 6.0, (s - np.sin(s)) / (s * s * s))
    c1 = _cross(phi, tau)
    t = tau + a * c1 + b * _cross(phi, c1)
    return t, q


def _retr(xi, t, q):
    dt, dq = _exp(xi)
    return _rot(dq, t) + dt, _qmul(dq, q)


def _project(poses, patches_c, intr, ii, jj, kk):
    """GT reprojection of patch centres (float64): coords [E,2]."""
    fx, fy, cx, cy = intr
    qi_c = np.concatenate([-poses[ii, 3:6], poses[ii, 6:7]], -1)
    qij = _qmul(poses[jj, 3:], qi_c)
    tij = poses[jj, :3] - _rot(qij, poses[ii, :3])
    x, y, d = patches_c[kk, 0], patches_c[kk, 1], patches_c[kk, 2]
    Xi = np.stack([(x - cx) / fx, (y - cy) / fy, np.ones_like(x)], -1)
    Xj = _rot(qij, Xi) + d[:, None] * tij
    return np.stack([fx * Xj[:, 0] / Xj[:, 2] + cx, fy * Xj[:, 1] / Xj[:, 2] + cy], -1)


# ---------------------------------------------------------------- edge rules
def window_edges(F, M=M_DEFAULT, lifetime=13):
    """Edges in the order slam.py appends them (slam.py:528-541 via append_factors :331-337): for every new
    frame f, forward edges (patches of the previous lifetime-1 frames -> f) then backward edges (patches of f ->
    the last lifetime frames incl. f), both kk-major."""
    kk_l, jj_l = [], []
    for f in range(F):
        lo = max(f - (lifetime - 1), 0)
        if f > lo:                                           # __edges_forw
            k = np.arange(lo * M, f * M)
            kk_l.append(k)
            jj_l.append(np.full_like(k, f))
        k = np.arange(f * M, (f + 1) * M)                    # __edges_back
        j = np.arange(lo, f + 1)
        kk_l.append(np.repeat(k, len(j)))
        jj_l.append(np.tile(j, len(k)))
    kk = np.concatenate(kk_l)
    jj = np.concatenate(jj_l)
    return kk // M, jj, kk


def bipartite_edges(F, M=M_DEFAULT):
    k = np.arange(F * M)
    kk = np.repeat(k, F)
    jj = np.tile(np.arange(F), F * M)
    return kk // M, jj, kk


def global_edges(F, M=M_DEFAULT, n_loops=200, rng=None):
    """c4: temporal edges 1<=|frame(k)-j|<=2 plus n_loops loop-closure frame pairs (j - i > 30), all patches."""
    rng = rng or np.random.default_rng(0)
    kk_l, jj_l = [], []
    for dlt in (-2, -1, 1, 2):
        f = np.arange(max(0, -dlt), min(F, F - dlt))
        k = (f[:, None] * M + np.arange(M)[None, :]).ravel()
        kk_l.append(k)
        jj_l.append(np.repeat(f + dlt, M))
    seen = set()
    while len(seen) < n_loops:
        i = int(rng.integers(0, F - 31))
        j = int(rng.integers(i + 31, F))
        seen.add((i, j))
    for (i, j) in sorted(seen):
        kk_l.append(i * M + np.arange(M))
        jj_l.append(np.full(M, j))
    kk = np.concatenate(kk_l)
    jj = np.concatenate(jj_l)
    return kk // M, jj, kk


# ---------------------------------------------------------------- problems
@dataclass
class Problem:
    name: str
    poses: np.ndarray          # [F,7] initial (perturbed) state, float64
    patches: np.ndarray        # [F*M,3,3,3]
    intrinsics: np.ndarray     # [F,4]
    target: np.ndarray         # [E,2]
    weight: np.ndarray         # [E,2]
    lmbda: float
    ii: np.ndarray
    jj: np.ndarray
    kk: np.ndarray
    t0: int
    t1: int
    M: int
    eff_impl: bool = False
    ht: int = 120
    wd: int = 160
    gt_poses: np.ndarray = field(default=None, repr=False)
    gt_patches: np.ndarray = field(default=None, repr=False)

    @property
    def E(self):
        return len(self.ii)

    @property
    def N(self):
        return self.t1 - self.t0


def make_problem(name, F, edges, t0, t1, seed=1234, M=M_DEFAULT, ht=120, wd=160,
                 intr=(80.0, 80.0, 80.0, 60.0), eff_impl=False, noise_px=0.5, pose_sigma=0.01):
    rng = np.random.default_rng(seed)
    poses = np.zeros((F, 7))
    poses[0, 6] = 1.0
    for f in range(1, F):
        xi = np.concatenate([np.array([0.05, 0, 0]) + rng.normal(0, 0.01, 3), rng.normal(0, 0.005, 3)])
        poses[f, :3], poses[f, 3:] = _retr(xi, poses[f - 1, :3], poses[f - 1, 3:])
    cx_i = rng.integers(8, wd - 8, size=F * M).astype(np.float64)
    cy_i = rng.integers(8, ht - 8, size=F * M).astype(np.float64)
    depth = rng.uniform(0.2, 1.0, size=F * M)
    off = np.arange(-1, 2, dtype=np.float64)
    patches = np.empty((F * M, 3, P, P))
    patches[:, 0] = cx_i[:, None, None] + off[None, None, :]
    patches[:, 1] = cy_i[:, None, None] + off[None, :, None]
    patches[:, 2] = depth[:, None, None]

    ii, jj, kk = edges
    centre = patches[:, :, 1, 1]
    gt = _project(poses, centre, intr, ii, jj, kk)
    target = gt + rng.normal(0, noise_px, gt.shape)
    weight = rng.uniform(0, 1, gt.shape) ** 2

    init_poses = poses.copy()
    free = np.arange(F) >= t0
    xi = rng.normal(0, pose_sigma, (F, 6))
    tn, qn = _retr(xi, poses[:, :3], poses[:, 3:])
    init_poses[free, :3] = tn[free]
    init_poses[free, 3:] = qn[free]
    init_patches = patches.copy()
    init_patches[:, 2] *= rng.uniform(0.8, 1.2, size=F * M)[:, None, None]
    intrinsics = np.tile(np.asarray(intr, np.float64), (F, 1))
    return Problem(name, init_poses, init_patches, intrinsics, target, weight, 1e-4,
                   ii.astype(np.int64), jj.astype(np.int64), kk.astype(np.int64), t0, t1, M, eff_impl, ht, wd,
                   poses, patches)


def config_c1(seed=1234, M=M_DEFAULT):
    return make_problem("c1", 10, bipartite_edges(10, M), 1, 10, seed, M)


def config_c2(seed=1234, M=M_DEFAULT, F=22, t0=12):
    return make_problem("c2", F, window_edges(F, M), t0, F, seed, M)


def config_c4(seed=1234, M=M_DEFAULT, F=1000, n_loops=200):
    rng = np.random.default_rng(seed + 7)
    return make_problem("c4", F, global_edges(F, M, n_loops, rng), 1, F, seed, M, eff_impl=True)


EUROC_INTR = (458.654 / 4, 457.296 / 4, 367.215 / 4, 248.375 / 4)


def config_c5_window(s, M=M_DEFAULT, F=22, t0=12):
    """One of the 64 independent EuRoC-shaped windows of c5 (480x752 image -> 120x188 features)."""
    return make_problem("c5[%d]" % s, F, window_edges(F, M), t0, F, 1234 + s, M, ht=120, wd=188, intr=EUROC_INTR)


def small_problem(seed=0, F=6, M=8, t0=2, lifetime=4, noise_px=0.5):
    """Tiny window graph for unit tests."""
    return make_problem("small", F, window_edges(F, M, lifetime), t0, F, seed, M, noise_px=noise_px)


# ---------------------------------------------------------------- feature maps for the correlation configs (c3)
def make_fmaps(problem, C=24, seed=1234, n_mem=36, levels=(1, 4), dtype=np.float32):
    """gmap [F*M, C, 3, 3] and a 2-level pyramid [[n_mem, C, H, W], [n_mem, C, H/4, W/4]], values N(0,1)/4
    (the reference network divides its feature maps by 4: net_cdv.py:284); level 1 is the 4x4 average pool of
    level 0 (slam.py:681-682)."""
    rng = np.random.default_rng(seed + 99)
    F = problem.poses.shape[0]
    gmap = (rng.standard_normal((F * problem.M, C, 3, 3)) / 4).astype(dtype)
    f0 = (rng.standard_normal((n_mem, C, problem.ht, problem.wd)) / 4).astype(np.float32)
    pyr = []
    for lv in levels:
        if lv == 1:
            pyr.append(f0.astype(dtype))
        else:
            h, w = problem.ht // lv, problem.wd // lv
            pyr.append(f0[:, :, :h * lv, :w * lv].reshape(n_mem, C, h, lv, w, lv).mean((3, 5)).astype(dtype))
    return gmap, pyr


def to_torch(p, device):
    """Problem -> dict of torch tensors in the reference's layouts (leading batch dim of 1) on `device`."""
    import torch
    f = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=device)
    l = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.int64, device=device)
    return dict(poses=f(p.poses)[None], patches=f(p.patches)[None], intrinsics=f(p.intrinsics)[None],
                target=f(p.target)[None], weight=f(p.weight)[None], lmbda=f([p.lmbda]), ii=l(p.ii), jj=l(p.jj),
                kk=l(p.kk))
