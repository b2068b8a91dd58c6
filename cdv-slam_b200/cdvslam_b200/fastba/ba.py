"""Mirror of the reference's cdvslam/fastba/ba.py (same names, same argument order)."""
import cuda_ba
from cdvslam_b200 import native

last_status = native.last_ba_status     # extension: PGBA_ST_* bits of the most recent BA call (0 = every edge processed)
last_plan_hits = native.last_ba_plan_hits   # extension: windows of the most recent call that reused their plan tables
neighbors = cuda_ba.neighbors
reproject = cuda_ba.reproject


def BA(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, M, iterations, eff_impl=False):
    """In-place patch-graph bundle adjustment; signature of cdvslam/fastba/ba.py:7-8."""
    return cuda_ba.forward(poses.data, patches, intrinsics, target, weight, lmbda, ii, jj, kk, M, t0, t1, iterations,
                           eff_impl)


def BA_host(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, M, iterations, eff_impl=False,
            device=None):
    """Extension: BA() for tensors that live in (pinned) host memory -- same argument order and in-place semantics;
    upload, solve and download are enqueued on the current CUDA stream of `device`."""
    return cuda_ba.forward_host(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, M, t0, t1, iterations,
                                eff_impl, device=device)
