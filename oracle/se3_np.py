"""numpy SE3 helpers (oracle; test infrastructure only).

Vectorised restatement of the device helpers in cdvslam/fastba/ba_cuda.cu:36-174 of the reference.
Quaternions are (qx, qy, qz, qw); poses are (t, q) world->camera; tangent vectors are (tau, phi).
All functions work on arrays with arbitrary leading batch dimensions and keep the input dtype.
"""
import numpy as np


def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def act_so3(q, X):
    """Rotate X by unit quaternion q  (ba_cuda.cu:36-46: uv = 2 q.vec x X ; Y = X + w uv + q.vec x uv)."""
    qv, qw = q[..., :3], q[..., 3:4]
    uv = 2.0 * _cross(qv, X)
    return X + qw * uv + _cross(qv, uv)


def act_se3(t, q, X4):
    """Homogeneous action on (X, Y, Z, W)  (ba_cuda.cu:48-55)."""
    Y = act_so3(q, X4[..., :3]) + X4[..., 3:4] * t
    return np.concatenate([Y, X4[..., 3:4]], axis=-1)


def quat_conj(q):
    return np.concatenate([-q[..., :3], q[..., 3:4]], axis=-1)


def quat_mul(a, b):
    """Hamilton product a (x) b in (x, y, z, w) layout (same expansion as ba_cuda.cu:165-168)."""
    ax, ay, az, aw = (a[..., i] for i in range(4))
    bx, by, bz, bw = (b[..., i] for i in range(4))
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def rel_se3(ti, qi, tj, qj):
    """Gij = Gj * Gi^-1  (ba_cuda.cu:74-85): qij = qj (x) conj(qi), tij = tj - R(qij) ti."""
    qij = quat_mul(qj, quat_conj(qi))
    tij = tj - act_so3(qij, ti)
    return tij, qij


def adj_se3(t, q, X6):
    """The map the reference calls adjSE3 (ba_cuda.cu:57-72), i.e. Ad(G)^T applied to a 6-vector:
    Y[:3] = R^T X[:3] ;  Y[3:] = R^T X[3:] + R^T (X[:3] x t)."""
    qinv = quat_conj(q)
    a = act_so3(qinv, X6[..., :3])
    b = act_so3(qinv, X6[..., 3:])
    u = _cross(X6[..., :3], t)     # u = -(t x X[:3]) as written at ba_cuda.cu:63-66
    v = act_so3(qinv, u)
    return np.concatenate([a, b + v], axis=-1)


def exp_so3(phi):
    """SO3 exponential to a quaternion with the reference's Taylor switch theta^2 < 1e-8 (ba_cuda.cu:88-110)."""
    theta_sq = np.sum(phi * phi, axis=-1, keepdims=True)
    theta_p4 = theta_sq * theta_sq
    theta = np.sqrt(theta_sq)
    small = theta_sq < 1e-8
    safe = np.where(small, 1.0, theta)
    imag = np.where(small, 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_p4,
                    np.sin(0.5 * safe) / safe)
    real = np.where(small, 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_p4,
                    np.cos(0.5 * safe))
    return np.concatenate([imag * phi, real], axis=-1).astype(phi.dtype)


def exp_se3(xi):
    """SE3 exponential (ba_cuda.cu:125-153): closed-form V(phi) tau only when theta > 1e-4, else t = tau."""
    tau, phi = xi[..., :3], xi[..., 3:]
    q = exp_so3(phi)
    theta_sq = np.sum(phi * phi, axis=-1, keepdims=True)
    theta = np.sqrt(theta_sq)
    big = theta > 1e-4
    safe_sq = np.where(big, theta_sq, 1.0)
    safe = np.where(big, theta, 1.0)
    a = (1.0 - np.cos(safe)) / safe_sq
    b = (safe - np.sin(safe)) / (safe * safe_sq)
    c1 = _cross(phi, tau)
    c2 = _cross(phi, c1)
    t = tau + np.where(big, a * c1 + b * c2, 0.0)
    return t.astype(xi.dtype), q


def retr_se3(xi, t, q):
    """Left retraction Exp(xi) * (t, q) without re-normalising q (ba_cuda.cu:156-174)."""
    dt, dq = exp_se3(xi)
    q1 = quat_mul(dq, q)
    t1 = act_so3(dq, t) + dt
    return t1, q1


def quat_to_matrix(q):
    """Rotation matrix of a unit quaternion (columns = images of the basis vectors under act_so3)."""
    eye = np.eye(3, dtype=q.dtype)
    cols = [act_so3(q, np.broadcast_to(eye[c], q.shape[:-1] + (3,))) for c in range(3)]
    return np.stack(cols, axis=-1)
