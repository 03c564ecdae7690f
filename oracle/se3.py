"""SE(3) helpers restated from the reference kernels (src/droid_kernels.cu:67-184, 886-904).

Pose layout everywhere: (tx, ty, tz, qx, qy, qz, qw), quaternions NOT renormalised (Q11).
All functions are batched over leading dimensions and work in the dtype of their inputs.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import torch

__all__ = ["act_so3", "act_se3", "adj_se3", "rel_se3", "exp_so3", "exp_se3", "retr_se3", "cross"]


def cross(a, b):
    return torch.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                        a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                        a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], dim=-1)


def act_so3(q, X):
    """src/droid_kernels.cu:67-77  Y = X + qw*uv + qv x uv, uv = 2 qv x X."""
    qv, qw = q[..., :3], q[..., 3:4]
    uv = 2.0 * cross(qv, X)
    return X + qw * uv + cross(qv, uv)


def act_se3(t, q, X):
    """src/droid_kernels.cu:79-86  homogeneous point (X,Y,Z,W): Y[:3] = R X[:3] + W t, Y[3] = W."""
    Y = act_so3(q, X[..., :3]) + X[..., 3:4] * t
    return torch.cat([Y, X[..., 3:4]], dim=-1)


def adj_se3(t, q, X):
    """src/droid_kernels.cu:88-103  transposed adjoint acting on a 6-vector X=(a,b):
    Y = (R^T a, R^T b + R^T (a x t))   (u = t[2]a[1]-t[1]a[2], ... = a x t)."""
    qinv = torch.cat([-q[..., :3], q[..., 3:4]], dim=-1)
    a, b = X[..., :3], X[..., 3:]
    Ya = act_so3(qinv, a)
    Yb = act_so3(qinv, b) + act_so3(qinv, cross(a, t))
    return torch.cat([Ya, Yb], dim=-1)


def rel_se3(ti, qi, tj, qj):
    """src/droid_kernels.cu:105-116  (tij, qij) = T_j * T_i^{-1}."""
    q = torch.stack([
        -qj[..., 3] * qi[..., 0] + qj[..., 0] * qi[..., 3] - qj[..., 1] * qi[..., 2] + qj[..., 2] * qi[..., 1],
        -qj[..., 3] * qi[..., 1] + qj[..., 1] * qi[..., 3] - qj[..., 2] * qi[..., 0] + qj[..., 0] * qi[..., 2],
        -qj[..., 3] * qi[..., 2] + qj[..., 2] * qi[..., 3] - qj[..., 0] * qi[..., 1] + qj[..., 1] * qi[..., 0],
        qj[..., 3] * qi[..., 3] + qj[..., 0] * qi[..., 0] + qj[..., 1] * qi[..., 1] + qj[..., 2] * qi[..., 2],
    ], dim=-1)
    tij = tj - act_so3(q, ti)
    return tij, q


def exp_so3(phi):
    """src/droid_kernels.cu:119-141."""
    theta_sq = (phi * phi).sum(-1, keepdim=True)
    theta_p4 = theta_sq * theta_sq
    theta = torch.sqrt(theta_sq)
    small = theta_sq < 1e-8
    th = torch.where(small, torch.ones_like(theta), theta)
    imag = torch.where(small, 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_p4, torch.sin(0.5 * th) / th)
    real = torch.where(small, 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_p4, torch.cos(0.5 * th))
    return torch.cat([imag * phi, real], dim=-1)


def exp_se3(xi):
    """src/droid_kernels.cu:156-184  xi = (tau, phi);  t = tau + a phi x tau + b phi x (phi x tau) if theta>1e-4."""
    tau, phi = xi[..., :3], xi[..., 3:]
    q = exp_so3(phi)
    theta_sq = (phi * phi).sum(-1, keepdim=True)
    theta = torch.sqrt(theta_sq)
    big = theta > 1e-4
    th = torch.where(big, theta, torch.ones_like(theta))
    thsq = torch.where(big, theta_sq, torch.ones_like(theta_sq))
    a = (1 - torch.cos(th)) / thsq
    b = (th - torch.sin(th)) / (th * thsq)
    c1 = cross(phi, tau)
    c2 = cross(phi, c1)
    t = torch.where(big, tau + a * c1 + b * c2, tau)
    return t, q


def retr_se3(xi, t, q):
    """src/droid_kernels.cu:886-904  T' = Exp(xi) * T (left multiplication), no renormalisation."""
    dt, dq = exp_se3(xi)
    q1 = torch.stack([
        dq[..., 3] * q[..., 0] + dq[..., 0] * q[..., 3] + dq[..., 1] * q[..., 2] - dq[..., 2] * q[..., 1],
        dq[..., 3] * q[..., 1] + dq[..., 1] * q[..., 3] + dq[..., 2] * q[..., 0] - dq[..., 0] * q[..., 2],
        dq[..., 3] * q[..., 2] + dq[..., 2] * q[..., 3] + dq[..., 0] * q[..., 1] - dq[..., 1] * q[..., 0],
        dq[..., 3] * q[..., 3] - dq[..., 0] * q[..., 0] - dq[..., 1] * q[..., 1] - dq[..., 2] * q[..., 2],
    ], dim=-1)
    t1 = act_so3(dq, t) + dt
    return t1, q1
