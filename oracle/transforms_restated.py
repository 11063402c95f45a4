"""Restatement of the three pytorch3d.transforms functions the reference uses
(slam/utils/opt_pose.py:8-10,69,100-101).  Real-first quaternions (w,x,y,z).

ORACLE / TEST INFRASTRUCTURE.  PARITY UNPINNED vs pytorch3d (not vendored, not
installed; the only in-repo pin is Frame.__init__'s round-trip assert,
slam/common/frame.py:40-43, atol 1e-3, which tests/test_pose.py replays).
"""
import torch


def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack(
        (
            1 - two_s * (j * j + k * k),
            two_s * (i * j - k * r),
            two_s * (i * k + j * r),
            two_s * (i * j + k * r),
            1 - two_s * (i * i + k * k),
            two_s * (j * k - i * r),
            two_s * (i * k - j * r),
            two_s * (j * k + i * r),
            1 - two_s * (i * i + j * j),
        ),
        -1,
    )
    return o.reshape(q.shape[:-1] + (3, 3))


def _sqrt_positive_part(x):
    return torch.sqrt(torch.clamp(x, min=0.0))


def matrix_to_quaternion(m):
    m00, m01, m02 = m[..., 0, 0], m[..., 0, 1], m[..., 0, 2]
    m10, m11, m12 = m[..., 1, 0], m[..., 1, 1], m[..., 1, 2]
    m20, m21, m22 = m[..., 2, 0], m[..., 2, 1], m[..., 2, 2]
    q_abs = _sqrt_positive_part(
        torch.stack([
            1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22,
            1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22
        ], -1))
    cand = torch.stack([
        torch.stack([q_abs[..., 0]**2, m21 - m12, m02 - m20, m10 - m01], -1),
        torch.stack([m21 - m12, q_abs[..., 1]**2, m10 + m01, m02 + m20], -1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2]**2, m12 + m21], -1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3]**2], -1),
    ], -2)
    cand = cand / (2.0 * q_abs[..., None].clamp(min=0.1))
    best = q_abs.argmax(-1)
    out = cand[..., best, :] if cand.dim() == 2 else torch.gather(
        cand, -2, best[..., None, None].expand(*best.shape, 1, 4)).squeeze(-2)
    # standardise to non-negative real part (recent pytorch3d)
    return torch.where(out[..., 0:1] < 0, -out, out)


def quaternion_to_axis_angle(q):
    norms = torch.norm(q[..., 1:], p=2, dim=-1, keepdim=True)
    half = torch.atan2(norms, q[..., :1])
    angles = 2 * half
    eps = 1e-6
    small = angles.abs() < eps
    k = torch.empty_like(angles)
    k[~small] = torch.sin(half[~small]) / angles[~small]
    k[small] = 0.5 - (angles[small] * angles[small]) / 48
    return q[..., 1:] / k
