"""Rotation conversions used by OptimizablePose: the three functions the
reference imports from pytorch3d.transforms (slam/utils/opt_pose.py:8-10),
written out here because pytorch3d is not a dependency of this package.
Real-first quaternions (w, x, y, z)."""
import torch


def quaternion_to_matrix(quaternions: torch.Tensor) -> torch.Tensor:
    w, x, y, z = torch.unbind(quaternions, -1)
    s = 2.0 / (quaternions * quaternions).sum(-1)
    rows = (1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w),
            s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w),
            s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y))
    return torch.stack(rows, -1).reshape(quaternions.shape[:-1] + (3, 3))


def matrix_to_quaternion(matrix: torch.Tensor) -> torch.Tensor:
    m = matrix
    t = torch.stack([
        1.0 + m[..., 0, 0] + m[..., 1, 1] + m[..., 2, 2],
        1.0 + m[..., 0, 0] - m[..., 1, 1] - m[..., 2, 2],
        1.0 - m[..., 0, 0] + m[..., 1, 1] - m[..., 2, 2],
        1.0 - m[..., 0, 0] - m[..., 1, 1] + m[..., 2, 2]
    ], -1)
    q_abs = torch.sqrt(torch.clamp(t, min=0.0))
    a, b, c = m[..., 2, 1] - m[..., 1, 2], m[..., 0, 2] - m[..., 2, 0], \
        m[..., 1, 0] - m[..., 0, 1]
    d, e, f = m[..., 1, 0] + m[..., 0, 1], m[..., 0, 2] + m[..., 2, 0], \
        m[..., 1, 2] + m[..., 2, 1]
    cands = torch.stack([
        torch.stack([q_abs[..., 0]**2, a, b, c], -1),
        torch.stack([a, q_abs[..., 1]**2, d, e], -1),
        torch.stack([b, d, q_abs[..., 2]**2, f], -1),
        torch.stack([c, e, f, q_abs[..., 3]**2], -1)
    ], -2) / (2.0 * q_abs[..., None].clamp(min=0.1))
    pick = q_abs.argmax(-1)
    idx = pick[..., None, None].expand(*pick.shape, 1, 4)
    q = torch.gather(cands, -2, idx).squeeze(-2)
    return torch.where(q[..., :1] < 0, -q, q)


def quaternion_to_axis_angle(quaternions: torch.Tensor) -> torch.Tensor:
    n = torch.norm(quaternions[..., 1:], p=2, dim=-1, keepdim=True)
    half = torch.atan2(n, quaternions[..., :1])
    ang = 2 * half
    small = ang.abs() < 1e-6
    k = torch.where(small, 0.5 - ang * ang / 48,
                    torch.sin(half) / torch.where(small, torch.ones_like(ang), ang))
    return quaternions[..., 1:] / k
