"""SE(3) pose parameters (host-side mirror of slam/utils/opt_pose.py:13-109):
[t, axis-angle] or [t, quaternion(w,x,y,z)] -> 4x4, differentiable so that the
in-kernel d loss / d rays reach the pose through torch autograd."""
from copy import deepcopy

import torch
import torch.nn as nn

from .transforms import (matrix_to_quaternion, quaternion_to_axis_angle,
                         quaternion_to_matrix)


class OptimizablePose(nn.Module):
    def __init__(self, init_pose, separate_LR=True, rot_rep='axis_angle'):
        super().__init__()
        self.separate_LR = separate_LR
        self.rot_rep = rot_rep
        if separate_LR:
            if rot_rep == 'axis_angle':
                self.register_parameter('data_r', nn.Parameter(init_pose[3:]))
            elif rot_rep == 'quat':
                self.register_parameter('data_q', nn.Parameter(init_pose[3:]))
            else:
                raise ValueError(f'unsupported rotation representation {rot_rep}')
            self.register_parameter('data_t', nn.Parameter(init_pose[:3]))
        else:
            self.register_parameter('data', nn.Parameter(init_pose))

    def copy_from(self, pose):
        for name in ('data_r', 'data_q', 'data_t', 'data'):
            if hasattr(self, name) and hasattr(pose, name):
                setattr(self, name, deepcopy(getattr(pose, name)))

    def _rot_params(self):
        if not self.separate_LR:
            return self.data[3:]
        return self.data_r if self.rot_rep == 'axis_angle' else self.data_q

    def rotation(self):
        p = self._rot_params()
        if self.rot_rep == 'axis_angle':
            return self.axis_angle_to_rotation_matrix(p)
        return quaternion_to_matrix(p)

    def translation(self):
        return self.data_t if self.separate_LR else self.data[:3]

    def matrix(self):
        Rt = torch.eye(4, device=self.translation().device)
        Rt[:3, :3] = self.rotation()
        Rt[:3, 3] = self.translation()
        return Rt

    @staticmethod
    def axis_angle_to_rotation_matrix(angle_axis):
        """Rodrigues (opt_pose.py:77-95)."""
        angle = torch.norm(angle_axis, dim=-1, keepdim=True)
        if torch.allclose(angle, torch.zeros_like(angle)):
            return torch.eye(3, device=angle_axis.device, dtype=angle_axis.dtype)
        w0, w1, w2 = (angle_axis / angle).unbind(dim=-1)
        z = torch.zeros_like(w0)
        wx = torch.stack([
            torch.stack([z, -w2, w1], dim=-1),
            torch.stack([w2, z, -w0], dim=-1),
            torch.stack([-w1, w0, z], dim=-1)
        ], dim=-2)
        eye = torch.eye(3, device=angle_axis.device, dtype=angle_axis.dtype)
        return eye + wx * torch.sin(angle) + (1. - torch.cos(angle)) * (wx @ wx)

    @classmethod
    def from_matrix(cls, Rt, separate_LR=True, rot_rep='axis_angle'):
        R, u = Rt[:3, :3], Rt[:3, 3]
        quat = matrix_to_quaternion(R)
        if rot_rep == 'axis_angle':
            rot = quaternion_to_axis_angle(quat)
        elif rot_rep == 'quat':
            rot = quat
        else:
            raise ValueError(rot_rep)
        return OptimizablePose(torch.cat([u, rot], dim=-1).detach().clone(),
                               separate_LR=separate_LR, rot_rep=rot_rep)


def pose_matrices(poses, detach=None):
    """[n,4,4] c2w of a list of OptimizablePose in ONE batched evaluation (the per-frame
    `matrix()` costs ~25 small host ops each; bundle adjustment evaluates 5-20 per
    iteration).  `detach`: optional list of bools (coslam.py:181-182 fixes frame 0)."""
    n = len(poses)
    rep, sep = poses[0].rot_rep, poses[0].separate_LR
    if any(p.rot_rep != rep or p.separate_LR != sep for p in poses):
        M = torch.stack([p.matrix() for p in poses])
    else:
        rot = torch.stack([p._rot_params() for p in poses])
        t = torch.stack([p.translation() for p in poses])
        if rep == 'axis_angle':
            angle = torch.norm(rot, dim=-1, keepdim=True)
            zero = angle <= 1e-8  # torch.allclose(angle, 0) of the per-frame path
            safe = torch.where(zero, torch.ones_like(angle), angle)
            w0, w1, w2 = (rot / safe).unbind(dim=-1)
            z = torch.zeros_like(w0)
            wx = torch.stack([torch.stack([z, -w2, w1], dim=-1),
                              torch.stack([w2, z, -w0], dim=-1),
                              torch.stack([-w1, w0, z], dim=-1)], dim=-2)
            eye = torch.eye(3, dtype=rot.dtype, device=rot.device).expand(n, 3, 3)
            s, c = torch.sin(safe)[..., None], torch.cos(safe)[..., None]
            Rm = eye + wx * s + (1. - c) * (wx @ wx)
            Rm = torch.where(zero[..., None], eye, Rm)
        else:
            Rm = quaternion_to_matrix(rot)
        bottom = torch.tensor([0., 0., 0., 1.], dtype=rot.dtype, device=rot.device)
        M = torch.cat([torch.cat([Rm, t[..., None]], -1), bottom.expand(n, 1, 4)], -2)
    if detach is not None and any(detach):
        m = torch.tensor(detach, device=M.device)[:, None, None]
        M = torch.where(m, M.detach(), M)
    return M
