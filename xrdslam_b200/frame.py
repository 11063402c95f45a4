"""RGB-D frame + its optimisable pose (mirror of slam/common/frame.py:10-74)."""
from typing import List

import torch
import torch.nn as nn
from torch.nn import Parameter

from .opt_pose import OptimizablePose


class Frame(nn.Module):
    def __init__(self, fid, rgb, depth, init_pose=None, gt_pose=None,
                 separate_LR=False, rot_rep='axis_angle') -> None:
        super().__init__()
        self.fid = fid
        if depth is not None:
            self.h, self.w = depth.shape
        else:
            self.h, self.w = rgb.shape[0], rgb.shape[1]
        self.rgb = rgb
        self.depth = depth
        self.gt_pose = gt_pose
        self.separate_LR = separate_LR
        self.rot_rep = rot_rep
        self.is_final_frame = False
        self.pose = None
        if init_pose is not None:
            self.set_pose(init_pose, separate_LR, rot_rep)
            ref = torch.as_tensor(init_pose, dtype=torch.float32)
            # frame.py:40-43 round-trip consistency check
            if not torch.allclose(ref, self.pose.matrix().detach().cpu(),
                                  atol=1e-3):
                raise ValueError('Transformation inconsistency detected!', ref,
                                 self.pose.matrix())

    def set_pose(self, pose_np, separate_LR=False, rot_rep='axis_angle'):
        pose = torch.as_tensor(pose_np, dtype=torch.float32)
        self.pose = OptimizablePose.from_matrix(pose, separate_LR=separate_LR,
                                                rot_rep=rot_rep)

    def get_pose(self):
        return self.pose.matrix()

    def get_translation(self):
        return self.pose.translation()

    def get_rotation(self):
        return self.pose.rotation()

    def get_params(self) -> List[Parameter]:
        if self.pose is None:
            return []
        if self.separate_LR:
            rot = self.pose.data_q if self.rot_rep == 'quat' else self.pose.data_r
            return [rot, self.pose.data_t]
        return list(self.pose.parameters())
