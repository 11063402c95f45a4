"""Torch-CPU restatement of the NICE-SLAM render-and-optimise step.

ORACLE / TEST INFRASTRUCTURE -- never imported by xrdslam_b200/.

Follows (reference @ f0366f20, paths relative to /root/reference):
  slam/models/conv_onet.py:324-337   load_bound   (int32*float -> float32 hazard)
  slam/models/conv_onet.py:254-291   grid_init  + slam/model_components/feature_grid_nice.py
  slam/models/conv_onet.py:377-524   render_batch_ray (far from bbox, 32 uniform + 16 surface
                                     samples in float64, sort)
  slam/models/conv_onet.py:339-375   eval_points (out-of-bound -> occupancy logit 100)
  slam/model_components/decoder_nice.py:386-414  NICE.forward(stage)
  slam/model_components/decoder_nice.py:195-234  MLP.sample_grid_feature / forward
  slam/common/common.py:16-31        normalize_3d_coordinate
  slam/model_components/utils.py:189-244  raw2outputs_nerf_color (occupancy)
  slam/models/conv_onet.py:145-185   get_loss_dict
Pinned bit-identically against the reference's own ConvOnet, all three stages, by
tests/test_oracle_cpu.py::test_nice_oracle_matches_reference_class_live (stage 'color' runs on
the host as is; for 'middle' / 'fine' the hard-coded 'cuda:N' device string of
decoder_nice.py:388, SURVEY Q5, is neutralised by oracle/ref_harness.cuda_calls_are_noops) and
by the committed vectors tests/golden/nice_*.npz.

Grids are held in the reference layout [1, C, Z, Y, X]; F.grid_sample is torch's own.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class NiceCfg:
    c_dim: int = 32
    hidden: int = 32
    grid_len_coarse: float = 2.0
    coarse_bound_enlarge: int = 2
    grid_len_middle: float = 0.32
    grid_len_fine: float = 0.16
    grid_len_color: float = 0.16
    bound_divisible: float = 0.32
    n_samples: int = 32
    n_surface: int = 16
    tracking_w_color: float = 0.5
    mapping_w_color: float = 0.2
    handle_dynamic: bool = True
    use_color_in_tracking: bool = True
    points_batch_size: int = 500000


def load_bound(bounding_box, divisible=0.32):
    """conv_onet.py:324-330 with its exact dtype chain: (f64/py-float).int() -> int32;
    (int32 + 1) * python float -> float32 (!); + f64 -> stored into the f64 tensor."""
    bb = torch.as_tensor(np.asarray(bounding_box), dtype=torch.float64).clone()
    bb[:, 1] = (((bb[:, 1] - bb[:, 0]) / divisible).int() + 1) * divisible + bb[:, 0]
    return bb


def grid_shape(bb, grid_len):
    """feature_grid_nice.py:8-10 -> (Z, Y, X)."""
    xyz_len = bb[:, 1] - bb[:, 0]
    s = list(map(int, (xyz_len / grid_len).tolist()))
    s[0], s[2] = s[2], s[0]
    return s


class DecoderMLP(nn.Module):
    """decoder_nice.py:101-234 (pos_embedding 'fourier', hidden 32, 5 blocks, skip at 2)."""
    def __init__(self, c_dim, color, concat_feature, hidden=32, gen=None):
        super().__init__()
        self.color = color
        self.concat = concat_feature
        self.B = nn.Parameter(torch.randn(3, 93, generator=gen) * 25)
        self.fc_c = nn.ModuleList([nn.Linear(c_dim, hidden) for _ in range(5)])
        dims = [93, hidden, hidden, hidden + 93, hidden]
        self.pts = nn.ModuleList([nn.Linear(d, hidden) for d in dims])
        self.out = nn.Linear(hidden, 4 if color else 1)
        for lin in list(self.pts) + [self.out]:
            act = 'linear' if lin is self.out else 'relu'
            nn.init.xavier_uniform_(lin.weight, gain=nn.init.calculate_gain(act))
            nn.init.zeros_(lin.bias)

    def forward(self, p, c):
        e = torch.sin(p.float() @ self.B)
        h = e
        for i in range(5):
            h = F.relu(self.pts[i](h)) + self.fc_c[i](c)
            if i == 2:
                h = torch.cat([e, h], -1)
        out = self.out(h)
        return out if self.color else out.squeeze(-1)


class CoarseMLP(nn.Module):
    """MLP_no_xyz (decoder_nice.py:237-320): the grid feature is the only input; 5 blocks of
    width 32, the feature re-concatenated after block 2, no positional embedding, no fc_c."""
    def __init__(self, c_dim=32, hidden=32, gen=None):
        super().__init__()
        dims = [hidden, hidden, hidden, hidden + c_dim, hidden]
        self.pts = nn.ModuleList([nn.Linear(d, hidden) for d in dims])
        self.out = nn.Linear(hidden, 1)
        for lin in list(self.pts) + [self.out]:
            act = 'linear' if lin is self.out else 'relu'
            nn.init.xavier_uniform_(lin.weight, gain=nn.init.calculate_gain(act))
            nn.init.zeros_(lin.bias)

    def forward(self, c):
        h = c
        for i in range(5):
            h = F.relu(self.pts[i](h))
            if i == 2:
                h = torch.cat([c, h], -1)
        return self.out(h).squeeze(-1)


def normalize_3d(p, bound):
    p = p.reshape(-1, 3).clone()
    for d in range(3):
        p[:, d] = ((p[:, d] - bound[d, 0]) / (bound[d, 1] - bound[d, 0])) * 2 - 1.0
    return p


def sample_grid(p, grid, bound):
    """decoder_nice.py:195-205 -> [P, C]."""
    p_nor = normalize_3d(p, bound).unsqueeze(0)
    vgrid = p_nor[:, :, None, None].float()
    c = F.grid_sample(grid, vgrid, padding_mode='border', align_corners=True,
                      mode='bilinear').squeeze(-1).squeeze(-1)
    return c.transpose(1, 2).squeeze(0)


class NiceOracle(nn.Module):
    def __init__(self, bounding_box, cfg: NiceCfg = None, seed=0, coarse=False):
        super().__init__()
        self.cfg = cfg or NiceCfg()
        c = self.cfg
        self.bound = load_bound(bounding_box, c.bound_divisible)
        g = torch.Generator().manual_seed(seed)
        self.middle = DecoderMLP(c.c_dim, False, False, c.hidden, g)
        self.fine = DecoderMLP(2 * c.c_dim, False, True, c.hidden, g)
        self.color = DecoderMLP(c.c_dim, True, False, c.hidden, g)
        self.grids = nn.ParameterDict()
        for key, gl, std in (('grid_middle', c.grid_len_middle, 0.01),
                             ('grid_fine', c.grid_len_fine, 0.0001),
                             ('grid_color', c.grid_len_color, 0.01)):
            shp = [1, c.c_dim] + grid_shape(self.bound, gl)
            self.grids[key] = nn.Parameter(torch.zeros(shp).normal_(0, std, generator=g))
        # coarse level (conv_onet.py:256-275, :335-337): own decoder, grid over the bound
        # scaled by model_coarse_bound_enlarge, sampled with that scaled bound
        self.coarse = None
        if not coarse:
            return
        self.coarse = CoarseMLP(c.c_dim, c.hidden, g)
        xyz_len = (self.bound[:, 1] - self.bound[:, 0]) * c.coarse_bound_enlarge
        s = list(map(int, (xyz_len / c.grid_len_coarse).tolist()))
        s[0], s[2] = s[2], s[0]
        self.grids['grid_coarse'] = nn.Parameter(
            torch.zeros([1, c.c_dim] + s).normal_(0, 0.01, generator=g))
        self.coarse_bound = self.bound * c.coarse_bound_enlarge

    # --- NICE.forward (decoder_nice.py:386-414) --------------------------------
    def decode(self, p, stage):
        b = self.bound
        if stage == 'coarse':
            raw = torch.zeros(p.shape[0], 4)
            raw[..., -1] = self.coarse(sample_grid(p, self.grids['grid_coarse'], self.coarse_bound))
            return raw
        c_mid = sample_grid(p, self.grids['grid_middle'], b)
        mid = self.middle(p, c_mid)
        raw = torch.zeros(p.shape[0], 4)
        if stage == 'middle':
            raw[..., -1] = mid
            return raw
        c_fine = sample_grid(p, self.grids['grid_fine'], b)
        with torch.no_grad():
            c_mid_ng = sample_grid(p, self.grids['grid_middle'], b)
        fine = self.fine(p, torch.cat([c_fine, c_mid_ng], 1))
        if stage == 'fine':
            raw[..., -1] = fine + mid
            return raw
        c_col = sample_grid(p, self.grids['grid_color'], b)
        rawc = self.color(p, c_col)
        raw = torch.cat([rawc[:, :3], (fine + mid)[:, None]], -1)
        return raw

    def eval_points(self, p, stage):
        b = self.bound
        mask = ((p[:, 0] < b[0][1]) & (p[:, 0] > b[0][0]) & (p[:, 1] < b[1][1]) &
                (p[:, 1] > b[1][0]) & (p[:, 2] < b[2][1]) & (p[:, 2] > b[2][0]))
        ret = self.decode(p, stage)
        occ = torch.where(mask, ret[:, 3], torch.full_like(ret[:, 3], 100.0))
        return torch.cat([ret[:, :3], occ[:, None]], -1)

    # --- render_batch_ray (conv_onet.py:377-524, rendering_perturb = 0) ---------
    def sample_z(self, rays_o, rays_d, gt_depth):
        c = self.cfg
        N_samples, N_surface = c.n_samples, c.n_surface
        if gt_depth is None:  # stage 'coarse' (conv_onet.py:137-138, :397-402): 32 uniform samples
            with torch.no_grad():
                det_o = rays_o.clone().detach().unsqueeze(-1)
                det_d = rays_d.clone().detach().unsqueeze(-1)
                t = (self.bound.unsqueeze(0) - det_o) / det_d
                far_bb, _ = torch.min(torch.max(t, dim=2)[0], dim=1)
                far_bb = far_bb.unsqueeze(-1)
                far_bb += 0.01
            t_vals = torch.linspace(0., 1., steps=N_samples)
            return 0.01 * (1. - t_vals) + far_bb * t_vals
        gt_depth = gt_depth.reshape(-1, 1)
        near = gt_depth.repeat(1, N_samples) * 0.01
        with torch.no_grad():
            det_o = rays_o.clone().detach().unsqueeze(-1)
            det_d = rays_d.clone().detach().unsqueeze(-1)
            t = (self.bound.unsqueeze(0) - det_o) / det_d
            far_bb, _ = torch.min(torch.max(t, dim=2)[0], dim=1)
            far_bb = far_bb.unsqueeze(-1)
            far_bb += 0.01
        far = torch.clamp(far_bb, 0, torch.max(gt_depth * 1.2))
        nz = gt_depth > 0
        gt_nz = gt_depth[nz].unsqueeze(-1)
        t_s = torch.linspace(0., 1., steps=N_surface).double()
        z_nz = 0.95 * gt_nz.repeat(1, N_surface) * (1. - t_s) + \
            1.05 * gt_nz.repeat(1, N_surface) * t_s
        z_surf = torch.zeros(gt_depth.shape[0], N_surface).double()
        nzm = nz.squeeze(-1)
        z_surf[nzm, :] = z_nz
        z_zero = 0.001 * (1. - t_s) + torch.max(gt_depth) * t_s
        z_surf[~nzm, :] = z_zero
        t_vals = torch.linspace(0., 1., steps=N_samples)
        z_vals = near * (1. - t_vals) + far * t_vals
        z_vals, _ = torch.sort(torch.cat([z_vals, z_surf.double()], -1), -1)
        return z_vals

    def render(self, rays_o, rays_d, gt_depth, stage):
        N = rays_o.shape[0]
        z_vals = self.sample_z(rays_o, rays_d, None if stage == 'coarse' else gt_depth)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
        raw = self.eval_points(pts.reshape(-1, 3), stage).reshape(N, z_vals.shape[1], -1)
        rgb = raw[..., :-1]
        alpha = torch.sigmoid(10.0 * raw[..., -1])
        weights = alpha.float() * torch.cumprod(
            torch.cat([torch.ones((N, 1)), (1. - alpha + 1e-10).float()], -1).float(),
            -1)[:, :-1]
        rgb_map = torch.sum(weights[..., None] * rgb, -2)
        depth_map = torch.sum(weights * z_vals, -1)
        tmp = z_vals - depth_map.unsqueeze(-1)
        depth_var = torch.sum(weights * tmp * tmp, dim=1)
        return dict(rgb=rgb_map, depth=depth_map, uncertainty=depth_var, z_vals=z_vals,
                    weights=weights, raw=raw)

    # --- get_loss_dict (conv_onet.py:145-185) ------------------------------------
    def loss_dict(self, out, target_s, target_d, is_mapping, stage):
        c = self.cfg
        target_d = target_d.squeeze()
        depth, color = out['depth'], out['rgb']
        unc = out['uncertainty'].detach()
        d = {}
        if not is_mapping:
            if c.handle_dynamic:
                tmp = torch.abs(target_d - depth) / torch.sqrt(unc + 1e-10)
                mask = (tmp < 10 * tmp.median()) & (target_d > 0)
            else:
                mask = target_d > 0
            d['depth_loss'] = (torch.abs(target_d - depth) / torch.sqrt(unc + 1e-10))[mask].sum()
            if c.use_color_in_tracking:
                d['rgb_loss'] = c.tracking_w_color * torch.abs(target_s - color)[mask].sum()
        else:
            mask = target_d > 0
            d['depth_loss'] = torch.abs(target_d[mask] - depth[mask]).sum()
            if stage == 'color':
                d['rgb_loss'] = c.mapping_w_color * torch.abs(target_s - color).sum()
        return d

    def step(self, rays_o, rays_d, target_s, target_d, is_mapping, stage):
        out = self.render(rays_o, rays_d, target_d, stage)
        ld = self.loss_dict(out, target_s, target_d, is_mapping, stage)
        total = None
        for v in ld.values():
            total = v if total is None else total + v
        return out, ld, total
