"""Torch-CPU restatement of the Point-SLAM render-and-optimise step (stage 'geometry').

ORACLE / TEST INFRASTRUCTURE -- never imported by xrdslam_b200/.

Follows (reference @ f0366f20):
  slam/model_components/neural_point_cloud.py:223-282  find_neighbors_faiss -- faiss-gpu
      IndexIVFFlat(nlist 400, nprobe 4) is un-vendored and approximate: parity is DEFINED
      against the exact radius-independent 8-NN below (ties by lower id, faiss sentinels
      id -1 / D FLT_MAX) -> "parity unpinned vs real faiss" 
  slam/model_components/decoder_pointslam.py:162-273    MLP_geometry
  slam/models/conv_onet_pointslam.py:311-461, :144-195  render_batch_ray, get_loss_dict
  slam/model_components/utils.py:247-295                raw2outputs_nerf_color2
Pinned against the reference's own ConvOnet2 / POINT classes (run on CPU with the exact-kNN
faiss stand-in of oracle/ref_harness.py) by tests/test_oracle_cpu.py and tests/golden.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

FLT_MAX = float(np.finfo(np.float32).max)


def _exact_knn_sort(cloud, q, k):
    d = (cloud[None, :, :] - q[:, None, :])
    D = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    N = cloud.shape[0]
    if N < k:
        pad = torch.full((q.shape[0], k - N), FLT_MAX)
        D = torch.cat([D, pad], 1)
    vals, idx = torch.sort(D, dim=1, stable=True)
    vals, idx = vals[:, :k], idx[:, :k]
    idx = torch.where(idx < N, idx, torch.full_like(idx, -1))
    return vals, idx


def exact_knn(cloud, q, k=8, chunk=1024):
    """Exact k-NN (squared L2, float32, ((dx^2 + dy^2) + dz^2)), ties by lower id.
    Large problems run in query chunks with top-(k+1) selection; the selected k are then
    ordered by (distance, id) and any row whose k-th and (k+1)-th distances tie falls back to
    the full stable sort -- the result is the stable sort's for every row."""
    N = cloud.shape[0]
    if N <= 4 * k or q.shape[0] * N <= (1 << 22):
        return _exact_knn_sort(cloud, q, k)
    vals_all, idx_all = [], []
    for s in range(0, q.shape[0], chunk):
        qc = q[s:s + chunk]
        d = (cloud[None, :, :] - qc[:, None, :])
        D = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        v, i = torch.topk(D, k + 1, dim=1, largest=False)
        # order the candidates by (distance, id): two stable sorts
        o = torch.argsort(i, dim=1, stable=True)
        v, i = torch.gather(v, 1, o), torch.gather(i, 1, o)
        o = torch.argsort(v, dim=1, stable=True)
        v, i = torch.gather(v, 1, o), torch.gather(i, 1, o)
        tie = v[:, k] == v[:, k - 1]
        v, i = v[:, :k].clone(), i[:, :k].clone()
        for r in torch.nonzero(tie).flatten().tolist():
            vr, ir = _exact_knn_sort(cloud, qc[r:r + 1], k)
            v[r], i[r] = vr[0], ir[0]
        # equal distances inside the top k: the stable sort lists the lower id first -- done
        vals_all.append(v)
        idx_all.append(i)
    return torch.cat(vals_all), torch.cat(idx_all)


class GeoDecoder(nn.Module):
    """MLP_geometry (decoder_pointslam.py:77-273): sin(2 pi p B), 5 x 32 + fc_c, skip at 2."""
    def __init__(self, gen=None):
        super().__init__()
        self.B = nn.Parameter(torch.randn(3, 93, generator=gen) * 25)
        self.fc_c = nn.ModuleList([nn.Linear(32, 32) for _ in range(5)])
        self.pts = nn.ModuleList([nn.Linear(d, 32) for d in (93, 32, 32, 125, 32)])
        self.out = nn.Linear(32, 1)

    def forward(self, p, c):
        e = torch.sin((2 * math.pi * p.float()) @ self.B)
        h = e
        for i in range(5):
            h = F.relu(self.pts[i](h)) + self.fc_c[i](c)
            if i == 2:
                h = torch.cat([e, h], -1)
        return self.out(h).squeeze(-1)


class ColorDecoder(nn.Module):
    """MLP_color + MLP_col_neighbor (decoder_pointslam.py:276-292, 313-542): per-neighbour
    F_theta([sin, cos(2 pi rel B_rel), col_feat]) weighted by the inverse-distance weights,
    then a 5 x 128 softplus(beta=100) trunk on [sin, cos](2 pi p B) with fc_c(c) added after
    every block and the embedding re-concatenated after block 2; sigmoid output."""
    def __init__(self, gen=None):
        super().__init__()
        self.B = nn.Parameter(torch.randn(3, 20, generator=gen) * 32, requires_grad=False)
        self.B_rel = nn.Parameter(torch.randn(3, 10, generator=gen) * 32)
        self.nb1, self.nb2 = nn.Linear(52, 128), nn.Linear(128, 32)
        self.fc_c = nn.ModuleList([nn.Linear(32, 128) for _ in range(5)])
        self.pts = nn.ModuleList([nn.Linear(d, 128) for d in (40, 128, 128, 168, 128)])
        self.out = nn.Linear(128, 3)

    @staticmethod
    def embed(x, B):
        x = (2 * math.pi * x) @ B
        return torch.cat([torch.sin(x), torch.cos(x)], -1)

    def neighbor_feats(self, rel, feats):
        """rel [P,8,3], feats [P,8,32] -> [P,8,32]"""
        e = self.embed(rel.reshape(-1, 3), self.B_rel).reshape(rel.shape[0], -1, 20)
        return self.nb2(F.softplus(self.nb1(torch.cat([e, feats], -1)), beta=100))

    def forward(self, p, c):
        e = self.embed(p.float(), self.B)
        h = e
        for i in range(5):
            h = F.softplus(self.pts[i](h), beta=100) + self.fc_c[i](c)
            if i == 2:
                h = torch.cat([e, h], -1)
        return torch.sigmoid(self.out(h))


class PointOracle(nn.Module):
    def __init__(self, n_surface=5, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.geo = GeoDecoder(g)
        self.col = ColorDecoder(g)
        self.col_feats = None
        self.w_color_map, self.w_color_trk, self.use_color_trk = 0.1, 0.5, True
        self.n_surface = n_surface
        self.near_s, self.far_s, self.near_end, self.coef, self.min_nn = 0.98, 1.02, 0.3, 0.1, 2
        self.cloud = None
        self.geo_feats = None
        self.frustum_mask = None

    def set_cloud(self, pos, geo_feats, mask=None, col_feats=None):
        self.cloud = pos.clone()
        self.geo_feats = nn.Parameter(geo_feats.clone())
        if col_feats is not None:
            self.col_feats = nn.Parameter(col_feats.clone())
        self.frustum_mask = mask if mask is not None else torch.ones(pos.shape[0], 1, dtype=torch.bool)

    def feature_at(self, p, radius, rand_feat, color=False):
        D, I = exact_knn(self.cloud, p.detach(), 8)
        bound = radius.reshape(-1, 1)**2
        nn_num = (D < bound).sum(-1)
        Dg = torch.sum(torch.square(self.cloud[I] - p.reshape(-1, 1, 3)), dim=-1)  # is_tracker
        has = nn_num > self.min_nn - 1
        w = 1.0 / (Dg + 1e-10)
        w = torch.where(Dg > bound, torch.zeros_like(w), w)
        w = torch.where(I < 0, torch.zeros_like(w), w)
        w = F.normalize(w, p=1, dim=1).unsqueeze(-1)
        if color:  # colour feats are NOT frustum-masked (decoder_pointslam.py:493)
            nf = self.col.neighbor_feats(self.cloud[I] - p[:, None, :], self.col_feats[I])
        else:
            nf = (self.geo_feats * self.frustum_mask)[I]
        c = (w * nf).sum(1)
        c = torch.where(has[:, None], c, rand_feat[None, :].expand_as(c))
        return c, has

    def sample_z(self, target_d):
        S = self.n_surface
        gt = target_d.reshape(-1, 1)
        far = torch.minimum(5 * gt.mean(), torch.max(gt * 1.2))
        nz = (gt > 0).squeeze(-1)
        t = torch.linspace(0.0, 1.0, steps=S)
        z = torch.zeros(gt.shape[0], S)
        gs = gt[nz].repeat(1, S)
        z[nz] = self.near_s * gs * (1. - t) + self.far_s * gs * t
        if nz.sum() < gt.shape[0]:
            z[~nz] = torch.linspace(self.near_end, float(far), steps=S).repeat(int((~nz).sum()), 1)
        return z, nz

    def render(self, rays_o, rays_d, target_d, radius, rand_feat, stage='geometry',
               rand_feat_color=None):
        S = self.n_surface
        z, nz = self.sample_z(target_d)
        pts = (rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]).reshape(-1, 3)
        r = radius.reshape(-1, 1).repeat_interleave(S, dim=0)
        c, has = self.feature_at(pts, r, rand_feat)
        valid_ray = ~(torch.sum(has.view(-1, S), 1) < int(S / 2 + 1))
        occ = self.geo(pts, c)
        occ = torch.where(has, occ, torch.full_like(occ, -100.0))
        occ = occ.reshape(-1, S)
        alpha = torch.sigmoid(self.coef * occ)
        w = alpha * torch.cumprod(torch.cat([torch.ones(alpha.shape[0], 1),
                                             1. - alpha + 1e-10], -1), -1)[:, :-1]
        wsum = w.sum(-1, keepdim=True) + 1e-10
        depth = (w * z).sum(-1) / wsum.squeeze(-1)
        tmp = z - depth.unsqueeze(-1)
        var = (w * tmp * tmp).sum(1)
        depth = torch.where(nz, depth, torch.zeros_like(depth))
        rgb = torch.zeros(rays_o.shape[0], 3)
        if stage == 'color':
            cc, _ = self.feature_at(pts, r, rand_feat_color if rand_feat_color is not None
                                    else torch.zeros(32), color=True)
            rgb = (w[..., None] * self.col(pts, cc).reshape(-1, S, 3)).sum(-2) / wsum
        return dict(depth=depth, uncertainty=var, valid_ray_mask=valid_ray, z_vals=z, rgb=rgb,
                    stage=stage)

    def loss_dict(self, out, target_d, target_s, is_mapping, handle_dynamic=True):
        """get_loss_dict (conv_onet_pointslam.py:144-195) incl. the colour term."""
        td = target_d.squeeze()
        depth, unc = out['depth'], out['uncertainty']
        d = {}
        if not is_mapping:
            unc = unc.detach()
            nan_mask = (~torch.isnan(depth)) & (~torch.isnan(unc))
            tmp = torch.abs(td - depth) / torch.sqrt(unc + 1e-10) if handle_dynamic \
                else torch.abs(td - depth)
            mask = (tmp < 10 * tmp.median()) & (td > 0) & nan_mask
            d['geo_loss'] = torch.clamp(torch.abs(td - depth) / torch.sqrt(unc + 1e-10),
                                        min=0.0, max=1e3)[mask].sum()
            if self.use_color_trk:
                d['rgb_loss'] = self.w_color_trk * torch.abs(target_s - out['rgb'])[mask].sum()
            return d
        m = (td > 0) & out['valid_ray_mask'] & (~torch.isnan(depth))
        d['geo_loss'] = torch.abs(td[m] - depth[m]).sum()
        if out['stage'] == 'color':
            d['rgb_loss'] = self.w_color_map * torch.abs(target_s[m] - out['rgb'][m]).sum()
        return d

    def loss(self, out, target_d, is_mapping, handle_dynamic=True):
        td = target_d.squeeze()
        depth, unc = out['depth'], out['uncertainty']
        if not is_mapping:
            unc = unc.detach()
            nan_mask = (~torch.isnan(depth)) & (~torch.isnan(unc))
            tmp = torch.abs(td - depth) / torch.sqrt(unc + 1e-10) if handle_dynamic \
                else torch.abs(td - depth)
            mask = (tmp < 10 * tmp.median()) & (td > 0) & nan_mask
            return torch.clamp(torch.abs(td - depth) / torch.sqrt(unc + 1e-10),
                               min=0.0, max=1e3)[mask].sum()
        m = (td > 0) & out['valid_ray_mask'] & (~torch.isnan(depth))
        return torch.abs(td[m] - depth[m]).sum()
