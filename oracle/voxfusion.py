"""CPU restatement of the Vox-Fusion render-and-optimise step.

ORACLE / TEST INFRASTRUCTURE -- never imported by xrdslam_b200/.

Follows (reference @ f0366f20, paths relative to /root/reference):
  third_party/sparse_voxels/src/intersect_gpu.cu:75-140,191-270   RayAABBIntersection and
      svo_intersect_point_kernel -> ``svo_intersect`` (python loops, float32 scalars;
      the kernel's __fdividef(1, d) is a true division here: t values agree to a few ulp)
  slam/model_components/voxel_helpers_voxfusion.py:647-687        ray_intersect
  third_party/sparse_voxels/src/sample_gpu.cu:133-239             inverse_cdf_sampling_kernel
  slam/model_components/voxel_helpers_voxfusion.py:399-481,690-714  its [G=200, K] batching
      (rays padded with copies of ray 0) and ray_sample
  slam/model_components/voxel_helpers_voxfusion.py:97-166         get_features / trilinear_interp
  slam/model_components/decoder_voxfusion.py:76-149               Decoder
  slam/models/sparse_voxel.py:152-304, :103-143                    render_rays, sdf2weights, losses
Pinned: octree + map states against the reference's own svo.Octree (oracle/_ref/svo.so, built
from the reference sources by oracle/build_ref.py); on the GPU box the two kernels are
additionally checked bit-for-bit against the reference's own `grid` CUDA extension
(oracle/_ref/grid.so).  The torch part (features, decoder, weights, losses, gradients) is
pinned against the reference's own SparseVoxel.render_rays + get_loss_dict run on CPU with
these intersections / samples injected (tests/test_voxfusion_cpu.py, oracle/ref_harness.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

MAX_DEPTH = 10.0
f32 = np.float32


# ------------------------------------------------------------------ kernels ---
def _ray_aabb(o, d, c, half):
    f_low, f_high = f32(0), f32(100000.)
    for k in range(3):
        inv = f32(1.0) / d[k]
        lo = f32(f32(f32(c[k] - half) - o[k]) * inv)
        hi = f32(f32(f32(c[k] + half) - o[k]) * inv)
        if hi < lo:
            lo, hi = hi, lo
        if hi < f_low or lo > f_high:
            return f32(-1), f32(-1)
        f_low = lo if lo > f_low else f_low
        f_high = hi if hi < f_high else f_high
        if f_low > f_high:
            return f32(-1), f32(-1)
    return f_low, f_high


def svo_intersect(rays_o, rays_d, centres, children, voxel_size, n_max):
    """-> idx [R,n_max] int32 (-1 padded), tmin, tmax [R,n_max] f32, DFS visiting order."""
    ro, rd = rays_o.numpy().astype(f32), rays_d.numpy().astype(f32)
    cen, ch = centres.numpy().astype(f32), children.numpy().astype(np.int32)
    R = ro.shape[0]
    idx = -np.ones((R, n_max), np.int32)
    tmin = np.zeros((R, n_max), f32)
    tmax = np.zeros((R, n_max), f32)
    half_voxel = f32(f32(voxel_size) * f32(0.5))
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        for j in range(R):
            stack = [0]
            cnt = 0
            while stack and cnt < n_max:
                k = stack.pop()
                lo, hi = _ray_aabb(ro[j], rd[j], cen[k], f32(half_voxel * f32(ch[k, 8])))
                if lo > f32(-1.0):
                    if ch[k, 8] == 1:
                        idx[j, cnt], tmin[j, cnt], tmax[j, cnt] = k, lo, hi
                        cnt += 1
                        continue
                    for u in range(8):
                        if ch[k, u] > -1:
                            stack.append(int(ch[k, u]))
    return torch.from_numpy(idx), torch.from_numpy(tmin), torch.from_numpy(tmax)


def ray_intersect(rays_o, rays_d, centres, children, voxel_size, max_distance=10.0,
                  intersect_fn=svo_intersect):
    pts_idx, min_depth, max_depth = intersect_fn(rays_o, rays_d, centres, children, voxel_size, 50)
    min_depth = min_depth.clone().masked_fill_(pts_idx.eq(-1), max_distance)
    max_depth = max_depth.clone().masked_fill_(pts_idx.eq(-1), max_distance)
    min_depth, sorted_idx = min_depth.sort(dim=-1, stable=True)
    max_depth = max_depth.gather(-1, sorted_idx)
    pts_idx = pts_idx.gather(-1, sorted_idx.long()).clone()
    pts_idx[min_depth > max_distance] = -1
    min_depth.masked_fill_(pts_idx.eq(-1), max_distance)
    max_depth.masked_fill_(pts_idx.eq(-1), max_distance)
    max_hits = int(torch.max(pts_idx.ne(-1).sum(-1)))
    inter = {'min_depth': min_depth[..., :max_hits], 'max_depth': max_depth[..., :max_hits],
             'intersected_voxel_idx': pts_idx[..., :max_hits]}
    return inter, pts_idx.ne(-1).any(-1)


def _inverse_cdf_block(pts_idx, min_depth, max_depth, noise, probs, steps):
    """One block of the kernel: arrays [num_rays, max_hits] / [num_rays, max_steps]."""
    num_rays, max_hits = pts_idx.shape
    max_steps = noise.shape[1]
    s_idx = -np.ones((num_rays, max_steps), np.int32)
    s_depth = np.zeros((num_rays, max_steps), f32)
    s_dist = np.zeros((num_rays, max_steps), f32)
    flat_idx = pts_idx.reshape(-1)
    for j in range(num_rays):
        H = j * max_hits
        curr_bin, s = 0, 0
        cmin, cmax = min_depth[j, 0], max_depth[j, 0]
        cmin_cdf, cmax_cdf = f32(0), probs[j, 0]
        step_size = f32(1.0 / float(steps[j]))
        z_low = cmin
        total = int(np.ceil(steps[j]))
        done = False
        for cs in range(total):
            curr_cdf = f32(f32(f32(cs) + noise[j, cs]) * step_size)
            while curr_cdf > cmax_cdf:
                s_idx[j, s] = pts_idx[j, curr_bin]
                s_dist[j, s] = f32(cmax - z_low)
                s_depth[j, s] = f32(float(f32(cmax + z_low)) * .5)
                curr_bin += 1
                s += 1
                if curr_bin >= max_hits or pts_idx[j, curr_bin] == -1:
                    done = True
                    break
                cmin, cmax = min_depth[j, curr_bin], max_depth[j, curr_bin]
                cmin_cdf = cmax_cdf
                cmax_cdf = f32(cmax_cdf + probs[j, curr_bin])
                z_low = cmin
            if done:
                break
            u = f32(f32(curr_cdf - cmin_cdf) / f32(cmax_cdf - cmin_cdf))
            # nvcc contracts  cmin + u * (cmax - cmin)  into one fma
            z = f32(float(cmin) + float(u) * float(f32(cmax - cmin)))
            s_idx[j, s] = pts_idx[j, curr_bin]
            s_dist[j, s] = f32(z - z_low)
            s_depth[j, s] = f32(float(f32(z + z_low)) * .5)
            z_low = z
            s += 1
        # (curr_bin == max_hits happens when a ray consumed every bin above: the kernel then
        #  reads one element past its row -- undefined; the restatement stops)
        while (z_low < cmax) and (num_rays > (H + curr_bin)) and curr_bin < max_hits:
            s_idx[j, s] = pts_idx[j, curr_bin]
            s_dist[j, s] = f32(cmax - z_low)
            s_depth[j, s] = f32(float(f32(cmax + z_low)) * .5)
            curr_bin += 1
            s += 1
            if curr_bin >= max_hits or flat_idx[curr_bin] == -1:
                break
            cmin, cmax = min_depth[j, curr_bin], max_depth[j, curr_bin]
            z_low = cmin
    return s_idx, s_depth, s_dist


def inverse_cdf_sampling(pts_idx, min_depth, max_depth, probs, steps, noise_fn):
    """voxel_helpers_voxfusion.py:399-481: G = 200 blocks, rows padded with copies of ray 0.
    noise_fn(shape) supplies the uniform noise of the [G, K, max_steps] tensor (already
    clamped to [0.001, 0.999])."""
    G, N, P = 200, pts_idx.size(0), pts_idx.size(1)
    H = int(np.ceil(N / G)) * G
    if H > N:
        pts_idx = torch.cat([pts_idx, pts_idx[:1].expand(H - N, P)], 0)
        min_depth = torch.cat([min_depth, min_depth[:1].expand(H - N, P)], 0)
        max_depth = torch.cat([max_depth, max_depth[:1].expand(H - N, P)], 0)
        probs = torch.cat([probs, probs[:1].expand(H - N, P)], 0)
        steps = torch.cat([steps, steps[:1].expand(H - N)], 0)
    K = H // G
    max_steps = int(steps.ceil().long().max()) + P
    noise = noise_fn((G, K, max_steps))
    pi = pts_idx.reshape(G, K, P).numpy().astype(np.int32)
    mn = min_depth.reshape(G, K, P).numpy().astype(f32)
    mx = max_depth.reshape(G, K, P).numpy().astype(f32)
    pr = probs.reshape(G, K, P).numpy().astype(f32)
    st = steps.reshape(G, K).numpy().astype(f32)
    nz = noise.numpy().astype(f32)
    outs = [_inverse_cdf_block(pi[g], mn[g], mx[g], nz[g], pr[g], st[g]) for g in range(G)]
    s_idx = torch.from_numpy(np.stack([o[0] for o in outs])).reshape(H, -1)[:N]
    s_depth = torch.from_numpy(np.stack([o[1] for o in outs])).reshape(H, -1)[:N]
    s_dist = torch.from_numpy(np.stack([o[2] for o in outs])).reshape(H, -1)[:N]
    max_len = int(s_idx.ne(-1).sum(-1).max())
    return s_idx[:, :max_len], s_depth[:, :max_len], s_dist[:, :max_len], K


def ray_sample(inter, step_size, noise_fn):
    dists = (inter['max_depth'] - inter['min_depth']).masked_fill(
        inter['intersected_voxel_idx'].eq(-1), 0)
    probs = dists / dists.sum(dim=-1, keepdim=True)
    steps = dists.sum(-1) / step_size
    s_idx, s_depth, s_dist, K = inverse_cdf_sampling(inter['intersected_voxel_idx'],
                                                     inter['min_depth'], inter['max_depth'],
                                                     probs, steps, noise_fn)
    s_dist = s_dist.clamp(min=0.0)
    s_depth = s_depth.masked_fill(s_idx.eq(-1), MAX_DEPTH)
    s_dist = s_dist.masked_fill(s_idx.eq(-1), 0.0)
    return {'sampled_point_depth': s_depth, 'sampled_point_distance': s_dist,
            'sampled_point_voxel_idx': s_idx, 'probs': probs, 'steps': steps, 'K': K}


# ------------------------------------------------------------------- model ---
class VoxDecoder(nn.Module):
    """decoder_voxfusion.py:76-149 with depth=2, width=128, embedder='none'."""
    def __init__(self, width=128, in_dim=16, sdf_dim=128):
        super().__init__()
        self.pts_linears = nn.ModuleList([nn.Linear(in_dim, width), nn.Linear(width, width)])
        self.sdf_out = nn.Linear(width, 1 + sdf_dim)
        self.color_out = nn.Sequential(nn.Linear(sdf_dim + in_dim, width), nn.ReLU(),
                                       nn.Linear(width, 3), nn.Sigmoid())

    def forward(self, x):
        h = x
        for lin in self.pts_linears:
            h = F.relu(lin(h))
        so = self.sdf_out(h)
        sdf, feat = so[:, :1], so[:, 1:]
        rgb = self.color_out(torch.cat([feat, x], dim=-1))
        return rgb, sdf[:, 0]


def offset_points_q():
    c = torch.arange(1, 4, 2)
    ox, oy, oz = torch.meshgrid([c, c, c], indexing='ij')
    off = (torch.cat([ox.reshape(-1, 1), oy.reshape(-1, 1), oz.reshape(-1, 1)], 1).float() - 2) / 1.0
    return off * 0.5 + 0.5  # [8,3] in {0,1}


class VoxOracle(nn.Module):
    def __init__(self, voxel_size=0.2, step_size=0.01, trunc=0.05, max_depth=10.0,
                 w=(0.5, 1.0, 5000.0, 10.0), num_embeddings=20000, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.voxel_size, self.step_size, self.trunc, self.max_depth, self.w = \
            voxel_size, step_size, trunc, max_depth, w
        self.embeddings = nn.Parameter(torch.randn(num_embeddings, 16, generator=g) * 0.01)
        self.decoder = VoxDecoder()

    def set_map(self, voxels, children, features):
        """update_map_states (sparse_voxel.py:334-351) from get_centres_and_children output."""
        self.centres = ((voxels[:, :3] + voxels[:, -1:] / 2) * self.voxel_size).float()
        self.children = torch.cat([children, voxels[:, -1:]], -1).int()
        self.vertex_idx = features

    def march(self, rays_o, rays_d, noise_fn):
        inter, hits = ray_intersect(rays_o, rays_d, self.centres, self.children, self.voxel_size)
        if hits.sum() == 0:
            return None
        inter = {k: v[hits] for k, v in inter.items()}
        samples = ray_sample(inter, self.step_size, noise_fn)
        return inter, hits, samples

    def render(self, rays_o, rays_d, target_s, target_d, marched):
        inter, ray_mask, samples = marched
        ro, rd = rays_o[ray_mask], rays_d[ray_mask]
        depth_s = samples['sampled_point_depth']
        idx_s = samples['sampled_point_voxel_idx'].long()
        sample_mask = idx_s.ne(-1)
        xyz = ro.unsqueeze(1) + rd.unsqueeze(1) * depth_s.unsqueeze(2)
        xyz_v, idx_v = xyz[sample_mask], idx_s[sample_mask]
        point_xyz = F.embedding(idx_v, self.centres)
        feats = F.embedding(F.embedding(idx_v, self.vertex_idx).long(),
                            self.embeddings).view(point_xyz.size(0), -1)
        p = ((xyz_v - point_xyz) / self.voxel_size + 0.5).unsqueeze(1)
        q = offset_points_q().unsqueeze(0)
        wts = (p * q + (1 - p) * (1 - q)).prod(dim=-1, keepdim=True)
        emb = (wts * feats.view(feats.size(0), 8, -1)).sum(1).float()
        rgb_v, sdf_v = self.decoder(emb)
        B, Kk = sample_mask.size()
        sdf = sdf_v.new_ones(B, Kk).masked_scatter(sample_mask, sdf_v)
        colour = rgb_v.new_zeros(B, Kk, 3).masked_scatter(
            sample_mask.unsqueeze(-1).expand(B, Kk, 3), rgb_v)
        valid = sample_mask.float()
        z_vals = depth_s
        tr = self.trunc
        weights = torch.sigmoid(sdf / tr) * torch.sigmoid(-sdf / tr)
        signs = sdf[:, 1:] * sdf[:, :-1]
        mask = torch.where(signs < 0.0, torch.ones_like(signs), torch.zeros_like(signs))
        inds = torch.argmax(mask, dim=1)[..., None]
        z_min = torch.gather(z_vals, 1, inds)
        mask = torch.where(z_vals < z_min + tr, torch.ones_like(z_vals), torch.zeros_like(z_vals))
        weights = weights * mask * valid
        weights = weights / (torch.sum(weights, dim=-1, keepdim=True) + 1e-8)
        rgb = torch.sum(weights[..., None] * colour, dim=-2)
        depth = torch.sum(weights * z_vals, dim=-1)
        R = ray_mask.shape[0]
        depth_full = depth.new_zeros(R).masked_scatter(ray_mask, depth)
        rgb_full = rgb.new_zeros(R, 3).masked_scatter(ray_mask.unsqueeze(-1).expand(R, 3), rgb)
        out = dict(depth=depth_full, rgb=rgb_full, sdf=sdf, z_vals=z_vals, ray_mask=ray_mask,
                   weights=weights)
        if target_s is None:
            return out, None
        # get_loss_dict (sparse_voxel.py:103-143)
        td = target_d[ray_mask]
        tc = target_s[ray_mask]
        vmask = (td.squeeze() > 0.01) * (td.squeeze() < self.max_depth)
        wv = vmask.clone().unsqueeze(-1)
        rgb_loss = F.l1_loss(rgb_full[ray_mask] * wv, tc * wv)
        depth_loss = F.l1_loss(depth_full[ray_mask].squeeze()[vmask], td.squeeze()[vmask])
        front = torch.where(z_vals < (td - tr), torch.ones_like(z_vals), torch.zeros_like(z_vals))
        back = torch.where(z_vals > (td + tr), torch.ones_like(z_vals), torch.zeros_like(z_vals))
        dm = torch.where(td > 0.0, torch.ones_like(td), torch.zeros_like(td))
        sdf_mask = (1.0 - front) * (1.0 - back) * dm
        n_fs, n_sdf = torch.count_nonzero(front), torch.count_nonzero(sdf_mask)
        n = n_sdf + n_fs
        fs_w, sdf_w = 1.0 - n_fs / n, 1.0 - n_sdf / n
        fs_loss = F.mse_loss(sdf * front, torch.ones_like(sdf) * front) * fs_w
        sdf_loss = F.mse_loss((z_vals + sdf * tr) * sdf_mask, td * sdf_mask) * sdf_w
        ld = {'rgb_loss': rgb_loss * self.w[0], 'depth_loss': depth_loss * self.w[1],
              'sdf_loss': sdf_loss * self.w[2], 'fs_loss': fs_loss * self.w[3]}
        return out, ld
