"""Keyframe selection by view overlap and the frustum tests used for map-parameter
selection (mirrors of slam/common/common.py:343-426 keyframe_selection_overlap,
slam/model_components/utils.py:298-375 and slam/algorithms/point_slam.py:368-424
get_mask_from_c2w).  These run once per mapping call, on the
device with torch ops; no per-iteration cost.

cv2.remap(INTER_LINEAR) is restated: bilinear interpolation with the sampling position
quantised to 1/32 pixel (OpenCV's INTER_TAB_SIZE) and zero outside the image
(BORDER_CONSTANT).  OpenCV is not in this image: "parity unpinned vs real cv2" -- the masks
are sets of voxels / points, and differences are confined to depth-discontinuity pixels."""
from __future__ import annotations

import numpy as np
import torch

from .common import get_samples


def _project(camera, c2w, pts):
    """uv [N,2] (float32) and z [N] of world points in the frame of c2w (the reference's
    numpy float64 chain: w2c @ [p,1]; x *= -1; K @ .; z += 1e-5)."""
    w2c = torch.linalg.inv(c2w.detach().double())
    p = pts.double()
    cam = p @ w2c[:3, :3].T + w2c[:3, 3]
    x, y, z = -cam[:, 0], cam[:, 1], cam[:, 2]
    zz = z + 1e-5
    u = (camera.fx * x + camera.cx * z) / zz
    v = (camera.fy * y + camera.cy * z) / zz
    return torch.stack([u, v], -1).float(), zz


def remap_linear(img, uv):
    """cv2.remap(img, u, v, INTER_LINEAR, BORDER_CONSTANT=0) for a float32 [H,W] image."""
    H, W = img.shape
    u, v = uv[:, 0].double(), uv[:, 1].double()
    # OpenCV rounds the coordinates to 1/32 pixel: saturate_cast<int>(x * 32)
    ui, vi = torch.round(u * 32).long(), torch.round(v * 32).long()
    x0, y0 = ui >> 5, vi >> 5
    ax, ay = (ui & 31).float() / 32, (vi & 31).float() / 32

    def at(y, x):
        ok = (x >= 0) & (x < W) & (y >= 0) & (y < H)
        val = img[y.clamp(0, H - 1), x.clamp(0, W - 1)]
        return torch.where(ok, val, torch.zeros_like(val))

    return ((1 - ay) * ((1 - ax) * at(y0, x0) + ax * at(y0, x0 + 1)) +
            ay * ((1 - ax) * at(y0 + 1, x0) + ax * at(y0 + 1, x0 + 1)))


def frustum_mask(camera, c2w, pts, depth, edge=0, near_cam=None):
    """bool [N]: points that project inside the image (margin `edge`) in front of the
    measured surface + 0.5 m; `near_cam` adds a ball around the camera centre (NICE)."""
    dev = pts.device
    depth = torch.as_tensor(np.asarray(depth, dtype=np.float32)).to(dev)
    c2w = torch.as_tensor(c2w).to(dev)
    uv, z = _project(camera, c2w, pts)
    d = remap_linear(depth, uv)
    H, W = depth.shape
    mask = camera.inside(uv[:, 0], uv[:, 1], edge)
    d = torch.where(d == 0, d.max(), d)  # rays with depth == 0 get the maximum depth
    mask = mask & (0 <= -z) & (-z <= d.double() + 0.5)
    if near_cam is not None:
        dist = pts - c2w[:3, 3].to(pts.dtype)
        mask = mask | ((dist * dist).sum(-1) < near_cam * near_cam)
    return mask


@torch.no_grad()
def keyframe_selection_overlap(camera, cur_frame, keyframes_graph, k, N_samples=16,
                               pixs_per_image=100, use_ray_sample=True, device='cuda:0'):
    if not use_ray_sample:
        raise NotImplementedError('use_ray_sample=False is the splaTAM path (out of scope)')
    rays_o, rays_d, gt_depth, _ = get_samples(camera, pixs_per_image, cur_frame.get_pose(),
                                              cur_frame.depth, cur_frame.rgb, device=device,
                                              depth_filter=True)
    gt_depth = gt_depth.reshape(-1, 1).repeat(1, N_samples)
    t = torch.linspace(0., 1., steps=N_samples).to(device)
    z_vals = gt_depth * 0.8 * (1. - t) + (gt_depth + 0.5) * t
    pts = (rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]).reshape(-1, 3)
    H, W, edge = camera.height, camera.width, 20
    scored = []
    for kf in keyframes_graph:
        uv, z = _project(camera, kf.get_pose().to(device), pts)
        m = camera.inside(uv[:, 0], uv[:, 1], edge) & (z < 0)
        scored.append((kf, float(m.sum()) / max(1, uv.shape[0])))
    scored.sort(key=lambda e: e[1], reverse=True)
    sel = [kf for kf, p in scored if p > 0.0]
    order = np.random.permutation(len(sel))[:k]
    return [sel[i] for i in order]
