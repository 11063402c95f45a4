"""Pixel sampling and ray generation (mirror of slam/common/common.py:39-122,
188-227,288-310 and slam/utils/utils.py:28-65).  Frames stay resident on the
device once uploaded (the reference re-uploads the full image every call,
common.py:67-68 -- SURVEY row f1)."""
import numpy as np
import torch


def _as_dev(x, device, dtype=torch.float32):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(device=device, dtype=dtype)


def get_rays_from_uv(i, j, c2w, fx, fy, cx, cy, device):
    """rays for pixel coords (i: column, j: row); differentiable w.r.t. c2w."""
    c2w = _as_dev(c2w, device)
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)],
                       -1).to(device).reshape(-1, 1, 3)
    rays_d = torch.sum(dirs * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def get_sample_uv(H0, H1, W0, W1, n, depth, color, device='cuda:0',
                  indices=None):
    """n pixels (with replacement, Q7) from rows H0..H1, cols W0..W1.
    ``indices`` (optional, [n] int64 into the cropped region, row-major) makes
    the draw an explicit input for parity runs."""
    depth = _as_dev(depth, device)[H0:H1, W0:W1]
    color = _as_dev(color, device)[H0:H1, W0:W1]
    w = W1 - W0
    if indices is None:
        indices = torch.randint((H1 - H0) * w, (n, ), device=device)
    indices = indices.to(device)
    i = (indices % w + W0).to(torch.float32)
    j = (torch.div(indices, w, rounding_mode='floor') + H0).to(torch.float32)
    d = depth.reshape(-1, 1)[indices]
    c = color.reshape(-1, 3)[indices]
    return i, j, d, c


def get_samples(camera, n, c2w, depth, color, device, Hedge=0, Wedge=0,
                depth_filter=False, return_index=False, depth_limit=None,
                indices=None):
    i, j, sample_depth, sample_color = get_sample_uv(
        Hedge, camera.height - Hedge, Wedge, camera.width - Wedge, n, depth,
        color, device=device, indices=indices)
    rays_o, rays_d = get_rays_from_uv(i, j, c2w, camera.fx, camera.fy,
                                      camera.cx, camera.cy, device)
    if depth_filter:
        sample_depth = sample_depth.reshape(-1)
        mask = sample_depth > 0
        if depth_limit is not None:
            mask = mask & (sample_depth < depth_limit)
        rays_o, rays_d = rays_o[mask], rays_d[mask]
        sample_depth, sample_color = sample_depth[mask], sample_color[mask]
        i, j = i[mask], j[mask]
    if return_index:
        return (rays_o, rays_d, sample_depth, sample_color, i.to(torch.int64),
                j.to(torch.int64))
    return rays_o, rays_d, sample_depth, sample_color


def get_rays(camera, c2w, device):
    """All H*W rays of a frame (render_img)."""
    c2w = _as_dev(c2w, device)
    i, j = torch.meshgrid(
        torch.linspace(0, camera.width - 1, camera.width, device=device),
        torch.linspace(0, camera.height - 1, camera.height, device=device),
        indexing='xy')
    dirs = torch.stack([(i - camera.cx) / camera.fx,
                        -(j - camera.cy) / camera.fy, -torch.ones_like(i)], -1)
    dirs = dirs.reshape(camera.height, camera.width, 1, 3)
    rays_d = torch.sum(dirs * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def get_camera_rays(H, W, fx, fy=None, cx=None, cy=None, type='OpenGL'):
    """Camera-frame ray directions [H,W,3] (slam/utils/utils.py:28-65)."""
    i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32),
                          torch.arange(H, dtype=torch.float32), indexing='xy')
    if cx is None:
        cx, cy = 0.5 * W, 0.5 * H
    if fy is None:
        fy = fx
    if type == 'OpenGL':
        return torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)
    if type == 'OpenCV':
        return torch.stack([(i - cx) / fx, (j - cy) / fy, torch.ones_like(i)], -1)
    raise NotImplementedError(type)
