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


class _PosedRays(torch.autograd.Function):
    """csrc/rays.cu: rays from camera-frame directions + per-ray pose row, and the pose
    gradient reduction (one launch each; replaces torch's gather / index_put backward)."""
    @staticmethod
    def forward(ctx, dirs, ids, poses):
        import ctypes as C
        from . import _cabi
        dirs = dirs.detach().to(torch.float32).contiguous()
        P = poses.detach().to(torch.float32).contiguous()
        R = dirs.shape[0]
        rays_o = torch.empty(R, 3, device=dirs.device)
        rays_d = torch.empty(R, 3, device=dirs.device)
        with torch.cuda.device(dirs.device):
            st = _cabi.lib().xrd_rays_from_poses(
                R, _cabi.ptr(dirs), _cabi.ptr(ids), _cabi.ptr(P), P.shape[0], _cabi.ptr(rays_o),
                _cabi.ptr(rays_d), torch.cuda.current_stream(dirs.device).cuda_stream)
        _cabi.check('xrd_rays_from_poses', st)
        ctx.save_for_backward(dirs, ids)
        ctx.n_poses = P.shape[0]
        return rays_o, rays_d

    @staticmethod
    def backward(ctx, g_o, g_d):
        from . import _cabi
        dirs, ids = ctx.saved_tensors
        d_poses = torch.empty(ctx.n_poses, 4, 4, device=dirs.device)
        g_o, g_d = g_o.contiguous(), g_d.contiguous()
        with torch.cuda.device(dirs.device):
            st = _cabi.lib().xrd_rays_pose_grads(
                dirs.shape[0], _cabi.ptr(dirs), _cabi.ptr(ids), ctx.n_poses, _cabi.ptr(g_o),
                _cabi.ptr(g_d), _cabi.ptr(d_poses),
                torch.cuda.current_stream(dirs.device).cuda_stream)
        _cabi.check('xrd_rays_pose_grads', st)
        return None, None, d_poses


def rays_from_poses(dirs_cam, pose_ids, poses):
    """rays_o, rays_d [R,3] from camera-frame directions [R,3], per-ray pose rows `pose_ids`
    ([R] int64, negative = from the end, or None: pose 0) and c2w `poses` [n,4,4];
    differentiable w.r.t. poses (coslam.py:208-216 / common.py:39-53)."""
    if dirs_cam.is_cuda:
        if pose_ids is not None:
            pose_ids = pose_ids.to(dirs_cam.device, torch.int64).contiguous()
        return _PosedRays.apply(dirs_cam, pose_ids, poses.to(dirs_cam.device))
    # host tensors (dataset preparation, CPU tests of the host logic): plain torch
    ids = pose_ids if pose_ids is not None else torch.zeros(dirs_cam.shape[0], dtype=torch.int64)
    rays_d = torch.sum(dirs_cam[:, None, :] * poses[ids, :3, :3], -1)
    return poses[ids, :3, -1], rays_d


def sample_window(camera, depth_imgs, rgb_imgs, poses, n, Hedge=0, Wedge=0, indices=None,
                  return_index=False):
    """get_samples for every frame of a window in THREE launches (draw, gather, rays):
    depth_imgs / rgb_imgs: lists of device-resident [H,W] / [H,W,3] fp32 frames, poses:
    [F,4,4] c2w on the device (differentiable).  Frame-major outputs
    rays_o, rays_d [F*n,3], depth [F*n,1], colour [F*n,3] (+ i, j int64 column / row).
    Mirrors common.py:188-227 applied per frame and concatenated (nice_slam.py:141-171)."""
    import ctypes as C
    from . import _cabi
    dev = poses.device
    if dev.type != 'cuda':
        raise RuntimeError('xrdslam_b200 has no CPU path')
    F = len(depth_imgs)
    H, W = camera.height, camera.width
    H0, H1, W0, W1 = Hedge, H - Hedge, Wedge, W - Wedge
    if indices is None:
        indices = torch.randint((H1 - H0) * (W1 - W0), (F * n, ), device=dev)
    indices = indices.to(dev, torch.int64).contiguous()
    tot = F * n
    dirs = torch.empty(tot, 3, device=dev)
    depth = torch.empty(tot, 1, device=dev)
    rgb = torch.empty(tot, 3, device=dev)
    ids = torch.empty(tot, dtype=torch.int64, device=dev)
    ij = torch.empty(tot, 2, dtype=torch.int64, device=dev) if return_index else None
    cfg = _cabi.XrdPixelSampleCfg(F, n, H, W, H0, H1, W0, W1, camera.fx, camera.fy, camera.cx,
                                  camera.cy)
    dp = (C.c_void_p * F)(*[_cabi.ptr(t) for t in depth_imgs])
    cp = (C.c_void_p * F)(*[_cabi.ptr(t) for t in rgb_imgs])
    with torch.cuda.device(dev):
        st = _cabi.lib().xrd_sample_pixels(
            C.byref(cfg), dp, cp, _cabi.ptr(indices), _cabi.ptr(dirs), _cabi.ptr(depth),
            _cabi.ptr(rgb), _cabi.ptr(ids), _cabi.ptr(ij),
            torch.cuda.current_stream(dev).cuda_stream)
    _cabi.check('xrd_sample_pixels', st)
    rays_o, rays_d = _PosedRays.apply(dirs, ids, poses)
    if return_index:
        return rays_o, rays_d, depth, rgb, ij[:, 0], ij[:, 1]
    return rays_o, rays_d, depth, rgb


def get_rays_from_uv(i, j, c2w, fx, fy, cx, cy, device):
    """rays for pixel coords (i: column, j: row); differentiable w.r.t. c2w."""
    c2w = _as_dev(c2w, device)
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)],
                       -1).to(device).reshape(-1, 3)
    return rays_from_poses(dirs, None, c2w.reshape(1, 4, 4))


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


# ---- colour-gradient pixel selection (Point-SLAM, common.py:74-106, :230-285) ---------------
# skimage.color.rgb2gray / skimage.filters.sobel_h, sobel_v are restated on scipy.ndimage (what
# skimage itself calls): Y = 0.2125 R + 0.7154 G + 0.0721 B; sobel = 3x3 correlation-free
# convolution with [1,2,1]^T x [1,0,-1] / 4, boundary mode 'reflect'.  skimage is not in this
# image: "parity unpinned" at that boundary (float64 here); everything after the gradient image
# (argpartition, region mask, np.random.choice) is the reference's own arithmetic.
_HSOBEL = np.array([[1, 2, 1], [0, 0, 0], [-1, -2, -1]], dtype=np.float64) / 4.0


def rgb2gray_np(image):
    img = np.asarray(image, dtype=np.float64)
    return img[..., 0] * 0.2125 + img[..., 1] * 0.7154 + img[..., 2] * 0.0721


def sobel_magnitude_np(image):
    from scipy import ndimage
    gray = rgb2gray_np(image)
    grad_y = ndimage.convolve(gray, _HSOBEL, mode='reflect')       # sobel_h
    grad_x = ndimage.convolve(gray, _HSOBEL.T, mode='reflect')     # sobel_v
    return np.sqrt(grad_x**2 + grad_y**2)


def get_sample_uv_with_grad(H0, H1, W0, W1, n, image, ratio=15):
    """n flat pixel indices drawn (np.random.choice, no replacement) from the ratio*n pixels
    with the largest colour-gradient magnitude that fall inside rows H0..H1, cols W0..W1."""
    image = np.asarray(image)
    grad_mag = sobel_magnitude_np(image)
    img_size = (image.shape[0], image.shape[1])
    selected_index = np.argpartition(grad_mag, -ratio * n, axis=None)[-ratio * n:]
    indices_h, indices_w = np.unravel_index(selected_index, img_size)
    mask = (indices_h >= H0) & (indices_h < H1) & (indices_w >= W0) & (indices_w < W1)
    indices_h, indices_w = indices_h[mask], indices_w[mask]
    selected_index = np.ravel_multi_index(np.array((indices_h, indices_w)), img_size)
    samples = np.random.choice(range(0, indices_h.shape[0]), size=n, replace=False)
    return selected_index[samples]


def get_samples_with_pixel_grad(camera, n_color, c2w, depth, color, device, Hedge=0, Wedge=0,
                                depth_filter=True, return_index=True, depth_limit=None):
    """Rays through n_color pixels chosen by colour gradient (common.py:230-285); depth /
    colour are the frame's HOST arrays like in the reference."""
    H, W = camera.height, camera.width
    assert n_color > 0, 'invalid number of rays to sample.'
    color_np = np.asarray(color)
    index_color_grad = get_sample_uv_with_grad(Hedge, H - Hedge, Wedge, W - Wedge, n_color,
                                               color_np)
    merged = np.union1d(index_color_grad, [])
    jj, ii = np.unravel_index(merged.astype(int), (H, W))  # row, column
    i = torch.from_numpy(ii).to(device).float()
    j = torch.from_numpy(jj).to(device).float()
    rays_o, rays_d = get_rays_from_uv(i, j, c2w, camera.fx, camera.fy, camera.cx, camera.cy,
                                      device)
    i, j = i.long(), j.long()
    depth_t = _as_dev(depth, device)
    color_t = _as_dev(color, device)
    sample_depth = depth_t[j, i].reshape(-1)
    sample_color = color_t[j, i].reshape(-1, 3)
    if depth_filter:
        mask = sample_depth > 0
        if depth_limit is not None:
            mask = mask & (sample_depth < depth_limit)
        rays_o, rays_d = rays_o[mask], rays_d[mask]
        sample_depth, sample_color = sample_depth[mask], sample_color[mask]
        i, j = i[mask], j[mask]
    if return_index:
        return rays_o, rays_d, sample_depth, sample_color, i.to(torch.int64), j.to(torch.int64)
    return rays_o, rays_d, sample_depth, sample_color


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
