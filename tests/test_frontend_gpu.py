"""Device front-end (csrc/rays.cu) and fused Adam (csrc/adam.cu) against the torch
statements of the reference lines they replace (coslam.py:208-216, common.py:39-53;
torch.optim.Adam as called by slam/engine/optimizers.py:125-148)."""
import pytest
import torch


@pytest.mark.gpu
@pytest.mark.parametrize('n_poses', [1, 6, 100])
def test_rays_from_poses_matches_torch_gather(cuda_dev, n_poses):
    from xrdslam_b200.common import rays_from_poses
    g = torch.Generator().manual_seed(n_poses)
    R = 5000
    dirs = torch.randn(R, 3, generator=g)
    ids = torch.randint(-1, n_poses, (R,), generator=g)
    poses = torch.eye(4).repeat(n_poses, 1, 1)
    poses[:, :3, :] = torch.randn(n_poses, 3, 4, generator=g)
    w_o, w_d = torch.randn(R, 3, generator=g), torch.randn(R, 3, generator=g)
    # torch statement (float64 accumulate for the gradient reference)
    P = poses.clone().requires_grad_(True)
    rd = torch.sum(dirs[:, None, :] * P[ids, :3, :3], -1)
    ro = P[ids, :3, -1]
    ((rd * w_d).sum() + (ro * w_o).sum()).backward()
    Pd = poses.to(cuda_dev).requires_grad_(True)
    ro_k, rd_k = rays_from_poses(dirs.to(cuda_dev), ids.to(cuda_dev), Pd)
    assert torch.equal(ro_k.cpu(), ro.detach())
    assert (rd_k.cpu() - rd.detach()).abs().max() <= 1e-6 * rd.abs().max()
    ((rd_k * w_d.to(cuda_dev)).sum() + (ro_k * w_o.to(cuda_dev)).sum()).backward()
    assert (Pd.grad.cpu() - P.grad).abs().max() <= 2e-5 * P.grad.abs().max()
    assert Pd.grad[:, 3].abs().max() == 0


@pytest.mark.gpu
def test_get_rays_from_uv_single_pose(cuda_dev):
    from xrdslam_b200.common import get_rays_from_uv
    g = torch.Generator().manual_seed(0)
    i = torch.randint(0, 640, (300,), generator=g).float()
    j = torch.randint(0, 480, (300,), generator=g).float()
    c2w = torch.eye(4)
    c2w[:3, :] = torch.randn(3, 4, generator=g)
    c2w.requires_grad_(True)
    ro, rd = get_rays_from_uv(i.to(cuda_dev), j.to(cuda_dev), c2w, 320., 320., 319.5, 239.5,
                              cuda_dev)
    dirs = torch.stack([(i - 319.5) / 320., -(j - 239.5) / 320., -torch.ones_like(i)], -1)
    rd_t = torch.sum(dirs.reshape(-1, 1, 3) * c2w[:3, :3], -1)
    assert (rd.cpu() - rd_t).abs().max() < 1e-6
    assert torch.equal(ro.cpu(), c2w[:3, -1].expand(300, 3))
    (rd.sum() + 2 * ro.sum()).backward()
    g_k = c2w.grad.clone()
    c2w.grad = None
    (rd_t.sum() + 2 * c2w[:3, -1].expand(300, 3).sum()).backward()
    assert (g_k - c2w.grad).abs().max() <= 2e-5 * c2w.grad.abs().max()


@pytest.mark.gpu
@pytest.mark.parametrize('kw', [dict(lr=1e-2, eps=1e-15, betas=(0.9, 0.99)),
                                dict(lr=1e-2, weight_decay=1e-6, betas=(0.9, 0.99)),
                                dict(lr=1e-3)])
def test_fused_adam_matches_torch_adam(cuda_dev, kw):
    from xrdslam_b200.optimizers import AdamOptimizerConfig, FusedAdam
    g = torch.Generator().manual_seed(5)
    shapes = [(1640944,), (32, 80), (16, 32), (3, 7), (1,), (1025, 3)]
    ref = [torch.randn(s, generator=g).to(cuda_dev).requires_grad_(True) for s in shapes]
    mine = [t.detach().clone().requires_grad_(True) for t in ref]
    o_ref = torch.optim.Adam(ref, **kw)
    o_mine = AdamOptimizerConfig(**kw).setup(mine)
    assert isinstance(o_mine, FusedAdam)
    for it in range(12):
        for a, b in zip(ref, mine):
            gr = torch.randn(a.shape, generator=g).to(cuda_dev) * (0.0 if it == 3 else 1.0)
            if it == 5 and a.numel() > 1000:
                gr[::2] = 0  # sparse gradient rows (untouched hash entries)
            a.grad, b.grad = gr.clone(), gr.clone()
        o_ref.step()
        o_mine.step()
    for a, b in zip(ref, mine):
        assert (a - b).abs().max() <= 2e-6 * max(1.0, float(a.detach().abs().max()))
    sd = o_mine.state_dict()
    assert set(sd['state'][0].keys()) == {'step', 'exp_avg', 'exp_avg_sq'}
    torch.optim.Adam(mine, **kw).load_state_dict(sd)  # checkpoints interchange


@pytest.mark.gpu
def test_coslam_model_input_on_device(cuda_dev):
    """CoSLAM.get_model_input (pinned staging + device ray build) against the reference's
    statement evaluated in torch from the same sampled rows."""
    import bench
    algo, kfs, cur = bench.build_algorithm(cuda_dev, seed=3)
    frames = kfs + [cur]
    inp = algo.get_model_input(frames, True)
    rows, ids = algo._staging.cur[0]['rows'], algo._staging.cur[0]['ids']
    n = inp['rays_o'].shape[0]
    rows, ids = rows[:n].clone(), ids[:n].clone()
    assert (ids[-algo.config.min_sample_pixels:] == -1).all() or (ids == -1).sum() > 0
    poses = torch.stack([f.get_pose() for f in frames]).detach()
    rd = torch.sum(rows[:, None, :3] * poses[ids, :3, :3], -1)
    assert (inp['rays_d'].cpu() - rd).abs().max() < 1e-6
    assert torch.equal(inp['rays_o'].cpu(), poses[ids, :3, -1])
    assert torch.equal(inp['target_d'].cpu(), rows[:, 6:7])
    assert torch.equal(inp['target_s'].cpu(), rows[:, 3:6])
    # pose gradients reach the CPU pose parameters (bundle adjustment), frame 0 fixed
    (inp['rays_d'].sum() + inp['rays_o'].sum()).backward()
    g0 = frames[0].pose.data_t.grad  # frame 0 is fixed (coslam.py:181-182): no gradient
    assert frames[0].fid == 0 and (g0 is None or g0.abs().sum() == 0)
    assert cur.pose.data_t.grad is not None and cur.pose.data_r.grad.abs().sum() > 0


@pytest.mark.gpu
def test_sample_window_matches_per_frame_get_samples(cuda_dev):
    """xrd_sample_pixels + rays_from_poses for a 3-frame window == get_samples per frame with
    the same pixel indices (common.py:188-227 as called by nice_slam.py:141-171)."""
    import numpy as np
    from xrdslam_b200.common import get_samples, sample_window
    from xrdslam_b200.frame import Frame
    from xrdslam_b200.opt_pose import pose_matrices
    from xrdslam_b200.synthetic import make_sequence
    cam, poses, fr = make_sequence(3, width=160, height=120)
    frames = [Frame(k, fr[k][0], fr[k][1], init_pose=poses[k], rot_rep='quat') for k in range(3)]
    n, He, We = 257, 7, 11
    g = torch.Generator().manual_seed(0)
    idx = torch.randint((120 - 2 * He) * (160 - 2 * We), (3 * n,), generator=g)
    dimg = [torch.from_numpy(f.depth).to(cuda_dev) for f in frames]
    cimg = [torch.from_numpy(f.rgb).to(cuda_dev) for f in frames]
    P = pose_matrices([f.pose for f in frames])
    ro, rd, d, c, i, j = sample_window(cam, dimg, cimg, P.to(cuda_dev), n, He, We,
                                       indices=idx.to(cuda_dev), return_index=True)
    (rd.sum() + 3 * ro.sum()).backward()
    g_batched = [p.grad.clone() for f in frames for p in f.pose.parameters()]
    for f in frames:
        for p in f.pose.parameters():
            p.grad = None
    parts = [get_samples(cam, n, f.get_pose(), dimg[k], cimg[k], cuda_dev, Hedge=He, Wedge=We,
                         return_index=True, indices=idx[k * n:(k + 1) * n].to(cuda_dev))
             for k, f in enumerate(frames)]
    cat = lambda q: torch.cat([p[q] for p in parts])
    assert torch.equal(ro, cat(0)) and (rd - cat(1)).abs().max() < 1e-6
    assert torch.equal(d, cat(2)) and torch.equal(c, cat(3))
    assert torch.equal(i, cat(4)) and torch.equal(j, cat(5))
    (cat(1).sum() + 3 * cat(0).sum()).backward()
    g_frames = [p.grad for f in frames for p in f.pose.parameters()]
    for a, b in zip(g_batched, g_frames):
        assert (a - b).abs().max() <= 2e-5 * max(1.0, float(b.abs().max()))
