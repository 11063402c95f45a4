"""render_img parity: the full-frame inference path of every algorithm
(slam/algorithms/coslam.py:245-289, nice_slam.py:234, voxfusion.py:125, point_slam.py:274)
on a small (80x60) frame against the oracle's forward pass on the same rays, and the K-iteration
Co-SLAM loss trajectory against the oracle's own optimisation from identical parameters
(north_star: "at matched loss")."""
import random
import warnings

import numpy as np
import pytest
import torch

from helpers import BOUND, max_abs

pytestmark = pytest.mark.gpu
W, H = 80, 60  # small frames: the CPU oracles render them too


def _host_rays(cam, c2w):
    from xrdslam_b200.common import get_rays  # host mirror, pinned to the reference function
    ro, rd = get_rays(cam, torch.as_tensor(c2w), device='cpu')
    return ro.reshape(-1, 3), rd.reshape(-1, 3)


def test_coslam_render_img_vs_oracle(cuda_dev):
    from oracle.coslam import CoslamOracle
    from test_bench_shapes_gpu import philox_uniform
    from xrdslam_b200.coslam import CoSLAMConfig
    from xrdslam_b200.synthetic import make_sequence
    cam, poses, fr = make_sequence(1, width=W, height=H)
    algo = CoSLAMConfig().setup(camera=cam, device=cuda_dev)
    ora = CoslamOracle(BOUND)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        ora.embed_fn.params.copy_((torch.rand(ora.embed_fn.params.shape, generator=g) * 2 - 1) * 0.3)
        for lin in (ora.sdf0, ora.sdf1, ora.col0, ora.col1):
            lin.weight.copy_(torch.randn(lin.weight.shape, generator=g) / np.sqrt(lin.weight.shape[1]))
        m = algo.model
        m.embed_fn.params.copy_(ora.embed_fn.params)
        for w, lin in zip(m._weights(), (ora.sdf0, ora.sdf1, ora.col0, ora.col1)):
            w.copy_(lin.weight)
    rgb, depth = fr[0]
    color_g, depth_g = algo.render_img(poses[0], gt_depth=depth)
    assert color_g.shape == (H, W, 3) and depth_g.shape == (H, W)
    R = W * H
    assert R <= algo.config.ray_batch_size  # one chunk -> one Philox seed
    noise = torch.from_numpy(philox_uniform(algo.model._last_seed, R * 43).reshape(R, 43))
    ro, rd = _host_rays(cam, poses[0])
    with torch.no_grad():
        out = ora.render_rays(ro, rd, torch.from_numpy(depth).reshape(-1, 1), noise)
    assert max_abs(color_g.reshape(-1, 3), out['rgb']) < 2e-5
    assert max_abs(depth_g.reshape(-1), out['depth']) < 2e-5


def test_nice_render_img_vs_oracle(cuda_dev):
    from test_nice_gpu import nice_pair
    from xrdslam_b200.nice_slam import NiceSLAMConfig
    from xrdslam_b200.synthetic import make_sequence
    cam, poses, fr = make_sequence(1, width=W, height=H)
    ora, model = nice_pair(cuda_dev)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        algo = NiceSLAMConfig(coarse=False, mapping_bound=[[-2.0, 2.0]] * 3).setup(camera=cam,
                                                                                  device=cuda_dev)
    model.camera = cam
    algo.model = model  # the pair's parameters (identical to the oracle's)
    # a camera inside the [-2,2]^3 test bound looking at its content
    c2w = poses[0].copy()
    c2w[:3, 3] = [0.3, -0.2, 0.1]
    depth = np.clip(fr[0][1], 0, 1.6).astype(np.float32)
    color_g, depth_g = algo.render_img(c2w, gt_depth=depth)
    ro, rd = _host_rays(cam, c2w)
    with torch.no_grad():
        out = ora.render(ro, rd, torch.from_numpy(depth).reshape(-1, 1), 'color')
    assert max_abs(color_g.reshape(-1, 3), out['rgb']) < 2e-4
    assert max_abs(depth_g.reshape(-1), out['depth']) < 2e-4


def test_voxfusion_render_img_vs_oracle(cuda_dev):
    """Chunked exactly like VoxFusion.render_img (ray_batch_size rays per model call: the
    inverse-CDF sampler's G = 200 batching quirk depends on the chunk), noise by hit rank."""
    from oracle.voxfusion import VoxOracle
    from xrdslam_b200.frame import Frame
    from xrdslam_b200.synthetic import make_sequence
    from xrdslam_b200.voxfusion import VoxFusionConfig
    w, h = 64, 50  # the CPU oracle marches in python loops; 3200 rays = 2 chunks of ray_batch_size
    cam, poses, fr = make_sequence(2, width=w, height=h, offset=(25.6,) * 3)
    algo = VoxFusionConfig().setup(camera=cam, device=cuda_dev)
    frames = [Frame(k, fr[k][0], fr[k][1], init_pose=poses[k]) for k in range(2)]
    for f in frames:
        algo.create_voxels(f)
    model = algo.model
    ora = VoxOracle()
    with torch.no_grad():
        g = torch.Generator().manual_seed(9)
        model.embeddings.copy_(torch.randn(model.embeddings.shape, generator=g) * 0.3)
        ora.embeddings.copy_(model.embeddings.cpu())
        ora.decoder.load_state_dict(model.decoder.state_dict())
    ora.set_map(*model.export_octree())
    R = w * h
    ms = model.config.max_samples_per_ray
    gen = torch.Generator().manual_seed(5)
    noise_rank = torch.rand(R, ms, generator=gen).clamp(0.001, 0.999)
    depth = fr[1][1]
    ro, rd = _host_rays(cam, poses[1])
    bs = algo.config.ray_batch_size
    assert R > bs  # more than one chunk
    rgb_o, dep_o = [], []
    noise = torch.full((R, ms), 0.5)
    for s in range(0, R, bs):
        o, d, nr = ro[s:s + bs], rd[s:s + bs], noise_rank[s:s + bs]

        def noise_fn(shape, nr=nr):
            G, K, st = shape
            out = torch.full((G * K, st), 0.5)
            n = min(G * K, nr.shape[0])
            out[:n] = nr[:n, :st]
            return out.reshape(G, K, st)
        marched = ora.march(o, d, noise_fn)
        assert marched is not None
        with torch.no_grad():
            out, _ = ora.render(o, d, None, None, marched)
        rgb_o.append(out['rgb']); dep_o.append(out['depth'])
        hits = marched[1]
        rank = torch.cumsum(hits.long(), 0) - 1
        chunk = noise[s:s + bs]
        chunk[hits] = nr[rank[hits]]
    color_g, depth_g = algo._render_full(poses[1], depth, per_pixel={'noise': noise.to(cuda_dev)})
    rgb_o, dep_o = torch.cat(rgb_o), torch.cat(dep_o)
    # sample->voxel assignment differs on a few 1e-3 of the samples (true division in the CPU
    # oracle vs __fdividef in the kernels, see test_voxfusion_gpu.py): robust + bulk checks
    dc = (torch.from_numpy(color_g.reshape(-1, 3)) - rgb_o).abs().max(1)[0]
    dd = (torch.from_numpy(depth_g.reshape(-1)).float() - dep_o).abs()
    assert float((dc > 2e-4).float().mean()) < 5e-3 and float((dd > 2e-4).float().mean()) < 5e-3
    assert float(dc.median()) < 2e-5 and float(dd.median()) < 2e-5


def test_pointslam_render_img_vs_oracle(cuda_dev):
    from helpers import pointslam_from_golden
    from xrdslam_b200.frame import Frame
    from xrdslam_b200.point_slam import PointSLAMConfig
    from xrdslam_b200.synthetic import make_sequence
    cam, poses, fr = make_sequence(2, width=W, height=H)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        # 80x60 = 4800 pixels: the colour-gradient pass takes the top 15 n of them
        algo = PointSLAMConfig(pixels_adding=2000, mapping_pixels_based_on_color_grad=200).setup(
            camera=cam, device=cuda_dev)
    frames = [Frame(k, fr[k][0], fr[k][1], init_pose=poses[k], separate_LR=True,
                    rot_rep='axis_angle') for k in range(2)]
    np.random.seed(0)
    torch.manual_seed(0)
    for f in frames:
        algo.pre_precessing(f, True)
    model = algo.model
    npc = model.neural_point_cloud
    with torch.no_grad():  # non-trivial decoders
        for p in model.decoder.parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.05)
    # an oracle with the model's decoders and cloud (same key mapping as the golden fixture)
    g = {'dec.' + k: v.detach().cpu().numpy() for k, v in model.decoder.geo_decoder.state_dict().items()}
    g.update({'cdec.' + k: v.detach().cpu().numpy()
              for k, v in model.decoder.color_decoder.state_dict().items()})
    g['cdec.embedder._B'] = model.decoder.color_decoder.embedder._B.detach().cpu().numpy()
    g.update(cloud_pos=npc.cloud_pos().cpu().numpy(), geo_feats=npc.geo_feats.detach().cpu().numpy(),
             col_feats=npc.col_feats.detach().cpu().numpy())
    ora = pointslam_from_golden(g, 'oracle')
    ora.frustum_mask = npc.frustum_mask.bool().cpu().reshape(-1, 1)  # set by pre_precessing
    gen = torch.Generator().manual_seed(3)
    rf, rfc = torch.randn(32, generator=gen) * 0.01, torch.randn(32, generator=gen) * 0.01
    depth = fr[1][1]
    R = W * H
    r_query = algo.dynamic_r_query_allkeyframe['1'].reshape(-1).float()
    color_g, depth_g = algo._render_full(
        poses[1], depth, extra={'stage': 'color', 'rand_feat': rf.to(cuda_dev),
                                'rand_feat_color': rfc.to(cuda_dev)},
        per_pixel={'batch_dynamic_r': r_query})
    ro, rd = _host_rays(cam, poses[1])
    td = torch.from_numpy(depth).reshape(-1, 1)
    bs = algo.config.ray_batch_size
    rgb_o, dep_o = [], []
    with torch.no_grad():
        for s in range(0, R, bs):
            out = ora.render(ro[s:s + bs], rd[s:s + bs], td[s:s + bs], r_query[s:s + bs].cpu(), rf,
                             'color', rfc)
            rgb_o.append(out['rgb']); dep_o.append(out['depth'])
    # the device builds the rays with its own rounding (1 ulp in rays_d): a neighbour sitting
    # exactly on a query radius can flip in or out -> a handful of rays change discretely;
    # everything else agrees to fp32 level
    dc = (torch.from_numpy(color_g.reshape(-1, 3)) - torch.cat(rgb_o)).abs().max(1)[0]
    dd = (torch.from_numpy(depth_g.reshape(-1)).float() - torch.cat(dep_o)).abs()
    print('point render_img: max', float(dc.max()), float(dd.max()), 'outliers',
          int((dc > 5e-5).sum()), int((dd > 5e-5).sum()), 'of', dc.numel())
    assert float((dc > 5e-5).float().mean()) < 5e-3 and float((dd > 5e-5).float().mean()) < 5e-3
    assert float(dc.median()) < 5e-6 and float(dd.median()) < 5e-6


def test_coslam_loss_trajectory_matches_oracle(cuda_dev):
    """6 captured mapping iterations (precision mode 0, smoothness on, Adam on table + decoder)
    against the oracle optimised with torch.optim.Adam on the same batches, Philox noise and
    smoothness offsets: the loss TRAJECTORY agrees to 1e-3 relative."""
    import bench
    from oracle.coslam import CoslamOracle
    from test_bench_shapes_gpu import philox_uniform
    from xrdslam_b200.common import rays_from_poses
    from xrdslam_b200.opt_pose import OptimizablePose, pose_matrices
    random.seed(31)
    algo, kfs, cur = bench.build_algorithm(cuda_dev, seed=31)
    algo.model.config.precision = 0
    frames = kfs + [cur]
    algo.config.mapping_sample = 512
    algo.config.min_sample_pixels = 512
    algo.bundle_adjust = False
    model = algo.model
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        model.embed_fn.params.copy_(((torch.rand(model.embed_fn.params.shape, generator=g) * 2 - 1)
                                     * 0.05).to(cuda_dev))
    ora = CoslamOracle(BOUND)
    with torch.no_grad():
        ora.embed_fn.params.copy_(model.embed_fn.params.cpu())
        for lin, w in zip((ora.sdf0, ora.sdf1, ora.col0, ora.col1), model._weights()):
            lin.weight.copy_(w.detach().cpu())
    opt = torch.optim.Adam([
        {'params': [ora.embed_fn.params], 'lr': 1e-2, 'eps': 1e-15, 'betas': (0.9, 0.99)},
        {'params': [ora.sdf0.weight, ora.sdf1.weight, ora.col0.weight, ora.col1.weight],
         'lr': 1e-2, 'weight_decay': 1e-6, 'betas': (0.9, 0.99)}])
    algo.setup_optimizers(6, frames, True)
    sess = algo.mapping_session(frames)
    assert not sess.ba
    sess.begin(frames)
    poses = [OptimizablePose(torch.cat([sess.trans[i].cpu(), sess.rot[i].cpu()]).clone(),
                             separate_LR=True, rot_rep='axis_angle') for i in range(len(frames))]
    with torch.no_grad():
        c2w = pose_matrices(poses)
    R, S = sess.R, 43
    l_gpu, l_cpu = [], []
    for it in range(6):
        l_gpu.append(float(sess.step(it, frames)))
        rows, ids = sess.rows.cpu(), sess.ids.cpu()
        dyn = sess.dyn.cpu().numpy()
        seed = int(dyn[0:8].view(np.uint64)[0])
        smooth = torch.from_numpy(dyn[8:32].view(np.float32).copy()).reshape(2, 3)
        noise = torch.from_numpy(philox_uniform(seed, R * S).reshape(R, S))
        ro, rd = rays_from_poses(rows[:, :3], ids, c2w)
        opt.zero_grad(set_to_none=True)
        _, _, tot = ora.step(ro, rd, rows[:, 3:6], rows[:, 6:7], noise, True, False,
                             smooth_rand=smooth)
        tot.backward()
        opt.step()
        l_cpu.append(float(tot.detach()))
    sess.end(frames)
    assert l_gpu[-1] < l_gpu[0]
    for a, b in zip(l_gpu, l_cpu):
        assert abs(a - b) <= 1e-3 * abs(b), (l_gpu, l_cpu)
