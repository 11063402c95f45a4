"""Size-independent properties at the full configuration sizes (640x480 frames, the default
batch sizes of SURVEY section 8): sortedness, ranges, idempotence, additivity, determinism --
where an oracle comparison would take too long."""
import numpy as np
import pytest
import torch


def _scene(n_frames, **kw):
    from xrdslam_b200.frame import Frame
    from xrdslam_b200.synthetic import make_sequence
    cam, poses, fr = make_sequence(n_frames, **kw)
    return cam, poses, fr


@pytest.mark.gpu
def test_coslam_full_batch_properties(cuda_dev):
    """4096 rays x 43 samples, default hash grid: z sorted per ray and inside [near, far];
    weights normalised (acc <= 1); the forward is bit-deterministic for a fixed seed; the
    gradient of a batch with fixed normalisers is additive over its halves."""
    import bench
    algo, kfs, cur = bench.build_algorithm(cuda_dev, seed=2)
    model = algo.model
    frames = kfs + [cur]
    algo.config.graph_mapping = False
    algo.config.min_sample_pixels = 2048
    inp = algo.get_model_input(frames, True)
    inp = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in inp.items()}
    assert inp['rays_o'].shape[0] == 4096
    w = model._weights()
    tab = model.embed_fn.params

    def run(sl=slice(None), grads=False, seed=77):
        return model._launch(inp['rays_o'][sl], inp['rays_d'][sl], tab, *w, inp['target_s'][sl],
                             inp['target_d'][sl], None, with_grads=grads, seed=seed)
    o1, _ = run()
    o2, _ = run()
    z = o1['z_vals']
    assert torch.equal(z, o2['z_vals']) and torch.equal(o1['rgb'], o2['rgb'])
    assert torch.equal(o1['depth'], o2['depth'])
    assert (z[:, 1:] >= z[:, :-1]).all()
    cfg = model.config
    td = inp['target_d'].squeeze(-1)
    assert z.min() >= min(cfg.cam_near, float(td[td > 0].min()) - cfg.training_range_d) - 1e-5
    assert z.max() <= max(cfg.cam_far, float(td.max()) + cfg.training_range_d) + 1e-5
    assert torch.isfinite(o1['rgb']).all() and torch.isfinite(o1['depth']).all()
    assert (o1['acc_map'] <= 1.0 + 1e-5).all() and (o1['acc_map'] >= 0).all()
    assert ((o1['rgb'] >= 0) & (o1['rgb'] <= 1 + 1e-5)).all()
    # perturbation changes z but keeps the order
    o3, _ = run(seed=78)
    assert not torch.equal(o3['z_vals'], z) and (o3['z_vals'][:, 1:] >= o3['z_vals'][:, :-1]).all()


@pytest.mark.gpu
def test_pointslam_knn_properties_large(cuda_dev):
    """200 k points, 100 k queries: rows ascending in D, ids unique and in range, the count
    equals the number of D < r^2, sentinels exactly where nothing is in range."""
    from xrdslam_b200.neural_point_cloud import NeuralPointCloud
    g = torch.Generator().manual_seed(0)
    N, Q = 200_000, 100_000
    pos = (torch.rand(N, 3, generator=g) - 0.5) * torch.tensor([6.0, 6.5, 4.5])
    npc = NeuralPointCloud(device=cuda_dev)
    npc.set_cloud(pos, torch.zeros(N, 32))
    q = (torch.rand(Q, 3, generator=g) - 0.5) * torch.tensor([6.4, 6.9, 4.9])
    r = torch.rand(Q, generator=g) * 0.12 + 0.04
    D, I, n = npc.find_neighbors(q, r)
    D, I, n, r2 = D.cpu(), I.cpu().long(), n.cpu().long(), (r * r)[:, None]
    assert (D[:, 1:] >= D[:, :-1]).all()
    found = I >= 0
    assert torch.equal(found, D < 3e38)
    assert (D[found] <= r2.expand_as(D)[found]).all()
    assert torch.equal(n, (D < r2).sum(1))
    assert (I[found] < N).all()
    srt = torch.sort(torch.where(found, I, torch.arange(8).neg() - 1), 1).values
    assert (srt[:, 1:] != srt[:, :-1]).all()  # no id twice in a row
    # distances are what the ids say
    rows = torch.nonzero(found)
    sub = rows[torch.randperm(rows.shape[0], generator=g)[:20000]]
    d = ((pos[I[sub[:, 0], sub[:, 1]]] - q[sub[:, 0]])**2)
    d = (d[:, 0] + d[:, 1]) + d[:, 2]
    assert torch.equal(d, D[sub[:, 0], sub[:, 1]])


@pytest.mark.gpu
def test_voxfusion_octree_and_march_properties(cuda_dev):
    """Full 640x480 frame (~300 k back-projected points): inserting the same voxels twice is
    idempotent; every sample lies inside the hit interval of a leaf the ray intersects and the
    samples of a ray are ordered in depth."""
    from xrdslam_b200.frame import Frame
    from xrdslam_b200.voxfusion import VoxFusionConfig
    off = (10.0, 10.0, 10.0)
    cam, poses, fr = _scene(1, offset=off)
    algo = VoxFusionConfig().setup(camera=cam, device=cuda_dev)
    f0 = Frame(0, fr[0][0], fr[0][1], init_pose=poses[0], rot_rep='quat')
    algo.create_voxels(f0)
    n1 = algo.model.map_states['voxel_center_xyz'].shape[0]
    vi = algo.model.map_states['voxel_vertex_idx']
    algo.create_voxels(f0)
    assert algo.model.map_states['voxel_center_xyz'].shape[0] == n1
    assert torch.equal(vi, algo.model.map_states['voxel_vertex_idx'])
    leaf = (vi >= 0).all(1)
    assert leaf.any() and vi[leaf].max() < algo.model.config.num_embeddings
    inp = algo.get_model_input([f0], True)
    m = algo.model.march(inp['rays_o'].detach().contiguous(), inp['rays_d'].detach().contiguous())
    assert m['n_hit_rays'] > 0 and m['n_points'] > 0
    cnt = m['smp_count'].cpu().long()
    dep, idx = m['smp_depth'].cpu(), m['smp_idx'].cpu().long()
    hit_idx, tmin, tmax = m['hit_idx'].cpu().long(), m['hit_tmin'].cpu(), m['hit_tmax'].cpu()
    assert int(cnt.sum()) == m['n_points']
    rows = torch.nonzero(cnt > 0).flatten()[:400]
    for r in rows.tolist():
        c = int(cnt[r])
        d = dep[r, :c]
        assert (d[1:] >= d[:-1]).all()
        for k in (0, c // 2, c - 1):
            h = torch.nonzero(hit_idx[r] == idx[r, k]).flatten()
            assert h.numel() >= 1
            assert tmin[r, h[0]] - 1e-4 <= d[k] <= tmax[r, h[0]] + 1e-4


@pytest.mark.gpu
def test_nice_full_batch_properties(cuda_dev):
    """1000 rays x 48 samples through the fused NICE step (stage color): finite outputs, depth
    within the sampled range, uncertainty >= 0, finite non-empty grid gradients, pose gradients for every frame."""
    from xrdslam_b200.frame import Frame
    from xrdslam_b200.nice_slam import NiceSLAMConfig
    cam, poses, fr = _scene(5)
    algo = NiceSLAMConfig(mapping_bound=[[-3.2, 3.2], [-4.2, 2.7], [-2.2, 2.7]]).setup(
        camera=cam, device=cuda_dev)
    frames = [Frame(k, fr[k][0], fr[k][1], init_pose=poses[k], rot_rep='quat') for k in range(5)]
    algo.stage = 'color'
    inp = algo.get_model_input(frames, True)
    assert 900 <= inp['rays_o'].shape[0] <= 1000
    out = algo.model(inp)
    ld = algo.model.get_loss_dict(out, inp, True, 'color')
    sum(ld.values()).backward()
    assert torch.isfinite(out['depth']).all() and torch.isfinite(out['rgb']).all()
    assert (out['uncertainty'] >= 0).all()
    td = inp['target_d'].squeeze(-1)
    assert (out['depth'] >= 0).all() and (out['depth'] <= 1.2 * td.max() + 0.02).all()
    for k, g in algo.model.grids.items():
        if k == 'grid_coarse':  # not part of stage 'color'
            assert g.grad is None
            continue
        assert torch.isfinite(g.grad).all() and (g.grad.abs().sum(-1) > 0).any(), k
    for f in frames[1:]:
        assert any(p.grad is not None and p.grad.abs().sum() > 0 for p in f.pose.parameters())
