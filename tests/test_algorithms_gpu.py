"""End-to-end runs of the algorithm layers (host mirrors of slam/algorithms/{nice_slam,
voxfusion,point_slam,coslam}.py) on the synthetic sequence: mapping + tracking through
Algorithm.optimize_update -- the loop the reference's Mapper/Tracker call."""
import numpy as np
import pytest
import torch


def _frames(n, width=160, height=120, rot_rep='axis_angle', separate_LR=False, noise=0.0,
            offset=(0.0, 0.0, 0.0)):
    from xrdslam_b200.frame import Frame
    from xrdslam_b200.synthetic import make_sequence
    cam, poses, frames = make_sequence(n, width=width, height=height, offset=offset)
    out = []
    for k in range(n):
        pose = poses[k].copy()
        if noise and k > 0:
            pose[:3, 3] += np.float32(noise) * np.array([1.0, -0.5, 0.3], np.float32)
        out.append(Frame(k, frames[k][0], frames[k][1], init_pose=pose,
                         separate_LR=separate_LR, rot_rep=rot_rep))
    return cam, poses, out


def _losses(algo, frames, n_iters, is_mapping):
    """optimize_update with the loss of every iteration recorded."""
    rec = []
    orig = algo.get_loss

    def wrapped(*a, **k):
        loss = orig(*a, **k)
        rec.append(float(loss.detach()))
        return loss
    algo.get_loss = wrapped
    try:
        ret = algo.optimize_update(n_iters, frames, is_mapping=is_mapping)
    finally:
        algo.get_loss = orig
    return rec, ret


@pytest.mark.gpu
def test_nice_slam_mapping_tracking(cuda_dev):
    from xrdslam_b200.nice_slam import NiceSLAMConfig
    torch.manual_seed(0)
    np.random.seed(0)
    cam, poses, fr = _frames(3, rot_rep='quat')
    cfg = NiceSLAMConfig(mapping_bound=[[-3.2, 3.2], [-4.2, 2.7], [-2.2, 2.7]],
                         marching_cubes_bound=[[-3.2, 3.2], [-4.2, 2.7], [-2.2, 2.7]],
                         ray_batch_size=8000, tracking_Wedge=10, tracking_Hedge=10)
    algo = cfg.setup(camera=cam, device=cuda_dev)
    g_before = {k: v.detach().clone() for k, v in algo.model.grids.items()}
    algo.add_keyframe(fr[0])
    rec, _ = _losses(algo, [fr[0]], 60, True)
    assert all(np.isfinite(rec))
    # stage schedule: middle (<= 0.4 n), fine (<= 0.6 n), color
    assert np.mean(rec[18:24]) < np.mean(rec[0:6])  # middle stage depth loss decreases
    algo.set_initialized()
    # frustum feature selection: voxels outside the mask keep their values exactly
    for k, g in algo.model.grids.items():
        if k == 'grid_coarse':  # optimised whole, and only by the coarse mapper (do_mapping)
            assert torch.equal(g.detach(), g_before[k])
            continue
        m = algo.model.grid_opti_mask[k].bool()
        assert 0 < int(m.sum()) < m.numel()
        assert torch.equal(g.detach()[~m], g_before[k][~m])
        assert not torch.equal(g.detach()[m], g_before[k][m])
    # tracking on the next frame from a perturbed pose
    cam2, _, pert = _frames(3, rot_rep='quat', noise=0.03)
    trk = pert[1]
    p0 = trk.get_pose().detach().clone()
    rec_t, cand = _losses(algo, [trk], 10, False)
    assert all(np.isfinite(rec_t)) and cand is not None and cand.shape == (4, 4)
    assert not torch.equal(trk.get_pose().detach(), p0)
    color, depth = algo.render_img(poses[0], gt_depth=fr[0].depth)
    assert color.shape == (cam.height, cam.width, 3) and depth.shape == (cam.height, cam.width)
    assert np.isfinite(color).all() and np.isfinite(depth).all()


@pytest.mark.gpu
def test_voxfusion_mapping_tracking(cuda_dev):
    from xrdslam_b200.voxfusion import VoxFusionConfig
    torch.manual_seed(0)
    np.random.seed(0)
    off = (10.0, 10.0, 10.0)  # tracker init_pose_offset=10: octree coordinates stay positive
    cam, poses, fr = _frames(2, rot_rep='quat', offset=off)
    algo = VoxFusionConfig().setup(camera=cam, device=cuda_dev)
    algo.add_keyframe(fr[0])
    rec, _ = _losses(algo, [fr[0]], 30, True)
    assert all(np.isfinite(rec)) and np.mean(rec[-5:]) < np.mean(rec[:5])
    assert algo.model.map_states['voxel_center_xyz'].shape[0] > 100
    algo.set_initialized()
    cam2, _, pert = _frames(2, rot_rep='quat', noise=0.02, offset=off)
    rec_t, cand = _losses(algo, [pert[1]], 10, False)
    assert all(np.isfinite(rec_t)) and cand is not None
    color, depth = algo.render_img(poses[0])
    assert color.shape == (cam.height, cam.width, 3) and np.isfinite(depth).all()


@pytest.mark.gpu
def test_point_slam_mapping_tracking(cuda_dev):
    from xrdslam_b200.point_slam import PointSLAMConfig
    torch.manual_seed(0)
    np.random.seed(0)
    cam, poses, fr = _frames(2, separate_LR=True)
    cfg = PointSLAMConfig(pixels_adding=1500, mapping_sample=1000, tracking_sample=300,
                          tracking_Wedge=10, tracking_Hedge=10)
    algo = cfg.setup(camera=cam, device=cuda_dev)
    algo.add_keyframe(fr[0])
    rec, _ = _losses(algo, [fr[0]], 40, True)
    npc = algo.model.neural_point_cloud
    assert npc.pts_num() > 1000 and all(np.isfinite(rec))
    # geometry stage (first 40 %) reduces the depth loss; the colour stage then adds rgb_loss
    assert np.mean(rec[10:16]) < np.mean(rec[0:4])
    assert np.mean(rec[-4:]) < np.mean(rec[18:22])
    algo.set_initialized()
    cam2, _, pert = _frames(2, separate_LR=True, noise=0.01)
    trk = pert[1]
    r0 = trk.pose.data_t.detach().clone()
    rec_t, cand = _losses(algo, [trk], 8, False)
    assert all(np.isfinite(rec_t)) and cand is not None
    assert not torch.equal(trk.pose.data_t.detach(), r0)
    color, depth = algo.render_img(poses[0], gt_depth=fr[0].depth, idx=0)
    assert color.shape == (cam.height, cam.width, 3) and np.isfinite(color).all()


@pytest.mark.gpu
def test_coslam_mapping_tracking(cuda_dev):
    from xrdslam_b200.coslam import CoSLAMConfig
    torch.manual_seed(0)
    np.random.seed(0)
    cam, poses, fr = _frames(3, separate_LR=True)
    # 160x120 frames: the bank keeps 5 % = 960 rays per keyframe
    algo = CoSLAMConfig(mapping_sample=512, tracking_sample=256, tracking_Wedge=5,
                        tracking_Hedge=5).setup(camera=cam, device=cuda_dev)
    # first map: current frame only -- runs as the captured CUDA-graph iteration
    assert algo._graph_ok([fr[0]])
    algo.optimize_update(3, [fr[0]], True)
    sess = algo.mapping_session([fr[0]])
    l0 = float(sess.loss_total)
    algo.optimize_update(37, [fr[0]], True)
    l1 = float(sess.loss_total)
    assert np.isfinite(l0) and l1 < 0.5 * l0
    algo.add_keyframe(fr[0])
    algo.set_initialized()
    cam2, _, pert = _frames(3, separate_LR=True, noise=0.02)
    rec_t, cand = _losses(algo, [pert[1]], 10, False)
    assert all(np.isfinite(rec_t)) and cand is not None
    # bundle adjustment over [kf0, cur] with the global ray bank
    window = algo.select_optimize_frames(pert[1], 'all')
    t_before = pert[1].pose.data_t.detach().clone()
    algo.optimize_update(10, window, True)
    sess2 = algo.mapping_session(window)
    assert sess2.ba and not sess2.first and np.isfinite(float(sess2.loss_total))
    assert not torch.equal(pert[1].pose.data_t.detach(), t_before)  # BA moved the pose
    # the generic autograd path stays available (graph_mapping=False) and agrees in form
    algo.config.graph_mapping = False
    rec2, _ = _losses(algo, window, 5, True)
    assert all(np.isfinite(rec2))


@pytest.mark.gpu
def test_nice_slam_coarse_mapper(cuda_dev):
    """The coarse mapper of the reference run config (coarse=True, nice_slam.py:102-109):
    optimize_update(coarse=True) renders stage 'coarse' and moves ONLY the coarse grid."""
    from xrdslam_b200.nice_slam import NiceSLAMConfig
    torch.manual_seed(0)
    cam, poses, fr = _frames(2, rot_rep='quat')
    algo = NiceSLAMConfig(mapping_bound=[[-3.2, 3.2], [-4.2, 2.7], [-2.2, 2.7]],
                          marching_cubes_bound=[[-3.2, 3.2], [-4.2, 2.7], [-2.2, 2.7]]).setup(
        camera=cam, device=cuda_dev)
    assert algo.config.coarse and 'grid_coarse' in algo.model.grids
    g_before = {k: v.detach().clone() for k, v in algo.model.grids.items()}
    algo.add_keyframe(fr[0])
    rec = []
    orig = algo.get_loss

    def wrapped(*a, **k):
        loss = orig(*a, **k)
        rec.append(float(loss.detach()))
        assert algo.stage == 'coarse'
        return loss
    algo.get_loss = wrapped
    algo.optimize_update(40, [fr[0]], is_mapping=True, coarse=True)
    algo.get_loss = orig
    assert len(rec) == 40 and all(np.isfinite(rec))
    assert np.mean(rec[-5:]) < np.mean(rec[:5])
    for k, g in algo.model.grids.items():
        if k == 'grid_coarse':
            assert not torch.equal(g.detach(), g_before[k])
        else:
            assert torch.equal(g.detach(), g_before[k]), k
