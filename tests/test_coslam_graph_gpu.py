"""Device poses (csrc/pose.cu) and the CUDA-graph mapping iteration (coslam_graph.py) against
the generic autograd path (Algorithm.optimize_update, base_algorithm.py:239-275)."""
import random

import numpy as np
import pytest
import torch


@pytest.mark.gpu
def test_pose_matrices_and_grads_match_optimizable_pose(cuda_dev):
    import ctypes as C
    from xrdslam_b200 import _cabi
    from xrdslam_b200.opt_pose import OptimizablePose
    g = torch.Generator().manual_seed(0)
    n = 9
    rot = torch.randn(n, 3, generator=g) * torch.tensor([0.01, 0.3, 1.0, 2.0, 3.0, 0.5, 1e-3, 0.0, 1.5])[:, None]
    rot[7] = 0  # identity branch
    trans = torch.randn(n, 3, generator=g)
    W = torch.randn(n, 4, 4, generator=g)
    poses = [OptimizablePose(torch.cat([trans[i], rot[i]]).clone()) for i in range(n)]
    M_ref = torch.stack([p.matrix() for p in poses])
    (M_ref * W).sum().backward()
    lib = _cabi.lib()
    d = lambda t: t.to(cuda_dev).contiguous()
    rot_d, trans_d, W_d = d(rot), d(trans), d(W)
    M = torch.empty(n, 4, 4, device=cuda_dev)
    st = torch.cuda.current_stream(cuda_dev).cuda_stream
    _cabi.check('fwd', lib.xrd_pose_matrices(n, rot_d.data_ptr(), trans_d.data_ptr(), M.data_ptr(), st))
    assert (M.cpu() - M_ref.detach()).abs().max() < 1e-6
    d_rot, d_trans = torch.zeros(n, 3, device=cuda_dev), torch.zeros(n, 3, device=cuda_dev)
    fixed = torch.zeros(n, dtype=torch.uint8, device=cuda_dev)
    fixed[2] = 1
    _cabi.check('bwd', lib.xrd_pose_matrices_grads(n, rot_d.data_ptr(), W_d.data_ptr(), fixed.data_ptr(),
                                                   d_rot.data_ptr(), d_trans.data_ptr(), st))
    for i, p in enumerate(poses):
        if i == 2:
            assert d_rot[i].abs().sum() == 0 and d_trans[i].abs().sum() == 0
            continue
        gr = p.data_r.grad if p.data_r.grad is not None else torch.zeros(3)
        assert (d_rot[i].cpu() - gr).abs().max() <= 2e-5 * max(1.0, float(gr.abs().max())), i
        assert torch.allclose(d_trans[i].cpu(), p.data_t.grad, atol=1e-6)


def _run(dev, graph, n_iters, seed=11):
    import bench
    random.seed(seed)
    algo, kfs, cur = bench.build_algorithm(dev, seed=seed)
    algo.config.graph_mapping = graph
    algo.config.min_sample_pixels = 256
    frames = kfs + [cur]
    torch.manual_seed(seed)
    losses = []
    if graph:
        algo.setup_optimizers(n_iters, frames, True)
        sess = algo.mapping_session(frames)
        sess.begin(frames)
        for i in range(n_iters):
            losses.append(float(sess.step(i, frames)))
        sess.end(frames)
    else:
        orig = algo.get_loss

        def wrapped(*a, **k):
            loss = orig(*a, **k)
            losses.append(float(loss.detach()))
            return loss
        algo.get_loss = wrapped
        algo.optimize_update(n_iters, frames, True)
    return algo, frames, losses


@pytest.mark.gpu
def test_graph_mapping_matches_generic_path(cuda_dev):
    """Same samples, seeds and Adam state: the captured iteration and the autograd path give
    the same loss trajectory, decoder weights and bundle-adjusted poses (7 iterations: the
    mapping poses step once, at iteration 5, on the summed gradients)."""
    a_g, f_g, l_g = _run(cuda_dev, True, 7)
    a_e, f_e, l_e = _run(cuda_dev, False, 7)
    assert len(l_g) == len(l_e) == 7
    for x, y in zip(l_g, l_e):
        assert abs(x - y) <= 2e-3 * abs(y), (l_g, l_e)
    assert l_g[-1] < l_g[0]
    for p, q in zip(a_g.model.decoder.parameters(), a_e.model.decoder.parameters()):
        assert (p - q).abs().max() <= 2e-3 * q.abs().max()
    moved = 0
    for fg, fe in zip(f_g, f_e):
        assert (fg.pose.data_t - fe.pose.data_t).abs().max() < 2e-5
        assert (fg.pose.data_r - fe.pose.data_r).abs().max() < 2e-5
    import bench
    _, kfs0, cur0 = bench.build_algorithm(cuda_dev, seed=11)
    for f0, fg in zip(kfs0 + [cur0], f_g):
        moved += int((f0.pose.data_t - fg.pose.data_t).abs().max() > 1e-5)
    assert moved == len(f_g) - 1  # every pose but the fixed first one was optimised
    assert torch.equal(kfs0[0].pose.data_t, f_g[0].pose.data_t)


@pytest.mark.gpu
def test_graph_first_mapping_and_optimize_update(cuda_dev):
    """First map (no keyframes, no smoothness, no BA) through CoSLAM.optimize_update."""
    import bench
    from xrdslam_b200.coslam import CoSLAMConfig
    from xrdslam_b200.frame import Frame
    from xrdslam_b200.synthetic import make_sequence
    torch.manual_seed(0)
    cam, poses, fr = make_sequence(1)
    algo = CoSLAMConfig().setup(camera=cam, device=cuda_dev)
    f0 = Frame(0, fr[0][0], fr[0][1], init_pose=poses[0], separate_LR=True, rot_rep='axis_angle')
    t0 = algo.model.embed_fn.params.detach().clone()
    assert algo._graph_ok([f0])
    algo.optimize_update(20, [f0], True)
    sess = algo.mapping_session([f0])
    assert sess.first and not sess.ba
    l0 = float(sess.loss_total)
    algo.optimize_update(20, [f0], True)
    assert float(sess.loss_total) < l0 and np.isfinite(l0)
    assert not torch.equal(t0, algo.model.embed_fn.params.detach())


@pytest.mark.gpu
def test_tracking_graph_matches_generic_iteration(cuda_dev):
    """One tracking iteration (same pixel indices, same Philox seed): the captured pose-only
    iteration gives the autograd path's loss, pose gradient and Adam-updated pose."""
    import bench
    from xrdslam_b200.coslam_graph import TrackingGraphSession
    from xrdslam_b200.common import sample_window
    random.seed(3)
    algo, kfs, cur = bench.build_algorithm(cuda_dev, seed=3)
    algo.config.graph_mapping = False
    cam = algo.camera
    He, We, n = algo.config.tracking_Hedge, algo.config.tracking_Wedge, algo.config.tracking_sample
    g = torch.Generator().manual_seed(0)
    idx = torch.randint((cam.height - 2 * He) * (cam.width - 2 * We), (n,), generator=g).to(cuda_dev)
    # graph path (indices supplied from outside)
    sess = TrackingGraphSession(algo, external_indices=True)
    r0, t0 = cur.pose.data_r.detach().clone(), cur.pose.data_t.detach().clone()
    sess.begin(cur)
    sess.idx.copy_(idx)
    c0 = algo.model._step_count
    loss_g = float(sess.step())
    d_rot_g, d_trans_g = sess.d_rot.cpu().reshape(3), sess.d_trans.cpu().reshape(3)
    cand = sess.end(cur)
    r_g, t_g = cur.pose.data_r.detach().clone(), cur.pose.data_t.detach().clone()
    with torch.no_grad():
        cur.pose.data_r.copy_(r0)
        cur.pose.data_t.copy_(t0)
    # autograd path on the same pixels
    algo.model._step_count = c0
    algo.model.freeze_map_grads = True
    opt = algo.setup_optimizers(1, [cur], is_mapping=False)
    opt.zero_grad_all()
    from xrdslam_b200.opt_pose import pose_matrices
    ro, rd, d, c = sample_window(cam, [algo._frame_tensor(cur, 'depth')], [algo._frame_tensor(cur, 'rgb')],
                                 pose_matrices([cur.pose]).to(cuda_dev), n, He, We, indices=idx)
    inp = dict(rays_o=ro, rays_d=rd, target_s=c, target_d=d, first=False)
    ld = algo.model.get_loss_dict(algo.model(inp), inp, False, 0)
    loss = sum(ld.values())
    loss.backward()
    assert abs(loss_g - float(loss.detach())) <= 1e-5 * abs(float(loss.detach()))
    assert (d_rot_g - cur.pose.data_r.grad).abs().max() <= 1e-4 * cur.pose.data_r.grad.abs().max()
    assert (d_trans_g - cur.pose.data_t.grad).abs().max() <= 1e-4 * cur.pose.data_t.grad.abs().max()
    opt.optimizer_step_all(step=0)
    assert (r_g - cur.pose.data_r.detach()).abs().max() < 1e-6
    assert (t_g - cur.pose.data_t.detach()).abs().max() < 1e-6
    # one iteration: the candidate is the starting pose
    from xrdslam_b200.opt_pose import OptimizablePose
    assert np.allclose(cand, OptimizablePose(torch.cat([t0, r0])).matrix().detach().numpy(), atol=1e-6)


@pytest.mark.gpu
def test_tracking_graph_recovers_perturbed_pose(cuda_dev):
    """After a short mapping run, 30 captured tracking iterations from a 3 cm perturbed pose
    reduce the translation error; the returned candidate has the smallest loss seen."""
    import bench
    random.seed(4)
    algo, kfs, cur = bench.build_algorithm(cuda_dev, seed=4)
    frames = kfs + [cur]
    algo.optimize_update(60, frames, True)  # graph mapping: learn the scene
    gt_t = cur.pose.data_t.detach().clone()
    with torch.no_grad():
        cur.pose.data_t.add_(torch.tensor([0.03, -0.02, 0.01]))
    e0 = float((cur.pose.data_t.detach() - gt_t).norm())
    cand = algo.optimize_update(30, [cur], False)
    assert cand.shape == (4, 4) and np.isfinite(cand).all()
    e1 = float((torch.from_numpy(cand[:3, 3]) - gt_t).norm())
    assert e1 < e0, (e0, e1)
    assert float(algo.tracking_session().best_loss) < float('inf')
