"""Parity of the fused Co-SLAM CUDA path (through the C-ABI) against the CPU
oracle (oracle/coslam.py, itself pinned to the reference's JointEncoding)."""
import numpy as np
import pytest
import torch

from helpers import BOUND, coslam_pair, make_rays, max_abs, rel_err

pytestmark = pytest.mark.gpu

# fp32 kernel vs fp32 oracle: different summation order only
TOL_OUT = 2e-5     # abs, rgb in [0,1], depth in metres
TOL_LOSS = 2e-5    # rel
TOL_GRAD = 2e-4    # rel l2 over the whole gradient tensor


def test_hash_indices_bit_exact(cuda_dev):
    import ctypes as C
    from oracle.tcnn_restated import hashgrid_indices
    from xrdslam_b200 import _cabi
    ora, model = coslam_pair(cuda_dev)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(20000, 3, generator=g) * 3 - 1  # in and outside [0,1]
    x[:8] = torch.tensor([[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [1, 0, 1],
                          [-0.25, 0.3, 2.0], [1e-7, 1 - 1e-7, 0.999999],
                          [-1e-4, 0.2, 0.3], [3.1, -2.2, 0.5]])
    idx_o, _ = hashgrid_indices(x, ora.embed_fn.table)
    xd = x.to(cuda_dev)
    feat = torch.empty(x.shape[0], 32, device=cuda_dev)
    idx = torch.empty(x.shape[0], 16, 8, dtype=torch.int32, device=cuda_dev)
    grid = model._grid_struct(model.embed_fn.params.detach())
    st = _cabi.lib().xrd_hashgrid_encode(C.byref(grid), xd.data_ptr(),
                                         x.shape[0], feat.data_ptr(),
                                         idx.data_ptr(), None)
    _cabi.check('xrd_hashgrid_encode', st)
    torch.cuda.synchronize()
    got = idx.cpu().to(torch.int64) & 0xFFFFFFFF
    assert torch.equal(got, idx_o)  # bit-exact
    with torch.no_grad():
        f_o = ora.embed_fn(x)
    assert max_abs(feat, f_o) < 1e-6


# rays_per_tile: -1 = tile kernel (k_fused), -2 = grouped persistent kernel (k_fused_g); the
# library's automatic choice (0) picks between the two by batch size
@pytest.mark.parametrize('kernel', [-1, -2])
@pytest.mark.parametrize('R', [1, 5, 257, 1024])
def test_step_parity(cuda_dev, R, kernel):
    ora, model = coslam_pair(cuda_dev, rays_per_tile=kernel)
    rays_o, rays_d, ts, td, noise = make_rays(R, seed=R)
    rays_o.requires_grad_(True)
    rays_d.requires_grad_(True)
    out_o, ld_o, tot_o = ora.step(rays_o, rays_d, ts, td, noise, False, True)
    tot_o.backward()

    ro = rays_o.detach().to(cuda_dev).requires_grad_(True)
    rd = rays_d.detach().to(cuda_dev).requires_grad_(True)
    inp = dict(rays_o=ro, rays_d=rd, target_s=ts.to(cuda_dev),
               target_d=td.to(cuda_dev), first=True, noise=noise.to(cuda_dev))
    out = model(inp)
    ld = model.get_loss_dict(out, inp, False, 0)
    tot = sum(ld.values())
    tot.backward()
    torch.cuda.synchronize()

    assert torch.equal(out['z_vals'].cpu(), out_o['z_vals'])  # bit-exact
    for k in ('rgb', 'depth', 'acc_map', 'depth_var'):
        assert max_abs(out[k], out_o[k]) < TOL_OUT, k
    assert max_abs(out['raw'], out_o['raw']) < 5e-5
    for k in ld_o:
        a, b = float(ld[k].detach()), float(ld_o[k].detach())
        assert abs(a - b) <= TOL_LOSS * max(abs(b), 1e-6), (k, a, b)
    assert rel_err(model.embed_fn.params.grad, ora.embed_fn.params.grad) < TOL_GRAD
    pairs = [(model.decoder.sdf_net.model[0], ora.sdf0),
             (model.decoder.sdf_net.model[2], ora.sdf1),
             (model.decoder.color_net.model[0], ora.col0),
             (model.decoder.color_net.model[2], ora.col1)]
    for m, o in pairs:
        assert rel_err(m.weight.grad, o.weight.grad) < TOL_GRAD
    assert rel_err(ro.grad, rays_o.grad) < TOL_GRAD
    assert rel_err(rd.grad, rays_d.grad) < TOL_GRAD


def test_smoothness_parity(cuda_dev):
    ora, model = coslam_pair(cuda_dev)
    rnd = torch.tensor([0.3, 0.7, 0.1, 0.9, 0.2, 0.5])
    s_o = ora.smoothness(rnd.reshape(2, 3))
    s_o.backward()
    s = model.smoothness(32, 0.1, 0.05, rand=rnd)
    s.backward()
    torch.cuda.synchronize()
    assert abs(float(s) - float(s_o)) <= 2e-5 * abs(float(s_o))
    assert rel_err(model.embed_fn.params.grad, ora.embed_fn.params.grad) < TOL_GRAD


def test_mapping_step_with_smoothness(cuda_dev):
    ora, model = coslam_pair(cuda_dev)
    R = 300
    rays_o, rays_d, ts, td, noise = make_rays(R, seed=11)
    rnd = torch.tensor([0.11, 0.52, 0.93, 0.4, 0.6, 0.8])
    out_o, ld_o, tot_o = ora.step(rays_o, rays_d, ts, td, noise, True, False,
                                  smooth_rand=rnd.reshape(2, 3))
    tot_o.backward()
    inp = dict(rays_o=rays_o.to(cuda_dev), rays_d=rays_d.to(cuda_dev),
               target_s=ts.to(cuda_dev), target_d=td.to(cuda_dev), first=False,
               noise=noise.to(cuda_dev), smooth_rand=rnd)
    out = model(inp)
    ld = model.get_loss_dict(out, inp, True, 0)
    assert set(ld) == set(ld_o)
    import functools
    functools.reduce(torch.add, ld.values()).backward()
    torch.cuda.synchronize()
    for k in ld_o:
        a, b = float(ld[k].detach()), float(ld_o[k].detach())
        assert abs(a - b) <= TOL_LOSS * max(abs(b), 1e-9), (k, a, b)
    assert rel_err(model.embed_fn.params.grad, ora.embed_fn.params.grad) < TOL_GRAD


def test_render_only_no_depth(cuda_dev):
    """target_d=None -> 256 uniform samples (joint_encoding.py:281-284)."""
    ora, model = coslam_pair(cuda_dev, training_perturb=0)
    ora.cfg.perturb = 0
    rays_o, rays_d, _, _, _ = make_rays(37, seed=5)
    with torch.no_grad():
        out_o = ora.render_rays(rays_o, rays_d, None, None)
        out = model(dict(rays_o=rays_o.to(cuda_dev), rays_d=rays_d.to(cuda_dev),
                         target_s=None, target_d=None))
    assert torch.equal(out['z_vals'].cpu(), out_o['z_vals'])
    assert max_abs(out['rgb'], out_o['rgb']) < TOL_OUT
    assert max_abs(out['depth'], out_o['depth']) < 5e-5


def test_strict_nonunit_loss_grad(cuda_dev):
    ora, model = coslam_pair(cuda_dev, strict_loss_grad=True)
    rays_o, rays_d, ts, td, noise = make_rays(64, seed=2)
    out_o, ld_o, _ = ora.step(rays_o, rays_d, ts, td, noise, False, True)
    (2.0 * ld_o['rgb_loss'] + 0.5 * ld_o['sdf_loss'] + ld_o['fs_loss']).backward()
    inp = dict(rays_o=rays_o.to(cuda_dev), rays_d=rays_d.to(cuda_dev),
               target_s=ts.to(cuda_dev), target_d=td.to(cuda_dev), first=True,
               noise=noise.to(cuda_dev))
    ld = model.get_loss_dict(model(inp), inp, False, 0)
    (2.0 * ld['rgb_loss'] + 0.5 * ld['sdf_loss'] + ld['fs_loss']).backward()
    assert rel_err(model.embed_fn.params.grad, ora.embed_fn.params.grad) < TOL_GRAD


def test_step_parity_mixed_mode(cuda_dev):
    """precision=1 (bench default): 3xTF32 forward -> outputs and losses at the
    fp32 tolerance; plain-TF32 backward -> gradients within 5e-3 relative."""
    ora, model = coslam_pair(cuda_dev, precision=1)
    R = 512
    rays_o, rays_d, ts, td, noise = make_rays(R, seed=22)
    rays_o.requires_grad_(True)
    rays_d.requires_grad_(True)
    out_o, ld_o, tot_o = ora.step(rays_o, rays_d, ts, td, noise, False, True)
    tot_o.backward()
    ro = rays_o.detach().to(cuda_dev).requires_grad_(True)
    rd = rays_d.detach().to(cuda_dev).requires_grad_(True)
    inp = dict(rays_o=ro, rays_d=rd, target_s=ts.to(cuda_dev),
               target_d=td.to(cuda_dev), first=True, noise=noise.to(cuda_dev))
    out = model(inp)
    ld = model.get_loss_dict(out, inp, False, 0)
    sum(ld.values()).backward()
    assert torch.equal(out['z_vals'].cpu(), out_o['z_vals'])
    assert max_abs(out['rgb'], out_o['rgb']) < TOL_OUT
    assert max_abs(out['depth'], out_o['depth']) < TOL_OUT
    for k in ld_o:
        a, b = float(ld[k].detach()), float(ld_o[k].detach())
        assert abs(a - b) <= TOL_LOSS * max(abs(b), 1e-6), (k, a, b)
    assert rel_err(model.embed_fn.params.grad, ora.embed_fn.params.grad) < 5e-3
    assert rel_err(model.decoder.sdf_net.model[0].weight.grad, ora.sdf0.weight.grad) < 5e-3
    assert rel_err(model.decoder.color_net.model[0].weight.grad, ora.col0.weight.grad) < 5e-3
    assert rel_err(ro.grad, rays_o.grad) < 5e-3
    assert rel_err(rd.grad, rays_d.grad) < 5e-3


def test_step_parity_tf32_mode(cuda_dev):
    """precision=2: plain TF32 decoder GEMMs everywhere (10-bit mantissa operands,
    fp32 accumulate) -- the looser tolerance of SURVEY 8c.  The stress parameters
    here (|table| <= 0.3, unit-variance weights) are far larger than a trained map."""
    ora, model = coslam_pair(cuda_dev, precision=2)
    R = 512
    rays_o, rays_d, ts, td, noise = make_rays(R, seed=21)
    out_o, ld_o, tot_o = ora.step(rays_o, rays_d, ts, td, noise, False, True)
    tot_o.backward()
    inp = dict(rays_o=rays_o.to(cuda_dev), rays_d=rays_d.to(cuda_dev),
               target_s=ts.to(cuda_dev), target_d=td.to(cuda_dev), first=True,
               noise=noise.to(cuda_dev))
    out = model(inp)
    ld = model.get_loss_dict(out, inp, False, 0)
    sum(ld.values()).backward()
    assert torch.equal(out['z_vals'].cpu(), out_o['z_vals'])
    assert max_abs(out['rgb'], out_o['rgb']) < 5e-2
    assert max_abs(out['depth'], out_o['depth']) < 5e-2
    for k in ld_o:
        a, b = float(ld[k].detach()), float(ld_o[k].detach())
        assert abs(a - b) <= 3e-2 * max(abs(b), 1e-6), (k, a, b)
    ga, gb = model.embed_fn.params.grad.cpu().double(), ora.embed_fn.params.grad.double()
    cos = float((ga * gb).sum() / (ga.norm() * gb.norm()))
    assert cos > 0.995


@pytest.mark.parametrize('kernel', [-1, -2])
def test_tracking_pose_only(cuda_dev, kernel):
    """freeze_map_grads: only d loss / d rays is produced (tracking)."""
    ora, model = coslam_pair(cuda_dev, rays_per_tile=kernel)
    model.freeze_map_grads = True
    rays_o, rays_d, ts, td, noise = make_rays(200, seed=4)
    rays_o.requires_grad_(True)
    rays_d.requires_grad_(True)
    _, _, tot_o = ora.step(rays_o, rays_d, ts, td, noise, False, False)
    tot_o.backward()
    ro = rays_o.detach().to(cuda_dev).requires_grad_(True)
    rd = rays_d.detach().to(cuda_dev).requires_grad_(True)
    inp = dict(rays_o=ro, rays_d=rd, target_s=ts.to(cuda_dev),
               target_d=td.to(cuda_dev), first=False, noise=noise.to(cuda_dev))
    ld = model.get_loss_dict(model(inp), inp, False, 0)
    sum(ld.values()).backward()
    assert model.embed_fn.params.grad is None
    assert rel_err(ro.grad, rays_o.grad) < TOL_GRAD
    assert rel_err(rd.grad, rays_d.grad) < TOL_GRAD


@pytest.mark.parametrize('kernel', [-1, -2])
def test_cuda_matches_reference_golden(cuda_dev, kernel):
    """CUDA path vs vectors produced by the reference's own JointEncoding class."""
    from helpers import load_golden_coslam, set_coslam_params
    from xrdslam_b200.camera import Camera
    from xrdslam_b200.joint_encoding import JointEncodingConfig
    g = load_golden_coslam()
    model = JointEncodingConfig(rays_per_tile=kernel).setup(camera=Camera(320., 320., 319.5, 239.5, 640, 480),
                                        bounding_box=BOUND)
    set_coslam_params(model, g, 'model')
    model.to(cuda_dev)
    t = lambda k: torch.from_numpy(g[k]).to(cuda_dev)
    ro = t('rays_o').requires_grad_(True)
    rd = t('rays_d').requires_grad_(True)
    inp = dict(rays_o=ro, rays_d=rd, target_s=t('target_s'), target_d=t('target_d'),
               first=False, noise=t('noise'), smooth_rand=torch.from_numpy(g['smooth_rand']))
    out = model(inp)
    ld = model.get_loss_dict(out, inp, True, 0)
    sum(ld.values()).backward()
    assert np.array_equal(out['z_vals'].cpu().numpy(), g['z_vals'])
    assert np.abs(out['rgb'].detach().cpu().numpy() - g['rgb']).max() < TOL_OUT
    assert np.abs(out['depth'].detach().cpu().numpy() - g['depth']).max() < TOL_OUT
    assert np.abs(out['raw'].detach().cpu().numpy() - g['raw']).max() < 5e-5
    got = np.array([float(ld[k].detach()) for k in
                    ('rgb_loss', 'depth_loss', 'sdf_loss', 'fs_loss', 'smooth_loss')])
    assert np.allclose(got, g['losses'], rtol=TOL_LOSS, atol=0)
    assert rel_err(ro.grad, torch.from_numpy(g['d_rays_o'])) < TOL_GRAD
    assert rel_err(rd.grad, torch.from_numpy(g['d_rays_d'])) < TOL_GRAD
    assert rel_err(model.decoder.sdf_net.model[0].weight.grad, torch.from_numpy(g['d_w_sdf0'])) < TOL_GRAD
    assert rel_err(model.decoder.color_net.model[2].weight.grad, torch.from_numpy(g['d_w_col1'])) < TOL_GRAD
    gt = model.embed_fn.params.grad
    assert abs(float(gt.double().norm()) - float(g['d_table_norm'])) <= TOL_GRAD * float(g['d_table_norm'])
    assert rel_err(gt[:4096], torch.from_numpy(g['d_table_head'])) < TOL_GRAD


def test_mesher_queries_match_oracle(cuda_dev):
    """query_fn / color_func / query_sdf / query_color_sdf / run_network (mesher path, SURVEY f4;
    joint_encoding.py:408-507) against the oracle's query_color_sdf: exact fp32 decoders."""
    ora, model = coslam_pair(cuda_dev)
    g = torch.Generator().manual_seed(12)
    P = 3001
    world = torch.rand(P, 3, generator=g) * torch.tensor([7.0, 7.5, 5.5]) + torch.tensor([-3.5, -4.5, -2.5])
    bb = torch.as_tensor(BOUND, dtype=torch.float64)
    xn = ((world - bb[:, 0]) / (bb[:, 1] - bb[:, 0]))  # float64, like the reference
    with torch.no_grad():
        raw_o = ora.query_color_sdf(xn)
    sdf = model.query_fn(world.to(cuda_dev))
    col = model.color_func(world.to(cuda_dev))
    assert sdf.shape == (P, 1) and col.shape == (P, 1, 3)
    assert max_abs(sdf[:, 0], raw_o[:, 3]) < 2e-5
    assert max_abs(col[:, 0], torch.sigmoid(raw_o[:, :3])) < 2e-5
    raw = model.run_network(world[:1000].reshape(100, 10, 3).to(cuda_dev))
    assert raw.shape == (100, 10, 4) and max_abs(raw.reshape(-1, 4), raw_o[:1000]) < 2e-5
    # normalised-coordinate entry points
    xn32 = xn.float()
    with torch.no_grad():
        raw_n = ora.query_color_sdf(xn32)
    got = model.query_color_sdf(xn32.reshape(P, 1, 3).to(cuda_dev))
    assert got.shape == (P, 1, 4) and max_abs(got[:, 0], raw_n) < 2e-5
    s2, geo = model.query_sdf(xn32.reshape(P, 1, 3).to(cuda_dev), return_geo=True)
    assert s2.shape == (P, 1) and geo.shape == (P, 1, 15)
    assert max_abs(s2[:, 0], raw_n[:, 3]) < 2e-5
    emb = model.query_sdf(xn32.reshape(P, 1, 3).to(cuda_dev), embed=True)
    with torch.no_grad():
        assert max_abs(emb[:, 0], ora.embed_fn(xn32)) < 1e-6
