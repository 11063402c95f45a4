"""Generate the committed golden vectors by running the REFERENCE's own classes
(imported from /root/reference through oracle/ref_harness.py) on CPU.

    python tests/golden/make_golden.py

Only runs in the build container (the GPU box has no /root/reference).  The tinycudann
encodings inside the reference model are the restated ones (parity unpinned there, see
DESIGN.md section 2); everything else -- sampling, decoders, sdf2weights, raw2outputs, losses,
smoothness -- is the reference's code, executed unmodified.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from helpers import BOUND, make_rays  # noqa: E402
from oracle import ref_harness  # noqa: E402


def coslam(R=96, seed=7):
    bb = torch.from_numpy(BOUND)
    ref = ref_harness.ref_joint_encoding(bb)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        ref.embed_fn.params.copy_(
            (torch.rand(ref.embed_fn.params.shape, generator=g) * 2 - 1) * 0.3)
        for seq in (ref.decoder.sdf_net.model, ref.decoder.color_net.model):
            for lin in (seq[0], seq[2]):
                lin.weight.copy_(torch.randn(lin.weight.shape, generator=g) /
                                 np.sqrt(lin.weight.shape[1]))
    rays_o, rays_d, ts, td, noise = make_rays(R, seed=seed)
    rays_o.requires_grad_(True)
    rays_d.requires_grad_(True)
    # the reference draws torch.rand(z_vals.shape) then (mapping) rand(3), rand(1,1,1,3)
    torch.manual_seed(seed)
    noise = torch.rand(R, 43)
    r1 = torch.rand(3)
    r2 = torch.rand((1, 1, 1, 3))
    torch.manual_seed(seed)
    inp = dict(rays_o=rays_o, rays_d=rays_d, target_s=ts, target_d=td, first=False)
    out = ref(inp)
    ld = ref.get_loss_dict(out, inp, True, 0)
    total = sum(ld.values())
    total.backward()
    sd = ref.state_dict()
    np.savez_compressed(
        os.path.join(HERE, 'coslam_map_step.npz'),
        rays_o=rays_o.detach().numpy(), rays_d=rays_d.detach().numpy(),
        target_s=ts.numpy(), target_d=td.numpy(), noise=noise.numpy(),
        smooth_rand=torch.cat([r1, r2.reshape(3)]).numpy(),
        w_sdf0=sd['decoder.sdf_net.model.0.weight'].numpy(),
        w_sdf1=sd['decoder.sdf_net.model.2.weight'].numpy(),
        w_col0=sd['decoder.color_net.model.0.weight'].numpy(),
        w_col1=sd['decoder.color_net.model.2.weight'].numpy(),
        table_seed=np.int64(seed),  # the 6.5 MB table is regenerated from the seed
        table_checksum=np.float64(sd['embed_fn.params'].double().sum().item()),
        z_vals=out['z_vals'].detach().numpy(), raw=out['raw'].detach().numpy(),
        rgb=out['rgb'].detach().numpy(), depth=out['depth'].detach().numpy(),
        depth_var=out['depth_var'].detach().numpy(), acc=out['acc_map'].detach().numpy(),
        losses=np.array([float(ld[k].detach()) for k in
                         ('rgb_loss', 'depth_loss', 'sdf_loss', 'fs_loss', 'smooth_loss')]),
        d_rays_o=rays_o.grad.numpy(), d_rays_d=rays_d.grad.numpy(),
        d_w_sdf0=ref.decoder.sdf_net.model[0].weight.grad.numpy(),
        d_w_col1=ref.decoder.color_net.model[2].weight.grad.numpy(),
        d_table_norm=np.float64(ref.embed_fn.params.grad.double().norm().item()),
        d_table_nnz=np.int64((ref.embed_fn.params.grad != 0).sum().item()),
        d_table_head=ref.embed_fn.params.grad[:4096].numpy())
    print('wrote coslam_map_step.npz')


def nice(R=120, seed=3):
    """Reference ConvOnet (slam/models/conv_onet.py), stage 'color' -- the one stage that
    runs on CPU -- mapping and tracking losses, with gradients."""
    bound = np.array([[-2.0, 2.0], [-2.0, 2.0], [-2.0, 2.0]])
    torch.manual_seed(seed)
    ref = ref_harness.ref_conv_onet(bound)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for i, k in enumerate(sorted(ref.grid_c)):
            # regenerated from the seed by the tests (keeps the fixture small)
            gg = torch.Generator().manual_seed(1000 + i)
            ref.grid_c[k] = (torch.randn(ref.grid_c[k].shape, generator=gg) * 0.3
                             ).requires_grad_(True)
        for dec in (ref.decoder.middle_decoder, ref.decoder.fine_decoder,
                    ref.decoder.color_decoder):
            dec.embedder._B.mul_(0.2)
            for lin in list(dec.fc_c) + list(dec.pts_linears) + [dec.output_linear]:
                lin.bias.copy_(torch.randn(lin.bias.shape, generator=g) * 0.1)
    rays_o = ((torch.rand(R, 3, generator=g) - 0.5) * 0.6).requires_grad_(True)
    rays_d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g),
                                           dim=-1).requires_grad_(True)
    td = torch.rand(R, 1, generator=g) * 1.5 + 0.3
    td[3::7] = 0
    ts = torch.rand(R, 3, generator=g)
    blob = dict(bound=bound, rays_o=rays_o.detach().numpy(), rays_d=rays_d.detach().numpy(),
                target_s=ts.numpy(), target_d=td.numpy())
    sd = ref.decoder.state_dict()
    for k, v in sd.items():
        blob['dec.' + k] = v.numpy()
    for k, v in ref.grid_c.items():
        blob[k + '.shape'] = np.array(v.shape)
        blob[k + '.checksum'] = np.float64(v.detach().double().sum().item())
    for tag, is_mapping in (('map', True), ('trk', False)):
        for t in [rays_o, rays_d] + list(ref.grid_c.values()) + list(ref.decoder.parameters()):
            t.grad = None
        inp = dict(rays_o=rays_o, rays_d=rays_d, target_s=ts, target_d=td, stage='color')
        out = ref(inp)
        ld = ref.get_loss_dict(out, inp, is_mapping, 'color')
        sum(ld.values()).backward()
        blob[tag + '.rgb'] = out['rgb'].detach().numpy()
        blob[tag + '.depth'] = out['depth'].detach().numpy()
        blob[tag + '.uncertainty'] = out['uncertainty'].detach().numpy()
        blob[tag + '.losses'] = np.array([float(ld['depth_loss'].detach()),
                                          float(ld['rgb_loss'].detach())])
        blob[tag + '.d_rays_o'] = rays_o.grad.numpy().copy()
        blob[tag + '.d_rays_d'] = rays_d.grad.numpy().copy()
        gcg = ref.grid_c['grid_color'].grad
        blob[tag + '.d_grid_color_norm'] = np.float64(gcg.double().norm().item())
        blob[tag + '.d_grid_color_slice'] = gcg[0, :, 10:14, 10:14, 10:14].numpy().copy()
        blob[tag + '.d_grid_middle_norm'] = np.float64(
            ref.grid_c['grid_middle'].grad.double().norm().item())
        cd = ref.decoder.color_decoder
        blob[tag + '.d_B'] = cd.embedder._B.grad.numpy().copy()
        blob[tag + '.d_pts3_w'] = cd.pts_linears[3].weight.grad.numpy().copy()
        blob[tag + '.d_fcc0_w'] = cd.fc_c[0].weight.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, 'nice_color_step.npz'), **blob)
    print('wrote nice_color_step.npz')


def pointslam(R0=300, R=80, seed=11):
    """Reference ConvOnet2 (slam/models/conv_onet_pointslam.py) + NeuralPointCloud, stage
    'geometry', on CPU with the exact-kNN stand-in for faiss (oracle/ref_harness.py):
    point-adding, mapping and tracking losses with gradients."""
    torch.manual_seed(seed)
    ref = ref_harness.ref_conv_onet2()
    g = torch.Generator().manual_seed(seed)
    ro0 = torch.zeros(R0, 3)
    rd0 = torch.nn.functional.normalize(
        torch.randn(R0, 3, generator=g) * torch.tensor([0.3, 0.3, 0.05]) +
        torch.tensor([0, 0, -1.0]), dim=-1)
    d0 = torch.rand(R0, generator=g) * 0.5 + 1.5
    e = torch.zeros(0)
    ref.model_update(dict(batch_rays_o=ro0, batch_rays_d=rd0, batch_gt_depth=d0,
                          batch_gt_color=torch.rand(R0, 3, generator=g),
                          batch_dynamic_r=torch.full((R0,), 0.04),
                          batch_rays_o_grad=ro0[:0], batch_rays_d_grad=rd0[:0],
                          batch_gt_depth_grad=e, batch_gt_color_grad=torch.rand(0, 3),
                          batch_dynamic_r_grad=e))
    npc = ref.neural_point_cloud
    gd = ref.decoder.geo_decoder
    cd = ref.decoder.color_decoder
    with torch.no_grad():
        gd.embedder._B.mul_(0.05)
        for lin in list(gd.fc_c) + list(gd.pts_linears) + [gd.output_linear]:
            lin.bias.copy_(torch.randn(lin.bias.shape, generator=g) * 0.1)
        npc.geo_feats.mul_(5.0)
        # colour decoder: moderate sin() arguments, non-zero biases, wider activations so the
        # softplus(beta=100) kinks are exercised
        cd.embedder._B.mul_(0.05)
        cd.embedder_rel_pos._B.mul_(0.3)
        for lin in list(cd.pts_linears) + [cd.output_linear]:
            lin.bias.copy_(torch.randn(lin.bias.shape, generator=g) * 0.05)
        npc.col_feats.mul_(5.0)
    rays_o = (torch.randn(R, 3, generator=g) * 0.01).requires_grad_(True)
    rays_d = rd0[:R].clone().requires_grad_(True)
    td = (d0[:R] + torch.randn(R, generator=g) * 0.01).reshape(-1, 1)
    td[5::9] = 0
    radius = torch.rand(R, generator=g) * 0.08 + 0.04
    blob = dict(add_rays_d=rd0.numpy(), add_depth=d0.numpy(),
                cloud_pos=np.asarray(npc._cloud_pos, np.float32),
                geo_feats=npc.geo_feats.detach().numpy().copy(),
                rays_o=rays_o.detach().numpy(), rays_d=rays_d.detach().numpy(),
                target_d=td.numpy(), radius=radius.numpy())
    for k, v in gd.state_dict().items():
        blob['dec.' + k] = v.numpy().copy()
    for k, v in cd.state_dict().items():
        blob['cdec.' + k] = v.numpy().copy()
    blob['cdec.embedder._B'] = cd.embedder._B.numpy().copy()  # plain attribute, not in state_dict
    blob['col_feats'] = npc.col_feats.detach().numpy().copy()
    target_s = torch.rand(R, 3, generator=g)
    blob['target_s'] = target_s.numpy()
    for tag, is_mapping in (('map', True), ('trk', False)):
        for t in [rays_o, rays_d, npc.geo_feats] + list(gd.parameters()):
            t.grad = None
        inp = dict(rays_o=rays_o, rays_d=rays_d, target_d=td, target_s=torch.zeros(R, 3),
                   stage='geometry', batch_dynamic_r=radius)
        torch.manual_seed(777)  # the decoders' N(0, 0.01) features (Q6) = the draws below
        out = ref(inp)
        ld = ref.get_loss_dict(out, inp, is_mapping)
        ld['geo_loss'].backward()
        blob[tag + '.depth'] = out['depth'].detach().numpy()
        blob[tag + '.uncertainty'] = out['uncertainty'].detach().numpy()
        blob[tag + '.valid'] = out['valid_ray_mask'].numpy()
        blob[tag + '.loss'] = np.float32(ld['geo_loss'].item())
        blob[tag + '.d_geo_feats'] = npc.geo_feats.grad.numpy().copy()
        blob[tag + '.d_rays_o'] = rays_o.grad.numpy().copy()
        blob[tag + '.d_rays_d'] = rays_d.grad.numpy().copy()
    torch.manual_seed(777)  # geometry decoder draws first, then the colour decoder
    blob['rand_feat'] = torch.zeros(32).normal_(mean=0, std=0.01).numpy()
    blob['rand_feat_color'] = torch.zeros(32).normal_(mean=0, std=0.01).numpy()
    # stage 'color'
    for tag, is_mapping in (('cmap', True), ('ctrk', False)):
        ps = [rays_o, rays_d, npc.geo_feats, npc.col_feats] + list(cd.parameters())
        for t in ps:
            t.grad = None
        inp = dict(rays_o=rays_o, rays_d=rays_d, target_d=td, target_s=target_s, stage='color',
                   batch_dynamic_r=radius)
        torch.manual_seed(777)
        out = ref(inp)
        ld = ref.get_loss_dict(out, inp, is_mapping)
        sum(ld.values()).backward()
        blob[tag + '.depth'] = out['depth'].detach().numpy()
        blob[tag + '.rgb'] = out['rgb'].detach().numpy()
        blob[tag + '.losses'] = np.array([ld['geo_loss'].item(), ld['rgb_loss'].item()],
                                         np.float32)
        blob[tag + '.d_geo_feats'] = npc.geo_feats.grad.numpy().copy()
        blob[tag + '.d_col_feats'] = npc.col_feats.grad.numpy().copy()
        blob[tag + '.d_rays_o'] = rays_o.grad.numpy().copy()
        blob[tag + '.d_rays_d'] = rays_d.grad.numpy().copy()
        for k, v in cd.named_parameters():
            blob[tag + '.d_cdec.' + k] = v.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, 'pointslam_geo_step.npz'), **blob)
    print('wrote pointslam_geo_step.npz', npc.pts_num(), 'points')


if __name__ == '__main__':
    assert ref_harness.available(), 'needs /root/reference'
    which = sys.argv[1:] or ['coslam', 'nice', 'pointslam']
    for w in which:
        globals()[w]()
