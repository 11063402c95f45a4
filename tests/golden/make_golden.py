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


if __name__ == '__main__':
    assert ref_harness.available(), 'needs /root/reference'
    coslam()
