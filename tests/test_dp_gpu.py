"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): mapping rays sharded over 2 ranks +
one NCCL all-reduce of the flat gradient bucket == the single-GPU gradient of the whole
batch, for Co-SLAM (global loss normalisers via the sample/render phase split) and
NICE-SLAM (global max depth)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, ret):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    from helpers import coslam_pair, make_rays, rel_err
    from xrdslam_b200.dp import MappingDataParallel
    out = {}
    # ---------------- Co-SLAM ------------------------------------------------
    _, model = coslam_pair(dev)
    R = 514
    ro, rd, ts, td, noise = make_rays(R, seed=3)
    full = dict(rays_o=ro.to(dev), rays_d=rd.to(dev), target_s=ts.to(dev), target_d=td.to(dev),
                first=True, noise=noise.to(dev))
    params = [model.embed_fn.params] + list(model.decoder.parameters())
    ld = model.get_loss_dict(model(full), full, True, 0)
    sum(ld.values()).backward()
    ref = [p.grad.clone() for p in params]
    ref_loss = torch.stack([ld[k].detach() for k in ('rgb_loss', 'depth_loss', 'sdf_loss', 'fs_loss')])
    for p in params:
        p.grad = None
    dp = MappingDataParallel(params)
    model.dp = dp
    sl = dp.shard(R)
    assert (sl.stop - sl.start) * world == R
    part = {k: (v[sl] if torch.is_tensor(v) else v) for k, v in full.items()}
    ld = model.get_loss_dict(model(part), part, True, 0)
    sum(ld.values()).backward()
    loss = torch.stack([ld[k].detach() for k in ('rgb_loss', 'depth_loss', 'sdf_loss', 'fs_loss')])
    dp.all_reduce_grads(extra=[loss])
    out['coslam_grad'] = max(rel_err(p.grad, g) for p, g in zip(params, ref))
    out['coslam_loss'] = float((loss - ref_loss).abs().max() / ref_loss.abs().max())
    # ---------------- NICE-SLAM ----------------------------------------------
    from test_nice_gpu import nice_pair, rays
    torch.manual_seed(0)  # nn.Linear default inits draw from the global generator
    _, nm = nice_pair(dev)
    R = 256
    ro, rd, ts, td = rays(R, 17)
    full = dict(rays_o=ro.to(dev), rays_d=rd.to(dev), target_s=ts.to(dev), target_d=td.to(dev),
                stage='color', is_mapping=True)
    nparams = [nm.grids[k] for k in ('grid_middle', 'grid_fine', 'grid_color')] + \
        list(nm.decoder.color_decoder.parameters())
    ld = nm.get_loss_dict(nm(full), full, True, 'color')
    sum(ld.values()).backward()
    ref = [p.grad.clone() for p in nparams]
    ref_loss = torch.stack([ld['depth_loss'].detach(), ld['rgb_loss'].detach()])
    for p in nparams:
        p.grad = None
    ndp = MappingDataParallel(nparams)
    nm.dp = ndp
    sl = ndp.shard(R)
    part = {k: (v[sl] if torch.is_tensor(v) else v) for k, v in full.items()}
    ld = nm.get_loss_dict(nm(part), part, True, 'color')
    sum(ld.values()).backward()
    loss = torch.stack([ld['depth_loss'].detach(), ld['rgb_loss'].detach()])
    ndp.all_reduce_grads(extra=[loss])
    out['nice_grad'] = max(rel_err(p.grad, g) for p, g in zip(nparams, ref))
    out['nice_loss'] = float((loss - ref_loss).abs().max() / ref_loss.abs().max())
    ret[rank] = out
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_mapping_equals_single_gpu():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    print(dict(ret))
    for r in range(world):
        o = ret[r]
        assert o['coslam_grad'] < 1e-4 and o['coslam_loss'] < 1e-5, o
        assert o['nice_grad'] < 1e-4 and o['nice_loss'] < 1e-5, o
