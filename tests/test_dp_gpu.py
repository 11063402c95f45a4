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


def _worker_graph(rank, world, port, ret):
    """CUDA-graph mapping iteration, rays sharded over 2 ranks (ONE captured graph holding the 2
    NCCL all-reduces) == the single-GPU captured iteration on the whole batch."""
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    import bench
    from helpers import rel_err
    from xrdslam_b200.dp import MappingDataParallel

    def build(mapping_sample, with_dp):
        algo, kfs, cur = bench.build_algorithm(dev, seed=5)  # same seed: identical replicas
        algo.config.mapping_sample = mapping_sample
        algo.config.min_sample_pixels = mapping_sample
        algo.model.config.training_perturb = 0  # perturbation noise is keyed by the local ray index
        frames = kfs + [cur]
        if with_dp:
            params = [algo.model.embed_fn.params] + list(algo.model.decoder.parameters())
            algo.model.dp = MappingDataParallel(params)
        algo.setup_optimizers(4, frames, True)
        sess = algo.mapping_session(frames)
        sess._gen = torch.Generator().manual_seed(977)
        sess.begin(frames)
        return algo, frames, sess

    # the global batch: 2 x 4096 rays staged once by rank-identical host sampling
    algo_s, frames_s, sess_s = build(4096, False)
    assert sess_s.world == 1 and sess_s.R == 8192
    rows, ids = sess_s.make_resident_batch(frames_s)
    loss_s = float(sess_s.step_resident(0, (rows, ids)))
    flat_s = sess_s.flat.clone()
    algo_d, frames_d, sess_d = build(2048, True)
    assert sess_d.world == 2 and sess_d.R == 4096
    # one graph holding both NCCL all-reduces (3 segments only if NCCL refused capture)
    assert len(sess_d.graphs) == (1 if sess_d.single_graph else 3)
    sl = slice(rank * 4096, (rank + 1) * 4096)
    loss_d = float(sess_d.step_resident(0, (rows[sl].contiguous(), ids[sl].contiguous())))
    out = {'loss': abs(loss_d - loss_s) / abs(loss_s)}
    # same bucket layout: [table | decoder | pose grads | losses]; compare slot by slot
    for name, a, b in (('table', sess_d.grads[algo_d.model.embed_fn.params],
                        sess_s.grads[algo_s.model.embed_fn.params]),
                       ('d_rot', sess_d.d_rot_it, sess_s.d_rot_it),
                       ('d_trans', sess_d.d_trans_it, sess_s.d_trans_it)):
        out[name] = rel_err(a, b)
    out['decoder'] = max(rel_err(sess_d.grads[p], sess_s.grads[q]) for p, q in
                         zip(algo_d.model.decoder.parameters(), algo_s.model.decoder.parameters()))
    # after the (replicated) Adam step the replicas still agree with the single-GPU model
    out['param'] = max(float((p.detach() - q.detach()).abs().max()) for p, q in
                       zip(algo_d.model.decoder.parameters(), algo_s.model.decoder.parameters()))
    out['single_graph'] = bool(sess_d.single_graph)
    ret[rank] = out
    sess_d.release()  # drop the captured NCCL nodes before the communicator goes away
    sess_s.release()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_graph_mapping_equals_single_gpu():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_graph, args=(world, port, ret), nprocs=world, join=True)
    print(dict(ret))
    for r in range(world):
        o = ret[r]
        assert o['loss'] < 1e-5 and o['table'] < 1e-4 and o['decoder'] < 1e-4, o
        assert o['d_rot'] < 1e-3 and o['d_trans'] < 1e-3 and o['param'] < 1e-4, o
