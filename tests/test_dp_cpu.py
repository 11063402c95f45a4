"""Host logic of the data-parallel mapping path on CPU: world_size-2 gloo process group
(no GPU): ray sharding, flat-bucket gradient all-reduce, scalar reductions."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from xrdslam_b200.dp import MappingDataParallel
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(1000)), torch.nn.Parameter(torch.randn(32, 80)),
              torch.nn.Parameter(torch.randn(3))]
    dp = MappingDataParallel(params)
    assert dp.world == world and dp.rank == rank
    # shards tile the batch exactly
    n = 4097
    sl = dp.shard(n)
    lens = [torch.tensor([sl.stop - sl.start])]
    tot = dp.all_reduce_sum(lens[0].clone())
    assert int(tot) == n
    # broadcast makes replicas identical
    with torch.no_grad():
        for p in params:
            p.add_(rank)
    dp.broadcast_params(0)
    ref = [p.detach().clone() for p in params]
    # per-rank gradients g_r = (rank+1) * x ; all-reduce -> sum_r (r+1) x, in ONE bucket
    for p in params[:2]:
        p.grad = (rank + 1.0) * torch.ones_like(p)
    params[2].grad = None  # parameters without a gradient are skipped
    loss = torch.tensor([float(rank + 1)])
    dp.all_reduce_grads(extra=[loss])
    s = sum(r + 1.0 for r in range(world))
    ok = all(torch.allclose(p.grad, s * torch.ones_like(p)) for p in params[:2])
    ok = ok and float(loss) == s and params[2].grad is None
    ok = ok and all(torch.equal(a, b.detach()) for a, b in zip(ref, params))
    mx = dp.all_reduce_max(torch.tensor([float(rank)]))
    ok = ok and float(mx) == world - 1
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_dp_host_logic_gloo_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
