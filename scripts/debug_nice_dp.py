import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, torch.distributed as dist
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
from test_nice_gpu import nice_pair, rays
from helpers import rel_err
from xrdslam_b200.dp import MappingDataParallel
_, nm = nice_pair(dev)
R = 256
ro, rd, ts, td = rays(R, 17)
full = dict(rays_o=ro.to(dev), rays_d=rd.to(dev), target_s=ts.to(dev), target_d=td.to(dev), stage='color', is_mapping=True)
names = ['gm', 'gf', 'gc'] + [n for n, _ in nm.decoder.color_decoder.named_parameters()]
nparams = [nm.grids[k] for k in ('grid_middle', 'grid_fine', 'grid_color')] + list(nm.decoder.color_decoder.parameters())
def run(inp):
    for p in nparams: p.grad = None
    ld = nm.get_loss_dict(nm(inp), inp, True, 'color')
    sum(ld.values()).backward()
    torch.cuda.synchronize()
    return torch.stack([ld['depth_loss'].detach(), ld['rgb_loss'].detach()])
l_full = run(full)
ref = [p.grad.clone() for p in nparams]
# cross-rank check of the reference itself
r0 = [g.clone() for g in ref]
for g in r0: dist.broadcast(g, 0)
print(rank, 'ref vs rank0 ref', max(rel_err(a, b) for a, b in zip(ref, r0)), 'loss full', l_full.tolist(), flush=True)
ndp = MappingDataParallel(nparams); nm.dp = ndp
sl = ndp.shard(R)
part = {k: (v[sl] if torch.is_tensor(v) else v) for k, v in full.items()}
l_part = run(part)
print(rank, 'slice', sl, 'loss part', l_part.tolist(), 'grad None?', [n for n, p in zip(names, nparams) if p.grad is None], flush=True)
loss = l_part.clone()
ndp.all_reduce_grads(extra=[loss])
torch.cuda.synchronize()
errs = {n: round(rel_err(p.grad, g), 5) for n, p, g in zip(names, nparams, ref)}
print(rank, 'loss sum', loss.tolist(), 'errs', {k: v for k, v in errs.items() if v > 1e-4}, flush=True)
dist.destroy_process_group()
