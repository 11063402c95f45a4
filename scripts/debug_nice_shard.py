import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from test_nice_gpu import nice_pair, rays
from helpers import rel_err
dev = torch.device('cuda:0')
_, nm = nice_pair(dev)
R = 256
ro, rd, ts, td = rays(R, 17)
full = dict(rays_o=ro.to(dev), rays_d=rd.to(dev), target_s=ts.to(dev), target_d=td.to(dev), stage='color', is_mapping=True)
nparams = [nm.grids[k] for k in ('grid_middle', 'grid_fine', 'grid_color')] + list(nm.decoder.color_decoder.parameters())
def run(inp):
    for p in nparams: p.grad = None
    ld = nm.get_loss_dict(nm(inp), inp, True, 'color')
    sum(ld.values()).backward()
    torch.cuda.synchronize()
    return [p.grad.clone() for p in nparams], torch.stack([ld['depth_loss'].detach(), ld['rgb_loss'].detach()])
g1, l1 = run(full)
g2, l2 = run(full)
print('repeat full: grad', max(rel_err(a, b) for a, b in zip(g1, g2)), 'loss', (l1 - l2).abs().max().item())
class FakeDP:
    world = 2
    def __init__(s, m): s.m = m
    def all_reduce_max(s, t): t.fill_(s.m); return t
nm.dp = FakeDP(float(td.max()))
parts = []
for sl in (slice(0, 128), slice(128, 256)):
    part = {k: (v[sl] if torch.is_tensor(v) else v) for k, v in full.items()}
    parts.append(run(part))
gs = [a + b for a, b in zip(parts[0][0], parts[1][0])]
ls = parts[0][1] + parts[1][1]
print('shard-sum vs full: grad', [round(rel_err(a, b), 6) for a, b in zip(gs, g1)][:6], 'loss', ls.tolist(), l1.tolist())
