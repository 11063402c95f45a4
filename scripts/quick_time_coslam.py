import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from helpers import coslam_pair, make_rays
dev = torch.device('cuda:0')
import os
_, model = coslam_pair(dev, table_amp=1e-2, precision=int(os.environ.get("XRD_PREC", "0")))
for R in [1024, 4096, 16384, 65536]:
    for nr in ([0] if len(sys.argv) < 2 else [int(a) for a in sys.argv[1:]]):
        model.config.rays_per_tile = nr  # 0 auto, -1 tile kernel, -2 grouped kernel
        rays_o, rays_d, ts, td, noise = make_rays(R, seed=1)
        inp = dict(rays_o=rays_o.to(dev), rays_d=rays_d.to(dev), target_s=ts.to(dev), target_d=td.to(dev), first=True)
        w = model._weights()
        tab = model.embed_fn.params
        def run(grads=True):
            return model._launch(inp['rays_o'], inp['rays_d'], tab, *w, inp['target_s'], inp['target_d'], None, with_grads=grads)
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        e0.record()
        for _ in range(n): run(False)
        e1.record(); torch.cuda.synchronize()
        msf = e0.elapsed_time(e1) / n
        print(f'R={R} NR={nr}: fwd+bwd {ms*1e3:.1f} us  ({R/ms/1e3:.2f} Mrays/s, {R*88064/ms/1e6:.1f} GB/s algorithmic)   fwd-only {msf*1e3:.1f} us')
