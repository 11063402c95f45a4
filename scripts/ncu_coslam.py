import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from helpers import coslam_pair, make_rays
dev = torch.device('cuda:0')
import os
_, model = coslam_pair(dev, table_amp=1e-2, precision=int(os.environ.get("XRD_PREC", "0")))
R = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rays_o, rays_d, ts, td, noise = make_rays(R, seed=1)
w = model._weights(); tab = model.embed_fn.params
for _ in range(3):
    model._launch(rays_o.to(dev), rays_d.to(dev), tab, *w, ts.to(dev), td.to(dev), None, with_grads=True)
torch.cuda.synchronize()
