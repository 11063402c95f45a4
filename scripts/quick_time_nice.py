import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from xrdslam_b200.camera import Camera
from xrdslam_b200.conv_onet import ConvOnetConfig
dev = torch.device('cuda:0')
bound = np.array([[-5.5, 5.9], [-6.7, 5.4], [-4.7, 5.3]])
model = ConvOnetConfig(mapping_frustum_feature_selection=False).setup(
    camera=Camera(320., 320., 319.5, 239.5, 640, 480), bounding_box=bound).to(dev)
for R in [200, 1000, 8192]:
    g = torch.Generator().manual_seed(R)
    ro = ((torch.rand(R, 3, generator=g) - 0.5) * 2).to(dev)
    rd = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).to(dev)
    td = (torch.rand(R, 1, generator=g) * 3 + 0.5).to(dev); ts = torch.rand(R, 3, generator=g).to(dev)
    for stage in ['middle', 'fine', 'color']:
        def run():
            return model._launch(stage, True, ro, rd, ts, td, True, True, (True, True, True), True)
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f'NICE R={R} stage={stage}: fwd+bwd {ms*1e3:.0f} us ({R/ms/1e3:.3f} Mrays/s)')
