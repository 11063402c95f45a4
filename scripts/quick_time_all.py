"""Per-iteration timings of the four algorithms on the 640x480 synthetic sequence (GPU box).
iteration = Algorithm.get_loss (sampling + fused step) + backward + optimizer step, as in
Algorithm.optimize_update; device ms by CUDA events, host ms by perf_counter."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from xrdslam_b200.frame import Frame
from xrdslam_b200.synthetic import make_sequence

dev = torch.device('cuda:0')
from xrdslam_b200 import _cabi
if 'XRD_GEMM_MODE' in os.environ:
    _cabi.lib().xrd_debug_gemm_mode(int(os.environ['XRD_GEMM_MODE']))
ONLY = os.environ.get('QT_ONLY', '')


def frames(n, offset=(0., 0., 0.), rot_rep='quat', separate_LR=False):
    cam, poses, fr = make_sequence(n, offset=offset)
    return cam, [Frame(k, fr[k][0], fr[k][1], init_pose=poses[k], separate_LR=separate_LR,
                       rot_rep=rot_rep) for k in range(n)]


N_IT = int(os.environ.get('QT_N', '30'))
N_WARM = int(os.environ.get('QT_WARM', '5'))


def time_iters(algo, frs, is_mapping, n=None, warm=None, step0=0, n_iters=100):
    n = n or N_IT
    warm = N_WARM if warm is None else warm
    torch.cuda.nvtx.range_push(f'{type(algo).__name__}:{"map" if is_mapping else "trk"}:{step0}')
    opt = algo.setup_optimizers(n_iters, frs, is_mapping)
    rays = []
    def it(i):
        opt.zero_grad_all()
        loss = algo.get_loss(frs, is_mapping, step0 + i, n_iters)
        loss.backward()
        algo.post_processing(i, is_mapping)
        opt.optimizer_step_all(step=i)
        return loss
    for i in range(warm):
        it(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(n):
        it(i)
    e1.record(); torch.cuda.synchronize()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.nvtx.range_pop()
    if os.environ.get('QT_PROFILE'):
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for i in range(5):
                it(i)
            torch.cuda.synchronize()
        ev = [e for e in prof.key_averages() if e.device_type.name == 'CUDA' or e.self_device_time_total > 0]
        tot = sum(e.self_device_time_total for e in ev)
        print(f'--- {type(algo).__name__} mapping={is_mapping} step0={step0}: GPU busy {tot / 5 / 1e3:.3f} ms/iter, '
              f'{sum(e.count for e in ev) / 5:.0f} kernels/iter', file=sys.stderr)
        for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:10]:
            print(f'   {e.self_device_time_total / 5:9.1f} us/iter x{e.count / 5:5.1f}  {e.key[:90]}', file=sys.stderr)
    return {'ms_per_iter_wall': round(host, 3), 'ms_per_iter_device': round(e0.elapsed_time(e1) / n, 3)}


out = {}
torch.manual_seed(0); np.random.seed(0)


def run_nice():
    # NICE-SLAM (cfg 4): 5 frames x 200 rays mapping, 200 rays tracking
    from xrdslam_b200.nice_slam import NiceSLAMConfig
    cam, frs = frames(5)
    algo = NiceSLAMConfig(mapping_bound=[[-3.2, 3.2], [-4.2, 2.7], [-2.2, 2.7]]).setup(camera=cam, device=dev)
    for f in frs[:4]:
        algo.add_keyframe(f)
    algo.pre_precessing(frs[-1], True)
    algo.set_initialized()
    for name, s0 in (('middle', 0), ('fine', 50), ('color', 90)):
        out['nice_map_' + name] = dict(rays=1000, **time_iters(algo, frs, True, step0=s0))
    out['nice_track'] = dict(rays=200, **time_iters(algo, frs[-1:], False))


def run_vox():
    from xrdslam_b200.voxfusion import VoxFusionConfig
    cam, frs = frames(3, offset=(10., 10., 10.))
    algo = VoxFusionConfig().setup(camera=cam, device=dev)
    for f in frs:
        algo.create_voxels(f)
    algo.set_initialized()
    out['vox_nodes'] = int(algo.model.map_states['voxel_center_xyz'].shape[0])
    out['vox_map'] = dict(rays=3 * 1024, **time_iters(algo, frs, True))
    out['vox_track'] = dict(rays=1024, **time_iters(algo, frs[-1:], False))


def run_point():
    from xrdslam_b200.point_slam import PointSLAMConfig
    cam, frs = frames(3, rot_rep='axis_angle', separate_LR=True)
    algo = PointSLAMConfig().setup(camera=cam, device=dev)
    for f in frs:
        algo.pre_precessing(f, True)
    algo.set_initialized()
    out['point_pts'] = algo.model.neural_point_cloud.pts_num()
    out['point_map_geometry'] = dict(rays=5000, **time_iters(algo, frs, True, step0=0))
    out['point_map_color'] = dict(rays=5000, **time_iters(algo, frs, True, step0=90))
    out['point_track'] = dict(rays=1500, **time_iters(algo, frs[-1:], False))


for name, fn in (('nice', run_nice), ('vox', run_vox), ('point', run_point)):
    if not ONLY or name in ONLY:
        fn()
print(json.dumps(out))
