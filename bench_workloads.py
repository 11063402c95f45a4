"""Workloads of bench.py for the BASELINE.json configs other than Co-SLAM (cfg 2, bench.py
itself): NICE-SLAM (cfg 1 / 4), Vox-Fusion (cfg 3), Point-SLAM (cfg 5).

One *step* = one mapping iteration of the reference loop body
(slam/algorithms/base_algorithm.py:255-273): zero_grad -> get_loss (pixel sampling + fused
render / loss / backward on the GPU) -> backward -> post_processing -> optimizer step, on a
640x480 synthetic RGB-D sequence, at the reference's default batch sizes
(slam/configs/input_config.py).  Every workload exposes

    step(i)            one iteration, frames already resident in HBM          -> `value`
    e2e(K)             Algorithm.optimize_update(K, frames, True) with the current frame's
                       images in HOST memory (upload inside the timed region) -> `e2e`
    kernel_roofline()  the dominant kernel / kernel chain, timed with CUDA events by the
                       library around its launches (xrd_debug_kernel_events)
    cpu_step()         the same iteration in the CPU oracle port (oracle/*.py)

oracle/ is imported ONLY by the cpu_* functions (bench.py's cpu_baseline / --impl reference
legs)."""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
W_IMG, H_IMG = 640, 480


def _frames(n, offset=(0., 0., 0.), rot_rep='quat', separate_LR=False):
    from xrdslam_b200.frame import Frame
    from xrdslam_b200.synthetic import make_sequence
    cam, poses, fr = make_sequence(n, offset=offset)
    return cam, [Frame(k, fr[k][0], fr[k][1], init_pose=poses[k], separate_LR=separate_LR,
                       rot_rep=rot_rep) for k in range(n)]


def _drop_device_images(frame):
    for k in ('_dev_depth', '_dev_rgb', '_ray_table'):
        frame.__dict__.pop(k, None)


class Workload:
    """Common driver around an Algorithm: the autograd iteration of optimize_update."""
    name = ''
    workload = ''
    metric = ''
    n_iters_schedule = 60      # n_iters the stage / lr schedules are laid out over
    dtype = 'f32'
    precision = 'fp32'
    map_info = None

    def __init__(self, dev, rank=0, world=1):
        self.dev, self.rank, self.world = dev, rank, world
        self.algo = None
        self.frames = None
        self.opt = None
        self.dp = None

    # -- subclass API
    def build(self):
        raise NotImplementedError

    def rays_per_step(self):
        raise NotImplementedError

    def launches_per_step(self):
        return None

    # -- shared
    def attach_dp(self, params):
        if self.world > 1:
            from xrdslam_b200.dp import MappingDataParallel
            self.dp = MappingDataParallel(params)
            self.dp.broadcast_params(0)
            self.algo.model.dp = self.dp

    def begin(self, n_iters):
        self.n_iters = n_iters
        self.opt = self.algo.setup_optimizers(n_iters, self.frames, True)

    def step(self, i):
        a = self.algo
        self.opt.zero_grad_all()
        loss = a.get_loss(self.frames, True, i % self.n_iters, self.n_iters)
        loss.backward()
        if self.dp is not None:
            self.dp.all_reduce_grads()
        a.post_processing(i, True)
        self.opt.optimizer_step_all(step=i)
        self.opt.scheduler_step_all()
        return loss

    def e2e(self, K):
        """K iterations through the plugin call; returns (seconds, h2d_bytes, d2h_bytes) per
        call.  The current frame arrives as host arrays (what the dataset hands the
        pipeline): its upload is inside the timed region; the call ends with a D2H read."""
        cur = self.frames[-1]
        _drop_device_images(cur)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        self.algo.optimize_update(K, self.frames, True)
        chk = float(self.checksum())  # D2H read of the step's result
        dt = time.perf_counter() - t0
        h2d = W_IMG * H_IMG * 4 * 4  # depth + rgb fp32 of the current frame
        return dt, h2d, 4, chk

    def checksum(self):
        p = next(iter(self.algo.model.parameters()))
        return p.detach().float().abs().sum()

    def tracking(self, n_frames=10):
        a = self.algo
        cur = self.frames[-1]
        n_it = a.config.tracking_n_iters
        for _ in range(2):
            a.optimize_update(n_it, [cur], False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_frames):
            a.optimize_update(n_it, [cur], False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return {'tracking_iters_per_s': n_frames * n_it / dt,
                'tracking_frames_per_s': n_frames / dt, 'tracking_iters_per_frame': n_it,
                'tracking_rays': a.config.tracking_sample}


# ------------------------------------------------------------------ NICE-SLAM ---
class NiceWorkload(Workload):
    name = 'nice'
    metric = ('rays/s (nice-slam mapping iteration: sample+bbox filter+3-level dense-grid '
              'trilerp+5x32 MLPs+composite+loss+backward+masked Adam, 640x480 synthetic RGB-D)')
    BOUND = [[-3.2, 3.2], [-4.2, 2.7], [-2.2, 2.7]]
    workload = ('nice-slam 3-level dense grids (0.32/0.16/0.16 m) + 5x32 Fourier MLPs, '
                '640x480 synthetic room, mapping iteration over a 5-frame window x 200 '
                'rays = 1000 rays per GPU, 32+16 samples/ray, stage schedule '
                'middle 40% / fine 20% / color 40% of 60 iterations')

    def build(self):
        from xrdslam_b200.nice_slam import NiceSLAMConfig
        cam, frs = _frames(5)
        algo = NiceSLAMConfig(mapping_bound=self.BOUND).setup(camera=cam, device=self.dev)
        for f in frs[:4]:
            algo.add_keyframe(f)
        algo.pre_precessing(frs[-1], True)
        algo.set_initialized()
        self.algo, self.frames = algo, frs
        self.attach_dp([p for g in algo.model.get_param_groups().values() for p in g])

    def rays_per_step(self):
        return 1000

    def kernel_roofline(self, k_ms, peak_gbs):
        # the bracketed launch is the LAST k_decoder_fwd of the step (level = stage); in stage
        # 'color' it gathers 8 corners x 32 channels x 4 B = 1024 B per sample point
        P = self.rays_per_step() * 48
        by = P * 1024
        ach = by / (k_ms * 1e-3) / 1e9
        return {'bound': 'hbm', 'kernel': 'xrd::nice::k_decoder_fwd (colour level)',
                'achieved': ach, 'peak': peak_gbs, 'unit': 'GB/s', 'frac': ach / peak_gbs,
                'traffic': None, 'kernel_ms': k_ms, 'algorithmic_bytes_per_launch': by,
                'note': 'upper bound on P (rays dropped by the bbox pre-filter are not counted out)'}

    def roofline_steps(self):
        # stage 'color' iterations
        return [int(0.7 * self.n_iters) + j for j in range(8)]

    def cpu_step_factory(self, R=None):
        sys.path.insert(0, ROOT)
        from oracle.nice import NiceOracle
        from xrdslam_b200.common import get_samples
        cam, frs = _frames(5)
        ora = NiceOracle(np.array(self.BOUND))
        params = list(ora.grids.values()) + [p for d in (ora.color,) for p in d.parameters()]
        opt = torch.optim.Adam([{'params': params, 'lr': 5e-3}])
        n = (R or 1000) // len(frs)
        stages = ['middle'] * 2 + ['fine'] + ['color'] * 2  # the 40/20/40 schedule

        def step(i=[0]):
            ro, rd, td, ts = [], [], [], []
            for f in frs:
                o, d, dep, col = get_samples(cam, n, f.get_pose().detach(),
                                             torch.as_tensor(f.depth), torch.as_tensor(f.rgb),
                                             device='cpu')
                ro.append(o); rd.append(d); td.append(dep); ts.append(col)
            opt.zero_grad(set_to_none=True)
            _, _, tot = ora.step(torch.cat(ro).float(), torch.cat(rd).float(), torch.cat(ts).float(),
                                 torch.cat(td).float().reshape(-1, 1), True, stages[i[0] % 5])
            tot.backward()
            opt.step()
            i[0] += 1
            return float(tot.detach())
        return step, n * len(frs), 'oracle/nice.py torch-CPU port (sampling + 3-level render + loss + backward + Adam)'


# ----------------------------------------------------------------- Vox-Fusion ---
class VoxWorkload(Workload):
    name = 'vox'
    metric = ('rays/s (vox-fusion mapping iteration: sample+octree ray march+inverse-CDF sampling+'
              'trilerp+128-wide SDF/colour MLP+composite+loss+backward+Adam, 640x480 synthetic RGB-D)')
    N_FRAMES = 5
    workload = ('vox-fusion sparse octree (0.2 m voxels, 16-d vertex embeddings) + '
                '16-128-128-129 / 144-128-3 decoder, 640x480 synthetic room, mapping iteration '
                'over a 5-frame window x 1024 rays = 5120 rays per GPU, step 0.01 m '
                'inverse-CDF samples')

    def build(self):
        from xrdslam_b200.voxfusion import VoxFusionConfig
        cam, frs = _frames(self.N_FRAMES, offset=(10., 10., 10.))
        algo = VoxFusionConfig().setup(camera=cam, device=self.dev)
        for f in frs:
            algo.create_voxels(f)
        algo.set_initialized()
        self.algo, self.frames = algo, frs
        self.map_info = {'octree_nodes': int(algo.model.map_states['voxel_center_xyz'].shape[0])}
        self.attach_dp(list(algo.model.decoder.parameters()) + [algo.model.embeddings])

    def rays_per_step(self):
        return self.N_FRAMES * 1024

    def kernel_roofline(self, k_ms, peak_tf, n_points):
        # bracketed: the forward decoder chain of one step (gather + 5 GEMMs + epilogues)
        fl = n_points * 2.0 * (16 * 128 + 128 * 128 + 128 * 129 + 144 * 128 + 128 * 3)
        ach = fl / (k_ms * 1e-3) / 1e12
        return {'bound': 'tensor', 'kernel': 'vox decoder forward chain (k_vox_gather + 5 GEMMs)',
                'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf,
                'traffic': None, 'kernel_ms': k_ms, 'algorithmic_flops_per_launch': fl,
                'points': n_points,
                'note': 'TF32 3x-split GEMMs against the measured dense bf16 peak'}

    def cpu_step_factory(self, R=None):
        sys.path.insert(0, ROOT)
        from oracle.voxfusion import VoxOracle
        from xrdslam_b200.common import get_samples
        from xrdslam_b200.voxfusion import VoxFusionConfig
        cam, frs = _frames(self.N_FRAMES, offset=(10., 10., 10.))
        # the octree itself is built by the product's host C++ (csrc/octree.cpp); the oracle
        # renders it
        algo = VoxFusionConfig().setup(camera=cam, device=self.dev)
        for f in frs:
            algo.create_voxels(f)
        ora = VoxOracle()
        ora.set_map(*algo.model.export_octree())
        opt = torch.optim.Adam([{'params': list(ora.parameters()), 'lr': 5e-3}])
        n = (R or self.rays_per_step()) // len(frs)

        def step():
            ro, rd, td, ts = [], [], [], []
            for f in frs:
                o, d, dep, col = get_samples(cam, n, f.get_pose().detach(),
                                             torch.as_tensor(f.depth), torch.as_tensor(f.rgb),
                                             device='cpu')
                ro.append(o); rd.append(d); td.append(dep); ts.append(col)
            ro, rd = torch.cat(ro).float(), torch.cat(rd).float()
            opt.zero_grad(set_to_none=True)
            marched = ora.march(ro, rd, lambda s: torch.rand(s).clamp(0.001, 0.999))
            _, ld = ora.render(ro, rd, torch.cat(ts).float(), torch.cat(td).float().reshape(-1, 1),
                               marched)
            tot = sum(ld.values())
            tot.backward()
            opt.step()
            return float(tot.detach())
        return step, n * len(frs), 'oracle/voxfusion.py CPU port (python octree DFS + inverse-CDF sampling, torch render + loss + backward + Adam)'


# ----------------------------------------------------------------- Point-SLAM ---
class PointWorkload(Workload):
    name = 'point'
    metric = ('rays/s (point-slam mapping iteration: sample+exact 8-NN over the neural point '
              'cloud+feature interpolation+geometry/colour MLPs+composite+loss+backward+Adam, '
              '640x480 synthetic RGB-D)')
    n_iters_schedule = 300
    workload = ('point-slam neural point cloud (32-d geometry + 32-d colour features, dynamic '
                'radii), exact radius-limited 8-NN, 5x32 geometry MLP + 5x128 colour MLP, 640x480 '
                'synthetic room, mapping iteration over a 3-frame window = 5000 rays per GPU x 5 '
                'surface samples, stage schedule geometry 40% / color 60% of 300 iterations')

    def build(self):
        from xrdslam_b200.point_slam import PointSLAMConfig
        cam, frs = _frames(3, rot_rep='axis_angle', separate_LR=True)
        algo = PointSLAMConfig().setup(camera=cam, device=self.dev)
        for f in frs:
            algo.pre_precessing(f, True)
        algo.set_initialized()
        self.algo, self.frames = algo, frs
        self.map_info = {'neural_points': algo.model.neural_point_cloud.pts_num()}
        npc = algo.model.neural_point_cloud
        self.attach_dp([npc.geo_feats, npc.col_feats] +
                       list(algo.model.decoder.color_decoder.parameters()))

    def rays_per_step(self):
        return 5000

    def checksum(self):
        return self.algo.model.neural_point_cloud.geo_feats.detach().abs().sum()

    def kernel_roofline(self, k_ms, peak_gbs):
        # bracketed: k_knn -- per query 27 hash-grid cells are scanned; algorithmic bytes are
        # the 12-byte positions of the candidates (~n_points * (0.24 m)^3 / volume per query)
        P = self.rays_per_step() * 5
        return {'bound': 'hbm', 'kernel': 'xrd::point::k_knn', 'achieved': None, 'peak': peak_gbs,
                'unit': 'GB/s', 'frac': None, 'traffic': None, 'kernel_ms': k_ms,
                'queries': P, 'note': 'latency-bound irregular gather; queries/s reported',
                'queries_per_s': P / (k_ms * 1e-3)}

    def cpu_step_factory(self, R=None):
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        from oracle.pointslam import PointOracle
        from xrdslam_b200.common import get_samples
        cam, frs = _frames(3, rot_rep='axis_angle', separate_LR=True)
        npc = self.algo.model.neural_point_cloud
        ora = PointOracle(seed=3)
        ora.set_cloud(npc.cloud_pos().cpu(), npc.geo_feats.detach().cpu(),
                      col_feats=npc.col_feats.detach().cpu())
        params = [ora.geo_feats, ora.col_feats] + list(ora.col.parameters())
        opt = torch.optim.Adam([{'params': [p for p in params if p.requires_grad], 'lr': 5e-3}])
        n = (R or 5000) // len(frs)
        rf, rfc = torch.randn(32) * 0.01, torch.randn(32) * 0.01

        def step(i=[0]):
            ro, rd, td, ts = [], [], [], []
            for f in frs:
                o, d, dep, col = get_samples(cam, n, f.get_pose().detach(),
                                             torch.as_tensor(f.depth), torch.as_tensor(f.rgb),
                                             device='cpu', depth_filter=True)
                ro.append(o); rd.append(d); td.append(dep); ts.append(col)
            ro, rd = torch.cat(ro).float(), torch.cat(rd).float()
            td_, ts_ = torch.cat(td).float().reshape(-1, 1), torch.cat(ts).float()
            stage = 'geometry' if i[0] % 5 < 2 else 'color'
            opt.zero_grad(set_to_none=True)
            out = ora.render(ro, rd, td_, torch.full((ro.shape[0],), 0.08), rf, stage, rfc)
            tot = sum(ora.loss_dict(out, td_, ts_, True).values()) if stage == 'color' \
                else ora.loss(out, td_, True)
            tot.backward()
            opt.step()
            i[0] += 1
            return float(tot.detach())
        return step, n * len(frs), 'oracle/pointslam.py torch-CPU port (exact 8-NN, both decoders, loss, backward, Adam)'


WORKLOADS = {'nice': NiceWorkload, 'vox': VoxWorkload, 'point': PointWorkload}
