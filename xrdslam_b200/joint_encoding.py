"""Co-SLAM model behind the reference's ``Model`` plugin surface, B200-native.

Host-side mirror of slam/models/joint_encoding.py (reference @ f0366f20): same
class name, same config fields, same ``forward / get_outputs / get_loss_dict /
get_param_groups / smoothness / query_sdf / query_color / query_fn / color_func``
signatures, same parameter-group names (``decoder``, ``embed_fn``) and the same
state_dict keys (``embed_fn.params``, ``decoder.sdf_net.model.{0,2}.weight``,
``decoder.color_net.model.{0,2}.weight``).

Nothing is computed in PyTorch: ``forward`` enqueues the fused sm_100a kernels of
``xrd_coslam_step`` through the C-ABI (forward + loss + full backward in one
launch when targets are present and grad is enabled) and ``get_loss_dict``
hands the already-computed loss terms back, wired into autograd so that
``loss.backward()`` (slam/algorithms/base_algorithm.py:266) delivers the
in-kernel gradients to the hash table, the four weight matrices and -- through
``rays_o / rays_d`` -- to the pose.  No CPU / PyTorch fallback exists.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Type, Union

import numpy as np
import torch
from torch import nn
from torch.nn import Parameter

from . import _cabi
from ._cabi import (XrdCoslamCfg, XrdCoslamGrads, XrdCoslamMlp, XrdCoslamOut,
                    XrdHashGrid, XrdRays, check, ptr)
from .base_model import Model, ModelConfig, scale_grads, upstream_scale


@dataclass
class JointEncodingConfig(ModelConfig):
    """slam/models/joint_encoding.py:17-66 (field names and defaults kept)."""
    _target: Type = field(default_factory=lambda: JointEncoding)
    voxel_sdf: float = 0.02
    voxel_color: float = 0.08
    enc: str = 'HashGrid'
    pos_enc: str = 'OneBlob'
    pos_nbins: int = 16
    hashsize: int = 16
    oneGrid: bool = True
    geo_feat_dim: int = 15
    hidden_dim: int = 32
    num_layers: int = 2
    num_layers_color: int = 2
    hidden_dim_color: int = 32
    tcnn_network: bool = False
    tcnn_encoding: bool = True
    trainging_rgb_weight: float = 5.0
    trainging_depth_weight: float = 0.1
    trainging_sdf_weight: float = 1000
    trainging_fs_weight: float = 10
    trainging_smooth_weight: float = 0.000001
    trainging_smooth_pts: int = 32
    trainging_smooth_vox: float = 0.1
    trainging_smooth_margin: float = 0.05
    training_n_samples: int = 256
    training_n_sample_d: int = 32
    training_range_d: float = 0.1
    training_n_range_d: int = 11
    training_n_importance: int = 0
    training_perturb: int = 1
    training_white_bkgd: bool = False
    training_trunc: float = 0.1
    training_rgb_missing: float = 0.05
    data_sc_factor: int = 1
    data_translation: int = 0
    cam_near: float = 0.0
    cam_far: float = 5.0
    cam_depth_trunc: float = 100.0
    mesh_render_color: bool = False
    # --- B200 path knobs (not in the reference) ---
    seed: int = 0  # Philox seed for in-kernel jitter when no noise is passed
    strict_loss_grad: bool = False  # verify upstream d(total)/d(term) == 1
    rays_per_tile: int = 0   # 0 automatic, -1 tile kernel (k_fused), -2 grouped kernel (k_fused_g)
    # decoder GEMMs on tensor cores: 0 = 3xTF32 everywhere (fp32-level parity),
    # 1 = 3xTF32 forward + TF32 backward, 2 = TF32 everywhere
    precision: int = 0


class HashGridParams(nn.Module):
    """Stands in for ``tcnn.Encoding('HashGrid', dtype=float)``: owns the flat
    fp32 ``params`` tensor in tcnn's layout (encodings_coslam.py:39-53)."""
    def __init__(self, n_levels, log2_hashmap_size, base_resolution,
                 per_level_scale, seed=1337):
        super().__init__()
        g = XrdHashGrid()
        check('xrd_hashgrid_layout',
              _cabi.lib().xrd_hashgrid_layout(C.byref(g), n_levels,
                                              log2_hashmap_size,
                                              base_resolution,
                                              float(np.float32(per_level_scale))))
        self.layout = g
        self.n_levels = n_levels
        self.n_output_dims = 2 * n_levels
        gen = torch.Generator().manual_seed(seed)
        p = (torch.rand(int(g.n_entries) * 2, generator=gen) * 2 - 1) * 1e-4
        self.params = nn.Parameter(p.float())


class _Seq(nn.Module):
    """nn.Sequential-shaped holder so state_dict keys equal the reference's
    (decoder_coslam.py:40-56, :94-111: Linear, ReLU, Linear; bias=False)."""
    def __init__(self, d_in, d_hidden, d_out):
        super().__init__()
        self.model = nn.Sequential(nn.Linear(d_in, d_hidden, bias=False),
                                   nn.ReLU(inplace=True),
                                   nn.Linear(d_hidden, d_out, bias=False))


class ColorSDFNet_v2(nn.Module):
    """Parameter container of decoder_coslam.py:139-163 (no forward: the MLP runs
    inside the fused kernel)."""
    def __init__(self, config, input_ch, input_ch_pos):
        super().__init__()
        self.color_net = _Seq(input_ch_pos + config.geo_feat_dim,
                              config.hidden_dim_color, 3)
        self.sdf_net = _Seq(input_ch + input_ch_pos, config.hidden_dim,
                            1 + config.geo_feat_dim)


class _CoslamStep(torch.autograd.Function):
    """losses[4] = fused(rays, table, weights); backward hands out the gradients
    the kernel already produced."""
    @staticmethod
    def forward(ctx, rays_o, rays_d, table, w0, w1, wc0, wc1, model, target_s,
                target_d, noise):
        need = any(ctx.needs_input_grad[:7])
        need_rays = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        # tracking: only the pose is optimised -> skip scatter / weight-gradient tiles
        map_grads = any(ctx.needs_input_grad[2:7]) and not (
            model.freeze_map_grads and need_rays)
        outs, grads = model._launch(rays_o, rays_d, table, w0, w1, wc0, wc1,
                                    target_s, target_d, noise, with_grads=need,
                                    need_ray_grads=need_rays,
                                    map_grads=map_grads)
        ctx.grads = grads
        ctx.model = model
        ctx.strict = model.config.strict_loss_grad
        ctx.inputs = (rays_o, rays_d, table, w0, w1, wc0, wc1, target_s,
                      target_d, noise, model._last_seed)
        ret = (outs['losses'], outs['rgb'], outs['depth'], outs['disp_map'],
               outs['acc_map'], outs['depth_var'], outs['z_vals'], outs['raw'])
        ctx.mark_non_differentiable(*ret[1:])
        return ret

    @staticmethod
    def backward(ctx, g_losses, *_):
        grads = ctx.grads
        if grads is None:
            raise RuntimeError('backward through a forward-only Co-SLAM pass')
        if ctx.strict:
            g = g_losses.detach().float().cpu()
            if not bool(torch.all(g == 1.0)):
                # non-unit upstream gradient: re-run the fused pass with the
                # per-term scales (the kernel is linear in them)
                (rays_o, rays_d, table, w0, w1, wc0, wc1, ts, td, noise,
                 seed) = ctx.inputs
                _, grads = ctx.model._launch(
                    rays_o, rays_d, table, w0, w1, wc0, wc1, ts, td, noise,
                    with_grads=True, need_ray_grads=True,
                    loss_scale=[float(v) for v in g], seed=seed)
        else:
            # sync-free: total-loss scaling is honoured, unequal per-term scales poison
            # the gradients with NaN (base_model.upstream_scale)
            s = upstream_scale(g_losses)
            keys = ('d_rays_o', 'd_rays_d', 'd_table', 'd_w_sdf0', 'd_w_sdf1', 'd_w_col0',
                    'd_w_col1')
            grads = dict(zip(keys, scale_grads([grads[k] for k in keys], s)))
        return (grads['d_rays_o'], grads['d_rays_d'], grads['d_table'],
                grads['d_w_sdf0'], grads['d_w_sdf1'], grads['d_w_col0'],
                grads['d_w_col1'], None, None, None, None)


class JointEncoding(Model):
    """Model class (slam/models/joint_encoding.py:69-531)."""

    config: JointEncodingConfig

    def __init__(self, config: JointEncodingConfig, camera, bounding_box,
                 **kwargs) -> None:
        super().__init__(config=config, camera=camera,
                         bounding_box=bounding_box, **kwargs)

    # ------------------------------------------------------------- set-up ---
    def populate_modules(self):
        super().populate_modules()
        cfg = self.config
        if not cfg.oneGrid or cfg.tcnn_network or cfg.training_n_importance:
            raise NotImplementedError(
                'B200 path covers the reference defaults: oneGrid=True, '
                'tcnn_network=False, n_importance=0 (input_config.py:253-254)')
        if not cfg.tcnn_encoding:
            raise NotImplementedError('tcnn_encoding=True is the co-slam '
                                      'default (input_config.py:254)')
        if (cfg.hidden_dim, cfg.hidden_dim_color, cfg.geo_feat_dim,
                cfg.pos_nbins, cfg.num_layers, cfg.num_layers_color) != (
                    32, 32, 15, 16, 2, 2):
            raise NotImplementedError('decoder shape fixed to the reference '
                                      'default 80-32-16 / 63-32-3')
        self.bounding_box = torch.as_tensor(np.asarray(self.bounding_box),
                                            dtype=torch.float64)
        self.get_resolution()
        self.get_encoding()
        self.get_decoder()
        n = cfg.training_n_range_d
        ls = torch.linspace
        self.register_buffer('_lin_uniform',
                             ls(cfg.cam_near, cfg.cam_far,
                                cfg.training_n_sample_d), persistent=False)
        self.register_buffer('_lin_range',
                             ls(-cfg.training_range_d, cfg.training_range_d,
                                steps=n), persistent=False)
        self.register_buffer('_lin_nodepth',
                             ls(cfg.cam_near, cfg.cam_far, steps=n),
                             persistent=False)
        self.register_buffer('_lin_full',
                             ls(cfg.cam_near, cfg.cam_far,
                                cfg.training_n_samples), persistent=False)
        self._step_count = 0
        # set by the Algorithm around tracking (only pose params have optimizers
        # there, base_algorithm.py:168-181): the fused pass then produces d rays only
        self.freeze_map_grads = False
        self.dp = None  # xrdslam_b200.dp.MappingDataParallel when mapping rays are sharded

    def get_resolution(self):
        """joint_encoding.py:199-210."""
        dim_max = (self.bounding_box[:, 1] - self.bounding_box[:, 0]).max()
        if self.config.voxel_sdf > 10:
            self.resolution_sdf = self.config.voxel_sdf
        else:
            self.resolution_sdf = int(dim_max / self.config.voxel_sdf)
        if self.config.voxel_color > 10:
            self.resolution_color = self.config.voxel_color
        else:
            self.resolution_color = int(dim_max / self.config.voxel_color)

    def get_encoding(self):
        """joint_encoding.py:212-234 + encodings_coslam.py:39-53,66-75."""
        n_levels, base = 16, 16
        per_level_scale = np.exp2(
            np.log2(self.resolution_sdf / base) / (n_levels - 1))
        self.embed_fn = HashGridParams(n_levels, self.config.hashsize, base,
                                       per_level_scale)
        self.input_ch = self.embed_fn.n_output_dims
        self.input_ch_pos = 3 * self.config.pos_nbins  # OneBlob, no params

    def get_decoder(self):
        self.decoder = ColorSDFNet_v2(self.config, input_ch=self.input_ch,
                                      input_ch_pos=self.input_ch_pos)

    # ------------------------------------------------------ C-ABI plumbing ---
    def _grid_struct(self, table):
        g = XrdHashGrid()
        C.memmove(C.byref(g), C.byref(self.embed_fn.layout), C.sizeof(g))
        bb = self.bounding_box
        for d in range(3):
            g.bbox_min[d] = float(bb[d, 0])
            g.bbox_max[d] = float(bb[d, 1])
        g.table = ptr(table)
        return g

    def _weights(self):
        d = self.decoder
        return (d.sdf_net.model[0].weight, d.sdf_net.model[2].weight,
                d.color_net.model[0].weight, d.color_net.model[2].weight)

    def _launch(self, rays_o, rays_d, table, w0, w1, wc0, wc1, target_s,
                target_d, noise, with_grads, need_ray_grads=True,
                loss_scale=None, seed=None, map_grads=True):
        cfg = self.config
        dev = table.device
        if dev.type != 'cuda':
            raise RuntimeError('xrdslam_b200 has no CPU path: model must be on '
                               'a CUDA (sm_100) device')
        lib = _cabi.lib()
        f32 = dict(dtype=torch.float32, device=dev)
        rays_o = rays_o.detach().to(**f32).contiguous()
        rays_d = rays_d.detach().to(**f32).contiguous()
        R = rays_o.shape[0]
        has_d = target_d is not None
        S = (cfg.training_n_sample_d +
             cfg.training_n_range_d) if has_d else cfg.training_n_samples
        if has_d and cfg.training_n_sample_d <= 0:
            raise NotImplementedError('training_n_sample_d must be > 0')
        td = target_d.detach().to(**f32).reshape(-1).contiguous() if has_d \
            else None
        ts = target_s.detach().to(**f32).contiguous() \
            if target_s is not None else None
        if noise is not None:
            noise = noise.detach().to(**f32).contiguous()
            assert noise.shape == (R, S)
        o = dict(rgb=torch.empty(R, 3, **f32), depth=torch.empty(R, **f32),
                 disp_map=torch.empty(R, **f32), acc_map=torch.empty(R, **f32),
                 depth_var=torch.empty(R, **f32),
                 z_vals=torch.empty(R, S, **f32),
                 raw=torch.empty(R, S, 4, **f32),
                 losses=torch.zeros(4, **f32))
        rays = XrdRays(R, ptr(rays_o), ptr(rays_d), ptr(ts), ptr(td))
        grid = self._grid_struct(table.detach())
        mlp = XrdCoslamMlp(ptr(w0.detach()), ptr(w1.detach()),
                           ptr(wc0.detach()), ptr(wc1.detach()))
        if seed is None:
            self._step_count += 1
            seed = (cfg.seed << 32) + self._step_count
        # sharding applies to mapping only (tracking is per-frame sequential, single GPU)
        dp = self.dp if (self.dp is not None and self.dp.world > 1 and with_grads
                         and map_grads) else None
        c = XrdCoslamCfg(
            S, cfg.training_n_sample_d, cfg.training_n_range_d,
            int(cfg.training_perturb > 0),
            cfg.training_trunc * cfg.data_sc_factor, cfg.cam_depth_trunc,
            cfg.trainging_rgb_weight, cfg.trainging_depth_weight,
            cfg.trainging_sdf_weight, cfg.trainging_fs_weight,
            ptr(self._lin_uniform), ptr(self._lin_range),
            ptr(self._lin_nodepth), ptr(self._lin_full),
            seed, cfg.rays_per_tile, cfg.precision, 0, 0, None, None, None)
        out = XrdCoslamOut(ptr(o['rgb']), ptr(o['depth']), ptr(o['disp_map']),
                           ptr(o['acc_map']), ptr(o['depth_var']),
                           ptr(o['z_vals']), ptr(o['raw']), ptr(o['losses']))
        g = None
        gs = None
        if with_grads:
            if not has_d or ts is None:
                raise RuntimeError('gradients need target_s and target_d')
            z = torch.zeros_like
            g = dict(d_table=z(table, **f32) if map_grads else None,
                     d_w_sdf0=z(w0) if map_grads else None,
                     d_w_sdf1=z(w1) if map_grads else None,
                     d_w_col0=z(wc0) if map_grads else None,
                     d_w_col1=z(wc1) if map_grads else None,
                     d_rays_o=torch.empty(R, 3, **f32) if need_ray_grads else None,
                     d_rays_d=torch.empty(R, 3, **f32) if need_ray_grads else None)
            sc = loss_scale or [1.0, 1.0, 1.0, 1.0]
            gs = XrdCoslamGrads(ptr(g['d_table']), ptr(g['d_w_sdf0']),
                                ptr(g['d_w_sdf1']), ptr(g['d_w_col0']),
                                ptr(g['d_w_col1']), ptr(g['d_rays_o']),
                                ptr(g['d_rays_d']), (C.c_float * 4)(*sc))
        ws_bytes = lib.xrd_coslam_workspace_bytes(R, S)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            if dp is not None:
                # sharded mapping: the loss normalisers are batch-global:
                # sample -> all-reduce the three counters -> render with the global counts
                counts = torch.zeros(4, dtype=torch.int32, device=dev)
                c.phase, c.counts_out = 1, ptr(counts)
                st = lib.xrd_coslam_step(C.byref(rays), C.byref(grid), C.byref(mlp),
                                         C.byref(c), ptr(noise), C.byref(out), None,
                                         ptr(ws), ws_bytes, stream)
                check('xrd_coslam_step[sample]', st)
                dp.all_reduce_sum(counts)
                c.phase, c.counts_global = 2, ptr(counts)
                c.n_rays_global = R * dp.world  # equal shards
            st = lib.xrd_coslam_step(C.byref(rays), C.byref(grid), C.byref(mlp),
                                     C.byref(c), ptr(noise), C.byref(out),
                                     C.byref(gs) if gs is not None else None,
                                     ptr(ws), ws_bytes, stream)
        check('xrd_coslam_step', st)
        self._last_seed = seed
        return o, g

    def _launch_smooth(self, table, rand6, need_grad, weight=1.0):
        cfg = self.config
        lib = _cabi.lib()
        dev = table.device
        grid = self._grid_struct(table.detach())
        loss = torch.zeros(1, dtype=torch.float32, device=dev)
        d_table = torch.zeros_like(table) if need_grad else None
        r = (C.c_float * 6)(*[float(v) for v in rand6])
        nb = lib.xrd_coslam_smoothness_workspace_bytes(cfg.trainging_smooth_pts)
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            st = lib.xrd_coslam_smoothness(
                C.byref(grid), cfg.trainging_smooth_pts,
                cfg.trainging_smooth_vox, cfg.trainging_smooth_margin, weight,
                r, ptr(loss), ptr(d_table), 1.0, ptr(ws), nb,
                torch.cuda.current_stream(dev).cuda_stream)
        check('xrd_coslam_smoothness', st)
        return loss[0], d_table

    # --------------------------------------------------------- Model API ---
    def get_outputs(self, input) -> Dict[str, Union[torch.Tensor, List]]:
        """joint_encoding.py:158-163.  Extra optional key ``noise`` [R,S]
        replaces torch.rand(z_vals.shape) (:292) for parity runs."""
        return self.render_rays(input['rays_o'], input['rays_d'],
                                target_d=input.get('target_d'),
                                target_s=input.get('target_s'),
                                noise=input.get('noise'))

    def render_rays(self, rays_o, rays_d, target_d=None, target_s=None,
                    noise=None):
        table = self.embed_fn.params
        w0, w1, wc0, wc1 = self._weights()
        fused = (torch.is_grad_enabled() and target_d is not None
                 and target_s is not None)
        if fused:
            (losses, rgb, depth, disp, acc, var, z_vals,
             raw) = _CoslamStep.apply(rays_o, rays_d, table, w0, w1, wc0, wc1,
                                      self, target_s, target_d, noise)
            ret = dict(rgb=rgb, depth=depth, disp_map=disp, acc_map=acc,
                       depth_var=var, z_vals=z_vals, raw=raw)
            ret['_losses'] = losses
            return ret
        o, _ = self._launch(rays_o, rays_d, table, w0, w1, wc0, wc1, target_s,
                            target_d, noise, with_grads=False)
        o.pop('losses')
        return o

    def get_loss_dict(self, outputs, inputs, is_mapping,
                      stage=None) -> Dict[str, torch.Tensor]:
        """joint_encoding.py:94-147 -- the four terms were produced (already
        weighted) by the fused kernel during forward()."""
        if '_losses' not in outputs:
            raise RuntimeError('get_loss_dict needs outputs of a forward() '
                               'run with grad enabled and targets present')
        ls = outputs['_losses']
        loss_dict = {
            'rgb_loss': ls[0],
            'depth_loss': ls[1],
            'sdf_loss': ls[2],
            'fs_loss': ls[3],
        }
        if is_mapping and not inputs['first']:
            # weight folded into the kernel (loss and gradient), :140-145
            loss_dict['smooth_loss'] = self.smoothness(
                self.config.trainging_smooth_pts,
                self.config.trainging_smooth_vox,
                self.config.trainging_smooth_margin,
                rand=inputs.get('smooth_rand'),
                weight=self.config.trainging_smooth_weight)
        return loss_dict

    def get_param_groups(self) -> Dict[str, List[Parameter]]:
        return {
            'decoder': list(self.decoder.parameters()),
            'embed_fn': list(self.embed_fn.parameters()),
        }

    def smoothness(self, sample_points=256, voxel_size=0.1, margin=0.05,
                   rand=None, weight=None):
        """joint_encoding.py:165-197.  ``rand`` = [torch.rand(3), torch.rand(3)]
        (the two CPU draws of :176 / :179) -- drawn here when not given."""
        cfg = self.config
        if (sample_points, voxel_size, margin) != (cfg.trainging_smooth_pts,
                                                   cfg.trainging_smooth_vox,
                                                   cfg.trainging_smooth_margin):
            raise NotImplementedError
        if rand is None:
            rand = torch.cat([torch.rand(3), torch.rand((1, 1, 1, 3)).reshape(3)])
        rand6 = torch.as_tensor(rand, dtype=torch.float32).reshape(6).tolist()
        return _SmoothFn.apply(self.embed_fn.params, self, rand6, weight)

    # ---- mesher-facing queries (joint_encoding.py:408-481) ----------------
    def _query(self, pts, normalised, want_raw=True, want_geo=False, want_feat=False):
        """xrd_coslam_query on [..., 3] points -> dict of flat [P, k] tensors (inference)."""
        table = self.embed_fn.params
        dev = table.device
        if dev.type != 'cuda':
            raise RuntimeError('xrdslam_b200 has no CPU path: model must be on a CUDA device')
        flat = pts.detach().reshape(-1, 3).to(device=dev, dtype=torch.float32).contiguous()
        P = flat.shape[0]
        f32 = dict(dtype=torch.float32, device=dev)
        raw = torch.empty(P, 4, **f32) if want_raw else None
        geo = torch.empty(P, 15, **f32) if want_geo else None
        feat = torch.empty(P, 32, **f32) if want_feat else None
        grid = self._grid_struct(table.detach())
        mlp = XrdCoslamMlp(*(ptr(w.detach()) for w in self._weights()))
        with torch.cuda.device(dev):
            st = _cabi.lib().xrd_coslam_query(C.byref(grid), C.byref(mlp), ptr(flat), P,
                                              int(normalised), ptr(raw), ptr(geo), ptr(feat),
                                              torch.cuda.current_stream(dev).cuda_stream)
        check('xrd_coslam_query', st)
        return dict(raw=raw, geo=geo, feat=feat)

    def query_fn(self, pi):
        """:409-416: SDF at world points [N,3] -> [N,1]."""
        return self._query(pi, False)['raw'][:, 3:4]

    def color_func(self, pi):
        """:419-425: colour at world points [N,3] -> [N,1,3]."""
        return torch.sigmoid(self._query(pi, False)['raw'][:, None, :3])

    def query_sdf(self, query_points, return_geo=False, embed=False):
        """:427-456, points already normalised to the unit cube."""
        shp = list(query_points.shape[:-1])
        if embed:
            return self._query(query_points, True, False, False, True)['feat'].reshape(shp + [32])
        o = self._query(query_points, True, True, return_geo)
        sdf = o['raw'][:, 3].reshape(shp)
        if not return_geo:
            return sdf
        return sdf, o['geo'].reshape(shp + [15])

    def query_color_sdf(self, query_points):
        """:463-481 -> raw [..., 4] (rgb logits ++ sdf) at normalised points."""
        return self._query(query_points, True)['raw'].reshape(list(query_points.shape[:-1]) + [4])

    def query_color(self, query_points):
        """:458-461."""
        return torch.sigmoid(self.query_color_sdf(query_points)[..., :3])

    def run_network(self, inputs):
        """:483-507: raw [..., 4] at WORLD points (normalised inside, float64 like the
        reference's bounding-box arithmetic)."""
        return self._query(inputs, False)['raw'].reshape(list(inputs.shape[:-1]) + [4])


class _SmoothFn(torch.autograd.Function):
    """weight=None: plain TV loss, exact upstream scaling in backward.
    weight=w: loss and gradient come out of the kernel already multiplied by w
    and backward assumes the unit upstream gradient of the plugin contract
    (loss = reduce(add, loss_dict.values()); loss.backward(), coslam.py:242)
    unless config.strict_loss_grad."""
    @staticmethod
    def forward(ctx, table, model, rand6, weight):
        loss, d_table = model._launch_smooth(
            table, rand6, ctx.needs_input_grad[0],
            1.0 if weight is None else float(weight))
        ctx.d_table = d_table
        ctx.exact = weight is None or model.config.strict_loss_grad
        return loss

    @staticmethod
    def backward(ctx, g):
        d = ctx.d_table  # linear in the upstream gradient: exact for any g, no host sync
        return (d * g if d is not None else None), None, None, None
