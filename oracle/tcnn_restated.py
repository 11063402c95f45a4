"""Torch-CPU restatement of the two tinycudann encodings Co-SLAM uses.

ORACLE / TEST INFRASTRUCTURE -- not shipped, never imported by xrdslam_b200/.

PARITY UNPINNED: tinycudann is an un-vendored, un-pinned dependency of the
reference (requirements.txt:5 ``git+https://github.com/NVlabs/tiny-cuda-nn``),
it is not installed in the build container and the reference holds no golden
vectors for it.  What is restated here is tcnn's published algorithm
(include/tiny-cuda-nn/encodings/grid.h and oneblob.h, as of the 1.7 line):

  grid_scale(l)      = exp2f(l * log2f(per_level_scale)) * base_resolution - 1
  grid_resolution    = (uint32)ceilf(scale) + 1
  params_in_level    = min(next_multiple(res^3, 8), 1 << log2_hashmap_size)
  pos_fract          : pos = fmaf(scale, x, 0.5f); cell = floorf(pos);
                       pos_grid = (uint32)(int)cell; w = pos - cell
  grid_index         : dense stride walk while stride <= hashmap_size, else
                       coherent-prime hash (1, 2654435761, 805459861), % size
  N-linear interp    : corner bit d set -> weight w_d, cell+1
  OneBlob            : right_cdf - left_cdf of the periodic quartic kernel

Call sites in the reference that fix the configuration:
  slam/model_components/encodings_coslam.py:39-53 (HashGrid), :66-75 (OneBlob)
  slam/models/joint_encoding.py:199-234 (resolution / encoder set-up)
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.nn as nn

# tcnn evaluates std::log2(float) / exp2f on the host with the C library; numpy's
# float32 log2/exp2 are SIMD approximations that differ in the last bit (which
# flips ceil(scale_15) between 324 and 325 at the Co-SLAM default), so call libm.
_libm = ctypes.CDLL('libm.so.6')
_libm.log2f.restype = ctypes.c_float
_libm.log2f.argtypes = [ctypes.c_float]
_libm.exp2f.restype = ctypes.c_float
_libm.exp2f.argtypes = [ctypes.c_float]

PRIME_Y = 2654435761
PRIME_Z = 805459861
U32 = 0xFFFFFFFF


def hashgrid_level_table(n_levels=16,
                         n_features=2,
                         log2_hashmap_size=16,
                         base_resolution=16,
                         per_level_scale=2.0):
    """Per-level (scale f32, resolution, size, offset, hashed) following tcnn's
    GridEncodingTemplated constructor + grid_scale/grid_resolution."""
    pls = np.float32(per_level_scale)  # json number -> float
    log2_pls = np.float32(_libm.log2f(float(pls)))  # std::log2(float)
    scales, ress, sizes, offs, hashed = [], [], [], [], []
    off = 0
    for lvl in range(n_levels):
        e = np.float32(_libm.exp2f(float(np.float32(lvl) * log2_pls)))
        sc = e * np.float32(base_resolution) - np.float32(1.0)
        assert sc.dtype == np.float32
        res = int(np.ceil(sc)) + 1
        dense = res**3
        if float(res)**3 > float(U32 // 2):
            dense = U32 // 2
        n = (dense + 7) // 8 * 8
        n = min(n, 1 << log2_hashmap_size)
        # replay grid_index's stride walk to learn whether this level hashes
        stride = 1
        for _ in range(3):
            if stride > n:
                break
            stride *= res
        scales.append(sc)
        ress.append(res)
        sizes.append(n)
        offs.append(off)
        hashed.append(n < stride)
        off += n
    return dict(scale=np.array(scales, dtype=np.float32),
                resolution=np.array(ress, dtype=np.uint32),
                size=np.array(sizes, dtype=np.uint32),
                offset=np.array(offs, dtype=np.uint32),
                hashed=np.array(hashed, dtype=bool),
                n_entries=off,
                n_params=off * n_features,
                n_features=n_features)


def hashgrid_indices(x: torch.Tensor, table: dict):
    """Integer part of the encoding: for points x [P,3] f32 return
    (idx [P,L,8] int64 entry index incl. level offset, w [P,L,3] f32).
    Bit-exact target for the CUDA kernel's index math."""
    assert x.dtype == torch.float32
    L = len(table['scale'])
    P = x.shape[0]
    dev = x.device  # (constants follow the points: the bench's torch-on-GPU leg runs this on cuda)
    scale = torch.from_numpy(table['scale']).to(dev)  # [L] f32
    # fmaf(scale, x, 0.5f): f64 product of two f32 is exact, one rounding left
    pos = (scale.double()[None, :, None] * x.double()[:, None, :] +
           0.5).float()  # [P,L,3]
    cell = torch.floor(pos)
    w = pos - cell
    g = cell.to(torch.int64) & U32  # (uint32)(int)cell
    res = torch.from_numpy(table['resolution'].astype(np.int64))[None, :].to(dev)
    size = torch.from_numpy(table['size'].astype(np.int64))[None, :].to(dev)
    off = torch.from_numpy(table['offset'].astype(np.int64))[None, :].to(dev)
    hashed = torch.from_numpy(table['hashed'])[None, :].to(dev)
    idx = torch.empty(P, L, 8, dtype=torch.int64, device=dev)
    for c in range(8):
        gx = (g[..., 0] + ((c >> 0) & 1)) & U32
        gy = (g[..., 1] + ((c >> 1) & 1)) & U32
        gz = (g[..., 2] + ((c >> 2) & 1)) & U32
        # dense walk: index += g[d]*stride while stride <= size
        index = gx.clone()
        stride = res.clone()  # after dim 0
        use1 = stride <= size
        index = torch.where(use1, (index + gy * stride) & U32, index)
        stride2 = torch.where(use1, stride * res, stride)
        use2 = use1 & (stride2 <= size)
        index = torch.where(use2, (index + gz * stride2) & U32, index)
        h = (gx ^ ((gy * PRIME_Y) & U32) ^ ((gz * PRIME_Z) & U32)) & U32
        index = torch.where(hashed, h, index)
        idx[..., c] = index % size + off
    return idx, w


class HashGridRestated(nn.Module):
    """tcnn.Encoding(otype='HashGrid', dtype=torch.float) restated.

    params: one flat fp32 Parameter, level-major / entry-major / feature-minor,
    init U(-1e-4, 1e-4) (tcnn default for grids; tcnn's own RNG stream is not
    reproduced -- values are passed explicitly in every parity test).
    """
    def __init__(self,
                 n_input_dims=3,
                 n_levels=16,
                 n_features_per_level=2,
                 log2_hashmap_size=19,
                 base_resolution=16,
                 per_level_scale=2.0,
                 seed=1337):
        super().__init__()
        assert n_input_dims == 3
        self.table = hashgrid_level_table(n_levels, n_features_per_level,
                                          log2_hashmap_size, base_resolution,
                                          per_level_scale)
        self.n_levels = n_levels
        self.F = n_features_per_level
        self.n_output_dims = n_levels * n_features_per_level
        g = torch.Generator().manual_seed(seed)
        p = (torch.rand(self.table['n_params'], generator=g) * 2 - 1) * 1e-4
        self.params = nn.Parameter(p.float())

    def forward(self, x):
        x = x.to(torch.float32).contiguous()
        idx, w_exact = hashgrid_indices(x.detach(), self.table)
        scale = torch.from_numpy(self.table['scale'])[None, :, None].to(x.device)
        xs = x[:, None, :] * scale  # differentiable carrier, d w/d x = scale
        w = w_exact + (xs - xs.detach())  # [P,L,3]
        tab = self.params.view(-1, self.F)
        out = 0
        for c in range(8):
            wc = 1.0
            for d in range(3):
                wd = w[..., d]
                wc = wc * (wd if (c >> d) & 1 else (1 - wd))
            out = out + wc[..., None] * tab[idx[..., c]]  # [P,L,F]
        return out.reshape(x.shape[0], self.n_output_dims)


def quartic_cdf(u_in, inv_radius):
    u = u_in * inv_radius
    u2 = u * u
    u4 = u2 * u2
    v = (15.0 / 16.0) * u * (1 - (2.0 / 3.0) * u2 + (1.0 / 5.0) * u4) + 0.5
    return torch.clamp(v, 0.0, 1.0)


class OneBlobRestated(nn.Module):
    """tcnn.Encoding(otype='OneBlob', n_bins) restated: out[d*n_bins+b]."""
    def __init__(self, n_input_dims=3, n_bins=16):
        super().__init__()
        self.n_bins = n_bins
        self.n_input_dims = n_input_dims
        self.n_output_dims = n_bins * n_input_dims

    def forward(self, x):
        x = x.to(torch.float32)
        nb = self.n_bins
        edges = torch.arange(nb + 1, dtype=torch.float32, device=x.device) / nb  # scalbnf
        d = edges[None, None, :] - x[:, :, None]  # [P,3,nb+1]
        s = float(nb)
        cdf = quartic_cdf(d, s) + quartic_cdf(d - 1.0, s) + quartic_cdf(
            d + 1.0, s)
        out = cdf[..., 1:] - cdf[..., :-1]
        return out.reshape(x.shape[0], self.n_output_dims)


class Encoding:
    """Drop-in for ``tinycudann.Encoding`` used by ref_harness to run the
    reference's own JointEncoding class on CPU."""
    def __new__(cls, n_input_dims, encoding_config, dtype=torch.float):
        ot = encoding_config['otype']
        if ot == 'HashGrid':
            return HashGridRestated(
                n_input_dims,
                n_levels=encoding_config['n_levels'],
                n_features_per_level=encoding_config['n_features_per_level'],
                log2_hashmap_size=encoding_config['log2_hashmap_size'],
                base_resolution=encoding_config['base_resolution'],
                per_level_scale=encoding_config['per_level_scale'])
        if ot == 'OneBlob':
            return OneBlobRestated(n_input_dims, encoding_config['n_bins'])
        raise NotImplementedError(ot)
