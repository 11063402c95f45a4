"""CPU tests for the Vox-Fusion map structure: this package's octree (csrc/octree.cpp) against
the reference's own svo.Octree compiled from its sources (oracle/_ref/svo.so)."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SVO = os.path.join(ROOT, 'oracle', '_ref', 'svo.so')


def ref_octree():
    if not os.path.exists(SVO):
        pytest.skip('oracle/_ref/svo.so not built (python oracle/build_ref.py)')
    torch.classes.load_library(SVO)
    o = torch.classes.svo.Octree()
    o.init(256, 16, 0.2)
    return o


def test_octree_bit_exact_vs_reference_svo():
    from xrdslam_b200 import _cabi
    lib = _cabi.lib()
    ref = ref_octree()
    # the reference numbers nodes with a process-global counter: offset of this tree's root
    v0, _, _ = ref.get_centres_and_children()
    t = lib.xrd_octree_create(256)
    g = torch.Generator().manual_seed(4)
    try:
        for it in range(4):
            pts = (torch.rand(2500, 3, generator=g) * 25 + 30 + 7 * it).int().contiguous()
            if it == 0:
                pts[0] = torch.tensor([0, 0, 0])
                pts[1] = torch.tensor([254, 254, 254])
                pts[2] = pts[3]  # duplicates
            ref.insert(pts)
            n = lib.xrd_octree_insert(t, pts.data_ptr(), pts.shape[0])
            v, c, f = ref.get_centres_and_children()
            assert n == v.shape[0] == lib.xrd_octree_num_nodes(t)
            mv = torch.empty(n, 4)
            mc = torch.empty(n, 8)
            mf = torch.empty(n, 8, dtype=torch.int32)
            assert lib.xrd_octree_export(t, mv.data_ptr(), mc.data_ptr(), mf.data_ptr()) == n
            if int(v0.shape[0]) == 1:  # first tree of the process: ids coincide
                assert torch.equal(v, mv) and torch.equal(c, mc) and torch.equal(f, mf)
            else:
                assert torch.equal(v, mv)
    finally:
        lib.xrd_octree_destroy(t)


def test_model_map_states_match_reference_formulas():
    from xrdslam_b200.camera import Camera
    from xrdslam_b200.sparse_voxel import SparseVoxelConfig
    m = SparseVoxelConfig().setup(camera=Camera(320, 320, 319.5, 239.5, 640, 480))
    g = torch.Generator().manual_seed(0)
    pts = torch.rand(800, 3, generator=g) * 3 + 24.1
    m.insert_points(pts)
    ms = m.map_states
    N = ms['voxel_center_xyz'].shape[0]
    assert ms['voxel_structure'].shape == (N, 9) and ms['voxel_vertex_idx'].shape == (N, 8)
    leaf = ms['voxel_structure'][:, 8] == 1
    surf = leaf & (ms['voxel_vertex_idx'][:, 0] >= 0)
    # every inserted point lies inside a SURFACE leaf: centre = (floor(p/0.2) + 0.5) * 0.2
    want = {tuple(v) for v in torch.div(pts, 0.2, rounding_mode='floor').int().tolist()}
    have = {tuple(v) for v in torch.round(ms['voxel_center_xyz'][surf] / 0.2 - 0.5).int().tolist()}
    assert want <= have and len(have) == len(want)
    assert int(ms['voxel_vertex_idx'].max()) < m.config.num_embeddings
