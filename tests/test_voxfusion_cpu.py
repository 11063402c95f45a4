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


def test_vox_oracle_torch_part_matches_reference_python():
    """oracle/voxfusion.py's features / decoder / sdf2weights / losses against the reference's
    own SparseVoxel.render_rays + get_loss_dict run on CPU (its two CUDA ops replaced by the
    oracle's intersections and samples, which the GPU tests pin bit-for-bit against the
    reference's compiled kernels)."""
    import sys
    sys.path.insert(0, ROOT)
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip('needs /root/reference')
    from oracle.voxfusion import VoxOracle
    from xrdslam_b200.camera import Camera
    from xrdslam_b200.sparse_voxel import SparseVoxelConfig
    g = torch.Generator().manual_seed(3)
    # a wall of voxels at z ~ 12.0 m (offset world), rays from above.  The map comes from this
    # package's octree (bit-exact vs the reference's svo above, with per-tree node ids: the
    # reference's ids are offset by a process-global counter)
    xy = torch.rand(3000, 2, generator=g) * 2.0 + 11.0
    pts = torch.cat([xy, torch.full((3000, 1), 12.05) + torch.rand(3000, 1, generator=g) * 0.3], 1)
    builder = SparseVoxelConfig().setup(camera=Camera(320, 320, 319.5, 239.5, 640, 480))
    builder.insert_points(pts)
    voxels, children, features = builder.export_octree()
    ora = VoxOracle(seed=5)
    ora.set_map(voxels, children, features)
    with torch.no_grad():
        ora.embeddings.mul_(30.0)
    R = 96
    ro = torch.cat([torch.rand(R, 2, generator=g) * 1.6 + 11.2, torch.full((R, 1), 10.5)], 1)
    rd = torch.nn.functional.normalize(
        torch.cat([torch.randn(R, 2, generator=g) * 0.15, torch.ones(R, 1)], 1), dim=-1)
    rd[::13] = torch.tensor([0.0, 0.0, -1.0])  # rays that miss the map
    marched = ora.march(ro, rd, lambda shape: torch.rand(shape, generator=g))
    assert marched is not None
    inter, hits, samples = marched
    assert hits.any() and not hits.all()
    td = torch.full((R, 1), 1.62) + torch.rand(R, 1, generator=g) * 0.2
    td[5::11] = 0
    ts = torch.rand(R, 3, generator=g)
    # the reference model on the same map / decoder / embeddings
    ms = {'voxel_vertex_idx': ora.vertex_idx, 'voxel_center_xyz': ora.centres,
          'voxel_structure': ora.children}
    full_inter = {k: torch.zeros((R,) + v.shape[1:], dtype=v.dtype) for k, v in inter.items()}
    for k in full_inter:
        full_inter[k][hits] = inter[k]
    ref, sv = ref_harness.ref_sparse_voxel_cpu(ms, ora.embeddings.detach(), (full_inter, hits, samples))
    rsd = ref.decoder.state_dict()
    with torch.no_grad():
        od = ora.decoder
        od.pts_linears[0].weight.copy_(rsd['pts_linears.0.weight']); od.pts_linears[0].bias.copy_(rsd['pts_linears.0.bias'])
        od.pts_linears[1].weight.copy_(rsd['pts_linears.1.weight']); od.pts_linears[1].bias.copy_(rsd['pts_linears.1.bias'])
        od.sdf_out.weight.copy_(rsd['sdf_out.weight']); od.sdf_out.bias.copy_(rsd['sdf_out.bias'])
        od.color_out[0].weight.copy_(rsd['color_out.0.weight']); od.color_out[0].bias.copy_(rsd['color_out.0.bias'])
        od.color_out[2].weight.copy_(rsd['color_out.2.weight']); od.color_out[2].bias.copy_(rsd['color_out.2.bias'])
    out_o, ld_o = ora.render(ro, rd, ts, td, marched)
    with ref_harness.cuda_calls_are_noops():
        out_r = ref.render_rays(ro.unsqueeze(0), rd.unsqueeze(0), target_d=td.unsqueeze(0))
    ld_r = ref.get_loss_dict(out_r, {'target_d': td, 'target_s': ts}, True)
    assert torch.equal(out_r['ray_mask'], out_o['ray_mask'])
    assert (out_r['depth'] - out_o['depth']).abs().max() < 1e-6
    assert (out_r['rgb'] - out_o['rgb']).abs().max() < 1e-6
    assert (out_r['sdf'] - out_o['sdf']).abs().max() < 1e-6
    for k in ld_r:
        assert abs(float(ld_r[k]) - float(ld_o[k])) <= 1e-5 * max(1e-3, abs(float(ld_r[k]))), k
    # gradients of the summed loss w.r.t. embeddings and decoder
    sum(ld_o.values()).backward()
    sum(ld_r.values()).backward()
    ge_o, ge_r = ora.embeddings.grad, ref.embeddings.grad
    assert (ge_o - ge_r).abs().max() <= 1e-5 * ge_r.abs().max()
    assert (od.sdf_out.weight.grad - ref.decoder.sdf_out.weight.grad).abs().max() <= \
        1e-5 * ref.decoder.sdf_out.weight.grad.abs().max()


def test_device_octree_equals_host_octree_numbering():
    """octree_device.DeviceOctree (data-parallel build: unique / sort / searchsorted) produces
    the SAME node ids, codes, child tables and corner-leaf tables as the host C++ octree (which
    is bit-exact against the reference's compiled svo.Octree, test above) -- over duplicate
    voxels, two successive insert calls, coordinates at the grid edge (wrap quirk) and an
    empty second call.  Runs on the host here; the same tensor program runs on the GPU."""
    from xrdslam_b200 import _cabi
    from xrdslam_b200.octree_device import DeviceOctree
    lib = _cabi.lib()

    def host(vlist):
        t = lib.xrd_octree_create(256)
        for v in vlist:
            v = v.int().contiguous()
            lib.xrd_octree_insert(t, v.data_ptr(), v.shape[0])
        N = lib.xrd_octree_num_nodes(t)
        vo, ch, fe = torch.empty(N, 4), torch.empty(N, 8), torch.empty(N, 8, dtype=torch.int32)
        lib.xrd_octree_export(t, vo.data_ptr(), ch.data_ptr(), fe.data_ptr())
        lib.xrd_octree_destroy(t)
        return vo, ch, fe
    g = torch.Generator().manual_seed(0)
    for trial, (n1, n2) in enumerate([(50, 30), (2000, 1500), (1, 1), (5000, 5000), (300, 0)]):
        base = torch.randint(100, 150, (max(n1 // 4, 1), 3), generator=g)
        v1 = base[torch.randint(0, base.shape[0], (n1,), generator=g)] + \
            torch.randint(-2, 3, (n1, 3), generator=g)
        v2 = base[torch.randint(0, base.shape[0], (max(n2, 1),), generator=g)][:n2] + \
            torch.randint(-6, 7, (n2, 3), generator=g)
        if trial == 3:
            v1[:5] = torch.tensor([[255, 255, 255], [0, 0, 0], [255, 0, 128], [254, 255, 3],
                                   [128, 128, 128]])
        ref = host([v1, v2] if n2 else [v1])
        t = DeviceOctree(256)
        t.insert(v1)
        t.insert(v2)
        got = t.export()
        assert t.num_nodes() == ref[0].shape[0]
        for a, b in zip(ref, got):
            assert a.dtype == b.dtype and torch.equal(a, b), trial
