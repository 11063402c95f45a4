"""On-disk formats of SURVEY row f4 (tracker.py:258-278,388-420 -> scripts/eval.py): eval.tar
and the PLY artefacts; CPU only."""
import numpy as np
import torch

from xrdslam_b200 import io_formats as io


class _Algo:
    def __init__(self, gt, est):
        self.gt, self.est = gt, est

    def get_gt_c2w_list_ori(self):
        return self.gt

    def get_gt_c2w_list(self):
        return [g.clone() for g in self.gt]

    def get_estimate_c2w_list(self):
        return self.est


def _traj(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n):
        a = torch.randn(3, 3, generator=g)
        q, _ = torch.linalg.qr(a)
        if torch.det(q) < 0:
            q[:, 0] = -q[:, 0]
        T = torch.eye(4)
        T[:3, :3] = q
        T[:3, 3] = torch.tensor([0.05 * i, 0.3 * np.sin(0.2 * i), 0.1 * np.cos(0.1 * i)]).float()
        out.append(T)
    return out


def test_eval_tar_is_what_ds_eval_reads(tmp_path):
    gt = _traj(12)
    gt[5] = torch.full((4, 4), float('nan'))   # ScanNet-style invalid ground truth
    est = _traj(12, seed=1)
    path = io.save_eval_tar(_Algo(gt, est), str(tmp_path), 11)
    # scripts/eval.py:42-46, verbatim access pattern
    ckpt = torch.load(path, map_location=torch.device('cpu'), weights_only=False)
    assert set(ckpt) == {'gt_c2w_list_ori', 'gt_c2w_list', 'estimate_c2w_list', 'idx'}
    assert int(ckpt['idx']) == 11 and len(ckpt['estimate_c2w_list']) == 12
    assert all(t.shape == (4, 4) for t in ckpt['gt_c2w_list_ori'])
    with open(path, 'rb') as fh:            # legacy serialization, not a zip archive
        assert fh.read(2) != b'PK'
    e, g, n = io.load_eval_tar(path)
    assert n == 11 and torch.equal(e[3], est[3])
    m = io.valid_pose_mask(g, n)
    assert m.sum() == 10 and not m[5]


def test_ate_alignment_recovers_a_rigid_transform():
    gt = _traj(40)
    R = torch.tensor([[0., -1., 0.], [1., 0., 0.], [0., 0., 1.]])
    t = torch.tensor([0.5, -2.0, 1.0])
    est = []
    for T in gt:
        E = T.clone()
        E[:3, 3] = R.T @ (T[:3, 3] - t)    # est = R^T (gt - t)  =>  gt = R est + t
        est.append(E)
    r = io.ate_rmse(gt, est)
    assert r['rmse'] < 1e-6
    assert np.allclose(r['rot'], R.numpy(), atol=1e-6) and np.allclose(r['trans'], t.numpy(), atol=1e-6)
    noisy = [E.clone() for E in est]
    for E in noisy:
        E[:3, 3] += 0.01
    assert io.ate_rmse(gt, noisy)['rmse'] < 1e-6          # a constant offset is absorbed
    noisy[7][:3, 3] += torch.tensor([0.3, 0., 0.])
    assert 0.01 < io.ate_rmse(gt, noisy)['rmse'] < 0.3


def test_ply_mesh_and_cloud_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    v = rng.normal(size=(50, 3)).astype(np.float32)
    f = rng.integers(0, 50, size=(80, 3)).astype(np.int32)
    c = rng.random((50, 3))
    p = io.write_ply(str(tmp_path / 'final_mesh.ply'), v, f, c)
    head = open(p, 'rb').read(400).decode('ascii', 'ignore')
    assert 'element vertex 50' in head and 'property list uchar int vertex_indices' in head
    v2, f2, c2 = io.read_ply(p)
    assert np.array_equal(v, v2) and np.array_equal(f, f2)
    assert np.array_equal(c2, np.clip(np.round(c * 255), 0, 255).astype(np.uint8))
    p = io.write_ply(str(tmp_path / 'cloud' / '00003.ply'), v, None, (c * 255).astype(np.uint8))
    v3, f3, c3 = io.read_ply(p)
    assert f3 is None and np.array_equal(v3, v) and c3.shape == (50, 3)
    v4, f4, c4 = io.read_ply(io.write_ply(str(tmp_path / 'bare.ply'), v, f))
    assert c4 is None and np.array_equal(f4, f)
