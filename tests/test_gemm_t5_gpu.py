"""The Blackwell-native GEMM (csrc/gemm_t5.cuh: tcgen05.mma kind::tf32 + TMEM + TMA, 3xTF32)
against float64 matmul and against the mma.sync path, on the shapes of the Vox-Fusion decoder and
the Point-SLAM colour trunk (forward and transposed-weight backward GEMMs, ragged N, masks)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

ACT = {0: lambda x: x, 1: torch.relu, 3: torch.sigmoid}


def run(dev, M, N, K, transA, act, mode, mask=False, addend=False, seed=0):
    from xrdslam_b200 import _cabi
    lib = _cabi.lib()
    g = torch.Generator().manual_seed(seed)
    A = torch.randn((K, M) if transA else (M, K), generator=g) / K ** 0.5
    Np = (N + 63) // 64 * 64
    B = torch.randn(K, Np, generator=g)
    bias = torch.randn(M, generator=g) * 0.1
    mk = torch.randn(M, Np, generator=g) if mask else None
    ad = torch.randn(M, Np, generator=g) if addend else None
    d = lambda t: t.to(dev).contiguous() if t is not None else None
    A_d, B_d, b_d, mk_d, ad_d = d(A), d(B), d(bias), d(mk), d(ad)
    C = torch.full((M, Np), float('nan'), device=dev)
    _cabi.check('mode', lib.xrd_debug_gemm_mode(mode))
    try:
        st = lib.xrd_debug_gemm(M, N, K, A_d.data_ptr(), A.shape[1], int(transA), B_d.data_ptr(), Np,
                                C.data_ptr(), Np, b_d.data_ptr(), act,
                                mk_d.data_ptr() if mask else None, Np,
                                ad_d.data_ptr() if addend else None, Np, None)
        _cabi.check('xrd_debug_gemm', st)
        torch.cuda.synchronize()
    finally:
        lib.xrd_debug_gemm_mode(1)
    Am = (A.t() if transA else A).double()
    ref = ACT[act](Am @ B.double()[:, :N] + bias.double()[:, None])
    if mask:
        ref = torch.where(mk[:, :N] > 0, ref, torch.zeros_like(ref))
    if addend:
        ref = ref + ad[:, :N].double()
    return C.cpu()[:, :N].double(), ref, C.cpu()[:, N:]


@pytest.mark.parametrize('M,N,K,transA,act', [
    (128, 4096, 128, 0, 1),      # trunk layer
    (128, 5000, 16, 0, 1),       # vox layer 0 (K = 16, ragged N)
    (128, 1111, 144, 0, 1),      # colour layer: K = 144 -> two-stage pipeline
    (128, 2048, 129, 1, 0),      # backward through sdf_out (transposed weights, K = 129)
    (144, 3000, 128, 1, 0),      # 144 rows = 128 (tcgen05) + 16 (mma.sync)
    (64, 1024, 128, 0, 3),       # M < 128 (zero-padded rows)
    (32, 3000, 128, 0, 0),       # Point-SLAM 128 -> 32 neighbour layer
    (128, 150000, 128, 0, 1),    # benchmark-sized point count
])
def test_gemm_t5_matches_float64(cuda_dev, M, N, K, transA, act):
    got, ref, pad = run(cuda_dev, M, N, K, transA, act, mode=1)
    err = (got - ref).abs().max().item()
    assert err < 5e-6 * max(1.0, ref.abs().max().item()), err   # fp32-level (3xTF32)
    assert torch.isnan(pad).all()                               # columns >= N untouched
    legacy, _, _ = run(cuda_dev, M, N, K, transA, act, mode=3)
    assert (got - legacy).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())


def test_gemm_t5_mask_and_addend(cuda_dev):
    got, ref, _ = run(cuda_dev, 128, 3333, 128, 1, 0, mode=1, mask=True, addend=True)
    assert (got - ref).abs().max().item() < 5e-6 * max(1.0, ref.abs().max().item())


def test_gemm_t5_sass_is_blackwell_native():
    """The shipped library contains tcgen05 MMAs, TMEM loads and TMA loads (SASS mnemonics)."""
    import os
    import shutil
    import subprocess
    from xrdslam_b200.build import LIB_PATH
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([cuobjdump, '-sass', LIB_PATH], capture_output=True, text=True).stdout
    for mn in ('UTCHMMA', 'LDTM', 'UTMALDG'):
        assert mn in sass, mn
