"""A/B timing of the tcgen05 GEMM revisions (csrc/gemm_t5.cuh) on the Vox-Fusion / Point-SLAM
layer shapes: variant 0 = revision 2 (default), variant 2 = revision 1; mode 3 = mma.sync 3xTF32."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xrdslam_b200 import _cabi

lib = _cabi.lib()
dev = torch.device('cuda:0')
N = int(os.environ.get('T5_N', 133540))
Np = (N + 63) // 64 * 64
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(M, K, transA, act, mask, addend, mode, variant, iters=10):
    g = torch.Generator().manual_seed(0)
    A = (torch.randn((K, M) if transA else (M, K), generator=g) / K ** 0.5).to(dev)
    B = torch.randn(K, Np, generator=g).to(dev)
    bias = torch.zeros(M, device=dev)
    C = torch.empty(M, Np, device=dev)
    mk = torch.randn(M, Np, generator=g).to(dev) if mask else None
    ad = torch.randn(M, Np, generator=g).to(dev) if addend else None
    lib.xrd_debug_gemm_mode(mode)
    lib.xrd_debug_gemm_variant(variant)

    def run():
        _cabi.check('gemm', lib.xrd_debug_gemm(
            M, N, K, A.data_ptr(), A.shape[1], int(transA), B.data_ptr(), Np, C.data_ptr(), Np,
            bias.data_ptr(), act, mk.data_ptr() if mask else None, Np,
            ad.data_ptr() if addend else None, Np, None))
    for _ in range(3):
        run()
    ms = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    lib.xrd_debug_gemm_mode(1)
    lib.xrd_debug_gemm_variant(0)
    ref = (A.t() if transA else A).double() @ B[:, :4096].double()
    got = C[:, :4096].double()
    if act == 1:
        ref = torch.relu(ref)
    if mask:
        ref = torch.where(mk[:, :4096] > 0, ref, torch.zeros_like(ref))
    if addend:
        ref = ref + ad[:, :4096].double()
    err = (got - ref).abs().max().item()
    us = ms / iters * 1e3
    gb = (K + M * (1 + mask + addend)) * N * 4 / 1e9
    return us, err, gb / (us * 1e-6)


if os.environ.get('T5_NCU'):
    # one launch per revision, no warm-up (for `ncu -k regex:k_gemm_t5 -c 4`)
    for (M, K, transA, act, mask, addend) in [(128, 128, 0, 1, 0, 0), (128, 128, 1, 0, 1, 0)]:
        for variant in (0, 2):
            g = torch.Generator().manual_seed(0)
            A = (torch.randn((K, M) if transA else (M, K), generator=g) / K ** 0.5).to(dev)
            B = torch.randn(K, Np, generator=g).to(dev)
            C = torch.empty(M, Np, device=dev)
            mk = torch.randn(M, Np, generator=g).to(dev) if mask else None
            lib.xrd_debug_gemm_variant(variant)
            _cabi.check('gemm', lib.xrd_debug_gemm(
                M, N, K, A.data_ptr(), A.shape[1], int(transA), B.data_ptr(), Np, C.data_ptr(), Np,
                None, act, mk.data_ptr() if mask else None, Np, None, Np, None))
            torch.cuda.synchronize()
    lib.xrd_debug_gemm_variant(0)
    sys.exit(0)

for (M, K, transA, act, mask, addend) in [(128, 128, 0, 1, 0, 0), (128, 16, 0, 1, 0, 0),
                                           (128, 144, 0, 1, 0, 0), (128, 128, 1, 0, 1, 0),
                                           (128, 129, 1, 0, 1, 1)]:
    for name, mode, variant in (('t5 rev2', 1, 0), ('t5 rev1', 1, 2), ('mma.sync 3xTF32', 3, 0)):
        us, err, gbs = bench(M, K, transA, act, mask, addend, mode, variant)
        print(f'M={M} K={K} N={N} transA={transA} mask={mask} addend={addend}  {name:16s} '
              f'{us:8.1f} us  {gbs:7.0f} GB/s (algorithmic)  max err {err:.2e}', flush=True)
