"""Bring-up probe of the tcgen05 GEMM: error of every descriptor variant on a few shapes, then
structured inputs (identity weights, index-coded activations) that expose layout mix-ups."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xrdslam_b200 import _cabi
lib = _cabi.lib()
dev = torch.device('cuda:0')


def gemm(A, B, N, mode, transA=0, variant=0):
    M = A.shape[1] if transA else A.shape[0]
    K = A.shape[0] if transA else A.shape[1]
    Np = B.shape[1]
    C = torch.full((M, Np), float('nan'), device=dev)
    lib.xrd_debug_gemm_mode(mode)
    lib.xrd_debug_gemm_variant(variant)
    st = lib.xrd_debug_gemm(M, N, K, A.data_ptr(), A.shape[1], transA, B.data_ptr(), Np, C.data_ptr(), Np,
                            None, 0, None, 0, None, 0, None)
    torch.cuda.synchronize()
    lib.xrd_debug_gemm_mode(1)
    lib.xrd_debug_gemm_variant(0)
    assert st == 0, st
    return C[:, :N]


g = torch.Generator().manual_seed(0)
for variant in (0, 1):
    for (M, N, K) in ((128, 1024, 16), (128, 1024, 128), (128, 1024, 144)):
        A = (torch.randn(M, K, generator=g) / K ** 0.5).to(dev)
        B = torch.randn(K, N, generator=g).to(dev)
        try:
            C = gemm(A, B, N, 1, variant=variant)
            ref = A.double() @ B.double()
            e = (C.double() - ref).abs().max().item()
            print(f'variant {variant} M{M} N{N} K{K}: max err {e:.3e}  (nan: {int(torch.isnan(C).sum())})', flush=True)
        except Exception as ex:  # noqa
            print(f'variant {variant} M{M} N{N} K{K}: EXC {ex!r}', flush=True)
            sys.exit(0)

# structured probes, variant 0: A = [I_16 ; 0]  ->  C[m][n] = B[m][n] (m < 16)
M, N, K = 128, 512, 16
A = torch.zeros(M, K, device=dev); A[:K, :K] = torch.eye(K, device=dev)
B = (torch.arange(K, device=dev)[:, None] * 1000 + torch.arange(N, device=dev)[None, :]).float()
for variant in (0, 1):
    C = gemm(A, B, N, 1, variant=variant)
    print(f'probe identity, variant {variant}: C[0:4, 0:8] =\n', C[:4, :8].cpu())
    print('  C[0:4, 256:260] =\n', C[:4, 256:260].cpu())
    print('  C[16:18, 0:4] =\n', C[16:18, :4].cpu())
    # which (k, n) does each output element hold?  value = 1000 k + n
    kk = (C[:16, :N] / 1000).floor().long().cpu(); nn = (C[:16, :N] % 1000).long().cpu()
    print('  rows -> k :', kk[:, 0].tolist())
    print('  cols 0..15 -> n (row 0):', nn[0, :16].tolist())
# A probe: B = [I ; ...]: C[m][n] = A[m][n] for n < K
A = (torch.arange(M, device=dev)[:, None] * 100 + torch.arange(K, device=dev)[None, :]).float()
B = torch.zeros(K, N, device=dev); B[:, :K] = torch.eye(K, device=dev)
C = gemm(A, B, N, 1)
print('probe A: C[0:3, 0:8] =\n', C[:3, :8].cpu(), '\n C[64:66, 0:4]=\n', C[64:66, :4].cpu())
