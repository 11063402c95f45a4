"""Host model of mma.sync.m16n8k8 (tf32) fragment layouts, used to validate the register
layer-chaining scheme planned for k_fused v5 without a GPU.

Fragment layouts (lane = 4 g + t, g = 0..7, t = 0..3; PTX ISA, m16n8k8 .tf32):
    A (16 x 8, row):  a0 = A[g][t]      a1 = A[g+8][t]    a2 = A[g][t+4]    a3 = A[g+8][t+4]
    B ( 8 x 8, col):  b0 = B[t][g]      b1 = B[t+4][g]
    C (16 x 8):       c0 = C[g][2t]     c1 = C[g][2t+1]   c2 = C[g+8][2t]   c3 = C[g+8][2t+1]

Claim checked here: the C fragments of H = X W1 can be used AS the A fragments of Y = H W2
without moving data between lanes, if k-step s of the second GEMM pairs physical k index t
with logical column 8 s + 2 t and physical k index t + 4 with logical column 8 s + 2 t + 1,
i.e. a0 = c0, a1 = c2, a2 = c1, a3 = c3 of tile s and the B fragment is read as
b0 = W2[8 s + 2 t][n], b1 = W2[8 s + 2 t + 1][n]."""
import numpy as np

rng = np.random.default_rng(0)
LANES = [(g, t) for g in range(8) for t in range(4)]


def mma(c, a, b):
    """c, a, b: per-lane fragment dicts -> c + A B, through the documented layouts."""
    A = np.zeros((16, 8)); B = np.zeros((8, 8)); C = np.zeros((16, 8))
    for (g, t) in LANES:
        A[g, t], A[g + 8, t], A[g, t + 4], A[g + 8, t + 4] = a[(g, t)]
        B[t, g], B[t + 4, g] = b[(g, t)]
        C[g, 2 * t], C[g, 2 * t + 1], C[g + 8, 2 * t], C[g + 8, 2 * t + 1] = c[(g, t)]
    D = C + A @ B
    return {(g, t): (D[g, 2 * t], D[g, 2 * t + 1], D[g + 8, 2 * t], D[g + 8, 2 * t + 1])
            for (g, t) in LANES}


def a_frag(X, k0):
    return {(g, t): (X[g, k0 + t], X[g + 8, k0 + t], X[g, k0 + t + 4], X[g + 8, k0 + t + 4])
            for (g, t) in LANES}


def b_frag(W, k0, n0):
    return {(g, t): (W[k0 + t, n0 + g], W[k0 + t + 4, n0 + g]) for (g, t) in LANES}


def gemm_tiles(X, W):
    """C fragments (one per 8-wide n tile) of X W for a 16-row tile, standard K order."""
    K, N = W.shape
    out = []
    for nt in range(N // 8):
        c = {l: (0.0, 0.0, 0.0, 0.0) for l in LANES}
        for ks in range(K // 8):
            c = mma(c, a_frag(X, 8 * ks), b_frag(W, 8 * ks, 8 * nt))
        out.append(c)
    return out


def chained(h_tiles, W2):
    """Y = H W2 with the C fragments of H reused as A fragments (permuted K order)."""
    K, N = W2.shape
    out = []
    for nt in range(N // 8):
        c = {l: (0.0, 0.0, 0.0, 0.0) for l in LANES}
        for s in range(K // 8):
            a = {l: (h_tiles[s][l][0], h_tiles[s][l][2], h_tiles[s][l][1], h_tiles[s][l][3])
                 for l in LANES}
            b = {(g, t): (W2[8 * s + 2 * t, 8 * nt + g], W2[8 * s + 2 * t + 1, 8 * nt + g])
                 for (g, t) in LANES}
            c = mma(c, a, b)
        out.append(c)
    return out


def to_matrix(tiles):
    M = np.zeros((16, 8 * len(tiles)))
    for nt, c in enumerate(tiles):
        for (g, t), v in c.items():
            M[g, 8 * nt + 2 * t], M[g, 8 * nt + 2 * t + 1] = v[0], v[1]
            M[g + 8, 8 * nt + 2 * t], M[g + 8, 8 * nt + 2 * t + 1] = v[2], v[3]
    return M


X = rng.normal(size=(16, 80))       # 16 points x 80 inputs (hash 32 + OneBlob 48)
W1 = rng.normal(size=(80, 32))
W2 = rng.normal(size=(32, 16))
h = gemm_tiles(X, W1)
assert np.allclose(to_matrix(h), X @ W1)
relu = [{l: tuple(max(v, 0.0) for v in c[l]) for l in LANES} for c in h]   # elementwise: layout-free
y = chained(relu, W2)
assert np.allclose(to_matrix(y), np.maximum(X @ W1, 0) @ W2)
# bank check of the permuted B reads: row stride ld = 4 (mod 32) is conflict-free
for ld in (36, 40):
    banks = sorted(((2 * t) * ld + g) % 32 for (g, t) in LANES)
    print('ld', ld, 'b0 banks distinct:', len(set(banks)) == 32)
print('C-fragment -> A-fragment chaining with the (t <-> 2t, t+4 <-> 2t+1) K order: OK')
