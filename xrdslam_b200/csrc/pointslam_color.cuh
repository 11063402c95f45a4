// Point-SLAM stage 'color' (SURVEY rows P5 colour half, P6, P7, P8): per-neighbour MLP,
// 128-wide softplus trunk, colour compositing/loss and the full backward.  Included by
// pointslam.cu inside namespace xrd::point.  Activations live in HBM as [feature][point]
// rows (leading dimension Pp = P rounded up to 4); neighbour columns are ordered
// n = j * Pp + p (j = neighbour slot) so the 8-way reduction strides by Pp.
// Layers are plain GEMMs (gemm.cuh); weight gradients use the shared k_dw (dw.cuh).

constexpr int CE = 20;    // colour position embedding: sin/cos of 20 frequencies -> 40
constexpr int CR = 10;    // relative-position embedding: 10 frequencies -> 20
constexpr int CNI = 2 * CR + CD;  // 52 neighbour-MLP inputs
constexpr int CW = 128;   // hidden width
constexpr int CX3 = 2 * CE + CW;  // 168 = cat[embedding, h] after block 2
constexpr float TWO_PI = 6.283185307179586f;

struct ColorP {
  KnnParams K;
  int Pp;
  int min_nn;
  const float* col_feats;
  const float* rand_feat;
  XrdPointColorDecoder dec;
  // forward buffers
  float* wn;    // [8][Pp]
  float* Xn;    // [52][8 Pp]
  float* Fn;    // [32][8 Pp]
  float* cc;    // [32][Pp]
  float* X3;    // [168][Pp]
  const unsigned char* has_nb;
  // backward
  const float* dcc;   // [32][Pp]
  float* dFn;         // [32][8 Pp]
  const float* dXn;   // [52][8 Pp]
  const float* dX3;   // [168][Pp] rows 0..39: d embedding
  float* dp;          // [3][P] (leading dimension P, shared with the geometry path)
  int need_dp;
  float* d_col_feats;
  float* d_B_rel;
};

__device__ __forceinline__ float rel_angle(const float rel[3], const float* B, int n, int m) {
  // (2 pi x) @ B : scale each coordinate first (decoder_pointslam.py:41)
  return fmaf(TWO_PI * rel[2], B[2 * n + m], fmaf(TWO_PI * rel[1], B[n + m], (TWO_PI * rel[0]) * B[m]));
}

// neighbour weights (same expression as nb_weights) + inputs of the per-neighbour MLP
__global__ void __launch_bounds__(128) k_nb_build(const ColorP C) {
  __shared__ float sB[3 * CR];
  for (int i = threadIdx.x; i < 3 * CR; i += blockDim.x) sB[i] = C.dec.B_rel[i];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= C.K.P) return;
  float q[3];
  query_point(C.K, p, q);
  const float r = C.K.radius[p / C.K.radius_div];
  const float r2 = r * r;
  float a[KNN]; int id[KNN];
  float A = 0.f;
#pragma unroll
  for (int j = 0; j < KNN; ++j) {
    id[j] = C.K.I[(size_t)p * KNN + j];
    float aj = 0.f;
    if (id[j] >= 0) {
      const float d2 = sqdist(q, C.K.ix.pos + (size_t)id[j] * 3);
      aj = (d2 > r2) ? 0.f : 1.0f / (d2 + 1e-10f);
    }
    a[j] = aj;
    A += aj;
  }
  A = fmaxf(A, 1e-12f);
  const size_t Np = (size_t)KNN * C.Pp;
#pragma unroll 1
  for (int j = 0; j < KNN; ++j) {
    const size_t n = (size_t)j * C.Pp + p;
    const float w = a[j] / A;
    C.wn[n] = w;
    if (w != 0.f) {
      const float* x = C.K.ix.pos + (size_t)id[j] * 3;
      const float rel[3] = {x[0] - q[0], x[1] - q[1], x[2] - q[2]};
#pragma unroll
      for (int m = 0; m < CR; ++m) {
        float sn, cs;
        sincosf(rel_angle(rel, sB, CR, m), &sn, &cs);
        C.Xn[(size_t)m * Np + n] = sn;
        C.Xn[(size_t)(CR + m) * Np + n] = cs;
      }
      const float4* f = reinterpret_cast<const float4*>(C.col_feats + (size_t)id[j] * CD);
#pragma unroll
      for (int m4 = 0; m4 < CD / 4; ++m4) {
        const float4 v = __ldg(&f[m4]);
        C.Xn[(size_t)(2 * CR + 4 * m4) * Np + n] = v.x; C.Xn[(size_t)(2 * CR + 4 * m4 + 1) * Np + n] = v.y;
        C.Xn[(size_t)(2 * CR + 4 * m4 + 2) * Np + n] = v.z; C.Xn[(size_t)(2 * CR + 4 * m4 + 3) * Np + n] = v.w;
      }
    } else {
      // zero-weight slot (missing neighbour / beyond the radius): contributes exactly 0
      for (int m = 0; m < CNI; ++m) C.Xn[(size_t)m * Np + n] = 0.f;
    }
  }
}

// c = sum_k w_k f_k (random feature for samples with < min_nn neighbours, Q6) + embedding
__global__ void __launch_bounds__(128) k_nb_reduce(const ColorP C) {
  __shared__ float sB[3 * CE];
  for (int i = threadIdx.x; i < 3 * CE; i += blockDim.x) sB[i] = C.dec.B[i];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= C.K.P) return;
  const size_t Np = (size_t)KNN * C.Pp;
  const bool has = C.has_nb[p];
  float w[KNN];
#pragma unroll
  for (int j = 0; j < KNN; ++j) w[j] = C.wn[(size_t)j * C.Pp + p];
  for (int m = 0; m < CD; ++m) {
    float v;
    if (has) {
      v = 0.f;
#pragma unroll
      for (int j = 0; j < KNN; ++j) v = fmaf(w[j], C.Fn[(size_t)m * Np + (size_t)j * C.Pp + p], v);
    } else {
      v = C.rand_feat ? C.rand_feat[m] : 0.f;
    }
    C.cc[(size_t)m * C.Pp + p] = v;
  }
  float q[3];
  query_point(C.K, p, q);
#pragma unroll
  for (int m = 0; m < CE; ++m) {
    float sn, cs;
    sincosf(rel_angle(q, sB, CE, m), &sn, &cs);
    C.X3[(size_t)m * C.Pp + p] = sn;
    C.X3[(size_t)(CE + m) * C.Pp + p] = cs;
  }
}

// d c -> d f_k (= w_k dc) and d w_k -> d p   (w = a / sum a, a = 1 / (D + eps))
__global__ void __launch_bounds__(128) k_nb_reduce_bwd(const ColorP C) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= C.K.P) return;
  const size_t Np = (size_t)KNN * C.Pp;
  const bool has = C.has_nb[p];
  float w[KNN], dw[KNN];
#pragma unroll
  for (int j = 0; j < KNN; ++j) { w[j] = has ? C.wn[(size_t)j * C.Pp + p] : 0.f; dw[j] = 0.f; }
  for (int m = 0; m < CD; ++m) {
    const float g = has ? C.dcc[(size_t)m * C.Pp + p] : 0.f;
#pragma unroll
    for (int j = 0; j < KNN; ++j) {
      const size_t n = (size_t)m * Np + (size_t)j * C.Pp + p;
      dw[j] = fmaf(g, C.Fn[n], dw[j]);
      C.dFn[n] = w[j] * g;
    }
  }
  if (!C.need_dp || !has) return;
  float q[3];
  query_point(C.K, p, q);
  float sw = 0.f, A = 0.f, a[KNN];
  const float r = C.K.radius[p / C.K.radius_div];
  const float r2 = r * r;
#pragma unroll
  for (int j = 0; j < KNN; ++j) {
    const int id = C.K.I[(size_t)p * KNN + j];
    a[j] = 0.f;
    if (id >= 0) {
      const float d2 = sqdist(q, C.K.ix.pos + (size_t)id * 3);
      a[j] = (d2 > r2) ? 0.f : 1.0f / (d2 + 1e-10f);
    }
    A += a[j];
    sw += w[j] * dw[j];
  }
  A = fmaxf(A, 1e-12f);
  float g[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < KNN; ++j) {
    if (a[j] == 0.f) continue;
    const int id = C.K.I[(size_t)p * KNN + j];
    const float dD = -((dw[j] - sw) / A) * a[j] * a[j];
    const float* x = C.K.ix.pos + (size_t)id * 3;
#pragma unroll
    for (int d = 0; d < 3; ++d) g[d] += dD * (-2.f) * (x[d] - q[d]);
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) C.dp[(size_t)d * C.K.P + p] += g[d];
}

// d X_n -> d col_feats, d B_rel, d p (relative positions); d embedding -> d p
__global__ void __launch_bounds__(128) k_nb_build_bwd(const ColorP C) {
  __shared__ float sB[3 * CR], sBe[3 * CE], sdB[3 * CR];
  for (int i = threadIdx.x; i < 3 * CR; i += blockDim.x) { sB[i] = C.dec.B_rel[i]; sdB[i] = 0.f; }
  for (int i = threadIdx.x; i < 3 * CE; i += blockDim.x) sBe[i] = C.dec.B[i];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  // d B_rel partial sums stay in registers (3 x CR per thread) and are reduced per warp at the
  // end: one shared-memory atomic per (warp, entry) instead of one per (thread, neighbour, entry)
  // -- every thread of the CTA used to hammer the same 30 shared-memory words
  float dB[3][CR];
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int m = 0; m < CR; ++m) dB[d][m] = 0.f;
  if (p < C.K.P) {
    float q[3];
    query_point(C.K, p, q);
    const size_t Np = (size_t)KNN * C.Pp;
    float g[3] = {0.f, 0.f, 0.f};
    if (C.has_nb[p]) {
#pragma unroll 1
      for (int j = 0; j < KNN; ++j) {
        const size_t n = (size_t)j * C.Pp + p;
        if (C.wn[n] == 0.f) continue;
        const int id = C.K.I[(size_t)p * KNN + j];
        if (C.d_col_feats)
#pragma unroll
          for (int m4 = 0; m4 < CD / 4; ++m4)
            red_add_v4(C.d_col_feats + (size_t)id * CD + 4 * m4,
                       C.dXn[(size_t)(2 * CR + 4 * m4) * Np + n], C.dXn[(size_t)(2 * CR + 4 * m4 + 1) * Np + n],
                       C.dXn[(size_t)(2 * CR + 4 * m4 + 2) * Np + n], C.dXn[(size_t)(2 * CR + 4 * m4 + 3) * Np + n]);
        if (!C.d_B_rel && !C.need_dp) continue;
        const float* x = C.K.ix.pos + (size_t)id * 3;
        const float rel[3] = {x[0] - q[0], x[1] - q[1], x[2] - q[2]};
#pragma unroll
        for (int m = 0; m < CR; ++m) {
          float sn, cs;
          sincosf(rel_angle(rel, sB, CR, m), &sn, &cs);
          const float da = C.dXn[(size_t)m * Np + n] * cs - C.dXn[(size_t)(CR + m) * Np + n] * sn;
          if (da == 0.f) continue;
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            dB[d][m] += TWO_PI * rel[d] * da;
            g[d] -= TWO_PI * sB[d * CR + m] * da;  // rel = x - p
          }
        }
      }
    }
    if (C.need_dp) {
#pragma unroll
      for (int m = 0; m < CE; ++m) {
        float sn, cs;
        sincosf(rel_angle(q, sBe, CE, m), &sn, &cs);
        const float da = C.dX3[(size_t)m * C.Pp + p] * cs - C.dX3[(size_t)(CE + m) * C.Pp + p] * sn;
#pragma unroll
        for (int d = 0; d < 3; ++d) g[d] += TWO_PI * sBe[d * CE + m] * da;
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) C.dp[(size_t)d * C.K.P + p] += g[d];
    }
  }
  if (C.d_B_rel) {
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int m = 0; m < CR; ++m) {
        const float v = warp_sum(dB[d][m]);
        if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(&sdB[d * CR + m], v);
      }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * CR; i += blockDim.x)
      if (sdB[i] != 0.f) atomicAdd(C.d_B_rel + i, sdB[i]);
  }
}

// dPre = dH * softplus'(pre) recovered from the activation output: sigmoid(beta x) = 1 - exp(-beta y)
__global__ void __launch_bounds__(256) k_dsoftplus(size_t n, const float* dH, const float* act, float* dPre) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dPre[i] = dH[i] * -expm1f(-100.f * act[i]);
}
