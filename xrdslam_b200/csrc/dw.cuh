// Generic weight-gradient kernel shared by the per-point MLP pipelines (NICE-SLAM, Vox-Fusion,
// Point-SLAM): activations and their gradients live in HBM in [feature][point] order; a job
// accumulates  out[j*sj + i*si] += sum_p B_j[p] * A_i[p]  (and bias[j] += sum_p B_j[p]) with a
// shared-memory tiled SIMT GEMM over a chunk of points per CTA and red.global.add at the end.
#pragma once
#include "common.cuh"

namespace xrd {
// out[j*sj + i*si] += sum_p Brow_j[p] (masked) * Arow_i[p];  bias[j] += sum_p Brow_j[p] (masked)
struct DwJob {
  const float* A; int nA;          // rows of Pp floats
  const float* B; int nB;
  const uint32_t* mask;            // per-point relu mask word (bit j) or NULL
  float* out; int sj, si;
  float* bias;                     // or NULL
};
constexpr int DW_MAX_JOBS = 32;
struct DwParams {
  DwJob jobs[DW_MAX_JOBS];
  int n_jobs, P, Pp, chunk;
};

static __global__ void __launch_bounds__(256) k_dw(const DwParams Q) {
  __shared__ float As[128][65];
  __shared__ float Bs[32][65];
  const int tid = threadIdx.x, ta = tid >> 3, tb = tid & 7;
  const int p_lo = blockIdx.x * Q.chunk, p_hi = min(Q.P, p_lo + Q.chunk);
  if (p_lo >= p_hi) return;
  {  // one (chunk of points, job) pair per CTA: grid = (chunks, n_jobs)
    const DwJob J = Q.jobs[blockIdx.y];
    for (int a0 = 0; a0 < J.nA; a0 += 128) {
      const int na = min(128, J.nA - a0);
      float acc[4][4];
      float bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int p0 = p_lo; p0 < p_hi; p0 += 64) {
        const int np = min(64, p_hi - p0);
        __syncthreads();
        for (int q = tid; q < 128 * 64; q += 256) {
          const int row = q >> 6, pp = q & 63;
          As[row][pp] = (row < na && pp < np) ? J.A[(size_t)(a0 + row) * Q.Pp + p0 + pp] : 0.f;
        }
        for (int q = tid; q < 32 * 64; q += 256) {
          const int row = q >> 6, pp = q & 63;
          float v = 0.f;
          if (row < J.nB && pp < np) {
            v = J.B[(size_t)row * Q.Pp + p0 + pp];
            if (J.mask && !((J.mask[p0 + pp] >> row) & 1u)) v = 0.f;
          }
          Bs[row][pp] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int pp = 0; pp < 64; ++pp) {
          float av[4], bv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) av[i] = As[4 * ta + i][pp];
#pragma unroll
          for (int j = 0; j < 4; ++j) bv[j] = Bs[4 * tb + j][pp];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
          if (ta == 0)
#pragma unroll
            for (int j = 0; j < 4; ++j) bacc[j] += bv[j];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int ia = a0 + 4 * ta + i, jbk = 4 * tb + j;
          if (4 * ta + i < na && jbk < J.nB && acc[i][j] != 0.f)
            red_add(J.out + (size_t)jbk * J.sj + (size_t)ia * J.si, acc[i][j]);
        }
      if (ta == 0 && a0 == 0 && J.bias)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (4 * tb + j < J.nB) red_add(J.bias + 4 * tb + j, bacc[j]);
    }
  }
}

// chunk: as many points per CTA as still give >= ~4 CTAs per SM over (chunks x jobs); a
// multiple of the 64-point staging tile.  Fewer, longer chunks mean fewer red.global.add.
static inline cudaError_t launch_dw(DwParams& Q, cudaStream_t stream) {
  if (Q.n_jobs <= 0 || Q.P <= 0) return cudaSuccess;
  const long long want = 4LL * 148;
  long long chunks = (want + Q.n_jobs - 1) / Q.n_jobs;
  if (chunks < 1) chunks = 1;
  long long chunk = ((long long)Q.P + chunks - 1) / chunks;
  chunk = (chunk + 63) / 64 * 64;
  if (chunk < 128) chunk = 128;
  if (chunk > 4096) chunk = 4096;
  Q.chunk = (int)chunk;
  dim3 grid((unsigned)((Q.P + Q.chunk - 1) / Q.chunk), (unsigned)Q.n_jobs);
  k_dw<<<grid, 256, 0, stream>>>(Q);
  return cudaGetLastError();
}

}  // namespace xrd
