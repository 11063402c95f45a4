// Generic weight-gradient kernel shared by the per-point MLP pipelines (NICE-SLAM, Vox-Fusion,
// Point-SLAM): activations and their gradients live in HBM in [feature][point] order; a job
// accumulates  out[j*sj + i*si] += sum_p B_j[p] * A_i[p]  (and bias[j] += sum_p B_j[p]) with a
// shared-memory tiled GEMM over a chunk of points per CTA and red.global.add at the end.
//   k_dw_tc (default): tensor cores, mma.sync m16n8k8 TF32 with the 3xTF32 split (fp32-level
//            accuracy): both operands are K-major as they lie in HBM (K = points), so the
//            fragments come straight out of the staged [row][64 points] tiles, conflict-free
//            with a row pitch of 68 floats.  32 x 128 outputs per CTA, 8 warps x (2 x 2) tiles.
//   k_dw    (xrd_debug_gemm_mode 0): the fp32 SIMT reference version.
#pragma once
#include "common.cuh"

namespace xrd {
extern thread_local int g_gemm_mode;  // cabi.cu (0 SIMT, 1 default, 2 TF32, 3 3xTF32 mma.sync)
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

// ---- tensor-core version --------------------------------------------------------------------
__device__ __forceinline__ void dw_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
constexpr int DW_LD = 68;  // floats per staged row: 64 points + 4 (== 4 mod 32: conflict-free fragments)

template <bool PREC3>
static __global__ void __launch_bounds__(256) k_dw_tc(const DwParams Q) {
  __shared__ __align__(16) float As[128 * DW_LD];  // rows of A (the MMA N side)
  __shared__ __align__(16) float Bs[32 * DW_LD];   // rows of B (the MMA M side), relu-masked
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int p_lo = blockIdx.x * Q.chunk, p_hi = min(Q.P, p_lo + Q.chunk);
  if (p_lo >= p_hi) return;
  const DwJob J = Q.jobs[blockIdx.y];
  const bool vec = ((Q.Pp & 3) == 0) && ((((uintptr_t)J.A) & 15) == 0) && ((((uintptr_t)J.B) & 15) == 0);
  for (int a0 = 0; a0 < J.nA; a0 += 128) {
    const int na = min(128, J.nA - a0);
    const bool warp_on = warp * 16 < na;
    float acc[2][2][4];
    float bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
    for (int p0 = p_lo; p0 < p_hi; p0 += 64) {
      __syncthreads();
      // stage 128 rows of A and 32 rows of B, 64 points each (16 float4 per row)
#pragma unroll
      for (int q = 0; q < 10; ++q) {
        const int e = tid + q * 256;           // 0..2047: A, 2048..2559: B
        const bool isB = e >= 2048;
        const int row = isB ? (e - 2048) >> 4 : e >> 4;
        const int c4 = (e & 15) * 4, pp = p0 + c4;
        const bool row_ok = isB ? row < J.nB : row < na;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_ok && pp < p_hi) {
          const float* src = (isB ? J.B + (size_t)row * Q.Pp : J.A + (size_t)(a0 + row) * Q.Pp) + pp;
          if (vec && pp + 3 < p_hi) {
            v = *reinterpret_cast<const float4*>(src);
          } else {
            v.x = src[0];
            if (pp + 1 < p_hi) v.y = src[1];
            if (pp + 2 < p_hi) v.z = src[2];
            if (pp + 3 < p_hi) v.w = src[3];
          }
          if (isB && J.mask) {
            if (!((J.mask[pp] >> row) & 1u)) v.x = 0.f;
            if (pp + 1 < p_hi && !((J.mask[pp + 1] >> row) & 1u)) v.y = 0.f;
            if (pp + 2 < p_hi && !((J.mask[pp + 2] >> row) & 1u)) v.z = 0.f;
            if (pp + 3 < p_hi && !((J.mask[pp + 3] >> row) & 1u)) v.w = 0.f;
          }
        }
        *reinterpret_cast<float4*>((isB ? Bs : As) + row * DW_LD + c4) = v;
      }
      __syncthreads();
      if (a0 == 0 && J.bias) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float* b = Bs + (warp * 4 + r) * DW_LD;
          bacc[r] += b[lane] + b[lane + 32];
        }
      }
      if (!warp_on) continue;
#pragma unroll 2
      for (int ks = 0; ks < 64; ks += 8) {
        uint32_t ab[2][4], as_[2][4], bb[2][2], bs[2][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const float* a = Bs + (mt * 16 + g) * DW_LD + ks + t;
          const float av[4] = {a[0], a[8 * DW_LD], a[4], a[8 * DW_LD + 4]};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (PREC3) {
              ab[mt][i] = __float_as_uint(av[i]) & 0xffffe000u;
              as_[mt][i] = __float_as_uint(av[i] - __uint_as_float(ab[mt][i]));
            } else {
              ab[mt][i] = __float_as_uint(av[i]);
            }
          }
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const float* b = As + (warp * 16 + nt * 8 + g) * DW_LD + ks + t;
          const float v0 = b[0], v1 = b[4];
          if (PREC3) {
            bb[nt][0] = __float_as_uint(v0) & 0xffffe000u;
            bb[nt][1] = __float_as_uint(v1) & 0xffffe000u;
            bs[nt][0] = __float_as_uint(v0 - __uint_as_float(bb[nt][0]));
            bs[nt][1] = __float_as_uint(v1 - __uint_as_float(bb[nt][1]));
          } else {
            bb[nt][0] = __float_as_uint(v0); bb[nt][1] = __float_as_uint(v1);
          }
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            if (PREC3) {
              dw_mma(acc[mt][nt], as_[mt], bb[nt]);
              dw_mma(acc[mt][nt], ab[mt], bs[nt]);
            }
            dw_mma(acc[mt][nt], ab[mt], bb[nt]);
          }
      }
    }
    if (warp_on) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = mt * 16 + g + ((i & 2) ? 8 : 0);
            const int col = warp * 16 + nt * 8 + 2 * t + (i & 1);
            const float v = acc[mt][nt][i];
            if (j < J.nB && col < na && v != 0.f)
              red_add(J.out + (size_t)j * J.sj + (size_t)(a0 + col) * J.si, v);
          }
    }
    if (a0 == 0 && J.bias) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float sum = warp_sum(bacc[r]);
        if (lane == 0 && warp * 4 + r < J.nB) red_add(J.bias + warp * 4 + r, sum);
      }
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
  if (g_gemm_mode == 0) k_dw<<<grid, 256, 0, stream>>>(Q);
  else if (g_gemm_mode == 2) k_dw_tc<false><<<grid, 256, 0, stream>>>(Q);
  else k_dw_tc<true><<<grid, 256, 0, stream>>>(Q);
  return cudaGetLastError();
}

}  // namespace xrd
