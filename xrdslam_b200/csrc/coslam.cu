// Co-SLAM render-and-optimise step, fused forward + backward (sm_100a).
//
// Replaces (reference @ f0366f20): slam/models/joint_encoding.py:250-344
// render_rays, :483-507 run_network, :463-481 query_color_sdf (tcnn HashGrid +
// OneBlob, slam/model_components/encodings_coslam.py:39-75; decoders
// slam/model_components/decoder_coslam.py:139-163), :346-406 sdf2weights /
// raw2outputs, :94-147 get_loss_dict (+ slam/model_components/utils.py:100-186)
// and the autograd backward of the whole chain.
//
// Kernels
//   k_sample   one warp per ray: merge the 32 uniform + 11 depth-guided samples,
//              stratified jitter, batch-global counts n_fs / n_sdf / n_valid.
//   k_fused    persistent CTAs over tiles of NR rays (NR*S points, one thread per
//              point for the gathers).  Per-point fp32 record in shared memory:
//                [feat32 | blob48 | geo15 | sdf | h1_32 | c1_32 | raw4]   (stride 164)
//              The decoder runs on the tensor cores: every warp pushes its own 32
//              points through mma.sync m16n8k8 TF32 tiles whose A fragments are read
//              straight from the records (stride 164 = 4 mod 32 -> conflict-free) and
//              whose B fragments are the transposed weights staged in shared memory;
//              3xTF32 error compensation keeps fp32-level parity (cfg.precision).
//              forward -> per-ray composite + loss gradient (one warp per ray) ->
//              layer-by-layer backward that overwrites activations with their
//              gradients in place; weight gradients are 16x8 MMA tiles with the
//              points on the K axis, accumulated in registers for the whole launch;
//              the hash backward runs in the MMA fragment layout (lane (g,t) owns
//              level 4*nt+t of rows g, g+8) and scatters with red.global.add.v2.f32.
//   k_finalize loss accumulators -> the 4 weighted loss terms.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace xrd {
namespace coslam {

constexpr int kL = XRD_MAX_LEVELS;
constexpr int kBins = 16;
// record layout (floats); stride == 4 (mod 32) keeps 128-bit LDS conflict-free
constexpr int R_FEAT = 0, R_BLOB = 32, R_GEO = 80, R_SDF = 95, R_H1 = 96, R_C1 = 128,
              R_RAW = 160, REC = 164;
struct GridDev {
  float scale[kL];
  uint32_t res[kL], size[kL], offset[kL], hashed[kL];
  uint32_t magic[kL];  // floor(2^32 / size): q = umulhi(index, magic) <= index/size, off by <= 1
  int n_levels;
  double bmin[3], binv[3];  // 1/(max-min): gradients only; the forward divides (index parity)
  double bmax[3];
};

struct Params {
  // rays
  int R, S, Rg;  // Rg: rays of the global (all-rank) batch the loss means run over
  const float *rays_o, *rays_d, *target_s, *target_d;
  const float* z_vals;  // [R,S] (written by k_sample)
  // grid + mlp
  GridDev g;
  const float* table;
  const float *w_sdf0, *w_sdf1, *w_col0, *w_col1;
  // cfg
  float trunc, depth_trunc, w_rgb, w_depth, w_sdf, w_fs;
  float ls[4];
  // outputs
  float *rgb, *depth, *disp, *acc, *depth_var, *raw;
  // grads
  float *d_table, *d_w_sdf0, *d_w_sdf1, *d_w_col0, *d_w_col1, *d_rays_o, *d_rays_d;
  // workspace
  const int* counts;  // n_fs, n_sdf, n_valid
  double* loss_acc;   // 4 sums
  int NR;             // rays per tile
  int n_tiles;
  float* jac;         // k_fused_g: [R*S][2 halves][8 levels][6] d feat / d xn, written by the forward
                      // gather when position gradients are wanted (replaces the backward re-gather)
};

// ------------------------------------------------------------------ sample ---
struct SampleParams {
  int R, S, n_a, n_b, perturb, has_depth;
  const float *target_d, *lin_uniform, *lin_range, *lin_nodepth, *lin_full, *noise;
  float trunc, depth_trunc;
  uint64_t seed;
  const uint64_t* seed_dev;
  float* z_vals;
  int* counts;
};

__global__ void __launch_bounds__(128) k_sample(SampleParams p) {
  __shared__ float zs[4][260];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + warp;
  int c_fs = 0, c_sdf = 0, c_valid = 0;
  if (r < p.R) {
    float* z = zs[warp];
    const int S = p.S;
    float d = 0.f;
    if (p.has_depth) {
      d = p.target_d[r];
      const bool nodepth = d <= 0.f;  // z_samples[target_d <= 0] = linspace(near, far)
      // rank-merge of the two sorted lists (values equal torch.sort's output)
      for (int k = lane; k < S; k += 32) {
        float v;
        int rank;
        if (k < p.n_a) {
          v = p.lin_uniform[k];
          int c = 0;
          for (int j = 0; j < p.n_b; ++j) {
            float b = nodepth ? p.lin_nodepth[j] : __fadd_rn(p.lin_range[j], d);
            c += (b < v);
          }
          rank = k + c;
        } else {
          int j = k - p.n_a;
          v = nodepth ? p.lin_nodepth[j] : __fadd_rn(p.lin_range[j], d);
          int c = 0;
          for (int i = 0; i < p.n_a; ++i) c += (p.lin_uniform[i] <= v);
          rank = j + c;
        }
        z[rank] = v;
      }
    } else {
      for (int k = lane; k < S; k += 32) z[k] = p.lin_full[k];
    }
    __syncwarp();
    for (int k = lane; k < S; k += 32) {
      float zv = z[k];
      if (p.perturb) {
        float lo = (k == 0) ? zv : __fmul_rn(0.5f, __fadd_rn(zv, z[k - 1]));
        float hi = (k == S - 1) ? zv : __fmul_rn(0.5f, __fadd_rn(z[k + 1], zv));
        float u;
        if (p.noise) {
          u = p.noise[(size_t)r * S + k];
        } else {
          float q[4];
          philox4(p.seed_dev ? *p.seed_dev : p.seed, (uint64_t)r * S + k, q);
          u = q[0];
        }
        zv = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), u));
      }
      p.z_vals[(size_t)r * S + k] = zv;
      if (p.has_depth) {
        const bool front = zv < __fsub_rn(d, p.trunc);
        const bool back = zv > __fadd_rn(d, p.trunc);
        c_fs += front;
        c_sdf += (!front && !back && d > 0.f);
      }
    }
    if (p.has_depth && lane == 0) c_valid = (d > 0.f && d < p.depth_trunc);
  }
  if (p.has_depth) {
    c_fs = warp_sum_i(c_fs);
    c_sdf = warp_sum_i(c_sdf);
    if (lane == 0) {
      if (c_fs) atomicAdd(&p.counts[0], c_fs);
      if (c_sdf) atomicAdd(&p.counts[1], c_sdf);
      if (c_valid) atomicAdd(&p.counts[2], c_valid);
    }
  }
}

// --------------------------------------------------------------- encoding ---

// Per-level constants, staged in SHARED memory: the level index differs between the lanes
// of a warp in the hash backward (fragment layout: l = 4 nt + t) and in the smoothness
// kernels, and a lane-varying index into the kernel-parameter constant bank serialises
// (ncu source view: ~20 % of the samples of k_fused sat on these loads).
struct __align__(16) Lv {
  float scale;
  uint32_t res, size, offset;
  uint32_t hashed, magic, pad0, pad1;
};

__device__ __forceinline__ void load_levels(Lv* s, const GridDev& g) {
  if (threadIdx.x < kL) {
    const int l = threadIdx.x;
    Lv v;
    v.scale = g.scale[l]; v.res = g.res[l]; v.size = g.size[l]; v.offset = g.offset[l];
    v.hashed = g.hashed[l]; v.magic = g.magic[l]; v.pad0 = v.pad1 = 0;
    s[l] = v;
  }
}

__device__ __forceinline__ uint32_t grid_index(const Lv& L, uint32_t x, uint32_t y, uint32_t z) {
  const uint32_t size = L.size;
  uint32_t index;
  if (L.hashed) {
    index = x ^ (y * 2654435761u) ^ (z * 805459861u);
  } else {
    // dense stride walk (stride <= size is true for all three dims on a dense level)
    const uint32_t res = L.res;
    index = x + y * res + z * res * res;
  }
  // index % size: sizes of hashed levels are powers of two; in-range dense cells are < size
  if ((size & (size - 1)) == 0) {
    index &= size - 1;
  } else if (index >= size) {  // points outside the bound wrap around (tcnn does not clamp)
    index -= __umulhi(index, L.magic) * size;  // exact remainder or remainder + size
    if (index >= size) index -= size;
    if (index >= size) index -= size;
  }
  return index + L.offset;
}

__device__ __forceinline__ void pos_fract(float x, float scale, float& w, uint32_t& c) {
  float pos = fmaf(scale, x, 0.5f);
  float fl = floorf(pos);
  c = (uint32_t)(int)fl;
  w = pos - fl;
}

// normalised coordinate exactly as the reference: f32 pts -> f64 (p-min)/(max-min) -> f32
__device__ __forceinline__ float normalise(float p, double bmin, double bmax) {
  return (float)(((double)p - bmin) / (bmax - bmin));
}

// quartic kernel cdf / pdf in bin units t (|t| <= 1 inside the kernel support)
__device__ __forceinline__ float qcdf(float t) {
  float t2 = t * t;
  float v = 0.9375f * t * (1.f - (2.f / 3.f) * t2 + 0.2f * t2 * t2) + 0.5f;
  return fminf(fmaxf(v, 0.f), 1.f);
}
__device__ __forceinline__ float qpdf(float t) {
  float q = 1.f - t * t;
  return (fabsf(t) < 1.f) ? 0.9375f * q * q : 0.f;
}

// OneBlob, sparse form.  out[b] = sum_{s in -1,0,1} K(e_{b+1}-x+s) - K(e_b-x+s); for image s
// only bins b0-1, b0, b0+1 (b0 = floor(16 (x-s))) are non-zero:
//   cdf(-f), cdf(1-f)-cdf(-f), 1-cdf(1-f)   with f = 16 (x-s) - b0.
__device__ __forceinline__ void blob_forward(const float xn[3], float* __restrict__ rec) {
#pragma unroll
  for (int q = 0; q < 12; ++q)
    *reinterpret_cast<float4*>(rec + R_BLOB + 4 * q) = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int s = -1; s <= 1; ++s) {
      const float u = (xn[d] - (float)s) * (float)kBins;
      const float fl = floorf(u);
      if (fl < -1.f || fl > (float)kBins) continue;
      const int b0 = (int)fl;
      const float f = u - fl;
      const float c0 = qcdf(-f), c1 = qcdf(1.f - f);
      float* o = rec + R_BLOB + d * kBins;
      if (b0 - 1 >= 0 && b0 - 1 < kBins) o[b0 - 1] += c0;
      if (b0 >= 0 && b0 < kBins) o[b0] += c1 - c0;
      if (b0 + 1 >= 0 && b0 + 1 < kBins) o[b0 + 1] += 1.f - c1;
    }
  }
}

// d/dx of dimension d's 16 OneBlob outputs, contracted with dblob (read through the functor)
template <typename F>
__device__ __forceinline__ float blob_backward_dim(float x, int d, F dblob) {
  float acc = 0.f;
#pragma unroll
  for (int s = -1; s <= 1; ++s) {
    const float u = (x - (float)s) * (float)kBins;
    const float fl = floorf(u);
    if (fl < -1.f || fl > (float)kBins) continue;
    const int b0 = (int)fl;
    const float f = u - fl;
    const float p0 = qpdf(-f), p1 = qpdf(1.f - f);  // d cdf(-f)/dx = -16 p0, d cdf(1-f)/dx = -16 p1
    if (b0 - 1 >= 0 && b0 - 1 < kBins) acc -= dblob(d * kBins + b0 - 1) * p0;
    if (b0 >= 0 && b0 < kBins) acc += dblob(d * kBins + b0) * (p0 - p1);
    if (b0 + 1 >= 0 && b0 + 1 < kBins) acc += dblob(d * kBins + b0 + 1) * p1;
  }
  return acc * (float)kBins;
}

// dx[d] += sum_b dblob[d*16+b] * d out_b / dx
template <typename F>
__device__ __forceinline__ void blob_backward(const float xn[3], F dblob, float dx[3]) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    // (fmaf(acc, 16, dx) as one rounding: keeps the v4 kernel's bits)
    float acc = blob_backward_dim(xn[d], d, dblob) * (1.0f / (float)kBins);
    dx[d] = fmaf(acc, (float)kBins, dx[d]);
  }
}

__device__ __forceinline__ void encode_point(const Params& P, const Lv* __restrict__ lv,
                                             const float xn[3], float* __restrict__ rec) {
  const float2* __restrict__ tab = reinterpret_cast<const float2*>(P.table);
#pragma unroll 4
  for (int l = 0; l < kL; ++l) {
    if (l >= P.g.n_levels) {
      rec[R_FEAT + 2 * l] = 0.f;
      rec[R_FEAT + 2 * l + 1] = 0.f;
      continue;
    }
    float w[3];
    uint32_t c[3];
    const Lv L = lv[l];
    pos_fract(xn[0], L.scale, w[0], c[0]);
    pos_fract(xn[1], L.scale, w[1], c[1]);
    pos_fract(xn[2], L.scale, w[2], c[2]);
    float2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      uint32_t idx = grid_index(L, c[0] + (k & 1), c[1] + ((k >> 1) & 1),
                                c[2] + ((k >> 2) & 1));
      v[k] = __ldg(&tab[idx]);
    }
    float f0 = 0.f, f1 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float wk = ((k & 1) ? w[0] : 1.f - w[0]) * ((k & 2) ? w[1] : 1.f - w[1]) *
                 ((k & 4) ? w[2] : 1.f - w[2]);
      f0 = fmaf(wk, v[k].x, f0);
      f1 = fmaf(wk, v[k].y, f1);
    }
    *reinterpret_cast<float2*>(rec + R_FEAT + 2 * l) = make_float2(f0, f1);
  }
  blob_forward(xn, rec);
}

// NL (point, level) items of the hash backward per call (scatter (g0,g1), optional d/dx).
// Deliberately not inlined: instruction-cache footprint matters for this kernel (profiles/).
// NL (point, level) items per call: the 8 NL corner loads of all levels are issued before any
// is consumed (the kernel runs at 7 warps/SM, so memory-level parallelism has to come from
// within the thread), then the 8 NL scatters.  on[h]: item present (level exists and its
// gradient is non-zero).  Levels are l0 + 4 h (fragment layout: lane t owns levels 4 nt + t).
template <int NL>
__device__ __noinline__ float3 hash_backward_multi(const Params& P, const Lv* __restrict__ lv,
                                                   int l0, const bool (&on)[NL], float x0,
                                                   float x1, float x2, const float (&g)[NL][2],
                                                   bool need_dx, bool scatter, int lstride = 4) {
  const float2* __restrict__ tab = reinterpret_cast<const float2*>(P.table);
  const float x[3] = {x0, x1, x2};
  float w[NL][3], sc[NL];
  uint32_t idx[NL][8];
#pragma unroll
  for (int h = 0; h < NL; ++h) {
    const Lv L = lv[on[h] ? l0 + lstride * h : 0];
    sc[h] = L.scale;
    uint32_t c[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) pos_fract(x[d], sc[h], w[h][d], c[d]);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      idx[h][k] = grid_index(L, c[0] + (k & 1), c[1] + ((k >> 1) & 1), c[2] + ((k >> 2) & 1));
  }
  float dx[3] = {0.f, 0.f, 0.f};
  if (need_dx) {
    float t[NL][8];
#pragma unroll
    for (int h = 0; h < NL; ++h)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float2 v = on[h] ? __ldg(&tab[idx[h][k]]) : make_float2(0.f, 0.f);
        t[h][k] = v.x * g[h][0] + v.y * g[h][1];
      }
#pragma unroll
    for (int h = 0; h < NL; ++h) {
      const float w0 = w[h][0], w1 = w[h][1], w2 = w[h][2];
      const float* tt = t[h];
      const float d0 = (1 - w1) * (1 - w2) * (tt[1] - tt[0]) + w1 * (1 - w2) * (tt[3] - tt[2]) +
                       (1 - w1) * w2 * (tt[5] - tt[4]) + w1 * w2 * (tt[7] - tt[6]);
      const float d1 = (1 - w0) * (1 - w2) * (tt[2] - tt[0]) + w0 * (1 - w2) * (tt[3] - tt[1]) +
                       (1 - w0) * w2 * (tt[6] - tt[4]) + w0 * w2 * (tt[7] - tt[5]);
      const float d2 = (1 - w0) * (1 - w1) * (tt[4] - tt[0]) + w0 * (1 - w1) * (tt[5] - tt[1]) +
                       (1 - w0) * w1 * (tt[6] - tt[2]) + w0 * w1 * (tt[7] - tt[3]);
      if (on[h]) {
        dx[0] = fmaf(sc[h], d0, dx[0]);
        dx[1] = fmaf(sc[h], d1, dx[1]);
        dx[2] = fmaf(sc[h], d2, dx[2]);
      }
    }
  }
  if (scatter) {
#pragma unroll
    for (int h = 0; h < NL; ++h) {
      if (!on[h]) continue;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float wk = ((k & 1) ? w[h][0] : 1.f - w[h][0]) * ((k & 2) ? w[h][1] : 1.f - w[h][1]) *
                         ((k & 4) ? w[h][2] : 1.f - w[h][2]);
        red_add_v2(P.d_table + 2 * (size_t)idx[h][k], wk * g[h][0], wk * g[h][1]);
      }
    }
  }
  return make_float3(dx[0], dx[1], dx[2]);
}

// ------------------------------------------------------- tensor-core GEMMs ---
// mma.sync m16n8k8 tf32 (fp32 accumulate).  PREC3 = 3xTF32 error-compensated split:
// x = big + small, d += a_s*b_b + a_b*b_s + a_b*b_b  -> fp32-level accuracy.
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4],
                                         const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// Operand preparation.  The tensor core reads only the upper 19 bits of a .tf32 register, so
// plain TF32 mode feeds the fp32 bits unchanged (truncation, |err| <= 2^-10 rel).  In 3xTF32
// mode big = x with the low 13 mantissa bits cleared (exactly what the hardware would see) and
// small = x - big (exact in fp32); small's own truncation error is second order (2^-20).
template <bool PREC3>
struct FragA {
  uint32_t big[4], small[4];
  __device__ __forceinline__ void set(const float (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (PREC3) {
        big[i] = __float_as_uint(v[i]) & 0xffffe000u;
        small[i] = __float_as_uint(v[i] - __uint_as_float(big[i]));
      } else {
        big[i] = __float_as_uint(v[i]);
      }
    }
  }
};
template <bool PREC3>
struct FragB {
  uint32_t big[2], small[2];
  __device__ __forceinline__ void set(float v0, float v1) {
    if (PREC3) {
      big[0] = __float_as_uint(v0) & 0xffffe000u;
      big[1] = __float_as_uint(v1) & 0xffffe000u;
      small[0] = __float_as_uint(v0 - __uint_as_float(big[0]));
      small[1] = __float_as_uint(v1 - __uint_as_float(big[1]));
    } else {
      big[0] = __float_as_uint(v0);
      big[1] = __float_as_uint(v1);
    }
  }
};
template <bool PREC3>
__device__ __forceinline__ void mma(float (&d)[4], const FragA<PREC3>& a, const FragB<PREC3>& b) {
  if (PREC3) {
    mma_tf32(d, a.small, b.big);
    mma_tf32(d, a.big, b.small);
  }
  mma_tf32(d, a.big, b.big);
}

// C[mt][nt] += A * B for the 32 points of this warp (2 m-tiles of 16 rows).
//   A[row][k]   = wrec[row*REC + aoff + k]              (activations / gradients, records)
//   B[k][n]     = TRANS ? w[(noff+n)*ld + koff + k] : w[(koff+k)*ld + noff + n]
template <int KS, int NT, bool TRANS, bool PREC3>
__device__ __forceinline__ void warp_gemm(const float* __restrict__ wrec, int aoff,
                                          const float* __restrict__ w, int ld, int koff,
                                          int noff, float (&c)[2][NT][4]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll 1
  for (int ks = 0; ks < KS; ++ks) {
    FragA<PREC3> a[2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const float* r0 = wrec + (mt * 16 + g) * REC + aoff + ks * 8 + t;
      const float v[4] = {r0[0], r0[8 * REC], r0[4], r0[8 * REC + 4]};
      a[mt].set(v);
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      FragB<PREC3> b;
      const int k0 = koff + ks * 8 + t, n0 = noff + nt * 8 + g;
      if (TRANS) b.set(w[n0 * ld + k0], w[n0 * ld + k0 + 4]);
      else b.set(w[k0 * ld + n0], w[(k0 + 4) * ld + n0]);
      mma<PREC3>(c[0][nt], a[0], b);
      mma<PREC3>(c[1][nt], a[1], b);
    }
  }
}

template <int NT>
__device__ __forceinline__ void zero_c(float (&c)[2][NT][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) c[mt][nt][i] = 0.f;
}

// store C (32 rows x 8*NT cols) into the records at column offset coff (float2 per row pair)
template <int NT, bool RELU>
__device__ __forceinline__ void store_c(float* __restrict__ wrec, int coff,
                                        const float (&c)[2][NT][4]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float v0 = c[mt][nt][0], v1 = c[mt][nt][1], v2 = c[mt][nt][2], v3 = c[mt][nt][3];
      if (RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
      float* r0 = wrec + (mt * 16 + g) * REC + coff + nt * 8 + 2 * t;
      *reinterpret_cast<float2*>(r0) = make_float2(v0, v1);
      *reinterpret_cast<float2*>(r0 + 8 * REC) = make_float2(v2, v3);
    }
}

// weights in shared memory: [in][out] with padded row strides (conflict-free B fragments)
constexpr int LD0 = 40, LD1 = 24, LDC0 = 40, LDC1 = 8;
constexpr int SW0 = 0;                     // [80][40]  w_sdf0^T
constexpr int SW1 = SW0 + 80 * LD0;        // [32][24]  w_sdf1^T, cols = (geo0..14, sdf)
constexpr int SWC0 = SW1 + 32 * LD1;       // [64][40]  w_col0^T, row 63 = 0
constexpr int SWC1 = SWC0 + 64 * LDC0;     // [32][8]   w_col1^T, cols 3..7 = 0
constexpr int SW_TOTAL = SWC1 + 32 * LDC1; // 6784 floats

// weight-gradient tiles (16 in-features x 8 out-units each), global order
//   [0,2) d w_col1   [2,18) d w_col0   [18,22) d w_sdf1   [22,42) d w_sdf0
constexpr int DW_TILES = 42;

template <bool PREC3>
__device__ __forceinline__ void dw_tile(const float* __restrict__ recs, int npts, int aoff,
                                        int boff, float (&acc)[4]) {
  // acc[16 x 8] += sum_p A[p][aoff + m] * B[p][boff + n]; k-slot t <-> point 2t, t+4 <-> 2t+1
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll 2
  for (int k0 = 0; k0 < npts; k0 += 8) {
    const float* p0 = recs + (k0 + 2 * t) * REC;
    const float* p1 = p0 + REC;
    const float va[4] = {p0[aoff + g], p0[aoff + g + 8], p1[aoff + g], p1[aoff + g + 8]};
    FragA<PREC3> a;
    a.set(va);
    FragB<PREC3> b;
    b.set(p0[boff + g], p1[boff + g]);
    mma<PREC3>(acc, a, b);
  }
}

// ----------------------------------------------------------------- fused ---
template <bool BWD, bool PREC3, bool BPREC3, int SLOTS>
__global__ void __launch_bounds__(256) k_fused(const Params P) {
  extern __shared__ __align__(16) float smem[];
  float* sw = smem;                            // SW_TOTAL
  float* recs = sw + SW_TOTAL;                 // (blockDim.x + 8) * REC
  float* zbuf = recs + (blockDim.x + 8) * REC; // NR*S z values
  float* sgb = zbuf + P.NR * P.S;              // NR*S sigma(sdf/trunc)
  float* ub = sgb + P.NR * P.S;                // NR*S unnormalised weights
  float* xnb = ub + P.NR * P.S;                // NR*S*3 normalised coordinates
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int nwarps = blockDim.x >> 5;
  const int S = P.S;

  __shared__ Lv s_lv[kL];
  load_levels(s_lv, P.g);
  // stage transposed weights (zero padding included)
  for (int q = tid; q < SW_TOTAL; q += blockDim.x) sw[q] = 0.f;
  __syncthreads();
  for (int q = tid; q < 80 * 32; q += blockDim.x) {
    int i = q / 32, j = q % 32;
    sw[SW0 + i * LD0 + j] = P.w_sdf0[j * 80 + i];
  }
  for (int q = tid; q < 32 * 16; q += blockDim.x) {
    int i = q / 16, jp = q % 16;  // jp: geo0..14 -> torch out 1..15, jp 15 -> out 0 (sdf)
    int jt = (jp + 1) & 15;
    sw[SW1 + i * LD1 + jp] = P.w_sdf1[jt * 32 + i];
  }
  for (int q = tid; q < 63 * 32; q += blockDim.x) {
    int i = q / 32, j = q % 32;
    sw[SWC0 + i * LDC0 + j] = P.w_col0[j * 63 + i];
  }
  for (int q = tid; q < 32 * 3; q += blockDim.x) {
    int i = q / 3, k = q % 3;
    sw[SWC1 + i * LDC1 + k] = P.w_col1[k * 32 + i];
  }
  // the 8 pad records after the tile stay zero for the whole kernel (k-step overrun of dw_tile)
  for (int q = tid; q < 8 * REC; q += blockDim.x) recs[blockDim.x * REC + q] = 0.f;

  float fs_w = 0.f, sdf_w = 0.f, inv_nvalid = 0.f;
  if (BWD) {
    const float n_fs = (float)P.counts[0], n_sdf = (float)P.counts[1];
    const float n = (float)(P.counts[0] + P.counts[1]);
    fs_w = 1.0f - n_fs / n;
    sdf_w = 1.0f - n_sdf / n;
    inv_nvalid = 1.0f / (float)P.counts[2];
  }
  double l_rgb = 0.0, l_depth = 0.0, l_sdf = 0.0, l_fs = 0.0;
  float dwacc[SLOTS][4];
#pragma unroll
  for (int j = 0; j < SLOTS; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) dwacc[j][i] = 0.f;
  __syncthreads();

  float* rec = recs + tid * REC;
  float* wrec = recs + (warp * 32) * REC;  // this warp's 32 records
  const bool need_dx = BWD && ((P.d_rays_o != nullptr) || (P.d_rays_d != nullptr));
  const bool map_grads = BWD && (P.d_table != nullptr);

  for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
    const int r0 = tile * P.NR;
    const int nr = min(P.NR, P.R - r0);
    const int npts = nr * S;
    const bool active = tid < npts;
    const bool warp_active = warp * 32 < npts;
    const int rl = active ? tid / S : 0;
    const int k = active ? tid - rl * S : 0;
    const int r = r0 + rl;
    float xn[3] = {0.f, 0.f, 0.f};
    float zv = 0.f;
    // ---------------- phase 1: forward -------------------------------------
    if (active) {
      zv = P.z_vals[(size_t)r * S + k];
      zbuf[tid] = zv;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float pt = __fadd_rn(P.rays_o[r * 3 + d], __fmul_rn(P.rays_d[r * 3 + d], zv));
        xn[d] = normalise(pt, P.g.bmin[d], P.g.bmax[d]);
        xnb[tid * 3 + d] = xn[d];
      }
      encode_point(P, s_lv, xn, rec);
    } else {
      // inactive rows must read as zero in every GEMM (k-dimension of the dW tiles)
#pragma unroll 1
      for (int q = 0; q < REC; q += 4)
        *reinterpret_cast<float4*>(rec + q) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    if (warp_active) {
      {  // h1 = relu(x W0^T)
        float c[2][4][4];
        zero_c<4>(c);
        warp_gemm<10, 4, false, PREC3>(wrec, R_FEAT, sw + SW0, LD0, 0, 0, c);
        store_c<4, true>(wrec, R_H1, c);
      }
      __syncwarp();
      {  // [geo, sdf] = h1 W1^T
        float c[2][2][4];
        zero_c<2>(c);
        warp_gemm<4, 2, false, PREC3>(wrec, R_H1, sw + SW1, LD1, 0, 0, c);
        store_c<2, false>(wrec, R_GEO, c);
      }
      __syncwarp();
      {  // c1 = relu([blob, geo, (sdf: zero weight row)] Wc0^T)
        float c[2][4][4];
        zero_c<4>(c);
        warp_gemm<8, 4, false, PREC3>(wrec, R_BLOB, sw + SWC0, LDC0, 0, 0, c);
        store_c<4, true>(wrec, R_C1, c);
      }
      __syncwarp();
      {  // rgb logits = c1 Wc1^T (cols 0..2 of an 8-wide tile)
        float c[2][1][4];
        zero_c<1>(c);
        warp_gemm<4, 1, false, PREC3>(wrec, R_C1, sw + SWC1, LDC1, 0, 0, c);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          float* q0 = wrec + (mt * 16 + g) * REC + R_RAW;
          if (t == 0) {
            q0[0] = c[mt][0][0]; q0[1] = c[mt][0][1];
            q0[8 * REC] = c[mt][0][2]; q0[8 * REC + 1] = c[mt][0][3];
          } else if (t == 1) {
            q0[2] = c[mt][0][0];
            q0[8 * REC + 2] = c[mt][0][2];
          }
        }
      }
      __syncwarp();
    }
    if (active) {
      rec[R_RAW + 3] = rec[R_SDF];
      if (P.raw) {
        float4 rw = *reinterpret_cast<const float4*>(rec + R_RAW);
        *reinterpret_cast<float4*>(P.raw + ((size_t)r * S + k) * 4) = rw;
      }
    }
    __syncthreads();
    // ---------------- phase 2: per-ray composite, loss, d loss / d raw ----
    // per sample the five sigmoids are evaluated once: colours overwrite the rgb logits in
    // RAW[0..2]; sigma(s/tr) and the unnormalised weight u live in sgb / ub.
    for (int q = warp; q < nr; q += nwarps) {
      const int rr = r0 + q;
      const float* zr = zbuf + q * S;
      float* sg_r = sgb + q * S;
      float* u_r = ub + q * S;
      float* rq = recs + (size_t)(q * S) * REC;
      const float tr = P.trunc;
      int first = 0x7fffffff;
      for (int kk = lane; kk < S; kk += 32) {
        float* rp = rq + kk * REC + R_RAW;
        const float s = rp[3];
        const float sg = sigmoidf_acc(s / tr);
        sg_r[kk] = sg;
        u_r[kk] = sg * sigmoidf_acc((-s) / tr);
        rp[0] = sigmoidf_acc(rp[0]);
        rp[1] = sigmoidf_acc(rp[1]);
        rp[2] = sigmoidf_acc(rp[2]);
        if (kk < S - 1 && rq[(kk + 1) * REC + R_RAW + 3] * s < 0.f) first = min(first, kk);
      }
      first = warp_min_i(first);
      if (first == 0x7fffffff) first = 0;
      __syncwarp();
      const float zlim = zr[first] + P.trunc;
      float usum = 0.f;
      for (int kk = lane; kk < S; kk += 32) {
        const float u = (zr[kk] < zlim) ? u_r[kk] : 0.f;
        u_r[kk] = u;
        usum += u;
      }
      usum = warp_sum(usum);
      const float W = usum + 1e-8f;
      float o_r = 0.f, o_g = 0.f, o_b = 0.f, o_d = 0.f, o_acc = 0.f;
      for (int kk = lane; kk < S; kk += 32) {
        const float* rp = rq + kk * REC + R_RAW;
        const float w = u_r[kk] / W;
        o_r = fmaf(w, rp[0], o_r);
        o_g = fmaf(w, rp[1], o_g);
        o_b = fmaf(w, rp[2], o_b);
        o_d = fmaf(w, zr[kk], o_d);
        o_acc += w;
      }
      o_r = warp_sum(o_r); o_g = warp_sum(o_g); o_b = warp_sum(o_b);
      o_d = warp_sum(o_d); o_acc = warp_sum(o_acc);
      if (P.depth_var) {
        float v = 0.f;
        for (int kk = lane; kk < S; kk += 32) {
          const float dz = zr[kk] - o_d;
          v = fmaf(u_r[kk] / W, dz * dz, v);
        }
        v = warp_sum(v);
        if (lane == 0) P.depth_var[rr] = v;
      }
      if (lane == 0) {
        if (P.rgb) { P.rgb[rr * 3] = o_r; P.rgb[rr * 3 + 1] = o_g; P.rgb[rr * 3 + 2] = o_b; }
        if (P.depth) P.depth[rr] = o_d;
        if (P.acc) P.acc[rr] = o_acc;
        if (P.disp) P.disp[rr] = 1.0f / fmaxf(1e-10f, o_d / o_acc);
      }
      if (BWD) {
        const float td = P.target_d[rr];
        const float tr_ = P.target_s[rr * 3], tg_ = P.target_s[rr * 3 + 1],
                    tb_ = P.target_s[rr * 3 + 2];
        const bool valid = (td > 0.f) && (td < P.depth_trunc);
        const float RS = (float)P.Rg * (float)S;
        const float c_rgb = P.ls[0] * P.w_rgb * 2.0f / (3.0f * (float)P.Rg);
        const float g_r = c_rgb * (o_r - tr_), g_g = c_rgb * (o_g - tg_), g_b = c_rgb * (o_b - tb_);
        const float g_d = valid ? P.ls[1] * P.w_depth * 2.0f * (o_d - td) * inv_nvalid : 0.f;
        if (lane == 0) {
          l_rgb += (double)((o_r - tr_) * (o_r - tr_) + (o_g - tg_) * (o_g - tg_) +
                            (o_b - tb_) * (o_b - tb_));
          if (valid) l_depth += (double)((o_d - td) * (o_d - td));
        }
        float qw = 0.f;
        for (int kk = lane; kk < S; kk += 32) {
          const float* rp = rq + kk * REC + R_RAW;
          const float q_ = g_r * rp[0] + g_g * rp[1] + g_b * rp[2] + g_d * zr[kk];
          qw = fmaf(q_, u_r[kk] / W, qw);
        }
        qw = warp_sum(qw);
        float a_fs = 0.f, a_sdf = 0.f;
        const float c_fs = P.ls[3] * P.w_fs * fs_w * 2.0f / RS;
        const float c_sdf = P.ls[2] * P.w_sdf * sdf_w * 2.0f / RS;
        for (int kk = lane; kk < S; kk += 32) {
          float* rp = rq + kk * REC + R_RAW;
          const float z = zr[kk];
          const float s = rp[3];
          const float sg = sg_r[kk];
          const float u = u_r[kk];  // a * mask
          const float w = u / W;
          const float c0 = rp[0], c1 = rp[1], c2 = rp[2];
          const float q_ = g_r * c0 + g_g * c1 + g_b * c2 + g_d * z;
          float ds = (q_ - qw) / W * u * (1.f - 2.f * sg) / tr;
          const bool front = z < __fsub_rn(td, P.trunc);
          const bool back = z > __fadd_rn(td, P.trunc);
          if (front) {
            ds += c_fs * (s - 1.f);
            a_fs += (s - 1.f) * (s - 1.f);
          }
          if (!front && !back && td > 0.f) {
            float e = (z + s * P.trunc) - td;
            ds += c_sdf * e * P.trunc;
            a_sdf += e * e;
          }
          rp[0] = g_r * w * c0 * (1.f - c0);
          rp[1] = g_g * w * c1 * (1.f - c1);
          rp[2] = g_b * w * c2 * (1.f - c2);
          rp[3] = ds;
        }
        a_fs = warp_sum(a_fs);
        a_sdf = warp_sum(a_sdf);
        if (lane == 0) { l_fs += (double)a_fs; l_sdf += (double)a_sdf; }
      }
    }
    __syncthreads();
    if (!BWD) continue;
    // ---------------- phase 3: backward ----------------------------------
    // weight-gradient tiles owned by this warp: tile id = warp + j * nwarps
    auto dw_phase = [&](int lo, int hi, int aoff, int boff, int n_tiles_n) {
      if (!map_grads) return;
#pragma unroll
      for (int j = 0; j < SLOTS; ++j) {
        const int id = warp + j * nwarps;
        if (id >= lo && id < hi) {
          const int loc = id - lo, mt = loc / n_tiles_n, nt = loc % n_tiles_n;
          dw_tile<BPREC3>(recs, npts, aoff + 16 * mt, boff + 8 * nt, dwacc[j]);
        }
      }
    };
    // d w_col1 += c1^T draw        (8-wide n tile over raw[4] + 4 floats of the next record:
    dw_phase(0, 2, R_C1, R_RAW, 1);  //  columns 3..7 are never written out)
    __syncthreads();
    float draw3 = 0.f;
    if (active) {
      const float4 dr = *reinterpret_cast<const float4*>(rec + R_RAW);
      draw3 = dr.w;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float* w = sw + SWC1 + j * LDC1;
        const float c1 = rec[R_C1 + j];
        rec[R_C1 + j] = (c1 > 0.f) ? (dr.x * w[0] + dr.y * w[1] + dr.z * w[2]) : 0.f;
      }
    }
    __syncthreads();
    // d w_col0 += [blob, geo, sdf]^T dc1pre   (row 63 is discarded at write-out)
    dw_phase(2, 18, R_BLOB, R_C1, 4);
    __syncthreads();
    if (warp_active) {
      // dgeo = dc1pre Wc0[:, 48:64]  -> dH = [dgeo(15), dsdf]
      float c[2][2][4];
      zero_c<2>(c);
      warp_gemm<4, 2, true, BPREC3>(wrec, R_C1, sw + SWC0, LDC0, 0, 48, c);
      store_c<2, false>(wrec, R_GEO, c);
    }
    __syncwarp();
    if (active) rec[R_SDF] = draw3;
    __syncthreads();
    // d w_sdf1 += h1^T dH
    dw_phase(18, 22, R_H1, R_GEO, 2);
    __syncthreads();
    if (warp_active) {
      // dh1pre = (h1 > 0) * (dH W1)
      float c[2][4][4];
      zero_c<4>(c);
      warp_gemm<2, 4, true, BPREC3>(wrec, R_GEO, sw + SW1, LD1, 0, 0, c);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          float* q0 = wrec + (mt * 16 + g) * REC + R_H1 + nt * 8 + 2 * t;
          float2 h0 = *reinterpret_cast<float2*>(q0);
          float2 h8 = *reinterpret_cast<float2*>(q0 + 8 * REC);
          *reinterpret_cast<float2*>(q0) =
              make_float2(h0.x > 0.f ? c[mt][nt][0] : 0.f, h0.y > 0.f ? c[mt][nt][1] : 0.f);
          *reinterpret_cast<float2*>(q0 + 8 * REC) =
              make_float2(h8.x > 0.f ? c[mt][nt][2] : 0.f, h8.y > 0.f ? c[mt][nt][3] : 0.f);
        }
    }
    __syncthreads();
    // d w_sdf0 += x^T dh1pre
    dw_phase(22, 42, R_FEAT, R_H1, 4);
    // (readers of C1 / GEO slots are done: dW phases 2 and 3 finished before the last barrier)
    float hdx[2][2][3];  // hash-path d loss / d x for rows (mt, g / g+8), fragment layout
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) hdx[a][b][0] = hdx[a][b][1] = hdx[a][b][2] = 0.f;
    if (warp_active) {
      if (need_dx) {
        // dblob = dc1pre Wc0[:, 0:48] + dh1pre W0[:, 32:80]  -> scratch (C1 slots 0..31, GEO 32..47)
        float c[2][6][4];
        zero_c<6>(c);
        warp_gemm<4, 6, true, BPREC3>(wrec, R_C1, sw + SWC0, LDC0, 0, 0, c);
        warp_gemm<4, 6, true, BPREC3>(wrec, R_H1, sw + SW0, LD0, 0, 32, c);
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 6; ++nt) {
            const int col = nt * 8 + 2 * t;
            const int off = (col < 32) ? (R_C1 + col) : (R_GEO + col - 32);
            float* q0 = wrec + (mt * 16 + g) * REC + off;
            *reinterpret_cast<float2*>(q0) = make_float2(c[mt][nt][0], c[mt][nt][1]);
            *reinterpret_cast<float2*>(q0 + 8 * REC) = make_float2(c[mt][nt][2], c[mt][nt][3]);
          }
      }
      if (map_grads || need_dx) {
        // dfeat = dh1pre W0[:, 0:32]; lane (g,t) holds (f0,f1) of level 4*nt+t for rows g, g+8
        float c[2][4][4];
        zero_c<4>(c);
        warp_gemm<4, 4, true, BPREC3>(wrec, R_H1, sw + SW0, LD0, 0, 0, c);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int row = warp * 32 + mt * 16 + g + 8 * h;
            if (row < npts) {
              const float x3[3] = {xnb[row * 3], xnb[row * 3 + 1], xnb[row * 3 + 2]};
              {  // the 4 levels of this lane (t, 4 + t, 8 + t, 12 + t) in one call
                bool on[4];
                float gg[4][2];
                bool any = false;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                  gg[nt][0] = c[mt][nt][2 * h]; gg[nt][1] = c[mt][nt][2 * h + 1];
                  on[nt] = (4 * nt + t) < P.g.n_levels && (gg[nt][0] != 0.f || gg[nt][1] != 0.f);
                  any |= on[nt];
                }
                if (any) {
                  const float3 d3 = hash_backward_multi<4>(P, s_lv, t, on, x3[0], x3[1], x3[2], gg,
                                                           need_dx, map_grads);
                  hdx[mt][h][0] += d3.x; hdx[mt][h][1] += d3.y; hdx[mt][h][2] += d3.z;
                }
              }
            }
          }
      }
      if (need_dx) {
        // reduce the hash-path dx over the 4 lanes (t) that share a row, park it in RAW[0..2]
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              float v = hdx[mt][h][d];
              v += __shfl_xor_sync(0xffffffffu, v, 1);
              v += __shfl_xor_sync(0xffffffffu, v, 2);
              if (t == 0) wrec[(mt * 16 + g + 8 * h) * REC + R_RAW + d] = v;
            }
        __syncwarp();
        if (active) {
          float dxn[3] = {rec[R_RAW], rec[R_RAW + 1], rec[R_RAW + 2]};
          blob_backward(xn, [&](int i) { return (i < 32) ? rec[R_C1 + i] : rec[R_GEO + i - 32]; }, dxn);
          // park d loss / d pts and z * d loss / d pts in the (dead) C1 slots 0..5
          float dp[3];
#pragma unroll
          for (int d = 0; d < 3; ++d)
            dp[d] = (float)((double)dxn[d] * P.g.binv[d]);
          rec[R_C1 + 0] = dp[0]; rec[R_C1 + 1] = dp[1]; rec[R_C1 + 2] = dp[2];
          rec[R_C1 + 3] = dp[0] * zv; rec[R_C1 + 4] = dp[1] * zv; rec[R_C1 + 5] = dp[2] * zv;
        }
      }
    }
    __syncthreads();
    if (need_dx) {
      for (int q = warp; q < nr; q += nwarps) {
        float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int kk = lane; kk < S; kk += 32) {
          const float* rp = recs + (size_t)(q * S + kk) * REC + R_C1;
#pragma unroll
          for (int d = 0; d < 6; ++d) a[d] += rp[d];
        }
#pragma unroll
        for (int d = 0; d < 6; ++d) a[d] = warp_sum(a[d]);
        if (lane == 0) {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            if (P.d_rays_o) P.d_rays_o[(r0 + q) * 3 + d] = a[d];
            if (P.d_rays_d) P.d_rays_d[(r0 + q) * 3 + d] = a[3 + d];
          }
        }
      }
      __syncthreads();
    }
  }

  if (BWD) {
    if (lane == 0) {
      if (l_rgb != 0.0) atomicAdd(&P.loss_acc[0], l_rgb);
      if (l_depth != 0.0) atomicAdd(&P.loss_acc[1], l_depth);
      if (l_sdf != 0.0) atomicAdd(&P.loss_acc[2], l_sdf);
      if (l_fs != 0.0) atomicAdd(&P.loss_acc[3], l_fs);
    }
    if (map_grads) {
      // write out the weight-gradient tiles: C(16 x 8): rows = in-feature, cols = out-unit
#pragma unroll
      for (int j = 0; j < SLOTS; ++j) {
        const int id = warp + j * nwarps;
        if (id >= DW_TILES) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float v = dwacc[j][i];
          const int rloc = g + ((i & 2) ? 8 : 0), cloc = 2 * t + (i & 1);
          if (id < 2) {  // d w_col1 [3][32]
            const int in = 16 * id + rloc, out = cloc;
            if (out < 3) red_add(P.d_w_col1 + out * 32 + in, v);
          } else if (id < 18) {  // d w_col0 [32][63]
            const int loc = id - 2, in = 16 * (loc / 4) + rloc, out = 8 * (loc % 4) + cloc;
            if (in < 63) red_add(P.d_w_col0 + out * 63 + in, v);
          } else if (id < 22) {  // d w_sdf1 [16][32], stored column jp -> torch row (jp+1)&15
            const int loc = id - 18, in = 16 * (loc / 2) + rloc, jp = 8 * (loc % 2) + cloc;
            red_add(P.d_w_sdf1 + ((jp + 1) & 15) * 32 + in, v);
          } else {  // d w_sdf0 [32][80]
            const int loc = id - 22, in = 16 * (loc / 4) + rloc, out = 8 * (loc % 4) + cloc;
            red_add(P.d_w_sdf0 + out * 80 + in, v);
          }
        }
      }
    }
  }
}


// ----------------------------------------------------------- fused, grouped ---
// v5: the same record-based algorithm, re-organised for latency hiding.  One persistent CTA
// per SM holds NGROUPS independent groups of 6 warps; a group owns GP record slots and pulls
// units of NR rays from a global queue, so at any time one group is in its gather phase
// (L1/L2 latency), another in the tensor-core phases and a third in the scatter phase: the
// phases that were serialised behind __syncthreads() in k_fused now overlap on the SM.
// Two threads share one sample point in the gather phases (8 levels each), one warp owns one
// 16-row MMA tile; all intra-group synchronisation is a named barrier (bar.sync grp+1, 192).
constexpr int GT = 192;      // threads per group
constexpr int GW = 6;        // warps per group = m-tiles of 16 points per unit
constexpr int GP = 96;       // record slots per group
constexpr int NGROUPS = 3;
constexpr int G_SLOTS = 7;   // 42 weight-gradient tiles / 6 warps

__device__ __forceinline__ void group_sync(int grp) {
  asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(GT) : "memory");
}

// hash levels half, half+2, ... and the OneBlob dims of this half (0: x,y  1: z)
__device__ __forceinline__ void encode_half(const Params& P, const Lv* __restrict__ lv,
                                            const float xn[3], float* __restrict__ rec, int half,
                                            float* __restrict__ jac) {
  const float2* __restrict__ tab = reinterpret_cast<const float2*>(P.table);
#pragma unroll 2
  for (int i = 0; i < kL / 2; ++i) {
    const int l = half + 2 * i;
    if (l >= P.g.n_levels) {
      *reinterpret_cast<float2*>(rec + R_FEAT + 2 * l) = make_float2(0.f, 0.f);
      if (jac) {
        *reinterpret_cast<float2*>(jac + 6 * i) = make_float2(0.f, 0.f);
        *reinterpret_cast<float2*>(jac + 6 * i + 2) = make_float2(0.f, 0.f);
        *reinterpret_cast<float2*>(jac + 6 * i + 4) = make_float2(0.f, 0.f);
      }
      continue;
    }
    float w[3];
    uint32_t c[3];
    const Lv L = lv[l];
    pos_fract(xn[0], L.scale, w[0], c[0]);
    pos_fract(xn[1], L.scale, w[1], c[1]);
    pos_fract(xn[2], L.scale, w[2], c[2]);
    float2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      uint32_t idx = grid_index(L, c[0] + (k & 1), c[1] + ((k >> 1) & 1), c[2] + ((k >> 2) & 1));
      v[k] = __ldg(&tab[idx]);
    }
    float f0 = 0.f, f1 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float wk = ((k & 1) ? w[0] : 1.f - w[0]) * ((k & 2) ? w[1] : 1.f - w[1]) *
                 ((k & 4) ? w[2] : 1.f - w[2]);
      f0 = fmaf(wk, v[k].x, f0);
      f1 = fmaf(wk, v[k].y, f1);
    }
    *reinterpret_cast<float2*>(rec + R_FEAT + 2 * l) = make_float2(f0, f1);
    if (jac) {
      // d feat / d xn while the 8 corners are in registers: [f0: dx dy dz | f1: dx dy dz] * scale
      const float w0 = w[0], w1 = w[1], w2 = w[2], sc = L.scale;
      const float a00 = (1 - w1) * (1 - w2), a10 = w1 * (1 - w2), a01 = (1 - w1) * w2, a11 = w1 * w2;
      const float b00 = (1 - w0) * (1 - w2), b10 = w0 * (1 - w2), b01 = (1 - w0) * w2, b11 = w0 * w2;
      const float c00 = (1 - w0) * (1 - w1), c10 = w0 * (1 - w1), c01 = (1 - w0) * w1, c11 = w0 * w1;
      const float jx0 = a00 * (v[1].x - v[0].x) + a10 * (v[3].x - v[2].x) + a01 * (v[5].x - v[4].x) + a11 * (v[7].x - v[6].x);
      const float jy0 = b00 * (v[2].x - v[0].x) + b10 * (v[3].x - v[1].x) + b01 * (v[6].x - v[4].x) + b11 * (v[7].x - v[5].x);
      const float jz0 = c00 * (v[4].x - v[0].x) + c10 * (v[5].x - v[1].x) + c01 * (v[6].x - v[2].x) + c11 * (v[7].x - v[3].x);
      const float jx1 = a00 * (v[1].y - v[0].y) + a10 * (v[3].y - v[2].y) + a01 * (v[5].y - v[4].y) + a11 * (v[7].y - v[6].y);
      const float jy1 = b00 * (v[2].y - v[0].y) + b10 * (v[3].y - v[1].y) + b01 * (v[6].y - v[4].y) + b11 * (v[7].y - v[5].y);
      const float jz1 = c00 * (v[4].y - v[0].y) + c10 * (v[5].y - v[1].y) + c01 * (v[6].y - v[2].y) + c11 * (v[7].y - v[3].y);
      *reinterpret_cast<float2*>(jac + 6 * i) = make_float2(sc * jx0, sc * jy0);
      *reinterpret_cast<float2*>(jac + 6 * i + 2) = make_float2(sc * jz0, sc * jx1);
      *reinterpret_cast<float2*>(jac + 6 * i + 4) = make_float2(sc * jy1, sc * jz1);
    }
  }
  const int d_lo = half ? 2 : 0, d_hi = half ? 3 : 2;
  for (int d = d_lo; d < d_hi; ++d) {
    float* o = rec + R_BLOB + d * kBins;
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(o + 4 * q) = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = -1; s <= 1; ++s) {
      const float u = (xn[d] - (float)s) * (float)kBins;
      const float fl = floorf(u);
      if (fl < -1.f || fl > (float)kBins) continue;
      const int b0 = (int)fl;
      const float f = u - fl;
      const float c0 = qcdf(-f), c1 = qcdf(1.f - f);
      if (b0 - 1 >= 0 && b0 - 1 < kBins) o[b0 - 1] += c0;
      if (b0 >= 0 && b0 < kBins) o[b0] += c1 - c0;
      if (b0 + 1 >= 0 && b0 + 1 < kBins) o[b0 + 1] += 1.f - c1;
    }
  }
}

// one 16-row m-tile: C[nt] += A * B   (A rows = wrec[row*REC + aoff + k])
template <int KS, int NT, bool TRANS, bool PREC3>
__device__ __forceinline__ void warp_gemm1(const float* __restrict__ wrec, int aoff,
                                           const float* __restrict__ w, int ld, int koff,
                                           int noff, float (&c)[NT][4]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll 1
  for (int ks = 0; ks < KS; ++ks) {
    FragA<PREC3> a;
    {
      const float* r0 = wrec + g * REC + aoff + ks * 8 + t;
      const float v[4] = {r0[0], r0[8 * REC], r0[4], r0[8 * REC + 4]};
      a.set(v);
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      FragB<PREC3> b;
      const int k0 = koff + ks * 8 + t, n0 = noff + nt * 8 + g;
      if (TRANS) b.set(w[n0 * ld + k0], w[n0 * ld + k0 + 4]);
      else b.set(w[k0 * ld + n0], w[(k0 + 4) * ld + n0]);
      mma<PREC3>(c[nt], a, b);
    }
  }
}
template <int NT>
__device__ __forceinline__ void zero_c1(float (&c)[NT][4]) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) c[nt][i] = 0.f;
}
template <int NT, bool RELU>
__device__ __forceinline__ void store_c1(float* __restrict__ wrec, int coff, const float (&c)[NT][4]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    float v0 = c[nt][0], v1 = c[nt][1], v2 = c[nt][2], v3 = c[nt][3];
    if (RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
    float* r0 = wrec + g * REC + coff + nt * 8 + 2 * t;
    *reinterpret_cast<float2*>(r0) = make_float2(v0, v1);
    *reinterpret_cast<float2*>(r0 + 8 * REC) = make_float2(v2, v3);
  }
}

template <bool BWD, bool PREC3, bool BPREC3>
__global__ void __launch_bounds__(GT * NGROUPS, 1) k_fused_g(const Params P, int* __restrict__ queue) {
  extern __shared__ __align__(16) float smem[];
  float* sw = smem;  // SW_TOTAL
  const int tid = threadIdx.x;
  const int grp = tid / GT, gt = tid - grp * GT;
  const int gw = gt >> 5, lane = gt & 31;
  const int g = lane >> 2, t = lane & 3;
  float* recs = sw + SW_TOTAL + grp * (GP * REC);
  float* misc = sw + SW_TOTAL + NGROUPS * (GP * REC) + grp * (GP * 6);
  float* zbuf = misc;          // GP z values
  float* sgb = misc + GP;      // GP sigma(sdf/trunc)
  float* ub = misc + 2 * GP;   // GP unnormalised weights
  float* xnb = misc + 3 * GP;  // GP*3 normalised coordinates
  const int S = P.S;

  __shared__ Lv s_lv[kL];
  __shared__ int s_unit[NGROUPS];
  load_levels(s_lv, P.g);
  for (int q = tid; q < SW_TOTAL; q += blockDim.x) sw[q] = 0.f;
  for (int q = gt; q < GP * REC; q += GT) recs[q] = 0.f;
  __syncthreads();
  for (int q = tid; q < 80 * 32; q += blockDim.x) {
    int i = q / 32, j = q % 32;
    sw[SW0 + i * LD0 + j] = P.w_sdf0[j * 80 + i];
  }
  for (int q = tid; q < 32 * 16; q += blockDim.x) {
    int i = q / 16, jp = q % 16;  // jp: geo0..14 -> torch out 1..15, jp 15 -> out 0 (sdf)
    int jt = (jp + 1) & 15;
    sw[SW1 + i * LD1 + jp] = P.w_sdf1[jt * 32 + i];
  }
  for (int q = tid; q < 63 * 32; q += blockDim.x) {
    int i = q / 32, j = q % 32;
    sw[SWC0 + i * LDC0 + j] = P.w_col0[j * 63 + i];
  }
  for (int q = tid; q < 32 * 3; q += blockDim.x) {
    int i = q / 3, k = q % 3;
    sw[SWC1 + i * LDC1 + k] = P.w_col1[k * 32 + i];
  }

  float fs_w = 0.f, sdf_w = 0.f, inv_nvalid = 0.f;
  if (BWD) {
    const float n_fs = (float)P.counts[0], n_sdf = (float)P.counts[1];
    const float n = (float)(P.counts[0] + P.counts[1]);
    fs_w = 1.0f - n_fs / n;
    sdf_w = 1.0f - n_sdf / n;
    inv_nvalid = 1.0f / (float)P.counts[2];
  }
  double l_rgb = 0.0, l_depth = 0.0, l_sdf = 0.0, l_fs = 0.0;
  float dwacc[G_SLOTS][4];
#pragma unroll
  for (int j = 0; j < G_SLOTS; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) dwacc[j][i] = 0.f;
  __syncthreads();

  const int p = gt >> 1, half = gt & 1;       // gather phases: two threads per point
  float* rec = recs + p * REC;
  float* wrec = recs + (gw * 16) * REC;       // MMA phases: this warp's 16 records
  const bool need_dx = BWD && ((P.d_rays_o != nullptr) || (P.d_rays_d != nullptr));
  const bool map_grads = BWD && (P.d_table != nullptr);
  int prev_npts = 0;

  for (;;) {
    if (gt == 0) s_unit[grp] = atomicAdd(queue, 1);
    group_sync(grp);  // also: every reader of the previous unit's records is done
    const int unit = s_unit[grp];
    if (unit >= P.n_tiles) break;
    const int r0 = unit * P.NR;
    const int nr = min(P.NR, P.R - r0);
    const int npts = nr * S;
    const bool active = p < npts;
    const bool warp_active = gw * 16 < npts;
    const int rl = active ? p / S : 0;
    const int k = active ? p - rl * S : 0;
    const int r = r0 + rl;
    float xn[3] = {0.f, 0.f, 0.f};
    float zv = 0.f;
    // ---------------- phase 1: forward -------------------------------------
    if (active) {
      zv = P.z_vals[(size_t)r * S + k];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float pt = __fadd_rn(P.rays_o[r * 3 + d], __fmul_rn(P.rays_d[r * 3 + d], zv));
        xn[d] = normalise(pt, P.g.bmin[d], P.g.bmax[d]);
      }
      if (half == 0) {
        zbuf[p] = zv;
        xnb[p * 3] = xn[0]; xnb[p * 3 + 1] = xn[1]; xnb[p * 3 + 2] = xn[2];
      }
      encode_half(P, s_lv, xn, rec, half,
                  (need_dx && P.jac) ? P.jac + ((size_t)((size_t)r * S + k) * 2 + half) * 48 : nullptr);
    } else if (p < prev_npts) {
      // rows left dirty by a larger unit must read as zero in every GEMM
      float* z0 = rec + half * (REC / 2);
#pragma unroll 1
      for (int q = 0; q < REC / 2; q += 2) *reinterpret_cast<float2*>(z0 + q) = make_float2(0.f, 0.f);
    }
    prev_npts = npts;
    __syncwarp();
    if (warp_active) {
      {  // h1 = relu(x W0^T)
        float c[4][4];
        zero_c1<4>(c);
        warp_gemm1<10, 4, false, PREC3>(wrec, R_FEAT, sw + SW0, LD0, 0, 0, c);
        store_c1<4, true>(wrec, R_H1, c);
      }
      __syncwarp();
      {  // [geo, sdf] = h1 W1^T
        float c[2][4];
        zero_c1<2>(c);
        warp_gemm1<4, 2, false, PREC3>(wrec, R_H1, sw + SW1, LD1, 0, 0, c);
        store_c1<2, false>(wrec, R_GEO, c);
      }
      __syncwarp();
      {  // c1 = relu([blob, geo, (sdf: zero weight row)] Wc0^T)
        float c[4][4];
        zero_c1<4>(c);
        warp_gemm1<8, 4, false, PREC3>(wrec, R_BLOB, sw + SWC0, LDC0, 0, 0, c);
        store_c1<4, true>(wrec, R_C1, c);
      }
      __syncwarp();
      {  // rgb logits = c1 Wc1^T (cols 0..2 of an 8-wide tile)
        float c[1][4];
        zero_c1<1>(c);
        warp_gemm1<4, 1, false, PREC3>(wrec, R_C1, sw + SWC1, LDC1, 0, 0, c);
        float* q0 = wrec + g * REC + R_RAW;
        if (t == 0) {
          q0[0] = c[0][0]; q0[1] = c[0][1];
          q0[8 * REC] = c[0][2]; q0[8 * REC + 1] = c[0][3];
        } else if (t == 1) {
          q0[2] = c[0][0];
          q0[8 * REC + 2] = c[0][2];
        }
      }
      __syncwarp();
    }
    if (active && half == 0) {
      rec[R_RAW + 3] = rec[R_SDF];
      if (P.raw) {
        float4 rw = *reinterpret_cast<const float4*>(rec + R_RAW);
        *reinterpret_cast<float4*>(P.raw + ((size_t)r * S + k) * 4) = rw;
      }
    }
    group_sync(grp);
    // ---------------- phase 2: per-ray composite, loss, d loss / d raw ----
    for (int q = gw; q < nr; q += GW) {
      const int rr = r0 + q;
      const float* zr = zbuf + q * S;
      float* sg_r = sgb + q * S;
      float* u_r = ub + q * S;
      float* rq = recs + (size_t)(q * S) * REC;
      const float tr = P.trunc;
      int first = 0x7fffffff;
      for (int kk = lane; kk < S; kk += 32) {
        float* rp = rq + kk * REC + R_RAW;
        const float s = rp[3];
        const float sg = sigmoidf_acc(s / tr);
        sg_r[kk] = sg;
        u_r[kk] = sg * sigmoidf_acc((-s) / tr);
        rp[0] = sigmoidf_acc(rp[0]);
        rp[1] = sigmoidf_acc(rp[1]);
        rp[2] = sigmoidf_acc(rp[2]);
        if (kk < S - 1 && rq[(kk + 1) * REC + R_RAW + 3] * s < 0.f) first = min(first, kk);
      }
      first = warp_min_i(first);
      if (first == 0x7fffffff) first = 0;
      __syncwarp();
      const float zlim = zr[first] + P.trunc;
      float usum = 0.f;
      for (int kk = lane; kk < S; kk += 32) {
        const float u = (zr[kk] < zlim) ? u_r[kk] : 0.f;
        u_r[kk] = u;
        usum += u;
      }
      usum = warp_sum(usum);
      const float W = usum + 1e-8f;
      float o_r = 0.f, o_g = 0.f, o_b = 0.f, o_d = 0.f, o_acc = 0.f;
      for (int kk = lane; kk < S; kk += 32) {
        const float* rp = rq + kk * REC + R_RAW;
        const float w = u_r[kk] / W;
        o_r = fmaf(w, rp[0], o_r);
        o_g = fmaf(w, rp[1], o_g);
        o_b = fmaf(w, rp[2], o_b);
        o_d = fmaf(w, zr[kk], o_d);
        o_acc += w;
      }
      o_r = warp_sum(o_r); o_g = warp_sum(o_g); o_b = warp_sum(o_b);
      o_d = warp_sum(o_d); o_acc = warp_sum(o_acc);
      if (P.depth_var) {
        float v = 0.f;
        for (int kk = lane; kk < S; kk += 32) {
          const float dz = zr[kk] - o_d;
          v = fmaf(u_r[kk] / W, dz * dz, v);
        }
        v = warp_sum(v);
        if (lane == 0) P.depth_var[rr] = v;
      }
      if (lane == 0) {
        if (P.rgb) { P.rgb[rr * 3] = o_r; P.rgb[rr * 3 + 1] = o_g; P.rgb[rr * 3 + 2] = o_b; }
        if (P.depth) P.depth[rr] = o_d;
        if (P.acc) P.acc[rr] = o_acc;
        if (P.disp) P.disp[rr] = 1.0f / fmaxf(1e-10f, o_d / o_acc);
      }
      if (BWD) {
        const float td = P.target_d[rr];
        const float tr_ = P.target_s[rr * 3], tg_ = P.target_s[rr * 3 + 1],
                    tb_ = P.target_s[rr * 3 + 2];
        const bool valid = (td > 0.f) && (td < P.depth_trunc);
        const float RS = (float)P.Rg * (float)S;
        const float c_rgb = P.ls[0] * P.w_rgb * 2.0f / (3.0f * (float)P.Rg);
        const float g_r = c_rgb * (o_r - tr_), g_g = c_rgb * (o_g - tg_), g_b = c_rgb * (o_b - tb_);
        const float g_d = valid ? P.ls[1] * P.w_depth * 2.0f * (o_d - td) * inv_nvalid : 0.f;
        if (lane == 0) {
          l_rgb += (double)((o_r - tr_) * (o_r - tr_) + (o_g - tg_) * (o_g - tg_) +
                            (o_b - tb_) * (o_b - tb_));
          if (valid) l_depth += (double)((o_d - td) * (o_d - td));
        }
        float qw = 0.f;
        for (int kk = lane; kk < S; kk += 32) {
          const float* rp = rq + kk * REC + R_RAW;
          const float q_ = g_r * rp[0] + g_g * rp[1] + g_b * rp[2] + g_d * zr[kk];
          qw = fmaf(q_, u_r[kk] / W, qw);
        }
        qw = warp_sum(qw);
        float a_fs = 0.f, a_sdf = 0.f;
        const float c_fs = P.ls[3] * P.w_fs * fs_w * 2.0f / RS;
        const float c_sdf = P.ls[2] * P.w_sdf * sdf_w * 2.0f / RS;
        for (int kk = lane; kk < S; kk += 32) {
          float* rp = rq + kk * REC + R_RAW;
          const float z = zr[kk];
          const float s = rp[3];
          const float sg = sg_r[kk];
          const float u = u_r[kk];  // a * mask
          const float w = u / W;
          const float c0 = rp[0], c1 = rp[1], c2 = rp[2];
          const float q_ = g_r * c0 + g_g * c1 + g_b * c2 + g_d * z;
          float ds = (q_ - qw) / W * u * (1.f - 2.f * sg) / tr;
          const bool front = z < __fsub_rn(td, P.trunc);
          const bool back = z > __fadd_rn(td, P.trunc);
          if (front) {
            ds += c_fs * (s - 1.f);
            a_fs += (s - 1.f) * (s - 1.f);
          }
          if (!front && !back && td > 0.f) {
            float e = (z + s * P.trunc) - td;
            ds += c_sdf * e * P.trunc;
            a_sdf += e * e;
          }
          rp[0] = g_r * w * c0 * (1.f - c0);
          rp[1] = g_g * w * c1 * (1.f - c1);
          rp[2] = g_b * w * c2 * (1.f - c2);
          rp[3] = ds;
        }
        a_fs = warp_sum(a_fs);
        a_sdf = warp_sum(a_sdf);
        if (lane == 0) { l_fs += (double)a_fs; l_sdf += (double)a_sdf; }
      }
    }
    if (!BWD) continue;  // the barrier at the top of the loop orders the next unit
    group_sync(grp);
    // ---------------- phase 3: backward ----------------------------------
    const int npts8 = (npts + 7) & ~7;
    auto dw_phase = [&](int lo, int hi, int aoff, int boff, int n_tiles_n) {
      if (!map_grads) return;
#pragma unroll
      for (int j = 0; j < G_SLOTS; ++j) {
        const int id = gw + j * GW;
        if (id >= lo && id < hi) {
          const int loc = id - lo, mt = loc / n_tiles_n, nt = loc % n_tiles_n;
          dw_tile<BPREC3>(recs, npts8, aoff + 16 * mt, boff + 8 * nt, dwacc[j]);
        }
      }
    };
    // d w_col1 += c1^T draw   (8-wide n tile over raw[4] + 4 floats of the next record: columns
    dw_phase(0, 2, R_C1, R_RAW, 1);  //  3..7 are never written out; the last record's overrun
    group_sync(grp);                 //  stays inside the group's misc block)
    float draw3 = 0.f;
    if (active) {
      const float4 dr = *reinterpret_cast<const float4*>(rec + R_RAW);
      draw3 = dr.w;
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const int j = half * 16 + jj;
        const float* w = sw + SWC1 + j * LDC1;
        const float c1 = rec[R_C1 + j];
        rec[R_C1 + j] = (c1 > 0.f) ? (dr.x * w[0] + dr.y * w[1] + dr.z * w[2]) : 0.f;
      }
    }
    group_sync(grp);
    // d w_col0 += [blob, geo, sdf]^T dc1pre   (row 63 is discarded at write-out)
    dw_phase(2, 18, R_BLOB, R_C1, 4);
    group_sync(grp);
    if (warp_active) {
      // dgeo = dc1pre Wc0[:, 48:64]  -> dH = [dgeo(15), dsdf]
      float c[2][4];
      zero_c1<2>(c);
      warp_gemm1<4, 2, true, BPREC3>(wrec, R_C1, sw + SWC0, LDC0, 0, 48, c);
      store_c1<2, false>(wrec, R_GEO, c);
    }
    __syncwarp();
    if (active && half == 0) rec[R_SDF] = draw3;
    group_sync(grp);
    // d w_sdf1 += h1^T dH
    dw_phase(18, 22, R_H1, R_GEO, 2);
    group_sync(grp);
    if (warp_active) {
      // dh1pre = (h1 > 0) * (dH W1)
      float c[4][4];
      zero_c1<4>(c);
      warp_gemm1<2, 4, true, BPREC3>(wrec, R_GEO, sw + SW1, LD1, 0, 0, c);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float* q0 = wrec + g * REC + R_H1 + nt * 8 + 2 * t;
        float2 h0 = *reinterpret_cast<float2*>(q0);
        float2 h8 = *reinterpret_cast<float2*>(q0 + 8 * REC);
        *reinterpret_cast<float2*>(q0) =
            make_float2(h0.x > 0.f ? c[nt][0] : 0.f, h0.y > 0.f ? c[nt][1] : 0.f);
        *reinterpret_cast<float2*>(q0 + 8 * REC) =
            make_float2(h8.x > 0.f ? c[nt][2] : 0.f, h8.y > 0.f ? c[nt][3] : 0.f);
      }
    }
    group_sync(grp);
    // d w_sdf0 += x^T dh1pre
    dw_phase(22, 42, R_FEAT, R_H1, 4);
    // (readers of C1 / GEO slots are done: dW phases 2 and 3 finished before the last barrier)
    float cfeat[4][4];
    zero_c1<4>(cfeat);
    if (warp_active) {
      if (need_dx) {
        // dblob = dc1pre Wc0[:, 0:48] + dh1pre W0[:, 32:80]  -> scratch (C1 slots 0..31, GEO 32..47)
        float c[6][4];
        zero_c1<6>(c);
        warp_gemm1<4, 6, true, BPREC3>(wrec, R_C1, sw + SWC0, LDC0, 0, 0, c);
        warp_gemm1<4, 6, true, BPREC3>(wrec, R_H1, sw + SW0, LD0, 0, 32, c);
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
          const int col = nt * 8 + 2 * t;
          const int off = (col < 32) ? (R_C1 + col) : (R_GEO + col - 32);
          float* q0 = wrec + g * REC + off;
          *reinterpret_cast<float2*>(q0) = make_float2(c[nt][0], c[nt][1]);
          *reinterpret_cast<float2*>(q0 + 8 * REC) = make_float2(c[nt][2], c[nt][3]);
        }
      }
      // dfeat = dh1pre W0[:, 0:32]
      if (map_grads || need_dx) warp_gemm1<4, 4, true, BPREC3>(wrec, R_H1, sw + SW0, LD0, 0, 0, cfeat);
    }
    group_sync(grp);  // every warp is past the x^T dh1pre tiles: the FEAT slots are free
    if (warp_active && (map_grads || need_dx)) store_c1<4, false>(wrec, R_FEAT, cfeat);
    group_sync(grp);
    // ---- hash backward, point-parallel and BALANCED: thread (pp, half) owns levels half, half+2, ..
    // of point pp = every 6th point of the unit per warp (a ray's samples behind the surface carry
    // no gradient: ray-major assignment would leave whole warps idle).  d loss / d xn comes from
    // the Jacobian cached by the forward gather (no second gather); the table gradient is
    // scattered with red.global.add.v2.
    {
      const int qd = gt >> 1;
      const int pp = (qd & 15) * GW + (qd >> 4);
      const bool act2 = pp < npts;
      float* rec2 = recs + pp * REC;
      float dxh[3] = {0.f, 0.f, 0.f};
      float xq[3] = {0.f, 0.f, 0.f};
      if (act2) { xq[0] = xnb[pp * 3]; xq[1] = xnb[pp * 3 + 1]; xq[2] = xnb[pp * 3 + 2]; }
      if (act2 && (map_grads || need_dx)) {
        const float* jp = (need_dx && P.jac) ? P.jac + ((size_t)((size_t)r0 * S + pp) * 2 + half) * 48 : nullptr;
#pragma unroll 1
        for (int i2 = 0; i2 < kL / 2; i2 += 2) {  // two levels per call: half + 2 i2, half + 2 i2 + 2
          bool on[2];
          float gg[2][2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int l = half + 2 * (i2 + e);
            const float2 gv = *reinterpret_cast<const float2*>(rec2 + R_FEAT + 2 * l);
            gg[e][0] = gv.x; gg[e][1] = gv.y;
            on[e] = l < P.g.n_levels && (gv.x != 0.f || gv.y != 0.f);
            if (jp && on[e]) {
              const float2 j0 = *reinterpret_cast<const float2*>(jp + 6 * (i2 + e));
              const float2 j1 = *reinterpret_cast<const float2*>(jp + 6 * (i2 + e) + 2);
              const float2 j2 = *reinterpret_cast<const float2*>(jp + 6 * (i2 + e) + 4);
              dxh[0] = fmaf(gv.x, j0.x, fmaf(gv.y, j1.y, dxh[0]));
              dxh[1] = fmaf(gv.x, j0.y, fmaf(gv.y, j2.x, dxh[1]));
              dxh[2] = fmaf(gv.x, j1.x, fmaf(gv.y, j2.y, dxh[2]));
            }
          }
          if ((map_grads || (need_dx && !jp)) && (on[0] || on[1])) {
            const float3 d3 = hash_backward_multi<2>(P, s_lv, half + 2 * i2, on, xq[0], xq[1], xq[2],
                                                     gg, need_dx && !jp, map_grads, 2);
            dxh[0] += d3.x; dxh[1] += d3.y; dxh[2] += d3.z;
          }
        }
      }
      if (need_dx) {
        // the two halves of a point are adjacent lanes: sum their level shares
#pragma unroll
        for (int d = 0; d < 3; ++d) dxh[d] += __shfl_xor_sync(0xffffffffu, dxh[d], 1);
        float dp0 = 0.f, dp1 = 0.f;
        if (act2) {
          auto dblob = [&](int i) { return (i < 32) ? rec2[R_C1 + i] : rec2[R_GEO + i - 32]; };
          if (half == 0) {
            dp0 = (float)((double)(blob_backward_dim(xq[0], 0, dblob) + dxh[0]) * P.g.binv[0]);
            dp1 = (float)((double)(blob_backward_dim(xq[1], 1, dblob) + dxh[1]) * P.g.binv[1]);
          } else {
            dp0 = (float)((double)(blob_backward_dim(xq[2], 2, dblob) + dxh[2]) * P.g.binv[2]);
          }
        }
        __syncwarp();  // both halves have read the dblob scratch before C1[0..5] is reused
        if (act2) {
          const float zq = zbuf[pp];
          // park d loss / d pts and z * d loss / d pts in the (dead) C1 slots 0..5
          if (half == 0) {
            rec2[R_C1 + 0] = dp0; rec2[R_C1 + 3] = dp0 * zq;
            rec2[R_C1 + 1] = dp1; rec2[R_C1 + 4] = dp1 * zq;
          } else {
            rec2[R_C1 + 2] = dp0; rec2[R_C1 + 5] = dp0 * zq;
          }
        }
      }
    }
    if (need_dx) {
      group_sync(grp);
      for (int q = gw; q < nr; q += GW) {
        float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int kk = lane; kk < S; kk += 32) {
          const float* rp = recs + (size_t)(q * S + kk) * REC + R_C1;
#pragma unroll
          for (int d = 0; d < 6; ++d) a[d] += rp[d];
        }
#pragma unroll
        for (int d = 0; d < 6; ++d) a[d] = warp_sum(a[d]);
        if (lane == 0) {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            if (P.d_rays_o) P.d_rays_o[(r0 + q) * 3 + d] = a[d];
            if (P.d_rays_d) P.d_rays_d[(r0 + q) * 3 + d] = a[3 + d];
          }
        }
      }
    }
  }

  if (BWD) {
    if (lane == 0) {
      if (l_rgb != 0.0) atomicAdd(&P.loss_acc[0], l_rgb);
      if (l_depth != 0.0) atomicAdd(&P.loss_acc[1], l_depth);
      if (l_sdf != 0.0) atomicAdd(&P.loss_acc[2], l_sdf);
      if (l_fs != 0.0) atomicAdd(&P.loss_acc[3], l_fs);
    }
    if (map_grads) {
      // sum the weight-gradient tiles of the NGROUPS groups in shared memory (the record area is
      // dead now), then one red.global per element and CTA
      float* red = sw + SW_TOTAL;  // [42 tiles][4][32 lanes]
      __syncthreads();
      for (int gsel = 0; gsel < NGROUPS; ++gsel) {
        if (grp == gsel) {
#pragma unroll
          for (int j = 0; j < G_SLOTS; ++j) {
            const int id = gw + j * GW;
            if (id >= DW_TILES) continue;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float* q0 = red + (id * 4 + i) * 32 + lane;
              *q0 = (gsel == 0) ? dwacc[j][i] : (*q0 + dwacc[j][i]);
            }
          }
        }
        __syncthreads();
      }
      if (grp == 0) {
#pragma unroll
        for (int j = 0; j < G_SLOTS; ++j) {
          const int id = gw + j * GW;
          if (id >= DW_TILES) continue;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float v = red[(id * 4 + i) * 32 + lane];
            const int rloc = g + ((i & 2) ? 8 : 0), cloc = 2 * t + (i & 1);
            if (id < 2) {  // d w_col1 [3][32]
              const int in = 16 * id + rloc, out = cloc;
              if (out < 3) red_add(P.d_w_col1 + out * 32 + in, v);
            } else if (id < 18) {  // d w_col0 [32][63]
              const int loc = id - 2, in = 16 * (loc / 4) + rloc, out = 8 * (loc % 4) + cloc;
              if (in < 63) red_add(P.d_w_col0 + out * 63 + in, v);
            } else if (id < 22) {  // d w_sdf1 [16][32], stored column jp -> torch row (jp+1)&15
              const int loc = id - 18, in = 16 * (loc / 2) + rloc, jp = 8 * (loc % 2) + cloc;
              red_add(P.d_w_sdf1 + ((jp + 1) & 15) * 32 + in, v);
            } else {  // d w_sdf0 [32][80]
              const int loc = id - 22, in = 16 * (loc / 4) + rloc, out = 8 * (loc % 4) + cloc;
              red_add(P.d_w_sdf0 + out * 80 + in, v);
            }
          }
        }
      }
    }
  }
}

struct FinalizeParams {
  const double* loss_acc;
  const int* counts;
  float* losses;
  int R, S;
  float w_rgb, w_depth, w_sdf, w_fs;
};

__global__ void k_finalize(FinalizeParams p) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float n_fs = (float)p.counts[0], n_sdf = (float)p.counts[1];
  const float n = (float)(p.counts[0] + p.counts[1]);
  const float fs_w = 1.0f - n_fs / n, sdf_w = 1.0f - n_sdf / n;
  const double RS = (double)p.R * (double)p.S;
  p.losses[0] = (float)(p.loss_acc[0] / (3.0 * p.R)) * p.w_rgb;
  p.losses[1] = (float)(p.loss_acc[1] / (double)p.counts[2]) * p.w_depth;  // NaN if no valid depth, like torch
  p.losses[2] = (float)(p.loss_acc[2] / RS) * sdf_w * p.w_sdf;
  p.losses[3] = (float)(p.loss_acc[3] / RS) * fs_w * p.w_fs;
}

// ------------------------------------------------------------- smoothness ---
struct SmoothParams {
  GridDev g;
  const float* table;
  int n;  // lattice side (sample_points - 1)
  double voxel, off[3], rnd[3];
  const float* rand_dev;  // DEVICE [6] overriding off/rnd (graph replay)
  double offmax[3], margin;
  float* feat;    // [n^3, 32]
  float* d_table;
  double* loss_acc;
  float coef;     // grad_scale * weight / sample_points^3
};

__device__ __forceinline__ void smooth_xn(const SmoothParams& p, int ix, int iy, int iz,
                                          float xn[3]) {
  const int c[3] = {ix, iy, iz};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    // pts = (coords + rand) * voxel + bb_min + offset ; then (pts - bb_min)/(bb_max-bb_min), all f64
    const double rnd = p.rand_dev ? (double)p.rand_dev[3 + d] : p.rnd[d];
    const double off = p.rand_dev ? (double)p.rand_dev[d] * p.offmax[d] + p.margin : p.off[d];
    double pt = ((double)(float)c[d] + rnd) * p.voxel + p.g.bmin[d] + off;
    xn[d] = (float)((pt - p.g.bmin[d]) / (p.g.bmax[d] - p.g.bmin[d]));
  }
}

__global__ void __launch_bounds__(256) k_smooth_fwd(SmoothParams p) {
  const int n = p.n, N = n * n * n;
  __shared__ Lv s_lv[kL];
  load_levels(s_lv, p.g);
  __syncthreads();
  const int q = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (point, level)
  const int pt = q / kL, l = q % kL;
  if (pt >= N || l >= p.g.n_levels) return;
  const int iz = pt % n, iy = (pt / n) % n, ix = pt / (n * n);
  float xn[3];
  smooth_xn(p, ix, iy, iz, xn);
  const float2* tab = reinterpret_cast<const float2*>(p.table);
  float w[3];
  uint32_t c[3];
  const Lv L = s_lv[l];
  for (int d = 0; d < 3; ++d) pos_fract(xn[d], L.scale, w[d], c[d]);
  float f0 = 0.f, f1 = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint32_t idx = grid_index(L, c[0] + (k & 1), c[1] + ((k >> 1) & 1), c[2] + ((k >> 2) & 1));
    float wk = ((k & 1) ? w[0] : 1.f - w[0]) * ((k & 2) ? w[1] : 1.f - w[1]) *
               ((k & 4) ? w[2] : 1.f - w[2]);
    float2 v = __ldg(&tab[idx]);
    f0 = fmaf(wk, v.x, f0);
    f1 = fmaf(wk, v.y, f1);
  }
  reinterpret_cast<float2*>(p.feat)[(size_t)pt * kL + l] = make_float2(f0, f1);
}

__global__ void __launch_bounds__(256) k_smooth_bwd(SmoothParams p) {
  __shared__ Lv s_lv[kL];
  load_levels(s_lv, p.g);
  __syncthreads();
  const int n = p.n, N = n * n * n;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int pt = q / kL, l = q % kL;
  float part = 0.f;
  if (pt < N && l < p.g.n_levels) {
    const int iz = pt % n, iy = (pt / n) % n, ix = pt / (n * n);
    const float2* F = reinterpret_cast<const float2*>(p.feat);
    const float2 f = F[(size_t)pt * kL + l];
    float g0 = 0.f, g1 = 0.f;
    const int stride[3] = {n * n, n, 1};
    const int ci[3] = {ix, iy, iz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (ci[d] > 0) {
        float2 o = F[(size_t)(pt - stride[d]) * kL + l];
        float a = f.x - o.x, b = f.y - o.y;
        g0 += a; g1 += b;
        part += a * a + b * b;  // each difference counted once (from its upper end)
      }
      if (ci[d] < n - 1) {
        float2 o = F[(size_t)(pt + stride[d]) * kL + l];
        g0 -= o.x - f.x; g1 -= o.y - f.y;
      }
    }
    if (p.d_table) {
      g0 *= 2.f * p.coef; g1 *= 2.f * p.coef;
      float xn[3];
      smooth_xn(p, ix, iy, iz, xn);
      float w[3];
      uint32_t c[3];
      const Lv L = s_lv[l];
      for (int d = 0; d < 3; ++d) pos_fract(xn[d], L.scale, w[d], c[d]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        uint32_t idx = grid_index(L, c[0] + (k & 1), c[1] + ((k >> 1) & 1), c[2] + ((k >> 2) & 1));
        float wk = ((k & 1) ? w[0] : 1.f - w[0]) * ((k & 2) ? w[1] : 1.f - w[1]) *
                   ((k & 4) ? w[2] : 1.f - w[2]);
        red_add_v2(p.d_table + 2 * (size_t)idx, wk * g0, wk * g1);
      }
    }
  }
  // one double atomic per CTA (per warp it was ~15 k serialised atomics on one address)
  __shared__ float s_part[8];
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < 8; ++i) a += (double)s_part[i];
    if (a != 0.0) atomicAdd(p.loss_acc, a);
  }
}

__global__ void k_smooth_finalize(const double* acc, float* loss, float scale) {
  if (threadIdx.x == 0 && blockIdx.x == 0) loss[0] = (float)(acc[0] * (double)scale);
}

// hash encode only
__global__ void __launch_bounds__(256) k_encode(GridDev g, const float* table, const float* x,
                                                int n, float* feat, uint32_t* idx_out) {
  __shared__ Lv s_lv[kL];
  load_levels(s_lv, g);
  __syncthreads();
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int pt = q / kL, l = q % kL;
  if (pt >= n || l >= g.n_levels) return;
  const Lv L = s_lv[l];
  const float2* tab = reinterpret_cast<const float2*>(table);
  float w[3];
  uint32_t c[3];
  for (int d = 0; d < 3; ++d) pos_fract(x[pt * 3 + d], L.scale, w[d], c[d]);
  float f0 = 0.f, f1 = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint32_t idx = grid_index(L, c[0] + (k & 1), c[1] + ((k >> 1) & 1), c[2] + ((k >> 2) & 1));
    if (idx_out) idx_out[((size_t)pt * g.n_levels + l) * 8 + k] = idx;
    float wk = ((k & 1) ? w[0] : 1.f - w[0]) * ((k & 2) ? w[1] : 1.f - w[1]) *
               ((k & 4) ? w[2] : 1.f - w[2]);
    float2 v = __ldg(&tab[idx]);
    f0 = fmaf(wk, v.x, f0);
    f1 = fmaf(wk, v.y, f1);
  }
  if (feat) {
    feat[(size_t)pt * 2 * g.n_levels + 2 * l] = f0;
    feat[(size_t)pt * 2 * g.n_levels + 2 * l + 1] = f1;
  }
}

// ------------------------------------------------------- point queries (mesher) ---
// query_color_sdf / query_sdf at arbitrary points (joint_encoding.py:425-481): thread per point,
// hash + OneBlob encoding into a shared-memory record, the two decoders in plain fp32 FMAs
// (the mesher path is not hot; exact fp32 keeps it within 1e-6 of the oracle).
struct QueryParams {
  GridDev g;
  const float* table;
  const float *w_sdf0, *w_sdf1, *w_col0, *w_col1;
  const float* pts;   // [P,3]
  int P, normalised;  // normalised = 1: pts are already (p - min) / (max - min)
  float* raw;         // [P,4] rgb logits ++ sdf, or NULL
  float* geo;         // [P,15] or NULL
  float* feat;        // [P,32] hash features (query_sdf(embed=True)) or NULL
};

__global__ void __launch_bounds__(128) k_query(const QueryParams Q) {
  extern __shared__ __align__(16) float qrec[];  // 128 records
  __shared__ Lv s_lv[kL];
  load_levels(s_lv, Q.g);
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Q.P) return;
  float* rec = qrec + threadIdx.x * REC;
  float xn[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float v = Q.pts[(size_t)p * 3 + d];
    xn[d] = Q.normalised ? v : normalise(v, Q.g.bmin[d], Q.g.bmax[d]);
  }
  Params P{};  // encode_point reads table / n_levels only
  P.table = Q.table;
  P.g.n_levels = Q.g.n_levels;
  encode_point(P, s_lv, xn, rec);
  if (Q.feat)
#pragma unroll 4
    for (int j = 0; j < 32; ++j) Q.feat[(size_t)p * 32 + j] = rec[R_FEAT + j];
  if (!Q.raw && !Q.geo) return;
  // h1 = relu(W0 [feat, blob]);  [sdf, geo] = W1 h1  (torch rows: 0 = sdf, 1..15 = geo)
  float h1[32];
#pragma unroll 1
  for (int j = 0; j < 32; ++j) {
    float a = 0.f;
    const float* w = Q.w_sdf0 + j * 80;
    for (int i = 0; i < 80; ++i) a = fmaf(__ldg(w + i), rec[R_FEAT + i], a);  // FEAT|BLOB contiguous
    h1[j] = fmaxf(a, 0.f);
  }
  float so[16];
#pragma unroll 1
  for (int j = 0; j < 16; ++j) {
    float a = 0.f;
    const float* w = Q.w_sdf1 + j * 32;
#pragma unroll
    for (int i = 0; i < 32; ++i) a = fmaf(__ldg(w + i), h1[i], a);
    so[j] = a;
  }
  if (Q.geo)
    for (int j = 0; j < 15; ++j) Q.geo[(size_t)p * 15 + j] = so[1 + j];
  if (!Q.raw) return;
  // colour: c1 = relu(Wc0 [blob, geo]);  rgb logits = Wc1 c1
  float c1[32];
#pragma unroll 1
  for (int j = 0; j < 32; ++j) {
    float a = 0.f;
    const float* w = Q.w_col0 + j * 63;
    for (int i = 0; i < 48; ++i) a = fmaf(__ldg(w + i), rec[R_BLOB + i], a);
    for (int i = 0; i < 15; ++i) a = fmaf(__ldg(w + 48 + i), so[1 + i], a);
    c1[j] = fmaxf(a, 0.f);
  }
  float o3[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) a = fmaf(__ldg(Q.w_col1 + k * 32 + i), c1[i], a);
    o3[k] = a;
  }
  *reinterpret_cast<float4*>(Q.raw + (size_t)p * 4) = make_float4(o3[0], o3[1], o3[2], so[0]);
}

static int fill_grid(GridDev& g, const XrdHashGrid* h) {
  if (h->n_levels < 1 || h->n_levels > kL) return XRD_E_SHAPE;
  g.n_levels = h->n_levels;
  for (int l = 0; l < kL; ++l) {
    g.scale[l] = h->scale[l]; g.res[l] = h->resolution[l]; g.size[l] = h->size[l] ? h->size[l] : 1;
    g.offset[l] = h->offset[l]; g.hashed[l] = h->hashed[l];
    g.magic[l] = (uint32_t)(0x100000000ull / g.size[l]);
  }
  for (int d = 0; d < 3; ++d) {
    g.bmin[d] = h->bbox_min[d]; g.bmax[d] = h->bbox_max[d];
    g.binv[d] = 1.0 / (h->bbox_max[d] - h->bbox_min[d]);
  }
  return XRD_OK;
}

}  // namespace coslam
}  // namespace xrd

using namespace xrd;
using namespace xrd::coslam;

// Optional Jacobian cache of the grouped kernel (2 halves x 8 levels x 6 floats per sample point):
// the forward gather stores d feat / d xn so that the backward needs no second gather for the ray
// gradients.  Measured on B200 (profiles/r02_coslam_kernel_ab.txt) it LOSES 2-4 % against
// re-gathering (the table is L2-resident, the cache is 384 B/point of extra traffic), so it is
// off unless XRD_COSLAM_JAC is set; kept for tables that outgrow L2.
static constexpr size_t kJacBytesPerPoint = 2 * 48 * sizeof(float);
static constexpr size_t kJacMaxBytes = (size_t)512 << 20;  // beyond this the backward re-gathers
static bool jac_enabled() {
  static const bool on = getenv("XRD_COSLAM_JAC") != nullptr;
  return on;
}
static size_t jac_bytes(int n_rays, int n_samples) {
  if (!jac_enabled()) return 0;
  const size_t b = (size_t)n_rays * n_samples * kJacBytesPerPoint;
  return b <= kJacMaxBytes ? b : 0;
}
extern "C" size_t xrd_coslam_workspace_bytes(int n_rays, int n_samples) {
  // [counts int[4] | loss_acc double[4] | z_vals R*S floats | Jacobian cache]
  return 256 + align_up((size_t)n_rays * n_samples * sizeof(float), 256) + jac_bytes(n_rays, n_samples);
}

static int pick_rays_per_tile(int S, int requested) {
  if (requested > 0) return requested;
  // largest NR with NR*S <= 224 -> for S=43: 5 rays, 215 points, 224 threads
  int nr = 224 / S;
  return nr < 1 ? 1 : nr;  // S in (224, 256]: one ray per 256-thread tile
}

extern "C" int xrd_coslam_step(const XrdRays* rays, const XrdHashGrid* grid,
                               const XrdCoslamMlp* mlp, const XrdCoslamCfg* cfg,
                               const float* noise, XrdCoslamOut* out, XrdCoslamGrads* grads,
                               void* workspace, size_t workspace_bytes, void* stream_) {
  if (!rays || !grid || !mlp || !cfg || !out || !workspace) return XRD_E_NULL;
  if (!rays->rays_o || !rays->rays_d || !grid->table || !mlp->w_sdf0 || !mlp->w_sdf1 ||
      !mlp->w_col0 || !mlp->w_col1)
    return XRD_E_NULL;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int R = rays->n_rays, S = cfg->n_samples;
  if (R <= 0) return XRD_OK;
  if (S < 2 || S > 256) return XRD_E_SHAPE;
  const bool has_depth = rays->target_d != nullptr;
  if (has_depth && S != cfg->n_sample_d + cfg->n_range_d) return XRD_E_SHAPE;
  if (has_depth && (!cfg->lin_uniform || !cfg->lin_range || !cfg->lin_nodepth)) return XRD_E_NULL;
  if (!has_depth && !cfg->lin_full) return XRD_E_NULL;
  if (grads && (!has_depth || !rays->target_s || !out->losses)) return XRD_E_NULL;
  if (grads) {
    // map gradients are all-or-nothing; d_table == NULL selects the pose-only (tracking) pass
    const int n_map = (grads->d_table != nullptr) + (grads->d_w_sdf0 != nullptr) +
                      (grads->d_w_sdf1 != nullptr) + (grads->d_w_col0 != nullptr) +
                      (grads->d_w_col1 != nullptr);
    if (n_map != 0 && n_map != 5) return XRD_E_NULL;
    if (n_map == 0 && !grads->d_rays_o && !grads->d_rays_d) return XRD_E_NULL;
  }
  if (workspace_bytes < xrd_coslam_workspace_bytes(R, S)) return XRD_E_WORKSPACE;

  int* counts = reinterpret_cast<int*>(workspace);  // n_fs, n_sdf, n_valid
  double* loss_acc = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 64);
  float* z_ws = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
  float* z_vals = out->z_vals ? out->z_vals : z_ws;
  const int phase = cfg->phase;  // 0: sample+render, 1: sample only, 2: render only
  if (phase < 0 || phase > 2) return XRD_E_SHAPE;
  if (phase != 0 && !out->z_vals) return XRD_E_NULL;
  if (phase == 1 && !cfg->counts_out) return XRD_E_NULL;
  XRD_CUDA_TRY(cudaMemsetAsync(workspace, 0, 256, stream));
  if (phase == 1) {
    counts = cfg->counts_out;
    XRD_CUDA_TRY(cudaMemsetAsync(counts, 0, 4 * sizeof(int), stream));
  }
  SampleParams sp;
  sp.R = R; sp.S = S; sp.n_a = cfg->n_sample_d; sp.n_b = cfg->n_range_d;
  sp.perturb = cfg->perturb; sp.has_depth = has_depth;
  sp.target_d = rays->target_d; sp.lin_uniform = cfg->lin_uniform; sp.lin_range = cfg->lin_range;
  sp.lin_nodepth = cfg->lin_nodepth; sp.lin_full = cfg->lin_full; sp.noise = noise;
  sp.trunc = cfg->trunc; sp.depth_trunc = cfg->depth_trunc; sp.seed = cfg->seed;
  sp.seed_dev = cfg->seed_dev;
  sp.z_vals = z_vals; sp.counts = counts;
  if (phase != 2) {
    k_sample<<<(R + 3) / 4, 128, 0, stream>>>(sp);
    XRD_LAUNCH_CHECK();
  }
  if (phase == 1) return XRD_OK;

  Params P;
  int st = fill_grid(P.g, grid);
  if (st != XRD_OK) return st;
  P.R = R; P.S = S;
  P.Rg = cfg->n_rays_global > 0 ? cfg->n_rays_global : R;
  P.rays_o = rays->rays_o; P.rays_d = rays->rays_d; P.target_s = rays->target_s; P.target_d = rays->target_d;
  P.z_vals = z_vals; P.table = grid->table;
  P.w_sdf0 = mlp->w_sdf0; P.w_sdf1 = mlp->w_sdf1; P.w_col0 = mlp->w_col0; P.w_col1 = mlp->w_col1;
  P.trunc = cfg->trunc; P.depth_trunc = cfg->depth_trunc;
  P.w_rgb = cfg->w_rgb; P.w_depth = cfg->w_depth; P.w_sdf = cfg->w_sdf; P.w_fs = cfg->w_fs;
  P.rgb = out->rgb; P.depth = out->depth; P.disp = out->disp; P.acc = out->acc;
  P.depth_var = out->depth_var; P.raw = out->raw;
  P.counts = cfg->counts_global ? cfg->counts_global : counts; P.loss_acc = loss_acc;
  for (int i = 0; i < 4; ++i) P.ls[i] = grads ? grads->loss_scale[i] : 0.f;
  if (grads) {
    P.d_table = grads->d_table; P.d_w_sdf0 = grads->d_w_sdf0; P.d_w_sdf1 = grads->d_w_sdf1;
    P.d_w_col0 = grads->d_w_col0; P.d_w_col1 = grads->d_w_col1;
    P.d_rays_o = grads->d_rays_o; P.d_rays_d = grads->d_rays_d;
  } else {
    P.d_table = P.d_w_sdf0 = P.d_w_sdf1 = P.d_w_col0 = P.d_w_col1 = P.d_rays_o = P.d_rays_d = nullptr;
  }
  P.jac = nullptr;
  const int sms = num_sms();
  // Kernel choice (rays_per_tile: 0 = automatic, -1 = tile kernel, -2 = grouped kernel, > 0 = tile
  // kernel with that many rays per tile).  Automatic: the grouped kernel wins the gradient pass
  // once every group has >= 1 unit in flight after the first wave (R = 4096: 436 vs 459 us,
  // 16384: 1461 vs 1607, 65536: 5629 vs 6230); the tile kernel wins small batches (tracking,
  // R = 1024: 182 vs 204 us) and forward-only rendering (160 vs 182 us at 4096 rays).
  bool grouped = false;
  if (S <= GP) {
    const int units = (R + GP / S - 1) / (GP / S);
    grouped = cfg->rays_per_tile == -2 || (cfg->rays_per_tile == 0 && grads && units >= 4 * sms);
  }
  if (grouped) {
    // grouped persistent kernel (k_fused_g): units of NR rays pulled from a queue
    P.NR = GP / S;
    P.n_tiles = (R + P.NR - 1) / P.NR;
    P.jac = (grads && (grads->d_rays_o || grads->d_rays_d) && jac_bytes(R, S))
                ? reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256 +
                                           align_up((size_t)R * S * sizeof(float), 256))
                : nullptr;
    int* queue = reinterpret_cast<int*>(reinterpret_cast<char*>(workspace) + 128);  // zeroed above
    const size_t smem_g = sizeof(float) * ((size_t)SW_TOTAL + (size_t)NGROUPS * GP * (REC + 6));
    int gridx = (P.n_tiles + NGROUPS - 1) / NGROUPS;
    if (gridx > sms) gridx = sms;
#define XRD_LAUNCH_G(KERNEL)                                                                      \
  do {                                                                                            \
    XRD_CUDA_TRY(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                      (int)smem_g));                                              \
    KernelTimer kt(stream);                                                                       \
    KERNEL<<<gridx, GT * NGROUPS, smem_g, stream>>>(P, queue);                                    \
  } while (0)
    if (grads) {
      switch (cfg->precision) {
        case 0: XRD_LAUNCH_G((k_fused_g<true, true, true>)); break;
        case 1: XRD_LAUNCH_G((k_fused_g<true, true, false>)); break;
        case 2: XRD_LAUNCH_G((k_fused_g<true, false, false>)); break;
        default: return XRD_E_SHAPE;
      }
      XRD_LAUNCH_CHECK();
      FinalizeParams fp{loss_acc, P.counts, out->losses, P.Rg, S, cfg->w_rgb, cfg->w_depth, cfg->w_sdf, cfg->w_fs};
      k_finalize<<<1, 32, 0, stream>>>(fp);
      XRD_LAUNCH_CHECK();
    } else {
      if (cfg->precision <= 1) XRD_LAUNCH_G((k_fused_g<false, true, false>));
      else XRD_LAUNCH_G((k_fused_g<false, false, false>));
      XRD_LAUNCH_CHECK();
    }
#undef XRD_LAUNCH_G
    return XRD_OK;
  }
  // tile kernel (k_fused): rays_per_tile > 0 selects its tile size, < 0 its default
  int NR = pick_rays_per_tile(S, cfg->rays_per_tile);
  if (NR < 1 || NR * S > 256) return XRD_E_SHAPE;
  int threads = (NR * S + 31) / 32 * 32;
  if (grads && threads < 224) threads = 224;  // 7+ warps own the 42 weight-gradient tiles (6 each)
  P.NR = NR;
  P.n_tiles = (R + NR - 1) / NR;
  const size_t smem = sizeof(float) * ((size_t)SW_TOTAL + (size_t)(threads + 8) * REC + 6 * (size_t)NR * S);
  const int nwarps = threads / 32;
  const int slots = (DW_TILES + nwarps - 1) / nwarps;
#define XRD_LAUNCH_FUSED(KERNEL)                                                                  \
  do {                                                                                            \
    XRD_CUDA_TRY(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                      (int)smem));                                                \
    int occ = 1;                                                                                  \
    XRD_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, KERNEL, threads, smem));     \
    if (occ < 1) return XRD_E_SHAPE;                                                              \
    int gridx = P.n_tiles < sms * occ ? P.n_tiles : sms * occ;                                    \
    KernelTimer kt(stream);                                                                       \
    KERNEL<<<gridx, threads, smem, stream>>>(P);                                                  \
  } while (0)
  if (grads) {
    if (slots != 6) return XRD_E_SHAPE;  // >= 7 warps own the 42 weight-gradient tiles
    switch (cfg->precision) {
      case 0: XRD_LAUNCH_FUSED((k_fused<true, true, true, 6>)); break;
      case 1: XRD_LAUNCH_FUSED((k_fused<true, true, false, 6>)); break;
      case 2: XRD_LAUNCH_FUSED((k_fused<true, false, false, 6>)); break;
      default: return XRD_E_SHAPE;
    }
    XRD_LAUNCH_CHECK();
    FinalizeParams fp{loss_acc, P.counts, out->losses, P.Rg, S, cfg->w_rgb, cfg->w_depth, cfg->w_sdf, cfg->w_fs};
    k_finalize<<<1, 32, 0, stream>>>(fp);
    XRD_LAUNCH_CHECK();
  } else {
    if (cfg->precision <= 1) XRD_LAUNCH_FUSED((k_fused<false, true, false, 1>));
    else XRD_LAUNCH_FUSED((k_fused<false, false, false, 1>));
    XRD_LAUNCH_CHECK();
  }
#undef XRD_LAUNCH_FUSED
  return XRD_OK;
}

extern "C" size_t xrd_coslam_smoothness_workspace_bytes(int sample_points) {
  size_t n = (size_t)(sample_points - 1);
  return 256 + n * n * n * 32 * sizeof(float);
}

static int smoothness_impl(const XrdHashGrid* grid, int sample_points, double voxel_size,
                           double margin, float weight, const float* smooth_rand, bool rand_on_device,
                           float* loss, float* d_table, float grad_scale,
                           void* workspace, size_t workspace_bytes, void* stream_) {
  if (!grid || !grid->table || !smooth_rand || !loss || !workspace) return XRD_E_NULL;
  if (sample_points < 3) return XRD_E_SHAPE;
  if (workspace_bytes < xrd_coslam_smoothness_workspace_bytes(sample_points)) return XRD_E_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  SmoothParams p;
  int st = fill_grid(p.g, grid);
  if (st != XRD_OK) return st;
  p.table = grid->table;
  p.n = sample_points - 1;
  // joint_encoding.py:171-181 -- float64 throughout (bounding_box is a float64 tensor); the
  // python floats voxel_size/margin enter as doubles, the two torch.rand draws as float32
  p.voxel = voxel_size;
  const double grid_size = (double)(sample_points - 1) * voxel_size;
  for (int d = 0; d < 3; ++d) {
    double offset_max = grid->bbox_max[d] - grid->bbox_min[d] - grid_size - 2.0 * margin;
    p.offmax[d] = offset_max;
    p.off[d] = rand_on_device ? 0.0 : (double)smooth_rand[d] * offset_max + margin;
    p.rnd[d] = rand_on_device ? 0.0 : (double)smooth_rand[3 + d];
  }
  p.margin = margin;
  p.rand_dev = rand_on_device ? smooth_rand : nullptr;
  p.loss_acc = reinterpret_cast<double*>(workspace);
  p.feat = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
  p.d_table = d_table;
  const double sp3 = (double)sample_points * sample_points * sample_points;
  p.coef = (float)((double)grad_scale * (double)weight / sp3);
  XRD_CUDA_TRY(cudaMemsetAsync(workspace, 0, 256, stream));
  const int N = p.n * p.n * p.n;
  const int blocks = (N * kL + 255) / 256;
  k_smooth_fwd<<<blocks, 256, 0, stream>>>(p);
  XRD_LAUNCH_CHECK();
  k_smooth_bwd<<<blocks, 256, 0, stream>>>(p);
  XRD_LAUNCH_CHECK();
  k_smooth_finalize<<<1, 32, 0, stream>>>(p.loss_acc, loss, (float)((double)weight / sp3));
  XRD_LAUNCH_CHECK();
  return XRD_OK;
}

extern "C" int xrd_coslam_smoothness(const XrdHashGrid* grid, int sample_points, double voxel_size,
                                     double margin, float weight, const float* smooth_rand,
                                     float* loss, float* d_table, float grad_scale,
                                     void* workspace, size_t workspace_bytes, void* stream_) {
  return smoothness_impl(grid, sample_points, voxel_size, margin, weight, smooth_rand, false, loss,
                         d_table, grad_scale, workspace, workspace_bytes, stream_);
}

extern "C" int xrd_coslam_smoothness_dev(const XrdHashGrid* grid, int sample_points, double voxel_size,
                                         double margin, float weight, const float* smooth_rand,
                                         float* loss, float* d_table, float grad_scale,
                                         void* workspace, size_t workspace_bytes, void* stream_) {
  return smoothness_impl(grid, sample_points, voxel_size, margin, weight, smooth_rand, true, loss,
                         d_table, grad_scale, workspace, workspace_bytes, stream_);
}

extern "C" int xrd_coslam_query(const XrdHashGrid* grid, const XrdCoslamMlp* mlp, const float* pts,
                                int n_points, int normalised, float* raw, float* geo, float* feat,
                                void* stream_) {
  if (!grid || !grid->table || !pts) return XRD_E_NULL;
  if ((raw || geo) && (!mlp || !mlp->w_sdf0 || !mlp->w_sdf1)) return XRD_E_NULL;
  if (raw && (!mlp->w_col0 || !mlp->w_col1)) return XRD_E_NULL;
  if (n_points <= 0) return XRD_OK;
  QueryParams Q;
  int st = fill_grid(Q.g, grid);
  if (st != XRD_OK) return st;
  Q.table = grid->table;
  Q.w_sdf0 = mlp ? mlp->w_sdf0 : nullptr; Q.w_sdf1 = mlp ? mlp->w_sdf1 : nullptr;
  Q.w_col0 = mlp ? mlp->w_col0 : nullptr; Q.w_col1 = mlp ? mlp->w_col1 : nullptr;
  Q.pts = pts; Q.P = n_points; Q.normalised = normalised; Q.raw = raw; Q.geo = geo; Q.feat = feat;
  const size_t smem = 128 * REC * sizeof(float);
  XRD_CUDA_TRY(cudaFuncSetAttribute(k_query, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_query<<<(n_points + 127) / 128, 128, smem, (cudaStream_t)stream_>>>(Q);
  XRD_LAUNCH_CHECK();
  return XRD_OK;
}

extern "C" int xrd_hashgrid_encode(const XrdHashGrid* grid, const float* x, int n_points,
                                   float* feat, uint32_t* idx, void* stream_) {
  if (!grid || !grid->table || !x) return XRD_E_NULL;
  if (n_points <= 0) return XRD_OK;
  GridDev g;
  int st = fill_grid(g, grid);
  if (st != XRD_OK) return st;
  const long long total = (long long)n_points * kL;
  k_encode<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(g, grid->table, x, n_points, feat, idx);
  XRD_LAUNCH_CHECK();
  return XRD_OK;
}
