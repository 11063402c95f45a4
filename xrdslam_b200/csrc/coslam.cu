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
//              point).  Per-point record in shared memory:
//                [feat32 | blob48 | geo15 | sdf | h1_32 | c1_32 | raw4]
//              forward -> per-ray composite + loss gradient (one warp per ray) ->
//              layer-by-layer backward that overwrites activations with their
//              gradients in place, a register-blocked tile GEMM per layer for the
//              weight gradients, red.global.add.v2.f32 scatter for the table.
//   k_finalize loss accumulators -> the 4 weighted loss terms.
#include <math.h>

#include "common.cuh"

namespace xrd {
namespace coslam {

constexpr int kL = XRD_MAX_LEVELS;
constexpr int kBins = 16;
// record layout (floats); stride == 4 (mod 32) keeps 128-bit LDS conflict-free
constexpr int R_FEAT = 0, R_BLOB = 32, R_GEO = 80, R_SDF = 95, R_H1 = 96, R_C1 = 128,
              R_RAW = 160, REC = 164;
// transposed weights in shared memory (floats)
constexpr int W0T = 0;                 // [80][32]  w_sdf0^T
constexpr int W1T = W0T + 80 * 32;     // [32][16]  w_sdf1^T, cols = (geo0..14, sdf)
constexpr int WC0T = W1T + 32 * 16;    // [64][32]  w_col0^T, row 63 = 0
constexpr int WC1T = WC0T + 64 * 32;   // [32][4]   w_col1^T, col 3 = 0
constexpr int W_TOTAL = WC1T + 32 * 4; // 5248 floats

struct GridDev {
  float scale[kL];
  uint32_t res[kL], size[kL], offset[kL], hashed[kL];
  int n_levels;
  double bmin[3], binv[3];  // 1/(max-min) is NOT used for the forward (division)
  double bmax[3];
};

struct Params {
  // rays
  int R, S;
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
};

// ------------------------------------------------------------------ sample ---
struct SampleParams {
  int R, S, n_a, n_b, perturb, has_depth;
  const float *target_d, *lin_uniform, *lin_range, *lin_nodepth, *lin_full, *noise;
  float trunc, depth_trunc;
  uint64_t seed;
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
          philox4(p.seed, (uint64_t)r * S + k, q);
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
__device__ __forceinline__ uint32_t grid_index(const GridDev& g, int l, uint32_t x,
                                               uint32_t y, uint32_t z) {
  const uint32_t res = g.res[l], size = g.size[l];
  uint32_t index;
  if (g.hashed[l]) {
    index = x ^ (y * 2654435761u) ^ (z * 805459861u);
  } else {
    // dense stride walk (stride <= size is true for all three dims on a dense level)
    index = x + y * res + z * res * res;
  }
  return index % size + g.offset[l];
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

__device__ __forceinline__ float quartic_cdf(float u_in, float s) {
  float u = u_in * s, u2 = u * u, u4 = u2 * u2;
  float v = 0.9375f * u * (1.f - (2.f / 3.f) * u2 + 0.2f * u4) + 0.5f;
  return fminf(fmaxf(v, 0.f), 1.f);
}
__device__ __forceinline__ float cdf3(float d, float s) {
  return quartic_cdf(d, s) + quartic_cdf(d - 1.f, s) + quartic_cdf(d + 1.f, s);
}
// derivative of cdf3 w.r.t. its argument (quartic kernel, zero where clamped)
__device__ __forceinline__ float quartic_pdf(float u_in, float s) {
  float u = u_in * s;
  float t = 1.f - u * u;
  return (fabsf(u) < 1.f) ? 0.9375f * s * t * t : 0.f;
}
__device__ __forceinline__ float pdf3(float d, float s) {
  return quartic_pdf(d, s) + quartic_pdf(d - 1.f, s) + quartic_pdf(d + 1.f, s);
}

__device__ __forceinline__ void encode_point(const Params& P, const float xn[3],
                                             float* __restrict__ rec) {
  const float2* __restrict__ tab = reinterpret_cast<const float2*>(P.table);
#pragma unroll 4
  for (int l = 0; l < kL; ++l) {
    if (l >= P.g.n_levels) break;
    float w[3];
    uint32_t c[3];
    pos_fract(xn[0], P.g.scale[l], w[0], c[0]);
    pos_fract(xn[1], P.g.scale[l], w[1], c[1]);
    pos_fract(xn[2], P.g.scale[l], w[2], c[2]);
    float2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      uint32_t idx = grid_index(P.g, l, c[0] + (k & 1), c[1] + ((k >> 1) & 1),
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
    rec[R_FEAT + 2 * l] = f0;
    rec[R_FEAT + 2 * l + 1] = f1;
  }
  // OneBlob: out[d*16+b] = cdf3(e_{b+1}-x) - cdf3(e_b-x)
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float prev = cdf3(0.f - xn[d], (float)kBins);
#pragma unroll
    for (int b = 0; b < kBins; ++b) {
      float e = (float)(b + 1) * (1.f / kBins);
      float cur = cdf3(e - xn[d], (float)kBins);
      rec[R_BLOB + d * kBins + b] = cur - prev;
      prev = cur;
    }
  }
}

// d loss / d x (normalised coords) through the OneBlob, given dblob[48] via a functor
template <typename F>
__device__ __forceinline__ void blob_backward(const float xn[3], F dblob, float dx[3]) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float prev = pdf3(0.f - xn[d], (float)kBins);
    float acc = 0.f;
#pragma unroll
    for (int b = 0; b < kBins; ++b) {
      float e = (float)(b + 1) * (1.f / kBins);
      float cur = pdf3(e - xn[d], (float)kBins);
      // d out_b / dx = -(pdf(e_{b+1}-x) - pdf(e_b - x))
      acc = fmaf(dblob(d * kBins + b), prev - cur, acc);
      prev = cur;
    }
    dx[d] += acc;
  }
}

// scatter dfeat into the table gradient; optionally d loss / d x through the trilerp
__device__ __forceinline__ void hash_backward(const Params& P, const float xn[3],
                                              const float* dfeat, bool need_dx,
                                              float dx[3]) {
  const float2* __restrict__ tab = reinterpret_cast<const float2*>(P.table);
#pragma unroll 2
  for (int l = 0; l < kL; ++l) {
    if (l >= P.g.n_levels) break;
    const float g0 = dfeat[2 * l], g1 = dfeat[2 * l + 1];
    float w[3];
    uint32_t c[3];
    const float sc = P.g.scale[l];
    pos_fract(xn[0], sc, w[0], c[0]);
    pos_fract(xn[1], sc, w[1], c[1]);
    pos_fract(xn[2], sc, w[2], c[2]);
    uint32_t idx[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      idx[k] = grid_index(P.g, l, c[0] + (k & 1), c[1] + ((k >> 1) & 1),
                          c[2] + ((k >> 2) & 1));
    if (need_dx) {
      float t[8];  // <table entry, dfeat>
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float2 v = __ldg(&tab[idx[k]]);
        t[k] = v.x * g0 + v.y * g1;
      }
      const float w0 = w[0], w1 = w[1], w2 = w[2];
      // d/dw0: sum over (y,z) corners of weight_yz * (t[x=1] - t[x=0])
      float d0 = (1 - w1) * (1 - w2) * (t[1] - t[0]) + w1 * (1 - w2) * (t[3] - t[2]) +
                 (1 - w1) * w2 * (t[5] - t[4]) + w1 * w2 * (t[7] - t[6]);
      float d1 = (1 - w0) * (1 - w2) * (t[2] - t[0]) + w0 * (1 - w2) * (t[3] - t[1]) +
                 (1 - w0) * w2 * (t[6] - t[4]) + w0 * w2 * (t[7] - t[5]);
      float d2 = (1 - w0) * (1 - w1) * (t[4] - t[0]) + w0 * (1 - w1) * (t[5] - t[1]) +
                 (1 - w0) * w1 * (t[6] - t[2]) + w0 * w1 * (t[7] - t[3]);
      dx[0] = fmaf(sc, d0, dx[0]);
      dx[1] = fmaf(sc, d1, dx[1]);
      dx[2] = fmaf(sc, d2, dx[2]);
    }
    if (g0 != 0.f || g1 != 0.f) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float wk = ((k & 1) ? w[0] : 1.f - w[0]) * ((k & 2) ? w[1] : 1.f - w[1]) *
                   ((k & 4) ? w[2] : 1.f - w[2]);
        red_add_v2(P.d_table + 2 * (size_t)idx[k], wk * g0, wk * g1);
      }
    }
  }
}

// ------------------------------------------------------------------- MLP ---
// h1 = relu(W0 x), h = W1 h1, c1 = relu(WC0 [blob, geo]), rgb = WC1 c1
__device__ __forceinline__ void mlp_forward(const float* __restrict__ sw,
                                            float* __restrict__ rec) {
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
#pragma unroll 2
  for (int i4 = 0; i4 < 80; i4 += 4) {
    const float4 xv = *reinterpret_cast<const float4*>(rec + R_FEAT + i4);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const float4* wr = reinterpret_cast<const float4*>(sw + W0T + (i4 + ii) * 32);
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 w = wr[j4];
        acc[4 * j4 + 0] = fmaf(xs[ii], w.x, acc[4 * j4 + 0]);
        acc[4 * j4 + 1] = fmaf(xs[ii], w.y, acc[4 * j4 + 1]);
        acc[4 * j4 + 2] = fmaf(xs[ii], w.z, acc[4 * j4 + 2]);
        acc[4 * j4 + 3] = fmaf(xs[ii], w.w, acc[4 * j4 + 3]);
      }
    }
  }
  float h[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) h[j] = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float a = fmaxf(acc[i], 0.f);
    rec[R_H1 + i] = a;
    const float4* wr = reinterpret_cast<const float4*>(sw + W1T + i * 16);
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      const float4 w = wr[j4];
      h[4 * j4 + 0] = fmaf(a, w.x, h[4 * j4 + 0]);
      h[4 * j4 + 1] = fmaf(a, w.y, h[4 * j4 + 1]);
      h[4 * j4 + 2] = fmaf(a, w.z, h[4 * j4 + 2]);
      h[4 * j4 + 3] = fmaf(a, w.w, h[4 * j4 + 3]);
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) rec[R_GEO + j] = h[j];  // geo0..14, sdf
  // colour net: inputs rec[32..96) = blob48, geo15, (sdf slot: weight row 63 == 0)
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
#pragma unroll 2
  for (int i4 = 0; i4 < 64; i4 += 4) {
    const float4 xv = *reinterpret_cast<const float4*>(rec + R_BLOB + i4);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const float4* wr = reinterpret_cast<const float4*>(sw + WC0T + (i4 + ii) * 32);
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 w = wr[j4];
        acc[4 * j4 + 0] = fmaf(xs[ii], w.x, acc[4 * j4 + 0]);
        acc[4 * j4 + 1] = fmaf(xs[ii], w.y, acc[4 * j4 + 1]);
        acc[4 * j4 + 2] = fmaf(xs[ii], w.z, acc[4 * j4 + 2]);
        acc[4 * j4 + 3] = fmaf(xs[ii], w.w, acc[4 * j4 + 3]);
      }
    }
  }
  float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float a = fmaxf(acc[i], 0.f);
    rec[R_C1 + i] = a;
    const float4 w = *reinterpret_cast<const float4*>(sw + WC1T + i * 4);
    r0 = fmaf(a, w.x, r0);
    r1 = fmaf(a, w.y, r1);
    r2 = fmaf(a, w.z, r2);
  }
  rec[R_RAW + 0] = r0;
  rec[R_RAW + 1] = r1;
  rec[R_RAW + 2] = r2;
  rec[R_RAW + 3] = h[15];
}

// dWT[(4*ib+ii)*ld + 4*jb+jj] += sum_p A[p][4*ib+ii] * B[p][4*jb+jj]
__device__ __forceinline__ void tile_gemm(const float* __restrict__ recs, int n_pts,
                                          int aoff, int nIb, int boff, int nJb,
                                          float* __restrict__ dwt, int ld) {
  for (int b = threadIdx.x; b < nIb * nJb; b += blockDim.x) {
    const int jb = b % nJb, ib = b / nJb;
    float acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.f;
    const float* pa = recs + aoff + 4 * ib;
    const float* pb = recs + boff + 4 * jb;
#pragma unroll 4
    for (int p = 0; p < n_pts; ++p) {
      const float4 a = *reinterpret_cast<const float4*>(pa + p * REC);
      const float4 bb = *reinterpret_cast<const float4*>(pb + p * REC);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[ii * 4 + jj] = fmaf(av[ii], bv[jj], acc[ii * 4 + jj]);
    }
    float* o = dwt + (4 * ib) * ld + 4 * jb;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) o[ii * ld + jj] += acc[ii * 4 + jj];
  }
}

// ----------------------------------------------------------------- fused ---
template <bool BWD>
__global__ void __launch_bounds__(256) k_fused(const Params P) {
  extern __shared__ __align__(16) float smem[];
  float* sw = smem;                              // W_TOTAL
  float* sdw = sw + W_TOTAL;                     // W_TOTAL (BWD only)
  float* recs = sdw + (BWD ? W_TOTAL : 0);       // blockDim.x * REC
  float* zbuf = recs + blockDim.x * REC;         // NR*S  z values
  float* rayacc = zbuf + P.NR * P.S;             // NR*8  d_rays accumulators
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwarps = blockDim.x >> 5;
  const int S = P.S;

  // stage transposed weights
  for (int q = tid; q < 80 * 32; q += blockDim.x) {
    int i = q / 32, j = q % 32;
    sw[W0T + q] = P.w_sdf0[j * 80 + i];
  }
  for (int q = tid; q < 32 * 16; q += blockDim.x) {
    int i = q / 16, jp = q % 16;           // jp: geo0..14 -> torch out 1..15, jp 15 -> out 0
    int jt = (jp + 1) & 15;
    sw[W1T + q] = P.w_sdf1[jt * 32 + i];
  }
  for (int q = tid; q < 64 * 32; q += blockDim.x) {
    int i = q / 32, j = q % 32;
    sw[WC0T + q] = (i < 63) ? P.w_col0[j * 63 + i] : 0.f;
  }
  for (int q = tid; q < 32 * 4; q += blockDim.x) {
    int i = q / 4, k = q % 4;
    sw[WC1T + q] = (k < 3) ? P.w_col1[k * 32 + i] : 0.f;
  }
  if (BWD)
    for (int q = tid; q < W_TOTAL; q += blockDim.x) sdw[q] = 0.f;

  // batch-global normalisers (utils.py:126-130), float32 like torch's int/int division
  float fs_w = 0.f, sdf_w = 0.f, inv_nvalid = 0.f;
  if (BWD) {
    const float n_fs = (float)P.counts[0], n_sdf = (float)P.counts[1];
    const float n = (float)(P.counts[0] + P.counts[1]);
    fs_w = 1.0f - n_fs / n;
    sdf_w = 1.0f - n_sdf / n;
    inv_nvalid = 1.0f / (float)P.counts[2];
  }
  double l_rgb = 0.0, l_depth = 0.0, l_sdf = 0.0, l_fs = 0.0;  // lane-0-of-warp partials
  __syncthreads();

  float* rec = recs + tid * REC;
  for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
    const int r0 = tile * P.NR;
    const int nr = min(P.NR, P.R - r0);
    const int npts = nr * S;
    const bool active = tid < npts;
    const int rl = active ? tid / S : 0;
    const int k = active ? tid - rl * S : 0;
    const int r = r0 + rl;
    float xn[3] = {0.f, 0.f, 0.f};
    float zv = 0.f;
    if (tid < nr * 8) rayacc[tid] = 0.f;
    // ---------------- phase 1: per-point forward --------------------------
    if (active) {
      zv = P.z_vals[(size_t)r * S + k];
      zbuf[tid] = zv;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float pt = __fadd_rn(P.rays_o[r * 3 + d], __fmul_rn(P.rays_d[r * 3 + d], zv));
        xn[d] = normalise(pt, P.g.bmin[d], P.g.bmax[d]);
      }
      encode_point(P, xn, rec);
      mlp_forward(sw, rec);
      if (P.raw) {
        float4 rw = *reinterpret_cast<const float4*>(rec + R_RAW);
        *reinterpret_cast<float4*>(P.raw + ((size_t)r * S + k) * 4) = rw;
      }
    }
    __syncthreads();
    // ---------------- phase 2: per-ray composite, loss, d loss / d raw ----
    for (int q = warp; q < nr; q += nwarps) {
      const int rr = r0 + q;
      const float* zr = zbuf + q * S;
      float* rq = recs + (size_t)(q * S) * REC;
      const float tr = P.trunc;
      // first zero crossing (argmax of the 0/1 mask -> first true, else 0)
      int first = 0x7fffffff;
      for (int kk = lane; kk < S - 1; kk += 32) {
        float s0 = rq[kk * REC + R_RAW + 3], s1 = rq[(kk + 1) * REC + R_RAW + 3];
        if (s1 * s0 < 0.f) first = min(first, kk);
      }
      first = warp_min_i(first);
      if (first == 0x7fffffff) first = 0;
      const float zlim = zr[first] + P.trunc;
      float usum = 0.f;
      for (int kk = lane; kk < S; kk += 32) {
        float s = rq[kk * REC + R_RAW + 3];
        float sg = sigmoidf_acc(s / tr);
        float a = sg * sigmoidf_acc((-s) / tr);
        float u = (zr[kk] < zlim) ? a : 0.f;
        usum += u;
      }
      usum = warp_sum(usum);
      const float W = usum + 1e-8f;
      float o_r = 0.f, o_g = 0.f, o_b = 0.f, o_d = 0.f, o_acc = 0.f;
      for (int kk = lane; kk < S; kk += 32) {
        const float* rp = rq + kk * REC + R_RAW;
        float s = rp[3];
        float sg = sigmoidf_acc(s / tr);
        float a = sg * sigmoidf_acc((-s) / tr);
        float w = ((zr[kk] < zlim) ? a : 0.f) / W;
        o_r = fmaf(w, sigmoidf_acc(rp[0]), o_r);
        o_g = fmaf(w, sigmoidf_acc(rp[1]), o_g);
        o_b = fmaf(w, sigmoidf_acc(rp[2]), o_b);
        o_d = fmaf(w, zr[kk], o_d);
        o_acc += w;
      }
      o_r = warp_sum(o_r); o_g = warp_sum(o_g); o_b = warp_sum(o_b);
      o_d = warp_sum(o_d); o_acc = warp_sum(o_acc);
      if (P.depth_var) {
        float v = 0.f;
        for (int kk = lane; kk < S; kk += 32) {
          float s = rq[kk * REC + R_RAW + 3];
          float a = sigmoidf_acc(s / tr) * sigmoidf_acc((-s) / tr);
          float w = ((zr[kk] < zlim) ? a : 0.f) / W;
          float dz = zr[kk] - o_d;
          v = fmaf(w, dz * dz, v);
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
        const float RS = (float)P.R * (float)S;
        // d total / d rgb, d depth  (mse means; Q1: rgb weight == 1 for every ray)
        const float c_rgb = P.ls[0] * P.w_rgb * 2.0f / (3.0f * (float)P.R);
        const float g_r = c_rgb * (o_r - tr_), g_g = c_rgb * (o_g - tg_), g_b = c_rgb * (o_b - tb_);
        const float g_d = valid ? P.ls[1] * P.w_depth * 2.0f * (o_d - td) * inv_nvalid : 0.f;
        if (lane == 0) {
          l_rgb += (double)((o_r - tr_) * (o_r - tr_) + (o_g - tg_) * (o_g - tg_) +
                            (o_b - tb_) * (o_b - tb_));
          if (valid) l_depth += (double)((o_d - td) * (o_d - td));
        }
        // sum_j q_j w_j
        float qw = 0.f;
        for (int kk = lane; kk < S; kk += 32) {
          const float* rp = rq + kk * REC + R_RAW;
          float s = rp[3];
          float a = sigmoidf_acc(s / tr) * sigmoidf_acc((-s) / tr);
          float w = ((zr[kk] < zlim) ? a : 0.f) / W;
          float q_ = g_r * sigmoidf_acc(rp[0]) + g_g * sigmoidf_acc(rp[1]) +
                     g_b * sigmoidf_acc(rp[2]) + g_d * zr[kk];
          qw = fmaf(q_, w, qw);
        }
        qw = warp_sum(qw);
        float a_fs = 0.f, a_sdf = 0.f;
        const float c_fs = P.ls[3] * P.w_fs * fs_w * 2.0f / RS;
        const float c_sdf = P.ls[2] * P.w_sdf * sdf_w * 2.0f / RS;
        for (int kk = lane; kk < S; kk += 32) {
          float* rp = rq + kk * REC + R_RAW;
          const float z = zr[kk];
          const float s = rp[3];
          const float sg = sigmoidf_acc(s / tr);
          const float a = sg * sigmoidf_acc((-s) / tr);
          const bool m = z < zlim;
          const float w = (m ? a : 0.f) / W;
          const float c0 = sigmoidf_acc(rp[0]), c1 = sigmoidf_acc(rp[1]), c2 = sigmoidf_acc(rp[2]);
          const float q_ = g_r * c0 + g_g * c1 + g_b * c2 + g_d * z;
          float ds = m ? (q_ - qw) / W * a * (1.f - 2.f * sg) / tr : 0.f;
          // free-space / sdf terms (utils.py:154-186)
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
    if (!BWD) { __syncthreads(); continue; }
    __syncthreads();
    // ---------------- phase 3: backward ----------------------------------
    // dW_col1 += c1^T draw
    tile_gemm(recs, npts, R_C1, 8, R_RAW, 1, sdw + WC1T, 4);
    __syncthreads();
    float draw3 = 0.f;
    if (active) {
      const float4 dr = *reinterpret_cast<const float4*>(rec + R_RAW);
      draw3 = dr.w;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(sw + WC1T + j * 4);
        const float c1 = rec[R_C1 + j];
        rec[R_C1 + j] = (c1 > 0.f) ? (dr.x * w.x + dr.y * w.y + dr.z * w.z) : 0.f;
      }
    }
    __syncthreads();
    // dW_col0 += [blob,geo]^T dc1pre
    tile_gemm(recs, npts, R_BLOB, 16, R_C1, 8, sdw + WC0T, 32);
    __syncthreads();
    float dxn[3] = {0.f, 0.f, 0.f};
    if (active) {
      float g[32];
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 v = *reinterpret_cast<const float4*>(rec + R_C1 + 4 * j4);
        g[4 * j4] = v.x; g[4 * j4 + 1] = v.y; g[4 * j4 + 2] = v.z; g[4 * j4 + 3] = v.w;
      }
      auto dot32 = [&](const float* wrow) {
        float a0 = 0.f, a1 = 0.f;
        const float4* wr = reinterpret_cast<const float4*>(wrow);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 w = wr[j4];
          a0 = fmaf(g[4 * j4], w.x, a0);
          a1 = fmaf(g[4 * j4 + 1], w.y, a1);
          a0 = fmaf(g[4 * j4 + 2], w.z, a0);
          a1 = fmaf(g[4 * j4 + 3], w.w, a1);
        }
        return a0 + a1;
      };
      // colour-path OneBlob gradient -> dx immediately (the blob slots stay live for dW_sdf0)
      blob_backward(xn, [&](int i) { return dot32(sw + WC0T + i * 32); }, dxn);
      float dgeo[15];
#pragma unroll
      for (int i = 0; i < 15; ++i) dgeo[i] = dot32(sw + WC0T + (48 + i) * 32);
      // (GEMM above already consumed geo) -> overwrite with dH = [dgeo, dsdf]
#pragma unroll
      for (int i = 0; i < 15; ++i) rec[R_GEO + i] = dgeo[i];
      rec[R_SDF] = draw3;
    }
    __syncthreads();
    // dW_sdf1 += h1^T dH
    tile_gemm(recs, npts, R_H1, 8, R_GEO, 4, sdw + W1T, 16);
    __syncthreads();
    if (active) {
      float dh[16];
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 v = *reinterpret_cast<const float4*>(rec + R_GEO + 4 * j4);
        dh[4 * j4] = v.x; dh[4 * j4 + 1] = v.y; dh[4 * j4 + 2] = v.z; dh[4 * j4 + 3] = v.w;
      }
#pragma unroll 4
      for (int i = 0; i < 32; ++i) {
        const float4* wr = reinterpret_cast<const float4*>(sw + W1T + i * 16);
        float a = 0.f;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 w = wr[j4];
          a = fmaf(dh[4 * j4], w.x, a);
          a = fmaf(dh[4 * j4 + 1], w.y, a);
          a = fmaf(dh[4 * j4 + 2], w.z, a);
          a = fmaf(dh[4 * j4 + 3], w.w, a);
        }
        const float h1 = rec[R_H1 + i];
        rec[R_H1 + i] = (h1 > 0.f) ? a : 0.f;
      }
    }
    __syncthreads();
    // dW_sdf0 += x^T dh1pre
    tile_gemm(recs, npts, R_FEAT, 20, R_H1, 8, sdw + W0T, 32);
    if (active) {
      float g[32];
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 v = *reinterpret_cast<const float4*>(rec + R_H1 + 4 * j4);
        g[4 * j4] = v.x; g[4 * j4 + 1] = v.y; g[4 * j4 + 2] = v.z; g[4 * j4 + 3] = v.w;
      }
      auto dot32 = [&](const float* wrow) {
        float a0 = 0.f, a1 = 0.f;
        const float4* wr = reinterpret_cast<const float4*>(wrow);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 w = wr[j4];
          a0 = fmaf(g[4 * j4], w.x, a0);
          a1 = fmaf(g[4 * j4 + 1], w.y, a1);
          a0 = fmaf(g[4 * j4 + 2], w.z, a0);
          a1 = fmaf(g[4 * j4 + 3], w.w, a1);
        }
        return a0 + a1;
      };
      blob_backward(xn, [&](int i) { return dot32(sw + W0T + (32 + i) * 32); }, dxn);
      float dfeat[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) dfeat[i] = dot32(sw + W0T + i * 32);
      const bool need_dx = (P.d_rays_o != nullptr) || (P.d_rays_d != nullptr);
      hash_backward(P, xn, dfeat, need_dx, dxn);
      if (need_dx) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          float dp = (float)((double)dxn[d] / (P.g.bmax[d] - P.g.bmin[d]));
          atomicAdd(&rayacc[rl * 8 + d], dp);
          atomicAdd(&rayacc[rl * 8 + 3 + d], dp * zv);
        }
      }
    }
    __syncthreads();
    if (tid < nr * 6) {
      const int q = tid / 6, c = tid % 6;
      const float v = rayacc[q * 8 + c];
      if (c < 3) { if (P.d_rays_o) P.d_rays_o[(r0 + q) * 3 + c] = v; }
      else       { if (P.d_rays_d) P.d_rays_d[(r0 + q) * 3 + c - 3] = v; }
    }
    __syncthreads();
  }

  if (BWD) {
    // flush loss partials and weight gradients
    if (lane == 0) {
      if (l_rgb != 0.0) atomicAdd(&P.loss_acc[0], l_rgb);
      if (l_depth != 0.0) atomicAdd(&P.loss_acc[1], l_depth);
      if (l_sdf != 0.0) atomicAdd(&P.loss_acc[2], l_sdf);
      if (l_fs != 0.0) atomicAdd(&P.loss_acc[3], l_fs);
    }
    __syncthreads();
    for (int q = tid; q < 80 * 32; q += blockDim.x) {
      int i = q / 32, j = q % 32;
      red_add(P.d_w_sdf0 + j * 80 + i, sdw[W0T + q]);
    }
    for (int q = tid; q < 32 * 16; q += blockDim.x) {
      int i = q / 16, jp = q % 16, jt = (jp + 1) & 15;
      red_add(P.d_w_sdf1 + jt * 32 + i, sdw[W1T + q]);
    }
    for (int q = tid; q < 63 * 32; q += blockDim.x) {
      int i = q / 32, j = q % 32;
      red_add(P.d_w_col0 + j * 63 + i, sdw[WC0T + q]);
    }
    for (int q = tid; q < 32 * 3; q += blockDim.x) {
      int i = q / 3, k = q % 3;
      red_add(P.d_w_col1 + k * 32 + i, sdw[WC1T + i * 4 + k]);
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
    double pt = ((double)(float)c[d] + p.rnd[d]) * p.voxel + p.g.bmin[d] + p.off[d];
    xn[d] = (float)((pt - p.g.bmin[d]) / (p.g.bmax[d] - p.g.bmin[d]));
  }
}

__global__ void __launch_bounds__(256) k_smooth_fwd(SmoothParams p) {
  const int n = p.n, N = n * n * n;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (point, level)
  const int pt = q / kL, l = q % kL;
  if (pt >= N || l >= p.g.n_levels) return;
  const int iz = pt % n, iy = (pt / n) % n, ix = pt / (n * n);
  float xn[3];
  smooth_xn(p, ix, iy, iz, xn);
  const float2* tab = reinterpret_cast<const float2*>(p.table);
  float w[3];
  uint32_t c[3];
  for (int d = 0; d < 3; ++d) pos_fract(xn[d], p.g.scale[l], w[d], c[d]);
  float f0 = 0.f, f1 = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint32_t idx = grid_index(p.g, l, c[0] + (k & 1), c[1] + ((k >> 1) & 1), c[2] + ((k >> 2) & 1));
    float wk = ((k & 1) ? w[0] : 1.f - w[0]) * ((k & 2) ? w[1] : 1.f - w[1]) *
               ((k & 4) ? w[2] : 1.f - w[2]);
    float2 v = __ldg(&tab[idx]);
    f0 = fmaf(wk, v.x, f0);
    f1 = fmaf(wk, v.y, f1);
  }
  reinterpret_cast<float2*>(p.feat)[(size_t)pt * kL + l] = make_float2(f0, f1);
}

__global__ void __launch_bounds__(256) k_smooth_bwd(SmoothParams p) {
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
      for (int d = 0; d < 3; ++d) pos_fract(xn[d], p.g.scale[l], w[d], c[d]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        uint32_t idx = grid_index(p.g, l, c[0] + (k & 1), c[1] + ((k >> 1) & 1), c[2] + ((k >> 2) & 1));
        float wk = ((k & 1) ? w[0] : 1.f - w[0]) * ((k & 2) ? w[1] : 1.f - w[1]) *
                   ((k & 4) ? w[2] : 1.f - w[2]);
        red_add_v2(p.d_table + 2 * (size_t)idx, wk * g0, wk * g1);
      }
    }
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0 && part != 0.f) atomicAdd(p.loss_acc, (double)part);
}

__global__ void k_smooth_finalize(const double* acc, float* loss, float scale) {
  if (threadIdx.x == 0 && blockIdx.x == 0) loss[0] = (float)(acc[0] * (double)scale);
}

// hash encode only
__global__ void __launch_bounds__(256) k_encode(GridDev g, const float* table, const float* x,
                                                int n, float* feat, uint32_t* idx_out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int pt = q / kL, l = q % kL;
  if (pt >= n || l >= g.n_levels) return;
  const float2* tab = reinterpret_cast<const float2*>(table);
  float w[3];
  uint32_t c[3];
  for (int d = 0; d < 3; ++d) pos_fract(x[pt * 3 + d], g.scale[l], w[d], c[d]);
  float f0 = 0.f, f1 = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint32_t idx = grid_index(g, l, c[0] + (k & 1), c[1] + ((k >> 1) & 1), c[2] + ((k >> 2) & 1));
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

static int fill_grid(GridDev& g, const XrdHashGrid* h) {
  if (h->n_levels < 1 || h->n_levels > kL) return XRD_E_SHAPE;
  g.n_levels = h->n_levels;
  for (int l = 0; l < kL; ++l) {
    g.scale[l] = h->scale[l]; g.res[l] = h->resolution[l]; g.size[l] = h->size[l] ? h->size[l] : 1;
    g.offset[l] = h->offset[l]; g.hashed[l] = h->hashed[l];
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

extern "C" size_t xrd_coslam_workspace_bytes(int n_rays, int n_samples) {
  // [counts int[4] | loss_acc double[4] | z_vals R*S floats]
  return 256 + align_up((size_t)n_rays * n_samples * sizeof(float), 256);
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
  if (grads && (!grads->d_table || !grads->d_w_sdf0 || !grads->d_w_sdf1 || !grads->d_w_col0 ||
                !grads->d_w_col1))
    return XRD_E_NULL;
  if (workspace_bytes < xrd_coslam_workspace_bytes(R, S)) return XRD_E_WORKSPACE;

  int* counts = reinterpret_cast<int*>(workspace);
  double* loss_acc = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 64);
  float* z_ws = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
  float* z_vals = out->z_vals ? out->z_vals : z_ws;
  XRD_CUDA_TRY(cudaMemsetAsync(workspace, 0, 256, stream));

  SampleParams sp;
  sp.R = R; sp.S = S; sp.n_a = cfg->n_sample_d; sp.n_b = cfg->n_range_d;
  sp.perturb = cfg->perturb; sp.has_depth = has_depth;
  sp.target_d = rays->target_d; sp.lin_uniform = cfg->lin_uniform; sp.lin_range = cfg->lin_range;
  sp.lin_nodepth = cfg->lin_nodepth; sp.lin_full = cfg->lin_full; sp.noise = noise;
  sp.trunc = cfg->trunc; sp.depth_trunc = cfg->depth_trunc; sp.seed = cfg->seed;
  sp.z_vals = z_vals; sp.counts = counts;
  k_sample<<<(R + 3) / 4, 128, 0, stream>>>(sp);
  XRD_LAUNCH_CHECK();

  Params P;
  int st = fill_grid(P.g, grid);
  if (st != XRD_OK) return st;
  P.R = R; P.S = S;
  P.rays_o = rays->rays_o; P.rays_d = rays->rays_d; P.target_s = rays->target_s; P.target_d = rays->target_d;
  P.z_vals = z_vals; P.table = grid->table;
  P.w_sdf0 = mlp->w_sdf0; P.w_sdf1 = mlp->w_sdf1; P.w_col0 = mlp->w_col0; P.w_col1 = mlp->w_col1;
  P.trunc = cfg->trunc; P.depth_trunc = cfg->depth_trunc;
  P.w_rgb = cfg->w_rgb; P.w_depth = cfg->w_depth; P.w_sdf = cfg->w_sdf; P.w_fs = cfg->w_fs;
  P.rgb = out->rgb; P.depth = out->depth; P.disp = out->disp; P.acc = out->acc;
  P.depth_var = out->depth_var; P.raw = out->raw;
  P.counts = counts; P.loss_acc = loss_acc;
  for (int i = 0; i < 4; ++i) P.ls[i] = grads ? grads->loss_scale[i] : 0.f;
  if (grads) {
    P.d_table = grads->d_table; P.d_w_sdf0 = grads->d_w_sdf0; P.d_w_sdf1 = grads->d_w_sdf1;
    P.d_w_col0 = grads->d_w_col0; P.d_w_col1 = grads->d_w_col1;
    P.d_rays_o = grads->d_rays_o; P.d_rays_d = grads->d_rays_d;
  } else {
    P.d_table = P.d_w_sdf0 = P.d_w_sdf1 = P.d_w_col0 = P.d_w_col1 = P.d_rays_o = P.d_rays_d = nullptr;
  }
  int NR = pick_rays_per_tile(S, cfg->rays_per_tile);
  if (NR < 1 || NR * S > 256) return XRD_E_SHAPE;
  P.NR = NR;
  P.n_tiles = (R + NR - 1) / NR;
  const int threads = (NR * S + 31) / 32 * 32;
  const size_t smem = sizeof(float) * ((size_t)W_TOTAL * (grads ? 2 : 1) + (size_t)threads * REC +
                                       (size_t)NR * S + (size_t)NR * 8);
  const int sms = num_sms();
  if (grads) {
    XRD_CUDA_TRY(cudaFuncSetAttribute(k_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    XRD_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fused<true>, threads, smem));
    if (occ < 1) return XRD_E_SHAPE;
    int gridx = P.n_tiles < sms * occ ? P.n_tiles : sms * occ;
    {
      KernelTimer kt(stream);
      k_fused<true><<<gridx, threads, smem, stream>>>(P);
    }
    XRD_LAUNCH_CHECK();
    FinalizeParams fp{loss_acc, counts, out->losses, R, S, cfg->w_rgb, cfg->w_depth, cfg->w_sdf, cfg->w_fs};
    k_finalize<<<1, 32, 0, stream>>>(fp);
    XRD_LAUNCH_CHECK();
  } else {
    XRD_CUDA_TRY(cudaFuncSetAttribute(k_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    XRD_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fused<false>, threads, smem));
    if (occ < 1) return XRD_E_SHAPE;
    int gridx = P.n_tiles < sms * occ ? P.n_tiles : sms * occ;
    k_fused<false><<<gridx, threads, smem, stream>>>(P);
    XRD_LAUNCH_CHECK();
  }
  return XRD_OK;
}

extern "C" size_t xrd_coslam_smoothness_workspace_bytes(int sample_points) {
  size_t n = (size_t)(sample_points - 1);
  return 256 + n * n * n * 32 * sizeof(float);
}

extern "C" int xrd_coslam_smoothness(const XrdHashGrid* grid, int sample_points, double voxel_size,
                                     double margin, float weight, const float* smooth_rand,
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
    p.off[d] = (double)smooth_rand[d] * offset_max + margin;
    p.rnd[d] = (double)smooth_rand[3 + d];
  }
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
