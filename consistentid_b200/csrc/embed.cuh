// Kernels of the embedding producers (ProjPlusModel functions.py:494-528, AttentionMLP functions.py:530-592,
// PerceiverAttention functions.py:407-455).  These run once per identity on a handful of latent rows: latency-bound, no tensor cores.
#pragma once
#include "elementwise.cuh"

namespace cid {

// LayerNorm with a grouped row mapping: logical row r = g * rows_per_group + i reads x row (g * x_group_rows + x_row0 + i) and writes
// y row (g * y_group_rows + y_row0 + i).  Lets LN1(image features) and LN2(latents) land directly in the concatenated
// [B, n_img + n_lat, dim] key/value input of PerceiverAttention (functions.py:437-444) without a cat.  Any C % 8 == 0; one warp per row.
__global__ void __launch_bounds__(256)
layernorm_rows_kernel(const uint16_t* __restrict__ x, long long ldx, long long x_group_rows, long long x_row0,
                      const uint16_t* __restrict__ gamma, const uint16_t* __restrict__ beta, uint16_t* __restrict__ y, long long ldy,
                      long long y_group_rows, long long y_row0, long long rows, long long rows_per_group, int C, float eps, int bf) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const long long g = row / rows_per_group, i = row % rows_per_group;
  const uint16_t* xr = x + (g * x_group_rows + x_row0 + i) * ldx;
  uint16_t* yr = y + (g * y_group_rows + y_row0 + i) * ldy;
  const int V = C / 8;
  float s = 0.f;
  for (int v = lane; v < V; v += 32) {
    float f[8]; unpack8(*reinterpret_cast<const uint4*>(xr + v * 8), f, bf);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += f[k];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
  for (int v = lane; v < V; v += 32) {
    float f[8]; unpack8(*reinterpret_cast<const uint4*>(xr + v * 8), f, bf);
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float d = f[k] - mean; q += d * d; }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + eps);
  for (int v = lane; v < V; v += 32) {
    float f[8], ga[8], be[8], o8[8];
    unpack8(*reinterpret_cast<const uint4*>(xr + v * 8), f, bf);
    unpack8(*reinterpret_cast<const uint4*>(gamma + v * 8), ga, bf);
    unpack8(*reinterpret_cast<const uint4*>(beta + v * 8), be, bf);
#pragma unroll
    for (int k = 0; k < 8; ++k) o8[k] = (f[k] - mean) * rstd * ga[k] + be[k];
    *reinterpret_cast<uint4*>(yr + v * 8) = pack8(o8, bf);
  }
}

// PerceiverAttention core (functions.py:446-453), dim_head = 64.  One CTA per (sample b, head h, latent query l):
//   w = softmax_fp32( r16( r16(q*s) . r16(k*s) ) ),  o = r16( r16(w) . v ),  s = 64^-1/4,  r16 = round to the 16-bit storage type
// (the roundings are those of the reference's 16-bit evaluation).  q [B*L, ldq], kv [B*n_kv, ldkv] with K in columns [0, inner) and V in
// [inner, 2*inner), out [B*L, ldo]; head h owns columns [64h, 64h+64).
__global__ void __launch_bounds__(128)
perceiver_attn_kernel(const uint16_t* __restrict__ q, long long ldq, const uint16_t* __restrict__ kv, long long ldkv,
                      uint16_t* __restrict__ out, long long ldo, int L, int n_kv, int heads, int bf) {
  extern __shared__ float sc[];                     // [n_kv] scores -> probabilities
  __shared__ float qs[64];
  __shared__ float red[4];
  __shared__ float part[2][64];
  const int l = blockIdx.x % L, h = (blockIdx.x / L) % heads, b = blockIdx.x / (L * heads);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int inner = heads * 64;
  const float s = 0.35355339059327379f;            // 64^-0.25
  auto r16 = [&](float v) { return bf ? __bfloat162float(__float2bfloat16_rn(v)) : __half2float(__float2half_rn(v)); };
  if (tid < 64) qs[tid] = r16(load16(q, (size_t)(b * L + l) * ldq + h * 64 + tid, bf) * s);
  __syncthreads();
  const uint16_t* kbase = kv + (size_t)b * n_kv * ldkv + h * 64;
  float lmax = -INFINITY;
  for (int t = tid; t < n_kv; t += 128) {
    const uint16_t* kr = kbase + (size_t)t * ldkv;
    float acc = 0.f;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      float f[8]; unpack8(*reinterpret_cast<const uint4*>(kr + v * 8), f, bf);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += qs[v * 8 + k] * r16(f[k] * s);
    }
    acc = r16(acc);
    sc[t] = acc;
    lmax = fmaxf(lmax, acc);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  const float m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float lsum = 0.f;
  for (int t = tid; t < n_kv; t += 128) { const float e = __expf(sc[t] - m); sc[t] = e; lsum += e; }
#pragma unroll
  for (int o = 16; o; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if (lane == 0) red[warp] = lsum;
  __syncthreads();
  const float inv = 1.f / (red[0] + red[1] + red[2] + red[3]);
  const int d = tid & 63, half = tid >> 6;
  const uint16_t* vbase = kbase + inner;
  float acc = 0.f;
  for (int t = half; t < n_kv; t += 2) acc += r16(sc[t] * inv) * load16(vbase, (size_t)t * ldkv + d, bf);
  part[half][d] = acc;
  __syncthreads();
  if (tid < 64) store16(out, (size_t)(b * L + l) * ldo + h * 64 + tid, part[0][tid] + part[1][tid], bf);
}

// In-place row softmax of a 16-bit matrix (fp32 math, one CTA per row, any cols % 8 == 0): the probabilities of the single-head,
// d = 512 attention of the VAE mid block, whose scores are produced / consumed by two cid_gemm launches (diffusers Attention with
// upcast_softmax).
__global__ void __launch_bounds__(256)
softmax_rows_kernel(uint16_t* __restrict__ x, long long ld, int cols, int bf) {
  __shared__ float red[8];
  uint16_t* row = x + (size_t)blockIdx.x * ld;
  const int V = cols / 8, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float m = -INFINITY;
  for (int v = threadIdx.x; v < V; v += 256) {
    float f[8]; unpack8(*reinterpret_cast<const uint4*>(row + v * 8), f, bf);
#pragma unroll
    for (int k = 0; k < 8; ++k) m = fmaxf(m, f[k]);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int v = threadIdx.x; v < V; v += 256) {
    float f[8]; unpack8(*reinterpret_cast<const uint4*>(row + v * 8), f, bf);
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += __expf(f[k] - m);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += red[w];
  const float inv = 1.f / sum;
  for (int v = threadIdx.x; v < V; v += 256) {
    float f[8]; unpack8(*reinterpret_cast<const uint4*>(row + v * 8), f, bf);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = __expf(f[k] - m) * inv;
    *reinterpret_cast<uint4*>(row + v * 8) = pack8(f, bf);
  }
}

}  // namespace cid
