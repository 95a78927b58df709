// tcgen05 GEMM / implicit-GEMM 3x3 convolution for the ConsistentID UNet hot path (sm_100a).
//
//   C[M, N] = epilogue( A[M, K] * B[N, K]^T )        16-bit inputs (fp16 | bf16), fp32 accumulate in TMEM
//
// One CTA = one 128 x BN output tile.  Warp roles (192 threads):
//   warp 0     TMA producer: cp.async.bulk.tensor loads of A / B k-blocks (64 elements = one 128-byte swizzled row)
//              into a STAGES-deep smem ring, completion on "full" mbarriers
//   warp 1     TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit releases smem stages ("empty")
//              and finally signals the accumulator ("acc_full")
//   warps 2-5  epilogue: tcgen05.ld accumulator rows (lane quarter = warp_id % 4), fused bias / per-sample bias
//              (time embedding) / residual add / GEGLU / QKV split with transposed V, 16-byte global stores
// Two CTAs are co-resident per SM (3 stages x 36 KB each, 256 TMEM columns each) so one CTA's epilogue overlaps
// the other's main loop.
//
// A-operand addressing modes:
//   GEMM        2-D map {K, M}; optional second source along K (virtual channel concat: 1x1 shortcut on cat([h, skip]))
//   CONV3x3     4-D map {C, W, H, N} over an NHWC activation: tap (ky,kx) = the same box shifted by (kx-1, ky-1);
//               TMA zero-fills out-of-bounds rows/cols = the conv's zero padding
//   CONV3x3 s2  5-D map {C, W/2, H/2, 4, N} over a phase-split copy of the input (see phase_split kernel)
#pragma once
#include "common.cuh"

namespace cid {

enum EpiMode : int { EPI_STORE = 0, EPI_GEGLU = 1, EPI_QKV = 2 };
enum AMode : int { A_GEMM = 0, A_CONV = 1, A_CONV_S2 = 2 };

struct GemmArgs {
  int M, N;
  int kblocks_a1, kblocks_a2;  // 64-wide k-blocks per tap taken from A1 / A2
  int taps;                    // 1 or 9
  int a_mode;
  int W, H, NB;                // conv: OUTPUT geometry
  int TW, TH, TN;              // conv tile (pixels) TW*TH*TN <= 128
  int tiles_x, tiles_y;
  void* C;
  long long ldc;
  const void* bias;
  const void* residual;
  long long ldr;
  const void* rowbias;         // [M / rows_per_group, ld_rowbias]
  int rows_per_group;
  long long ld_rowbias;
  int epi;
  int is_bf16;
  void* Vt;                    // EPI_QKV: V^T [B*heads, hdim, ntok]
  int n_split, heads, hdim, ntok;
  float out_scale;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;  // + barriers + alignment slack
};

// exact (erf) GELU, erf by Abramowitz-Stegun 7.1.28: erf(t) = 1 - (1 + a1 t + ... + a6 t^6)^-16, |err| <= 3e-7
// (far below 16-bit output rounding) - ~15 instructions instead of erff's ~40, the GEGLU epilogue is ALU-bound
__device__ __forceinline__ float gelu_erf(float x) {
  const float t = fabsf(x) * 0.70710678118654752f;
  float p = fmaf(t, 0.0000430638f, 0.0002765672f);
  p = fmaf(t, p, 0.0001520143f);
  p = fmaf(t, p, 0.0092705272f);
  p = fmaf(t, p, 0.0422820123f);
  p = fmaf(t, p, 0.0705230784f);
  p = fmaf(t, p, 1.0f);
  p = p * p; p = p * p; p = p * p; p = p * p;
  float rp;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rp) : "f"(p));
  const float e = 1.0f - rp;
  return 0.5f * x * (1.0f + copysignf(e, x));
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N");
  constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  using SM = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + SM::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t acc_bar = bar_base + 8u * (2 * STAGES);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + SM::BAR_OFF + 8 * (2 * STAGES + 1));

  const int warp = warp_id();
  const int lane = lane_id();
  const int n0 = blockIdx.x * BN;
  const int mt = blockIdx.y;
  const int kb_per_tap = g.kblocks_a1 + g.kblocks_a2;
  const int num_kb = g.taps * kb_per_tap;

  // conv tile origin
  int tile_n0 = 0, tile_y0 = 0, tile_x0 = 0;
  if (g.a_mode != A_GEMM) {
    int tx = mt % g.tiles_x;
    int rest = mt / g.tiles_x;
    int ty = rest % g.tiles_y;
    int tn = rest / g.tiles_y;
    tile_x0 = tx * g.TW; tile_y0 = ty * g.TH; tile_n0 = tn * g.TN;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    if (g.kblocks_a2 > 0) tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      mbar_init(acc_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      const uint32_t a_bytes = (g.a_mode == A_GEMM) ? uint32_t(SM::A_BYTES) : uint32_t(g.TW * g.TH * g.TN * 128);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + stage * SM::STAGE_BYTES;
        const uint32_t sb = sa + SM::A_BYTES;
        mbar_expect_tx(full_bar(stage), a_bytes + uint32_t(SM::B_BYTES));
        const int tap = kb / kb_per_tap;
        const int cb = kb - tap * kb_per_tap;
        if (g.a_mode == A_GEMM) {
          if (cb < g.kblocks_a1) tma_load_2d(sa, &tmA1, full_bar(stage), cb * GEMM_BK, mt * GEMM_BM);
          else tma_load_2d(sa, &tmA2, full_bar(stage), (cb - g.kblocks_a1) * GEMM_BK, mt * GEMM_BM);
        } else if (g.a_mode == A_CONV) {
          const int ky = tap / 3, kx = tap - ky * 3;
          tma_load_4d(sa, &tmA1, full_bar(stage), cb * GEMM_BK, tile_x0 + kx - 1, tile_y0 + ky - 1, tile_n0);
        } else {
          const int ky = tap / 3, kx = tap - ky * 3;
          const int py = (ky == 1) ? 0 : 1, dy = (ky == 0) ? -1 : 0;
          const int px = (kx == 1) ? 0 : 1, dx = (kx == 0) ? -1 : 0;
          tma_load_5d(sa, &tmA1, full_bar(stage), cb * GEMM_BK, tile_x0 + dx, tile_y0 + dy, py * 2 + px, tile_n0);
        }
        tma_load_2d(sb, &tmB, full_bar(stage), kb * GEMM_BK, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t idesc = make_idesc(GEMM_BM, BN, g.is_bf16);
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_base + stage * SM::STAGE_BYTES;
        const uint32_t sb = sa + SM::A_BYTES;
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          umma_ss(tmem_acc, make_desc_sw128(sa + k * 32), make_desc_sw128(sb + k * 32), idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));                   // frees this smem stage when the MMAs above retire
        if (kb == num_kb - 1) umma_commit(acc_bar);      // accumulator complete
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ================================================================ epilogue (warps 2..5)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                  // accumulator row == TMEM lane
    long long grow; bool row_ok;
    if (g.a_mode == A_GEMM) {
      grow = (long long)mt * GEMM_BM + r;
      row_ok = grow < g.M;
    } else {
      const int per_img = g.TW * g.TH;
      const int dn = r / per_img, rem = r - dn * per_img;
      const int dy = rem / g.TW, dx = rem - dy * g.TW;
      const int n = tile_n0 + dn, y = tile_y0 + dy, x = tile_x0 + dx;
      row_ok = (dn < g.TN) && (n < g.NB) && (y < g.H) && (x < g.W);
      grow = ((long long)n * g.H + y) * g.W + x;
    }
    const int bf = g.is_bf16;
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_acc + (uint32_t(quarter * 32) << 16);

    if (g.epi == EPI_GEGLU) {
      // tile columns [0, BN/2) = value half, [BN/2, BN) = gate half (weight rows pre-interleaved per tile)
      constexpr int HALF = BN / 2;
      const int out_col0 = blockIdx.x * HALF;
      static_assert(HALF % 16 == 0 || BN < 32, "GEGLU chunking");
#pragma unroll 1
      for (int c = 0; c < HALF; c += 16) {
        uint32_t a[16], b[16];
        tmem_ld_x16(t_row + c, a);
        tmem_ld_x16(t_row + HALF + c, b);
        tmem_ld_wait();
        if (row_ok) {
          uint32_t packed[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float v0 = __uint_as_float(a[j]), v1 = __uint_as_float(a[j + 1]);
            float g0 = __uint_as_float(b[j]), g1 = __uint_as_float(b[j + 1]);
            if (g.bias) {
              v0 += load16(g.bias, n0 + c + j, bf); v1 += load16(g.bias, n0 + c + j + 1, bf);
              g0 += load16(g.bias, n0 + HALF + c + j, bf); g1 += load16(g.bias, n0 + HALF + c + j + 1, bf);
            }
            // the reference materialises proj(x) in 16-bit before chunk/gelu: round both halves first
            float2 vr = unpack16(pack16(v0, v1, bf), bf), gr = unpack16(pack16(g0, g1, bf), bf);
            packed[j >> 1] = pack16(vr.x * gelu_erf(gr.x), vr.y * gelu_erf(gr.y), bf);
          }
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(g.C) + grow * g.ldc + out_col0 + c);
          dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        }
      }
    } else {
      constexpr int CH = (BN % 32 == 0) ? 32 : 16;
#pragma unroll 1
      for (int c = 0; c < BN; c += CH) {
        float v[CH];
        if constexpr (CH == 32) {
          uint32_t a[32];
          tmem_ld_x32(t_row + c, a);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(a[j]);
        } else {
          uint32_t a[16];
          tmem_ld_x16(t_row + c, a);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(a[j]);
        }
        const int col0 = n0 + c;
        if (!row_ok || col0 >= g.N) continue;
        const bool full = (col0 + CH <= g.N);
        if (g.bias) {
#pragma unroll
          for (int j = 0; j < CH; ++j) if (full || col0 + j < g.N) v[j] += load16(g.bias, col0 + j, bf);
        }
        if (g.rowbias) {
          const uint16_t* rb = reinterpret_cast<const uint16_t*>(g.rowbias) + (grow / g.rows_per_group) * g.ld_rowbias;
#pragma unroll
          for (int j = 0; j < CH; ++j) if (full || col0 + j < g.N) v[j] += load16(rb, col0 + j, bf);
        }
        if (g.epi == EPI_QKV && col0 >= g.n_split) {
          // V columns: store transposed, Vt[(b*heads + h), dd, tok]; lanes hold consecutive tokens -> coalesced
          const int b = int(grow / g.ntok), tok = int(grow - (long long)b * g.ntok);
          const size_t vC = (size_t)g.heads * g.hdim;
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int vc = col0 + j - g.n_split;
            if (full || col0 + j < g.N) {
              store16(g.Vt, ((size_t)b * vC + vc) * g.ntok + tok, v[j], bf);   // (b*heads + h)*hdim + dd == b*C + vc
            }
          }
          continue;
        }
        uint16_t* crow = reinterpret_cast<uint16_t*>(g.C) + grow * g.ldc + col0;
        const bool vec_ok = full && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0);
        if (g.residual) {
          const uint16_t* rrow = reinterpret_cast<const uint16_t*>(g.residual) + grow * g.ldr + col0;
          if (vec_ok && ((reinterpret_cast<uintptr_t>(rrow) & 15) == 0)) {
#pragma unroll
            for (int q = 0; q < CH / 8; ++q) {
              uint4 u = reinterpret_cast<const uint4*>(rrow)[q];
              uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) { float2 f = unpack16(w[e], bf); v[q * 8 + 2 * e] += f.x; v[q * 8 + 2 * e + 1] += f.y; }
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) if (full || col0 + j < g.N) v[j] += load16(rrow, j, bf);
          }
        }
        if (g.out_scale != 1.0f) {
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] *= g.out_scale;
        }
        if (vec_ok) {
#pragma unroll
          for (int q = 0; q < CH / 8; ++q) {
            uint4 u;
            u.x = pack16(v[q * 8 + 0], v[q * 8 + 1], bf); u.y = pack16(v[q * 8 + 2], v[q * 8 + 3], bf);
            u.z = pack16(v[q * 8 + 4], v[q * 8 + 5], bf); u.w = pack16(v[q * 8 + 6], v[q * 8 + 7], bf);
            reinterpret_cast<uint4*>(crow)[q] = u;
          }
        } else {
#pragma unroll
          for (int j = 0; j < CH; ++j) if (full || col0 + j < g.N) store16(crow, j, v[j], bf);
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_acc);
  }
}

}  // namespace cid
