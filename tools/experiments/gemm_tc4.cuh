// gemm_tc4_kernel - persistent tcgen05 GEMM / implicit-GEMM conv3x3 with KPI k-blocks per TMA instruction (v4 of the GEMM;
// same arguments, math, roles and epilogue as gemm_tc2.cuh).
//
// Measured with tools/microbench.cu on B200 (profiles/r01_microbench_*.log): a cp.async.bulk.tensor costs ~225 cycles fixed plus
// ~1 cycle per 200 bytes and the requests of one SM are serialised, so a pipeline that needs TWO TMA instructions per 64-wide
// k-block (A tile, B tile) advances one k-block per ~550-630 cycles whatever the tile size: 58 % tensor utilisation for the
// 128x160 tile even with every byte L2-resident.  Fetching KPI=2 k-blocks per instruction (3-D box {64, rows, 2} over a
// [rows, K] operand viewed as {64, rows, K/64}; 5-D box for the NHWC conv operand) brings that to ~406 cycles (79 %).
// A stage therefore holds KPI k-blocks: [A chunk 0 | A chunk 1 | B chunk 0 | B chunk 1]; odd k-block counts (K=320 -> 5)
// leave the last chunk zero-filled by TMA and the MMA issuer simply skips it.
#pragma once
#include <type_traits>
#include "gemm_tc2.cuh"

namespace cid {


template <int BN, int STAGES, int KPI>
struct Gemm4Smem {
  static constexpr int A_BYTES = GEMM_BM * 128;               // one 64-wide k-block of A (full 128-row tile)
  static constexpr int B_BYTES = BN * 128;                    // one k-block of B
  static constexpr int STAGE_BYTES = KPI * (A_BYTES + B_BYTES);
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFF = BAR_OFF + 256;             // 2 x BN floats
  static constexpr int TOTAL = BIAS_OFF + 2 * BN * 4 + 1024;
};


template <int BN, int STAGES, int KPI>
__global__ void __launch_bounds__(GEMM2_THREADS, 1)
gemm_tc4_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmB, const GemmArgs g, const int n_tiles, const int total_tiles) {
  static_assert(BN % 32 == 0 || BN == 16, "column split");
  constexpr int ACC_STRIDE = 256;                              // TMEM column offset between the two accumulators
  using SM = Gemm4Smem<BN, STAGES, KPI>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + SM::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto acc_full = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto acc_empty = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + SM::BAR_OFF + 8 * (2 * STAGES + 4));
  float* bias_s = reinterpret_cast<float*>(smem_gen + SM::BIAS_OFF);

  const int warp = warp_id();
  const int lane = lane_id();
  const int kb_per_tap = g.kblocks_a1 + g.kblocks_a2;
  const int num_kb = g.taps * kb_per_tap;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    if (g.kblocks_a2 > 0) tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), GEMM2_EPI_THREADS); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_origin = [&](int mt, int& tn0, int& ty0, int& tx0) {
    const int tx = mt % g.tiles_x;
    const int rest = mt / g.tiles_x;
    tx0 = tx * g.TW; ty0 = (rest % g.tiles_y) * g.TH; tn0 = (rest / g.tiles_y) * g.TN;
  };

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      const uint32_t a_bytes = (g.a_mode == A_GEMM) ? uint32_t(SM::A_BYTES) : uint32_t(g.TW * g.TH * g.TN * 128);
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % n_tiles, mt = tile / n_tiles;
        int tn0 = 0, ty0 = 0, tx0 = 0;
        if (g.a_mode != A_GEMM) tile_origin(mt, tn0, ty0, tx0);
        for (int tap = 0; tap < g.taps; ++tap) {
          const int ky = tap / 3, kx = tap - ky * 3;
          for (int src = 0; src < 2; ++src) {
            const int nk = src == 0 ? g.kblocks_a1 : g.kblocks_a2;
            const int kb0 = tap * kb_per_tap + (src ? g.kblocks_a1 : 0);        // first k-block of this (tap, source) in B
            for (int cb = 0; cb < nk; cb += KPI) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              const uint32_t sa = smem_base + stage * SM::STAGE_BYTES;
              const uint32_t sb = sa + KPI * SM::A_BYTES;
              mbar_expect_tx(full_bar(stage), uint32_t(KPI) * (a_bytes + uint32_t(SM::B_BYTES)));
              if (g.a_mode == A_GEMM) tma_load_3d(sa, src ? &tmA2 : &tmA1, full_bar(stage), 0, mt * GEMM_BM, cb);
              else tma_load_5d(sa, &tmA1, full_bar(stage), 0, tx0 + kx - 1, ty0 + ky - 1, tn0, cb);
              tma_load_3d(sb, &tmB, full_bar(stage), 0, nt * BN, kb0 + cb);
              if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t idesc = make_idesc(GEMM_BM, BN, g.is_bf16);
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    const uint32_t a_lo0 = desc_lo(smem_base);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      const uint32_t aphase = uint32_t(it >> 1) & 1u;
      mbar_wait(acc_empty(ab), aphase ^ 1u);              // epilogue has drained this accumulator (first use: free)
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + ab * ACC_STRIDE;
      if (lane == 0) {                                     // one thread runs the whole issue loop (no per-k-block warp sync)
        uint32_t first = 1u;
        const uint32_t a_chunk16 = ((g.a_mode == A_GEMM) ? uint32_t(SM::A_BYTES) : uint32_t(g.TW * g.TH * g.TN * 128)) / 16;
        for (int tap = 0; tap < g.taps; ++tap) {
          for (int src = 0; src < 2; ++src) {
            const int nk = src == 0 ? g.kblocks_a1 : g.kblocks_a2;
            for (int cb = 0; cb < nk; cb += KPI) {
              mbar_wait(full_bar(stage), phase);
              tc_fence_after();
              const uint32_t a_lo = a_lo0 + uint32_t(stage) * uint32_t(SM::STAGE_BYTES / 16);
              const uint32_t b_lo = a_lo + uint32_t(KPI * SM::A_BYTES / 16);
              const int nvalid = (nk - cb < KPI) ? (nk - cb) : KPI;              // trailing chunk of an odd count is zero-filled: skip it
#pragma unroll
              for (int q = 0; q < KPI; ++q) {
                if (q < nvalid) {
#pragma unroll
                  for (int k = 0; k < GEMM_BK / 16; ++k) {
                    umma_ss(tmem_acc, desc_make(a_lo + q * a_chunk16 + k * 2), desc_make(b_lo + q * (SM::B_BYTES / 16) + k * 2), idesc, first ? 0u : 1u);
                    first = 0u;
                  }
                }
              }
              umma_commit(empty_bar(stage));
              if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
          }
        }
        umma_commit(acc_full(ab));
      }
      stage = __shfl_sync(0xffffffffu, stage, 0); phase = __shfl_sync(0xffffffffu, phase, 0);
    }
  } else {
    // ================================================================ epilogue (warps 2..9)
    // (generic lambda: the 16-bit flavour becomes a compile-time constant, so pack/unpack fold to one conversion)
    auto epilogue = [&](auto bf_tag) {
    constexpr int bf = decltype(bf_tag)::value;
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;                              // which half of the tile's columns this warp drains
    const int r = quarter * 32 + lane;
    const int et = threadIdx.x - 64;                       // 0..255
    const bool geglu = g.epi == EPI_GEGLU;
    // column range [c_beg, c_end) in 16-column chunks (GEGLU: over the value half only)
    constexpr int NCHUNK = BN / 16;
    constexpr int NCHUNK_G = (BN / 2) / 16 > 0 ? (BN / 2) / 16 : 1;
    const int nch = geglu ? NCHUNK_G : NCHUNK;
    const int ch_beg = half == 0 ? 0 : (nch + 1) / 2;
    const int ch_end = half == 0 ? (nch + 1) / 2 : nch;
    constexpr int MAXCH = (NCHUNK + 1) / 2;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      const uint32_t aphase = uint32_t(it >> 1) & 1u;
      const int nt = tile % n_tiles, mt = tile / n_tiles;
      const int n0 = nt * BN;
      long long grow; bool row_ok;
      if (g.a_mode == A_GEMM) {
        grow = (long long)mt * GEMM_BM + r;
        row_ok = grow < g.M;
      } else {
        int tn0, ty0, tx0;
        tile_origin(mt, tn0, ty0, tx0);
        const int per_img = g.TW * g.TH;
        const int dn = r / per_img, rem = r - dn * per_img;
        const int dy = rem / g.TW, dx = rem - dy * g.TW;
        const int n = tn0 + dn, y = ty0 + dy, x = tx0 + dx;
        row_ok = (dn < g.TN) && (n < g.NB) && (y < g.H) && (x < g.W);
        grow = ((long long)n * g.H + y) * g.W + x;
      }
      // stage this tile's bias slice (fp32) in smem; buffer alternates with the accumulator
      float* bs = bias_s + ab * BN;
      for (int j = et; j < BN; j += GEMM2_EPI_THREADS) bs[j] = (g.bias && n0 + j < g.N) ? load16(g.bias, n0 + j, bf) : 0.f;
      // prefetch residual rows for this thread's chunks (latency overlaps the wait for the accumulator)
      uint4 res[MAXCH][2];
      const bool use_res = g.residual != nullptr && !geglu && row_ok;
      const uint16_t* rrow = use_res ? reinterpret_cast<const uint16_t*>(g.residual) + grow * g.ldr + n0 : nullptr;
      const bool res_vec = use_res && ((reinterpret_cast<uintptr_t>(rrow) & 15) == 0) && (n0 + BN <= g.N);
#pragma unroll
      for (int c = 0; c < MAXCH; ++c) {
        const int ch = ch_beg + c;
        if (res_vec && ch < ch_end) {
          res[c][0] = reinterpret_cast<const uint4*>(rrow + ch * 16)[0];
          res[c][1] = reinterpret_cast<const uint4*>(rrow + ch * 16)[1];
        }
      }
      epi_bar_sync();                                       // bias slice visible to all epilogue threads
      mbar_wait(acc_full(ab), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ab * ACC_STRIDE + (uint32_t(quarter * 32) << 16);

      if (BN >= 32 && geglu) {
        constexpr int HALF = BN / 2;
        const int out_col0 = nt * HALF;
#pragma unroll
        for (int c = 0; c < MAXCH; ++c) {
          const int ch = ch_beg + c;
          if (ch < ch_end) {
            uint32_t a[16], b[16];
            tmem_ld_x16(t_row + ch * 16, a);
            tmem_ld_x16(t_row + HALF + ch * 16, b);
            tmem_ld_wait();
            if (row_ok) {
              uint32_t packed[8];
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float v0 = __uint_as_float(a[j]) + bs[ch * 16 + j], v1 = __uint_as_float(a[j + 1]) + bs[ch * 16 + j + 1];
                const float g0 = __uint_as_float(b[j]) + bs[HALF + ch * 16 + j], g1 = __uint_as_float(b[j + 1]) + bs[HALF + ch * 16 + j + 1];
                packed[j >> 1] = pack16(v0 * gelu_erf(g0), v1 * gelu_erf(g1), bf);
              }
              uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(g.C) + grow * g.ldc + out_col0 + ch * 16);
              dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
              dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
            }
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < MAXCH; ++c) {
          const int ch = ch_beg + c;
          if (ch < ch_end) {
            uint32_t a[16];
            tmem_ld_x16(t_row + ch * 16, a);
            tmem_ld_wait();
            const int col0 = n0 + ch * 16;
            if (row_ok && col0 < g.N) {
              const bool full = (col0 + 16 <= g.N);
              float v[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(a[j]) + bs[ch * 16 + j];
              if (g.rowbias) {
                const uint16_t* rb = reinterpret_cast<const uint16_t*>(g.rowbias) + (grow / g.rows_per_group) * g.ld_rowbias + col0;
                if (full && ((reinterpret_cast<uintptr_t>(rb) & 15) == 0)) {
                  float f0[8], f1[8];
                  unpack8(reinterpret_cast<const uint4*>(rb)[0], f0, bf); unpack8(reinterpret_cast<const uint4*>(rb)[1], f1, bf);
#pragma unroll
                  for (int j = 0; j < 8; ++j) { v[j] += f0[j]; v[8 + j] += f1[j]; }
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j) if (full || col0 + j < g.N) v[j] += load16(rb, j, bf);
                }
              }
              if (g.epi == EPI_QKV && col0 >= g.n_split) {
                const int b = int(grow / g.ntok), tok = int(grow - (long long)b * g.ntok);
          const size_t vC = (size_t)g.heads * g.hdim;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int vc = col0 + j - g.n_split;
                  if (full || col0 + j < g.N) {
                    store16(g.Vt, ((size_t)b * vC + vc) * g.ntok + tok, v[j], bf);   // (b*heads + h)*hdim + dd == b*C + vc
                  }
                }
              } else {
                if (use_res) {
                  if (res_vec) {
                    float f0[8], f1[8];
                    unpack8(res[c][0], f0, bf); unpack8(res[c][1], f1, bf);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { v[j] += f0[j]; v[8 + j] += f1[j]; }
                  } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (full || col0 + j < g.N) v[j] += load16(rrow, ch * 16 + j, bf);
                  }
                }
                if (g.out_scale != 1.0f) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) v[j] *= g.out_scale;
                }
                uint16_t* crow = reinterpret_cast<uint16_t*>(g.C) + grow * g.ldc + col0;
                if (full && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
                  float lo[8], hi[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) { lo[j] = v[j]; hi[j] = v[8 + j]; }
                  reinterpret_cast<uint4*>(crow)[0] = pack8(lo, bf);
                  reinterpret_cast<uint4*>(crow)[1] = pack8(hi, bf);
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j) if (full || col0 + j < g.N) store16(crow, j, v[j], bf);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty(ab));                           // this thread no longer reads accumulator `ab`
    }
    };
    if (g.is_bf16) epilogue(std::integral_constant<int, 1>{}); else epilogue(std::integral_constant<int, 0>{});
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace cid
