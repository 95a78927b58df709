// gemm_tc5_kernel - gemm_tc2.cuh with a ROW-COALESCED epilogue (v5; same arguments, math and main loop).
//
// Why: for the small-K GEMMs that dominate SD1.5 (K = 320: five k-blocks) a tile took ~10k cycles against ~3k of main loop.
// In the v2 epilogue every lane owns one accumulator ROW, so each 16-byte store (and residual load) of a warp touches 32
// different 128-byte lines: 2560 LSU wavefronts per 128x160 tile each way.  v5 stages the tile through shared memory:
//   phase 1  (lane = row)    tcgen05.ld 16 columns, + bias (+ per-sample time-embedding bias, GEGLU), round to 16-bit exactly where
//                            the reference rounds its Linear output, write 32 bytes into a padded row-major smem tile;
//                            the TMEM accumulator is released right after this phase
//   phase 2  (lane = column) each warp takes 16 rows; lanes read consecutive 16-byte segments of a row from smem, load the residual
//                            row COALESCED, add, store COALESCED: 320 wavefronts per tile instead of 2560.
// V columns of the QKV epilogue keep their direct transposed stores (already lane-contiguous).  Tiles whose width is not a
// multiple of 8 columns (conv_out: 4 channels) use the v2 kernel.
#pragma once
#include <type_traits>
#include "gemm_tc2.cuh"

namespace cid {


template <int BN, int STAGES>
struct Gemm5Smem {
  static constexpr int A_BYTES = GEMM_BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFF = BAR_OFF + 256;             // 2 x BN floats
  static constexpr int OUT_PITCH = BN * 2 + 16;              // padded row pitch (bytes) of the output staging tile
  static constexpr int OUT_OFF = (BIAS_OFF + 2 * BN * 4 + 15) / 16 * 16;
  static constexpr int TOTAL = OUT_OFF + GEMM_BM * OUT_PITCH + 1024;
};


template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM2_THREADS, 1)
gemm_tc5_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmB, const GemmArgs g, const int n_tiles, const int total_tiles) {
  static_assert(BN % 32 == 0 || BN == 16, "column split");
  constexpr int ACC_STRIDE = 256;                              // TMEM column offset between the two accumulators
  using SM = Gemm5Smem<BN, STAGES>;
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
            tma_load_4d(sa, &tmA1, full_bar(stage), cb * GEMM_BK, tx0 + kx - 1, ty0 + ky - 1, tn0);
          } else {
            const int ky = tap / 3, kx = tap - ky * 3;
            const int py = (ky == 1) ? 0 : 1, dy = (ky == 0) ? -1 : 0;
            const int px = (kx == 1) ? 0 : 1, dx = (kx == 0) ? -1 : 0;
            tma_load_5d(sa, &tmA1, full_bar(stage), cb * GEMM_BK, tx0 + dx, ty0 + dy, py * 2 + px, tn0);
          }
          tma_load_2d(sb, &tmB, full_bar(stage), kb * GEMM_BK, nt * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
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
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + uint32_t(stage) * uint32_t(SM::STAGE_BYTES / 16);
          const uint32_t b_lo = a_lo + uint32_t(SM::A_BYTES / 16);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_ss(tmem_acc, desc_make(a_lo + k * 2), desc_make(b_lo + k * 2), idesc, (kb | k) ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (kb == num_kb - 1) umma_commit(acc_full(ab));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      stage = __shfl_sync(0xffffffffu, stage, 0); phase = __shfl_sync(0xffffffffu, phase, 0);
    }
  } else {
    // ================================================================ epilogue (warps 2..9)
    auto epilogue = [&](auto bf_tag) {
    constexpr int bf = decltype(bf_tag)::value;
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;                              // which half of the tile's columns this warp drains in phase 1
    const int r = quarter * 32 + lane;
    const int et = threadIdx.x - 64;                       // 0..255
    const bool geglu = g.epi == EPI_GEGLU;
    constexpr int NCHUNK = BN / 16;
    constexpr int NCHUNK_G = (BN / 2) / 16 > 0 ? (BN / 2) / 16 : 1;
    const int nch = geglu ? NCHUNK_G : NCHUNK;
    const int ch_beg = half == 0 ? 0 : (nch + 1) / 2;
    const int ch_end = half == 0 ? (nch + 1) / 2 : nch;
    constexpr int MAXCH = (NCHUNK + 1) / 2;
    uint8_t* out_s = smem_gen + SM::OUT_OFF;
    // global row of tile row `rr` (and whether it exists)
    auto row_of = [&](int mt, int rr, long long& grow) -> bool {
      if (g.a_mode == A_GEMM) { grow = (long long)mt * GEMM_BM + rr; return grow < g.M; }
      int tn0, ty0, tx0;
      tile_origin(mt, tn0, ty0, tx0);
      const int per_img = g.TW * g.TH;
      const int dn = rr / per_img, rem = rr - dn * per_img;
      const int dy = rem / g.TW, dx = rem - dy * g.TW;
      const int n = tn0 + dn, y = ty0 + dy, x = tx0 + dx;
      grow = ((long long)n * g.H + y) * g.W + x;
      return (dn < g.TN) && (n < g.NB) && (y < g.H) && (x < g.W);
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      const uint32_t aphase = uint32_t(it >> 1) & 1u;
      const int nt = tile % n_tiles, mt = tile / n_tiles;
      const int n0 = nt * BN;
      long long grow;
      const bool row_ok = row_of(mt, r, grow);
      // output-tile geometry: first global column, width staged through smem (QKV: only the non-V columns)
      const int out_col0 = geglu ? nt * (BN / 2) : n0;
      const int out_n = geglu ? g.N / 2 : g.N;
      int w_stage = geglu ? BN / 2 : BN;
      if (g.epi == EPI_QKV) { const int lim = g.n_split - n0; w_stage = lim < 0 ? 0 : (lim < BN ? lim : BN); }
      if (out_col0 + w_stage > out_n) w_stage = out_n - out_col0 > 0 ? out_n - out_col0 : 0;
      float* bs = bias_s + ab * BN;
      for (int j = et; j < BN; j += GEMM2_EPI_THREADS) bs[j] = (g.bias && n0 + j < g.N) ? load16(g.bias, n0 + j, bf) : 0.f;
      epi_bar_sync();                                       // bias slice visible; previous tile's phase 2 finished with out_s
      mbar_wait(acc_full(ab), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ab * ACC_STRIDE + (uint32_t(quarter * 32) << 16);
      uint8_t* my_row = out_s + r * SM::OUT_PITCH;
      // ---------------- phase 1: TMEM -> (+bias, activation) -> 16-bit -> smem tile (lane = row)
#pragma unroll
      for (int c = 0; c < MAXCH; ++c) {
        const int ch = ch_beg + c;
        if (ch < ch_end) {
          float v[16];
          if (BN >= 32 && geglu) {
            constexpr int HALF = BN / 2;
            uint32_t a[16], b[16];
            tmem_ld_x16(t_row + ch * 16, a);
            tmem_ld_x16(t_row + HALF + ch * 16, b);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
              v[j] = (__uint_as_float(a[j]) + bs[ch * 16 + j]) * gelu_erf(__uint_as_float(b[j]) + bs[HALF + ch * 16 + j]);
          } else {
            uint32_t a[16];
            tmem_ld_x16(t_row + ch * 16, a);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(a[j]) + bs[ch * 16 + j];
            const int col0 = n0 + ch * 16;
            if (g.rowbias && row_ok && col0 < g.N) {
              const uint16_t* rb = reinterpret_cast<const uint16_t*>(g.rowbias) + (grow / g.rows_per_group) * g.ld_rowbias + col0;
              if (col0 + 16 <= g.N && ((reinterpret_cast<uintptr_t>(rb) & 15) == 0)) {
                float f0[8], f1[8];
                unpack8(reinterpret_cast<const uint4*>(rb)[0], f0, bf); unpack8(reinterpret_cast<const uint4*>(rb)[1], f1, bf);
#pragma unroll
                for (int j = 0; j < 8; ++j) { v[j] += f0[j]; v[8 + j] += f1[j]; }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) if (col0 + j < g.N) v[j] += load16(rb, j, bf);
              }
            }
            if (g.epi == EPI_QKV && col0 >= g.n_split) {
              // V columns: transposed store Vt[(b*C + vc), tok]; lanes hold consecutive tokens -> already contiguous
              if (row_ok && col0 < g.N) {
                const int b = int(grow / g.ntok), tok = int(grow - (long long)b * g.ntok);
                const size_t vC = (size_t)g.heads * g.hdim;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (col0 + j < g.N) store16(g.Vt, ((size_t)b * vC + (col0 + j - g.n_split)) * g.ntok + tok, v[j], bf);
              }
              continue;
            }
          }
          float lo[8], hi[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { lo[j] = v[j]; hi[j] = v[8 + j]; }
          reinterpret_cast<uint4*>(my_row + ch * 32)[0] = pack8(lo, bf);
          reinterpret_cast<uint4*>(my_row + ch * 32)[1] = pack8(hi, bf);
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty(ab));                           // accumulator fully read: the next tile's MMAs may start
      epi_bar_sync();                                       // whole staging tile written
      // ---------------- phase 2: smem tile -> (+residual) -> global, row-coalesced (lane = 8-column segment)
      const int segs = w_stage >> 3;                        // 16-byte segments per row
      if (segs > 0) {
#pragma unroll 1
        for (int rr = ew; rr < GEMM_BM; rr += 8) {
          long long gr;
          if (!row_of(mt, rr, gr)) continue;
          const uint8_t* srow = out_s + rr * SM::OUT_PITCH;
          uint16_t* crow = reinterpret_cast<uint16_t*>(g.C) + gr * g.ldc + out_col0;
          const uint16_t* rrow = g.residual ? reinterpret_cast<const uint16_t*>(g.residual) + gr * g.ldr + out_col0 : nullptr;
          for (int sgi = lane; sgi < segs; sgi += 32) {
            uint4 u = *reinterpret_cast<const uint4*>(srow + sgi * 16);
            if (rrow || g.out_scale != 1.0f) {
              float f[8];
              unpack8(u, f, bf);
              if (rrow) {
                float q[8];
                unpack8(*reinterpret_cast<const uint4*>(rrow + sgi * 8), q, bf);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] += q[j];
              }
              if (g.out_scale != 1.0f) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] *= g.out_scale;
              }
              u = pack8(f, bf);
            }
            *reinterpret_cast<uint4*>(crow + sgi * 8) = u;
          }
        }
      }
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
