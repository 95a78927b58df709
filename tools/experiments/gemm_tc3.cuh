// 2-CTA (cta_group::2) persistent tcgen05 GEMM / implicit-GEMM conv3x3 (v3; same arguments and math as gemm_tc2.cuh).
//
// ncu on the 1-CTA kernel: tensor pipe ~50 % active with nothing saturated except the per-SM L2->SMEM ingest
// (a 128xBN SS tile needs (128+BN)*128 B per BN*2 MMA cycles = 96-115 B/clk/SM; ~55-65 B/clk is delivered).  Here a
// cluster of two CTAs computes a 256 x BN tile with ONE tcgen05.mma.cta_group::2 stream issued by the leader CTA: each
// CTA stages its own 128 rows of A but only HALF of the B tile (BN/2 rows), so the bytes each SM must pull per MMA cycle
// drop to 64 B/clk (BN=256) / 83 B/clk (BN=160).  Both CTAs' TMA loads complete on the LEADER's "full" barrier,
// tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs, and the peer's epilogue threads release the
// accumulator with remote mbarrier arrives.  Everything else is the v2 design:
//
// One CTA per SM loops over output tiles (tile = blockIdx.x + i*gridDim.x, N-tile fastest so CTAs running at the same
// time share the activation (A) tile in L2).  Roles (320 threads):
//   warp 0      TMA producer - keeps the STAGES-deep smem ring full ACROSS tile boundaries
//   warp 1      TMEM allocator (512 columns = two accumulators) + tcgen05.mma issuer; alternates accumulators so the
//               main loop of tile i+1 overlaps the epilogue of tile i
//   warps 2-9   epilogue: two warps per TMEM lane quarter, each draining half of the tile's columns in 16-column
//               tcgen05.ld chunks; residual rows are prefetched into registers BEFORE waiting for the accumulator and the
//               bias slice of the tile is staged once in smem, so no global-load latency sits between TMEM and the stores
#pragma once
#include "gemm_tc2.cuh"

namespace cid {


// ---- cluster / 2-CTA PTX
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_ss_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs retired) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// TMA loads whose completion bytes are credited to a barrier that may live in the peer (leader) CTA
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma2_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

template <int BN, int STAGES>
struct Gemm3Smem {
  static constexpr int A_BYTES = GEMM_BM * 128;
  static constexpr int B_BYTES = (BN / 2) * 128;                 // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFF = BAR_OFF + 256;             // 2 x BN floats
  static constexpr int TOTAL = BIAS_OFF + 2 * BN * 4 + 1024;
};


template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM2_THREADS, 1)
gemm_tc3_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmB, const GemmArgs g, const int n_tiles, const int total_tiles) {
  static_assert(BN % 32 == 0, "column split / B halves");
  constexpr int ACC_STRIDE = 256;                              // TMEM column offset between the two accumulators
  using SM = Gemm3Smem<BN, STAGES>;
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
  const uint32_t rank = cluster_ctarank();               // 0 = leader (issues the MMAs), 1 = peer
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
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
      for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 2 * GEMM2_EPI_THREADS); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2cta<512>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                     // peer barriers initialised before any remote signal
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
      for (int tile = pair; tile < total_tiles; tile += num_pairs) {
        const int nt = tile % n_tiles, mt = 2 * (tile / n_tiles) + int(rank);      // this CTA's 128-row half of the 256-row tile
        int tn0 = 0, ty0 = 0, tx0 = 0;
        if (g.a_mode != A_GEMM) tile_origin(mt, tn0, ty0, tx0);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * SM::STAGE_BYTES;
          const uint32_t sb = sa + SM::A_BYTES;
          const uint32_t lead_full = mapa_rank(full_bar(stage), 0);                // completion bytes go to the leader's barrier
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * (a_bytes + uint32_t(SM::B_BYTES)));
          const int tap = kb / kb_per_tap;
          const int cb = kb - tap * kb_per_tap;
          if (g.a_mode == A_GEMM) {
            if (cb < g.kblocks_a1) tma2_load_2d(sa, &tmA1, lead_full, cb * GEMM_BK, mt * GEMM_BM);
            else tma2_load_2d(sa, &tmA2, lead_full, (cb - g.kblocks_a1) * GEMM_BK, mt * GEMM_BM);
          } else if (g.a_mode == A_CONV) {
            const int ky = tap / 3, kx = tap - ky * 3;
            tma2_load_4d(sa, &tmA1, lead_full, cb * GEMM_BK, tx0 + kx - 1, ty0 + ky - 1, tn0);
          } else {
            const int ky = tap / 3, kx = tap - ky * 3;
            const int py = (ky == 1) ? 0 : 1, dy = (ky == 0) ? -1 : 0;
            const int px = (kx == 1) ? 0 : 1, dx = (kx == 0) ? -1 : 0;
            tma2_load_5d(sa, &tmA1, lead_full, cb * GEMM_BK, tx0 + dx, ty0 + dy, py * 2 + px, tn0);
          }
          tma2_load_2d(sb, &tmB, lead_full, kb * GEMM_BK, nt * BN + int(rank) * (BN / 2));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader CTA only)
    const uint32_t idesc = make_idesc(2 * GEMM_BM, BN, g.is_bf16);
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    const uint32_t a_lo0 = desc_lo(smem_base);
    if (rank == 0)
    for (int tile = pair; tile < total_tiles; tile += num_pairs, ++it) {
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
            umma_ss_2cta(tmem_acc, desc_make(a_lo + k * 2), desc_make(b_lo + k * 2), idesc, (kb | k) ? 1u : 0u);
          umma_commit_2cta(empty_bar(stage));
          if (kb == num_kb - 1) umma_commit_2cta(acc_full(ab));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      stage = __shfl_sync(0xffffffffu, stage, 0); phase = __shfl_sync(0xffffffffu, phase, 0);
    }
  } else {
    // ================================================================ epilogue (warps 2..9)
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;                              // which half of the tile's columns this warp drains
    const int r = quarter * 32 + lane;
    const int et = threadIdx.x - 64;                       // 0..255
    const int bf = g.is_bf16;
    const bool geglu = g.epi == EPI_GEGLU;
    // column range [c_beg, c_end) in 16-column chunks (GEGLU: over the value half only)
    constexpr int NCHUNK = BN / 16;
    constexpr int NCHUNK_G = (BN / 2) / 16 > 0 ? (BN / 2) / 16 : 1;
    const int nch = geglu ? NCHUNK_G : NCHUNK;
    const int ch_beg = half == 0 ? 0 : (nch + 1) / 2;
    const int ch_end = half == 0 ? (nch + 1) / 2 : nch;
    constexpr int MAXCH = (NCHUNK + 1) / 2;
    int it = 0;
    for (int tile = pair; tile < total_tiles; tile += num_pairs, ++it) {
      const int ab = it & 1;
      const uint32_t aphase = uint32_t(it >> 1) & 1u;
      const int nt = tile % n_tiles, mt = 2 * (tile / n_tiles) + int(rank);
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
                const float2 vr = unpack16(pack16(v0, v1, bf), bf), gr = unpack16(pack16(g0, g1, bf), bf);
                packed[j >> 1] = pack16(vr.x * gelu_erf(gr.x), vr.y * gelu_erf(gr.y), bf);
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
      mbar_arrive_cluster(mapa_rank(acc_empty(ab), 0));     // release accumulator `ab` to the leader's MMA warp
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                     // no CTA leaves while its pair may still touch its smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<512>(tmem_base);
  }
}

}  // namespace cid
