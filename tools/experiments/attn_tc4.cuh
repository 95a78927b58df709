// attn_self4_kernel - flash self-attention, v4 (same operands / results as attn_self3_kernel).
//
// What bounds v3 (profiles/r01_ncu_final_attn_self3_*, tools/microbench_softmax.cu): per 128 x 128 score tile the SM must
//   * read 64 KB of fp32 scores out of TMEM        (tcgen05.ld: a fixed bytes/clk per SM),
//   * evaluate 16 384 exponentials                 (MUFU: 16 / clk / SM = 1024 cycles; ex2.approx.f16x2 is TWO MUFU ops in SASS),
//   * and run ~450-580 tensor-pipe cycles (d = 40 ... 64),
// three different units - but v3 runs them one after the other inside each softmax thread (4 x LDTM, wait, row max, 128 exps),
// so a tile costs roughly the SUM of the first two.  v4 overlaps them inside every thread:
//   * the 128 scores of a row are processed in four 32-column chunks; the tcgen05.ld of chunk c+1 is in flight while chunk c goes
//     through max / FFMA / cvt / MUFU / st.shared (two 32-register buffers instead of v3's 128-register row);
//   * that needs the row maximum BEFORE the tile has been read, so P is computed OPTIMISTICALLY against the stale maximum m_use that
//     O and l already carry (FlashAttention-4's lazy rescaling makes this legal: any m_use within 2^8 of the true maximum is fine).
//     The chunk loop tracks the tile's true maximum; only if it exceeds m_use by more than the threshold (first tiles of a row block)
//     the warp rescales O / l and recomputes the tile's P from TMEM - S is released to the MMA warp only after that decision;
//   * the first key tile has no m_use yet: one extra max-only pass over S.
// Everything else (TMA producer, MMA issue order, O and l accumulated in TMEM, P.1 row sums, epilogue) is v3's.
#pragma once
#include "attn_common.cuh"

namespace cid {

__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr float ATTN_RESCALE_THRESHOLD = 8.0f;      // log2 units: P <= 2^8, exact range of fp16 / bf16

// max over one 32-column chunk (columns >= nvalid are masked to -inf first when the tile is ragged)
__device__ __forceinline__ float chunk_max(uint32_t (&v)[32], int nvalid, float mx) {
  if (nvalid < 32) {
#pragma unroll
    for (int i = 0; i < 32; ++i) if (i >= nvalid) v[i] = 0xff800000u;
  }
#pragma unroll
  for (int i = 0; i < 32; i += 2) mx = max3(mx, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
  return mx;
}
// P = 2^(s c - m c) for one 32-column chunk, packed 16-bit, straight into the 128B-swizzled A-operand tile (columns cc .. cc+31 of row r)
template <int BF>
__device__ __forceinline__ void chunk_exp_store(const uint32_t (&v)[32], float c, float nmc, uint8_t* sP, int r, int cc) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 32; i += 2)
    pk[i >> 1] = ex2_pack<BF>(fmaf(__uint_as_float(v[i]), c, nmc), fmaf(__uint_as_float(v[i + 1]), c, nmc));
  uint8_t* tile = sP + (cc >> 6) * 16384;
  const int col = cc & 63;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    st_sw128(tile, r, col + q * 8, make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]));
}

template <int D_PAD, int BF>
__global__ void __launch_bounds__(ATTN_THREADS, Attn2Cfg<D_PAD>::MIN_CTAS)
attn_self4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmVt, const AttnArgs a) {
  using C = Attn2Cfg<D_PAD>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if ((sbase & 1023u) != 0) { if (threadIdx.x == 0) printf("cid: attn smem base not 1024-aligned\n"); __trap(); }
  const uint32_t bar0 = sbase + C::OFF_BAR;
  const uint32_t q_full = bar0;
  auto k_full = [&](int s) { return bar0 + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bar0 + 8u * (1 + STAGES + s); };
  auto v_full = [&](int s) { return bar0 + 8u * (1 + 2 * STAGES + s); };
  auto v_empty = [&](int s) { return bar0 + 8u * (1 + 3 * STAGES + s); };
  const uint32_t s_full = bar0 + 8u * (1 + 4 * STAGES);
  const uint32_t s_free = s_full + 8u;
  const uint32_t p_full = s_full + 16u;
  const uint32_t pv_full = s_full + 24u;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + C::OFF_BAR + 8 * (6 + 4 * STAGES));

  const int warp = warp_id(), lane = lane_id();
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int T = (a.Nkv + 127) / 128;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmVt); }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int s = 0; s < STAGES; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); }
      mbar_init(s_full, 1); mbar_init(s_free, 128); mbar_init(p_full, 128); mbar_init(pv_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<C::TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  }
  if (warp == 2) {
    reinterpret_cast<uint32_t*>(smem_raw + C::OFF_ONES)[lane] = BF ? 0x3F803F80u : 0x3C003C00u;
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_S = tmem, tmem_O = tmem + 128, tmem_L = tmem + 128 + D_PAD;
  griddep_wait();                  // PDL: the prologue above overlaps the predecessor's tail

  if (warp == 0) {
    // ============================================================ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, C::Q_BYTES);
      for (int ch = 0; ch < C::NCH; ++ch) tma_load_4d(sbase + ch * 16384, &tmQ, q_full, ch * 64, q0, h, b);
      int stage = 0; uint32_t phase = 0;
      for (int j = 0; j < T; ++j) {
        const int k0 = j * 128;
        mbar_wait(k_empty(stage), phase ^ 1u);
        mbar_expect_tx(k_full(stage), C::K_BYTES);
        for (int ch = 0; ch < C::NCH; ++ch)
          tma_load_4d(sbase + C::OFF_K + stage * C::K_BYTES + ch * 16384, &tmK, k_full(stage), ch * 64, k0, h, b);
        mbar_wait(v_empty(stage), phase ^ 1u);
        mbar_expect_tx(v_full(stage), C::V_BYTES);
        for (int kc = 0; kc < 2; ++kc)
          tma_load_3d(sbase + C::OFF_V + stage * C::V_BYTES + kc * C::V_CHUNK, &tmVt, v_full(stage), k0 + kc * 64, 0, b * a.H + h);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      griddep_launch_dependents();
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    const uint32_t idesc_s = make_idesc(128, 128, BF);
    const uint32_t idesc_pv = make_idesc(128, D_PAD, BF);
    const uint32_t idesc_l = make_idesc(128, 16, BF);
    const uint64_t ones_desc = make_desc_alias128(sbase + C::OFF_ONES);
    const uint32_t q_lo = desc_lo(sbase), k_lo = desc_lo(sbase + C::OFF_K), v_lo = desc_lo(sbase + C::OFF_V), p_lo = desc_lo(sbase + C::OFF_P);
    auto issue_S = [&](int stage) {
      const uint32_t kl = k_lo + uint32_t(stage * C::K_BYTES) / 16;
#pragma unroll
      for (int ch = 0; ch < C::NCH; ++ch) {
        const int ksteps = (D_PAD - ch * 64 >= 64) ? 4 : (D_PAD - ch * 64) / 16;
#pragma unroll
        for (int kk = 0; kk < ksteps; ++kk)
          umma_ss(tmem_S, desc_make(q_lo + ch * 1024 + kk * 2), desc_make(kl + ch * 1024 + kk * 2), idesc_s, (ch | kk) ? 1u : 0u);
      }
    };
    mbar_wait(q_full, 0);
    int stage = 0; uint32_t phase = 0;
    int nstage = 0; uint32_t nphase = 0;
    mbar_wait(k_full(0), 0);
    tc_fence_after();
    if (lane == 0) { issue_S(0); umma_commit(s_full); umma_commit(k_empty(0)); }
    __syncwarp();
    if (++nstage == STAGES) { nstage = 0; nphase ^= 1u; }
    for (int j = 0; j < T; ++j) {
      if (j + 1 < T) {
        mbar_wait(k_full(nstage), nphase);
        mbar_wait(s_free, uint32_t(j & 1));              // every softmax thread is done reading S_j from TMEM
        tc_fence_after();
        if (lane == 0) { issue_S(nstage); umma_commit(s_full); umma_commit(k_empty(nstage)); }
        __syncwarp();
        if (++nstage == STAGES) { nstage = 0; nphase ^= 1u; }
      }
      mbar_wait(p_full, uint32_t(j & 1));                // P_j in smem, O / l rescaled if needed
      mbar_wait(v_full(stage), phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t vl = v_lo + uint32_t(stage * C::V_BYTES) / 16;
        const uint32_t acc0 = j > 0 ? 1u : 0u;
#pragma unroll
        for (int kc = 0; kc < 2; ++kc)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t pdesc = desc_make(p_lo + kc * 1024 + kk * 2);
            const uint32_t acc = (kc | kk) ? 1u : acc0;
            umma_ss(tmem_O, pdesc, desc_make(vl + kc * (C::V_CHUNK / 16) + kk * 2), idesc_pv, acc);
            umma_ss(tmem_L, pdesc, ones_desc, idesc_l, acc);
          }
        umma_commit(pv_full);
        umma_commit(v_empty(stage));
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ============================================================ softmax + output (warps 2..5, one query row per thread)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    const uint32_t tS = tmem_S + lane_off;
    uint8_t* sP = smem_raw + C::OFF_P;
    const float c = a.scale_log2;
    float m_use = -INFINITY;                              // row max baked into O, l and used for P
    uint32_t va[32], vb[32];

    for (int j = 0; j < T; ++j) {
      const int kvalid = a.Nkv - j * 128;                 // >= 1
      mbar_wait(s_full, uint32_t(j & 1));
      tc_fence_after();
      if (j == 0) {
        // first key tile: nothing to be stale against - exact row max with a max-only pass (the same double-buffered loads)
        float m0 = -INFINITY;
        tmem_ld_x32(tS, va);
        tmem_ld_wait(); tmem_ld_x32(tS + 32, vb); m0 = chunk_max(va, kvalid, m0);
        tmem_ld_wait(); tmem_ld_x32(tS + 64, va); m0 = chunk_max(vb, kvalid - 32, m0);
        tmem_ld_wait(); tmem_ld_x32(tS + 96, vb); m0 = chunk_max(va, kvalid - 64, m0);
        tmem_ld_wait(); m0 = chunk_max(vb, kvalid - 96, m0);
        m_use = m0;
      }
      // ---- optimistic pass: P against the stale m_use, chunk c+1 in flight while chunk c is exponentiated
      float nmc = -m_use * c;
      float mx = m_use;
      tmem_ld_x32(tS, va);
      tmem_ld_wait(); tmem_ld_x32(tS + 32, vb); mx = chunk_max(va, kvalid, mx);
      if (j > 0) { mbar_wait(pv_full, uint32_t((j - 1) & 1)); tc_fence_after(); }      // P.V_{j-1} retired: P smem reusable, O / l quiescent
      chunk_exp_store<BF>(va, c, nmc, sP, r, 0);
      tmem_ld_wait(); tmem_ld_x32(tS + 64, va); mx = chunk_max(vb, kvalid - 32, mx);
      chunk_exp_store<BF>(vb, c, nmc, sP, r, 32);
      tmem_ld_wait(); tmem_ld_x32(tS + 96, vb); mx = chunk_max(va, kvalid - 64, mx);
      chunk_exp_store<BF>(va, c, nmc, sP, r, 64);
      tmem_ld_wait(); mx = chunk_max(vb, kvalid - 96, mx);
      const bool need = (mx - m_use) * c > ATTN_RESCALE_THRESHOLD;
      if (!__any_sync(0xffffffffu, need)) {
        tc_fence_before();
        mbar_arrive(s_free);                              // S_j fully consumed: S_{j+1} may overwrite it
        chunk_exp_store<BF>(vb, c, nmc, sP, r, 96);
      } else {
        // ---- rare: the true max ran away from m_use (early tiles).  Rescale O / l, refresh m_use, recompute this tile's P from TMEM.
        const float f = need ? fast_exp2((m_use - mx) * c) : 1.0f;
        if (j > 0) {
#pragma unroll
          for (int cc = 0; cc < D_PAD + 16; cc += 16) {   // O columns then the 16 row-sum columns (contiguous in TMEM)
            uint32_t t[16];
            tmem_ld_x16(tmem_O + lane_off + cc, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * f);
            tmem_st_x16(tmem_O + lane_off + cc, t);
          }
          tmem_st_wait();
        }
        if (need) m_use = mx;
        nmc = -m_use * c;
        tmem_ld_x32(tS, va);
        tmem_ld_wait(); tmem_ld_x32(tS + 32, vb); chunk_max(va, kvalid, 0.f);            // (re-applies the ragged-tile mask)
        chunk_exp_store<BF>(va, c, nmc, sP, r, 0);
        tmem_ld_wait(); tmem_ld_x32(tS + 64, va); chunk_max(vb, kvalid - 32, 0.f);
        chunk_exp_store<BF>(vb, c, nmc, sP, r, 32);
        tmem_ld_wait(); tmem_ld_x32(tS + 96, vb); chunk_max(va, kvalid - 64, 0.f);
        chunk_exp_store<BF>(va, c, nmc, sP, r, 64);
        tmem_ld_wait(); chunk_max(vb, kvalid - 96, 0.f);
        tc_fence_before();
        mbar_arrive(s_free);
        chunk_exp_store<BF>(vb, c, nmc, sP, r, 96);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---- normalise and store
    mbar_wait(pv_full, uint32_t((T - 1) & 1));
    tc_fence_after();
    float l;
    {
      uint32_t t[16];
      tmem_ld_x16(tmem_L + lane_off, t);
      tmem_ld_wait();
      l = __uint_as_float(t[0]);
    }
    const float inv = 1.f / l;
    const bool row_ok = q0 + r < a.Nq;
    uint16_t* dst = reinterpret_cast<uint16_t*>(a.O) + ((long long)b * a.Nq + q0 + r) * a.ldo + h * a.d;
#pragma unroll
    for (int cc = 0; cc < D_PAD; cc += 16) {
      uint32_t t[16];
      tmem_ld_x16(tmem_O + lane_off + cc, t);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (cc + q * 8 < a.d) {
            uint4 u;
            u.x = pack16(__uint_as_float(t[q * 8 + 0]) * inv, __uint_as_float(t[q * 8 + 1]) * inv, BF);
            u.y = pack16(__uint_as_float(t[q * 8 + 2]) * inv, __uint_as_float(t[q * 8 + 3]) * inv, BF);
            u.z = pack16(__uint_as_float(t[q * 8 + 4]) * inv, __uint_as_float(t[q * 8 + 5]) * inv, BF);
            u.w = pack16(__uint_as_float(t[q * 8 + 6]) * inv, __uint_as_float(t[q * 8 + 7]) * inv, BF);
            *reinterpret_cast<uint4*>(dst + cc + q * 8) = u;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<C::TMEM_COLS>(tmem); }
}

}  // namespace cid
