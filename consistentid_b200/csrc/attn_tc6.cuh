// attn_self6_kernel - flash self-attention, v6: FlashAttention-4 layout for head dims <= 80 (same operands / results as v5).
//
// Why (phase trace of v5, profiles/r02_attn_phase_trace_*.txt, and tools/microbench_softmax.cu): a 128 x 128 score tile costs 1030 cycles of
// MUFU (16 ex2 / clk / SM, whatever the packing) - the floor - but v3 ... v5 sit at ~1750 - 2000 cycles per tile because (a) P travels
// through SHARED MEMORY (32 KB written by the softmax warps, read back by P.V and again by the row-sum MMA: together with the operand reads
// of S = Q K^T that is ~1.3 k cycles of the SM's 128 B/clk shared-memory port per tile) and (b) the two co-resident CTAs run in lock-step:
// both do their exponentials at the same time and then both wait for the tensor pipe.  v6 removes both:
//   * one CTA per SM works on TWO 128-row query tiles with two softmax warpgroups and one shared K / V stream;
//   * P never touches shared memory: the softmax thread of row r writes its packed 16-bit probabilities over the first 64 columns of its own
//     score row in TENSOR MEMORY (tcgen05.st) and P.V takes its A operand from there (tcgen05.mma with A in TMEM);
//   * V^T tiles carry 16 extra rows of ones, so the same MMA that accumulates O = sum P V also accumulates the row sums l = sum P 1 in the 16
//     columns next to O (no second pass over P);
//   * the two warpgroups run free: while one waits for its P.V_j / S_{j+1} on the tensor pipe and reloads / max-reduces its next score row,
//     the other has the MUFU alone; when both are in their exponentials they share it (two warps per scheduler reach 92 % of the MUFU
//     rate, one warp alone only 67 %, tools/microbench_softmax.cu - which is why a strict ping-pong token, -DCID_ATTN_TOKEN, measured slower).
// TMEM (512 columns): S0 | S1 (128 fp32 columns each; P_i aliases columns 0..63 of S_i) | O0, l0 | O1, l1 (D_PAD + 16 each).
// Ordering relies on tcgen05.mma instructions of one thread executing in issue order: P.V_i(j) (reads P_i) is issued before S_i(j+1)
// (overwrites it), and the commit that signals S_i(j+1) therefore also covers P.V_i(j) - the O / l rescale needs no extra wait.
#pragma once
#include "attn_common.cuh"

namespace cid {

#ifndef CID_TMEM_ST_DEFINED
#define CID_TMEM_ST_DEFINED
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
#endif

constexpr int ATTN6_THREADS = 352;                    // warp 0 TMA (Q, K), warp 1 MMA, warps 2-5 softmax of query tile 0, warps 6-9 of tile 1, warp 10 TMA (V^T)
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int D_PAD>
struct Attn6Cfg {
  static_assert(D_PAD % 16 == 0 && D_PAD <= 80, "v6 covers head dims <= 80");
  static constexpr int NCH = (D_PAD + 63) / 64;                 // 64-wide head-dim chunks of Q / K
  static constexpr int STAGES = 2;
  static constexpr int Q_BYTES = NCH * 16384;                   // per query tile
  static constexpr int K_BYTES = NCH * 16384;                   // per stage
  static constexpr int VN = D_PAD + 16;                         // rows of the V^T operand: d (padded) + 16 rows of ones
  static constexpr int V_CHUNK = VN * 128;                      // one 64-key chunk
  static constexpr int V_BYTES = 2 * V_CHUNK;                   // per stage
  static constexpr int V_TX = 2 * D_PAD * 128;                  // bytes TMA delivers per stage (the ones rows are static)
  static constexpr int OFF_K = 2 * Q_BYTES;
  static constexpr int OFF_V = OFF_K + STAGES * K_BYTES;
  static constexpr int OFF_BAR = OFF_V + STAGES * V_BYTES;
  static constexpr int TOTAL = OFF_BAR + 256;
  static constexpr int TM_S = 0, TM_O = 256, O_STRIDE = D_PAD + 16;      // TMEM columns
  static_assert(TM_O + 2 * O_STRIDE <= 512, "TMEM budget");
};

template <int D_PAD, int BF>
__global__ void __launch_bounds__(ATTN6_THREADS, 1)
attn_self6_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmVt, const AttnArgs a) {
  using C = Attn6Cfg<D_PAD>;
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
  const uint32_t bar1 = bar0 + 8u * (1 + 4 * STAGES);
  auto s_full = [&](int i) { return bar1 + 8u * i; };            // S_i(j) complete (MMA commit)
  auto p_full = [&](int i) { return bar1 + 16u + 8u * i; };      // P_i(j) in TMEM (128 softmax threads)
  auto o_full = [&](int i) { return bar1 + 32u + 8u * i; };      // last P.V_i retired
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + C::OFF_BAR + 8 * (7 + 4 * STAGES));
  static_assert(8 * (8 + 4 * STAGES) <= 256, "barrier block");

  const int warp = warp_id(), lane = lane_id();
  const int q0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int T = (a.Nkv + 127) / 128;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmVt); }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int s = 0; s < STAGES; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(s_full(i), 1); mbar_init(p_full(i), 128); mbar_init(o_full(i), 1); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  }
  // the 16 rows of ones under every V^T chunk (swizzle-invariant: every 16-byte chunk of a row is the same)
  for (int i = threadIdx.x; i < STAGES * 2 * (16 * 128 / 4); i += blockDim.x) {
    const int chunk = i / (16 * 128 / 4), w = i % (16 * 128 / 4);
    reinterpret_cast<uint32_t*>(smem_raw + C::OFF_V + (chunk >> 1) * C::V_BYTES + (chunk & 1) * C::V_CHUNK + D_PAD * 128)[w] = BF ? 0x3F803F80u : 0x3C003C00u;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  griddep_wait();                  // PDL: the prologue above overlaps the predecessor's tail

  // Producer / issuer warps are WARP-CONVERGED with the single-thread instructions under elect_one() (see gemm_tc2.cuh: `if (lane == 0)`
  // regions cost ~90 cycles per tcgen05.mma and ~225 per TMA instruction in R2UR moves and ELECT / BRA.U.ANY retry loops).
  if (warp == 0) {
    // ============================================================ TMA producer: Q0, Q1, then the K stream
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * C::Q_BYTES);
      for (int i = 0; i < 2; ++i)
        for (int ch = 0; ch < C::NCH; ++ch) tma_load_4d(sbase + i * C::Q_BYTES + ch * 16384, &tmQ, q_full, ch * 64, q0 + i * 128, h, b);
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int j = 0; j < T; ++j) {
      mbar_wait(k_empty(stage), phase ^ 1u);
      const uint32_t kb = k_full(stage), dst = sbase + C::OFF_K + stage * C::K_BYTES;
      if (elect_one()) {
        mbar_expect_tx(kb, C::K_BYTES);
#pragma unroll
        for (int ch = 0; ch < C::NCH; ++ch) tma_load_4d(dst + ch * 16384, &tmK, kb, ch * 64, j * 128, h, b);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 10) {
    // ============================================================ TMA producer: the V^T stream (its own warp: a V buffer is released only
    // when P.V of BOTH query tiles retired, and a K tile must never queue behind that)
    int stage = 0; uint32_t phase = 0;
    const int bh = b * a.H + h;
    for (int j = 0; j < T; ++j) {
      mbar_wait(v_empty(stage), phase ^ 1u);
      const uint32_t vb = v_full(stage), dst = sbase + C::OFF_V + stage * C::V_BYTES;
      if (elect_one()) {
        mbar_expect_tx(vb, C::V_TX);
        tma_load_3d(dst, &tmVt, vb, j * 128, 0, bh);
        tma_load_3d(dst + C::V_CHUNK, &tmVt, vb, j * 128 + 64, 0, bh);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
    if (elect_one()) griddep_launch_dependents();
    __syncwarp();
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    const uint32_t idesc_s = make_idesc(128, 128, BF);
    const uint32_t idesc_pv = make_idesc(128, C::VN, BF);
    const uint32_t q_lo = desc_lo(sbase), k_lo = desc_lo(sbase + C::OFF_K), v_lo = desc_lo(sbase + C::OFF_V);
    // S_i = Q_i K^T into TMEM columns [i * 128, +128); all operands are warp-uniform values computed by the whole warp
    auto issue_S = [&](int i, int stage, uint32_t bar_s, uint32_t bar_k, bool release_k) {
      const uint32_t ql = q_lo + uint32_t(i * C::Q_BYTES) / 16, kl = k_lo + uint32_t(stage * C::K_BYTES) / 16;
      const uint32_t d_tm = tmem + C::TM_S + i * 128;
      if (elect_one()) {
#pragma unroll
        for (int ch = 0; ch < C::NCH; ++ch) {
          const int ksteps = (D_PAD - ch * 64 >= 64) ? 4 : (D_PAD - ch * 64) / 16;
#pragma unroll
          for (int kk = 0; kk < ksteps; ++kk)
            umma_ss(d_tm, desc_make(ql + ch * 1024 + kk * 2), desc_make(kl + ch * 1024 + kk * 2), idesc_s, (ch | kk) ? 1u : 0u);
        }
        umma_commit(bar_s);
        if (release_k) umma_commit(bar_k);
      }
      __syncwarp();
    };
#ifdef CID_ATTN_TRACE
    const bool trm = a.trace != nullptr && blockIdx.y == 0 && blockIdx.z == 0 && blockIdx.x < 16 && lane == 0;
    auto mstamp = [&](int j_, int e) { if (trm && j_ < 64) a.trace[((size_t)(32 + blockIdx.x) * 64 + j_) * 8 + e] = clock64(); };
#else
    auto mstamp = [&](int, int) {};
#endif
    mbar_wait(q_full, 0);
    mbar_wait(k_full(0), 0);
    tc_fence_after();
    issue_S(0, 0, s_full(0), 0u, false);
    issue_S(1, 0, s_full(1), k_empty(0), true);
    int ks = (STAGES > 1) ? 1 : 0; uint32_t kph = (STAGES > 1) ? 0u : 1u;     // stage / phase of K_{j+1}
    int vs = 0; uint32_t vph = 0;                                               // stage / phase of V_j
    for (int j = 0; j < T; ++j) {
      const bool more = j + 1 < T;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        mbar_wait(p_full(i), uint32_t(j & 1));               // P_i(j) written over S_i, O_i / l_i rescaled if needed
        mstamp(j, i * 3 + 0);
        if (i == 0) mbar_wait(v_full(vs), vph);
        tc_fence_after();
        const uint32_t vl = v_lo + uint32_t(vs * C::V_BYTES) / 16;
        const uint32_t a_tm = tmem + C::TM_S + i * 128;       // P_i: 64 columns of packed 16-bit pairs, 8 columns per 16-key MMA step
        const uint32_t d_tm = tmem + C::TM_O + i * C::O_STRIDE;
        const uint32_t acc0 = j > 0 ? 1u : 0u;
        const uint32_t ob = o_full(i), vb = v_empty(vs);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_ts(d_tm, a_tm + kk * 8, desc_make(vl + (kk >> 2) * (C::V_CHUNK / 16) + (kk & 3) * 2), idesc_pv, kk ? 1u : acc0);
          if (!more) umma_commit(ob);
          if (i == 1) umma_commit(vb);
        }
        __syncwarp();
        mstamp(j, i * 3 + 1);
        if (more) {
          if (i == 0) { mbar_wait(k_full(ks), kph); tc_fence_after(); }
          issue_S(i, ks, s_full(i), k_empty(ks), i == 1);      // executes after P.V_i(j) (issue order): S_i may overwrite P_i
        }
        mstamp(j, i * 3 + 2);
      }
      if (++vs == STAGES) { vs = 0; vph ^= 1u; }
      if (++ks == STAGES) { ks = 0; kph ^= 1u; }
    }
  } else {
    // ============================================================ softmax warpgroups (warps 2-5: query tile 0, warps 6-9: tile 1)
    const int wg = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    const uint32_t tS = tmem + C::TM_S + wg * 128 + lane_off;
    const uint32_t tO = tmem + C::TM_O + wg * C::O_STRIDE + lane_off;
    const float c = a.scale_log2;
    float m_use = -INFINITY;                              // row max baked into O, l and used for P
    const int bar_mine = 1 + wg, bar_other = 2 - wg;      // named barriers 1 / 2: the exponential token
#ifdef CID_ATTN_TOKEN
    if (wg == 1) named_bar_arrive(1, 256);                // warpgroup 0 goes first
#endif
#ifdef CID_ATTN_TRACE
    const bool tr = a.trace != nullptr && blockIdx.y == 0 && blockIdx.z == 0 && (warp == 2 || warp == 6) && lane == 0 && blockIdx.x < 16;
    auto stamp = [&](int j_, int e) { if (tr && j_ < 64) a.trace[((size_t)(blockIdx.x * 2 + wg) * 64 + j_) * 8 + e] = clock64(); };
    if (tr) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); a.trace[((size_t)(blockIdx.x * 2 + wg) * 64) * 8 + 7] = smid; }
#else
    auto stamp = [&](int, int) {};
#endif

    for (int j = 0; j < T; ++j) {
      const int kvalid = a.Nkv - j * 128;
      stamp(j, 0);
      mbar_wait(s_full(wg), uint32_t(j & 1));             // S_i(j) done - and with it every MMA issued before it, P.V_i(j-1) included
      tc_fence_after();
      stamp(j, 1);
      uint32_t v[128];
      {
        uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
        uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
        uint32_t (&v2)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[64]);
        uint32_t (&v3)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[96]);
        tmem_ld_x32(tS + 0, v0);
        tmem_ld_x32(tS + 32, v1);
        tmem_ld_x32(tS + 64, v2);
        tmem_ld_x32(tS + 96, v3);
        tmem_ld_wait();
      }
      stamp(j, 2);
      if (kvalid < 128) {
#pragma unroll
        for (int i = 0; i < 128; ++i) if (i >= kvalid) v[i] = 0xff800000u;       // -inf
      }
      float m8[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) m8[q] = max3(__uint_as_float(v[q * 16]), __uint_as_float(v[q * 16 + 1]), __uint_as_float(v[q * 16 + 2]));
#pragma unroll
      for (int i = 3; i < 15; i += 2) {
#pragma unroll
        for (int q = 0; q < 8; ++q) m8[q] = max3(m8[q], __uint_as_float(v[q * 16 + i]), __uint_as_float(v[q * 16 + i + 1]));
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) m8[q] = fmaxf(m8[q], __uint_as_float(v[q * 16 + 15]));
      const float mx = fmaxf(max3(max3(m8[0], m8[1], m8[2]), max3(m8[3], m8[4], m8[5]), m8[6]), m8[7]);
      const float m_new = fmaxf(m_use, mx);
      const bool need = (m_new - m_use) * c > ATTN_RESCALE_THRESHOLD;            // also true on the first tile (m_use = -inf)
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        const float f = need ? fast_exp2((m_use - m_new) * c) : 1.0f;            // O_i, l_i quiescent: see the s_full wait above
#pragma unroll
        for (int cc = 0; cc < D_PAD + 16; cc += 16) {
          uint32_t t[16];
          tmem_ld_x16(tO + cc, t);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * f);
          tmem_st_x16(tO + cc, t);
        }
        tmem_st_wait();
      }
      if (need) m_use = m_new;
      const float nmc = -m_use * c;
      stamp(j, 3);
#ifdef CID_ATTN_TOKEN
      named_bar_sync(bar_mine, 256);                      // the exponential token: the other warpgroup has finished its MUFU phase
#endif
      stamp(j, 4);
      uint32_t pk[64];
#pragma unroll
      for (int i = 0; i < 128; i += 2)
        pk[i >> 1] = ex2_pack<BF>(fmaf(__uint_as_float(v[i]), c, nmc), fmaf(__uint_as_float(v[i + 1]), c, nmc));
#ifdef CID_ATTN_TOKEN
      if (!(wg == 1 && j == T - 1)) named_bar_arrive(bar_other, 256);           // (no dangling arrival after the last tile)
#endif
      stamp(j, 5);
      // P_i(j) over the first 64 columns of this row's scores, packed pairs in key order: the A operand of P.V
#pragma unroll
      for (int cc = 0; cc < 64; cc += 16) {
        uint32_t (&p16)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[cc]);
        tmem_st_x16(tS + cc, p16);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full(wg));
      stamp(j, 6);
    }
    // ---- normalise and store
    mbar_wait(o_full(wg), 0);
    tc_fence_after();
    float l;
    {
      uint32_t t[16];
      tmem_ld_x16(tO + D_PAD, t);
      tmem_ld_wait();
      l = __uint_as_float(t[0]);
    }
    const float inv = 1.f / l;
    const int row = q0 + wg * 128 + r;
    const bool row_ok = row < a.Nq;
    uint16_t* dst = reinterpret_cast<uint16_t*>(a.O) + ((long long)b * a.Nq + row) * a.ldo + h * a.d;
#pragma unroll
    for (int cc = 0; cc < D_PAD; cc += 16) {
      uint32_t t[16];
      tmem_ld_x16(tO + cc, t);
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
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

}  // namespace cid
