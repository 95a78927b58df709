// attn_self7_kernel - flash self-attention for head dims <= 64 (same operands / results as v5 / v6): v6's layout (two 128-row query
// tiles per CTA, two softmax warpgroups, P in tensor memory, TS-MMA for P.V) with the last serialisation removed.
//
// v6's phase trace (profiles/r02_attn_phase_trace_v6_sd15.txt): because P_i aliases the score columns S_i, S_i(j+1) = Q_i K_{j+1}^T can
// only be ISSUED after P.V_i(j), i.e. after the exponentials of tile j - every warpgroup then waits ~800 cycles per tile (barrier hand-off
// + P.V + S on the tensor pipe), both warpgroups at nearly the same time, with the MUFU idle.  Here:
//   * P_i has its own 64 TMEM columns, so S_i(j+1) is issued as soon as the warpgroup has pulled S_i(j) into registers (s_free) and runs
//     under the exponentials of tile j; the next tile's scores are ready long before they are needed, and P.V_i(j) is off the critical path;
//   * the TMEM budget for that (S0 S1 | P0 P1 | O0 O1 = 256 + 128 + 2 d <= 512) has no room for row-sum columns: exponentials are evaluated
//     in fp32 (ex2.approx.ftz.f32, the same MUFU rate as the packed 16-bit form, which is two MUFU ops anyway) and the row sum is a plain
//     fp32 FADD of the unrounded probabilities (four partial sums), kept in a register and rescaled together with O.
// Roles (352 threads): warp 0 TMA (Q0, Q1, K stream), warp 10 TMA (V^T stream), warp 1 MMA issuer, warps 2-5 / 6-9 softmax of tile 0 / 1;
// producer / issuer warps are warp-converged with elect_one() around the single-thread instructions (see gemm_tc2.cuh).
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

#ifndef CID_ATTN6_SHARED
#define CID_ATTN6_SHARED
constexpr int ATTN6_THREADS = 352;                    // warp 0 TMA (Q, K), warp 1 MMA, warps 2-5 softmax of query tile 0, warps 6-9 of tile 1, warp 10 TMA (V^T)
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
#endif

template <int D_PAD>
struct Attn7Cfg {
  static_assert(D_PAD % 16 == 0 && D_PAD <= 64, "v7 covers head dims <= 64 (TMEM: 256 S + 128 P + 2 d O columns)");
  static constexpr int NCH = 1;                                 // one 64-wide head-dim chunk of Q / K
  static constexpr int STAGES = 2;
  static constexpr int Q_BYTES = 16384;                         // per query tile
  static constexpr int K_BYTES = 16384;                         // per stage
  static constexpr int V_CHUNK = D_PAD * 128;                   // one 64-key chunk of V^T
  static constexpr int V_BYTES = 2 * V_CHUNK;                   // per stage
  static constexpr int OFF_K = 2 * Q_BYTES;
  static constexpr int OFF_V = OFF_K + STAGES * K_BYTES;
  static constexpr int OFF_BAR = OFF_V + STAGES * V_BYTES;
  static constexpr int TOTAL = OFF_BAR + 256;
  static constexpr int TM_S = 0, TM_P = 256, TM_O = 384;        // TMEM columns: S0 S1 | P0 P1 (64 each: packed 16-bit pairs) | O0 O1 (D_PAD each)
  static_assert(TM_O + 2 * D_PAD <= 512, "TMEM budget");
  static_assert(V_CHUNK % 1024 == 0, "128B-swizzle atoms need 1 KB aligned chunks");
};

template <int D_PAD, int BF>
__global__ void __launch_bounds__(ATTN6_THREADS, 1)
attn_self7_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmVt, const AttnArgs a) {
  using C = Attn7Cfg<D_PAD>;
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
  auto s_free = [&](int i) { return bar1 + 16u + 8u * i; };      // S_i(j) pulled into registers by its 128 softmax threads
  auto p_full = [&](int i) { return bar1 + 32u + 8u * i; };      // P_i(j) in TMEM (128 softmax threads)
  auto pv_done = [&](int i) { return bar1 + 48u + 8u * i; };     // P.V_i(j) retired (MMA commit): P_i reusable, O_i quiescent
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + C::OFF_BAR + 8 * (9 + 4 * STAGES));
  static_assert(8 * (10 + 4 * STAGES) <= 256, "barrier block");

  const int warp = warp_id(), lane = lane_id();
  const int q0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int T = (a.Nkv + 127) / 128;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmVt); }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int s = 0; s < STAGES; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(s_full(i), 1); mbar_init(s_free(i), 128); mbar_init(p_full(i), 128); mbar_init(pv_done(i), 1); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  griddep_wait();                  // PDL: the prologue above overlaps the predecessor's tail

  if (warp == 0) {
    // ============================================================ TMA producer: Q0, Q1, then the K stream
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * C::Q_BYTES);
      for (int i = 0; i < 2; ++i) tma_load_4d(sbase + i * C::Q_BYTES, &tmQ, q_full, 0, q0 + i * 128, h, b);
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int j = 0; j < T; ++j) {
      mbar_wait(k_empty(stage), phase ^ 1u);
      const uint32_t kb = k_full(stage), dst = sbase + C::OFF_K + stage * C::K_BYTES;
      if (elect_one()) {
        mbar_expect_tx(kb, C::K_BYTES);
        tma_load_4d(dst, &tmK, kb, 0, j * 128, h, b);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 10) {
    // ============================================================ TMA producer: the V^T stream
    int stage = 0; uint32_t phase = 0;
    const int bh = b * a.H + h;
    for (int j = 0; j < T; ++j) {
      mbar_wait(v_empty(stage), phase ^ 1u);
      const uint32_t vb = v_full(stage), dst = sbase + C::OFF_V + stage * C::V_BYTES;
      if (elect_one()) {
        mbar_expect_tx(vb, C::V_BYTES);
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
    const uint32_t idesc_pv = make_idesc(128, D_PAD, BF);
    const uint32_t q_lo = desc_lo(sbase), k_lo = desc_lo(sbase + C::OFF_K), v_lo = desc_lo(sbase + C::OFF_V);
    constexpr int KSTEPS = D_PAD / 16;
    // S_i = Q_i K^T into TMEM columns [i * 128, +128); operands are warp-uniform values computed by the whole warp
    auto issue_S = [&](int i, int stage, bool release_k) {
      const uint32_t ql = q_lo + uint32_t(i * C::Q_BYTES) / 16, kl = k_lo + uint32_t(stage * C::K_BYTES) / 16;
      const uint32_t d_tm = tmem + C::TM_S + i * 128, sb = s_full(i), kb = k_empty(stage);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < KSTEPS; ++kk) umma_ss(d_tm, desc_make(ql + kk * 2), desc_make(kl + kk * 2), idesc_s, kk ? 1u : 0u);
        umma_commit(sb);
        if (release_k) umma_commit(kb);
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
    issue_S(0, 0, false);
    issue_S(1, 0, true);
    int ks = (STAGES > 1) ? 1 : 0; uint32_t kph = (STAGES > 1) ? 0u : 1u;     // stage / phase of K_{j+1}
    int vs = 0; uint32_t vph = 0;                                               // stage / phase of V_j
    for (int j = 0; j < T; ++j) {
      if (j + 1 < T) {
        // next tile's scores: as soon as each warpgroup holds S_i(j) in registers - they then run under the exponentials of tile j
        mbar_wait(k_full(ks), kph);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          mbar_wait(s_free(i), uint32_t(j & 1));
          tc_fence_after();
          issue_S(i, ks, i == 1);
        }
      }
      mstamp(j, 0);
      mbar_wait(v_full(vs), vph);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        mbar_wait(p_full(i), uint32_t(j & 1));               // P_i(j) in TMEM, O_i rescaled if needed
        mstamp(j, 1 + i * 2);
        tc_fence_after();
        const uint32_t vl = v_lo + uint32_t(vs * C::V_BYTES) / 16;
        const uint32_t a_tm = tmem + C::TM_P + i * 64;        // P_i: 64 columns of packed 16-bit pairs, 8 columns per 16-key MMA step
        const uint32_t d_tm = tmem + C::TM_O + i * D_PAD;
        const uint32_t acc0 = j > 0 ? 1u : 0u;
        const uint32_t pb = pv_done(i), vb = v_empty(vs);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_ts(d_tm, a_tm + kk * 8, desc_make(vl + (kk >> 2) * (C::V_CHUNK / 16) + (kk & 3) * 2), idesc_pv, kk ? 1u : acc0);
          umma_commit(pb);
          if (i == 1) umma_commit(vb);
        }
        __syncwarp();
        mstamp(j, 2 + i * 2);
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
    const uint32_t tP = tmem + C::TM_P + wg * 64 + lane_off;
    const uint32_t tO = tmem + C::TM_O + wg * D_PAD + lane_off;
    const float c = a.scale_log2;
    float m_use = -INFINITY;                              // row max baked into O, l and used for P
    float l_run = 0.f;                                    // row sum of the (unrounded) probabilities, same scaling as O
    // (Measured and dropped: half a period of head start for one warpgroup, to put the two in anti-phase - the exponential phase is ~70 % of
    // a tile, so the warpgroups overlap in it whatever the offset; no change in kernel time, profiles/r02_attn_phase_trace_v7b_sd15.txt.)
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
      mbar_wait(s_full(wg), uint32_t(j & 1));
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
      tc_fence_before();
      mbar_arrive(s_free(wg));                            // S_i(j) lives in registers: S_i(j+1) may be computed
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
        mbar_wait(pv_done(wg), uint32_t((j - 1) & 1));   // O_i quiescent (rare path: the first tiles of a row block)
        tc_fence_after();
        const float f = need ? fast_exp2((m_use - m_new) * c) : 1.0f;
#pragma unroll
        for (int cc = 0; cc < D_PAD; cc += 16) {
          uint32_t t[16];
          tmem_ld_x16(tO + cc, t);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * f);
          tmem_st_x16(tO + cc, t);
        }
        tmem_st_wait();
        l_run *= f;
      }
      if (need) m_use = m_new;
      const float nmc = -m_use * c;
      // P.V_i(j-1) must have read P_i before it is overwritten below.  It was issued a whole tile ago: probe the barrier now (the probe's
      // latency hides under the exponentials) and only fall back to a blocking wait if it really has not retired yet.
      const bool pv_ok = (j == 0) || mbar_test(pv_done(wg), uint32_t((j - 1) & 1));
      stamp(j, 3);
      // P = 2^((s - m_use) c) in fp32, row sum in four partial sums, packed to 16 bits for the MMA
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[64];
#pragma unroll
      for (int i = 0; i < 128; i += 2) {
        const float e0 = fast_exp2(fmaf(__uint_as_float(v[i]), c, nmc)), e1 = fast_exp2(fmaf(__uint_as_float(v[i + 1]), c, nmc));
        ls[(i >> 1) & 3] += e0 + e1;
        pk[i >> 1] = pack16(e0, e1, BF);
      }
      l_run += (ls[0] + ls[1]) + (ls[2] + ls[3]);
      stamp(j, 4);
      if (!pv_ok) mbar_wait(pv_done(wg), uint32_t((j - 1) & 1));
      tc_fence_after();
      stamp(j, 5);
#pragma unroll
      for (int cc = 0; cc < 64; cc += 16) {
        uint32_t (&p16)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[cc]);
        tmem_st_x16(tP + cc, p16);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full(wg));
      stamp(j, 6);
    }
    // ---- normalise and store
    mbar_wait(pv_done(wg), uint32_t((T - 1) & 1));
    tc_fence_after();
    const float inv = 1.f / l_run;
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
