// attn_cross2_kernel - the decoupled text + id cross-attention of Consistent_IPAttProcessor (attention.py:259-279), same operands and
// results as attn_cross_kernel (attn_cross.cuh), restructured as a PERSISTENT, PIPELINED kernel for head dims <= 80.
//
// attn_cross_kernel runs one CTA per (128-query tile, head, sample) and that CTA walks a strictly serial chain (TMA Q/K/V -> S MMA ->
// softmax -> P through shared memory -> P.V MMA -> epilogue): ~6 000 SM cycles per tile against ~800 cycles of exponentials and
// ~900 cycles of HBM time - 75-110 TF/s, 14 % of HBM (profiles/r02_shapes_*).  Here every SM keeps ONE CTA that owns a contiguous,
// balanced range of the launch's (sample, head, query tile) units and overlaps them:
//   * warp 0 streams Q tiles through a 4-slot ring and K_cat / V_cat^T through a 2-slot ring (reloaded only when (sample, head) changes);
//   * warp 1 issues S(u) = Q K_cat^T into TMEM buffer u % 3 three units ahead of the P.V stream (S(x+3) right after P.V(x));
//   * three softmax warpgroups (warps 2-5 / 6-9 / 10-13, started a third of a period apart) own one TMEM buffer each: pull the 96 score
//     columns into registers, two masked softmaxes (text keys [0, n_text), id keys [80, 80 + n_ip)), NORMALISE in registers
//     (p / l_text, ip_scale * p / l_ip - both sums are known before P is written), write P as packed 16-bit pairs over the first 48 score
//     columns (TMEM is the A operand of the P.V MMAs - no shared-memory round trip, no CTA-wide barrier); ONE accumulator
//     O = P'_text V_text + P'_ip V_ip then needs no mixing epilogue: drain, convert, store.
//     (attention.py:264,276-279 rounds each branch to 16 bits before the mix; here the 16-bit rounding sits on the normalised
//     probabilities instead - same order of error, inside the parity tolerance of the processor goldens.)
// TMEM (3 x 160 columns): per buffer [S 96 (P aliases 0..47) | O D_PAD]; at D_PAD = 80 O starts at column 48 (over the dead half of S) and the
// next S of that buffer waits until O has been drained.
#pragma once
#include "attn_common.cuh"
#include "attn_tc7.cuh"          // tmem_st_x16 / tmem_st_wait

namespace cid {

constexpr int CROSS2_WGS = 3;
constexpr int CROSS2_THREADS = 64 + 128 * CROSS2_WGS;
#ifndef CID_CROSS2_STAGGER
#define CID_CROSS2_STAGGER 800
#endif
constexpr long long CROSS2_STAGGER_CYCLES = CID_CROSS2_STAGGER;
constexpr int CROSS_IP_OFF = 80;                       // first id key row of K_cat / V_cat (cid_pack_cross_kv, cid_attn_cross: ip_off is always 80)

template <int D_PAD>
struct Cross2Cfg {
  static_assert(D_PAD % 16 == 0 && D_PAD <= 80, "attn_cross2 covers head dims <= 80");
  static constexpr int NCH = (D_PAD + 63) / 64;
  static constexpr int NQ = (NCH == 1) ? 8 : 4;                  // Q ring slots: the S stream runs 3 units ahead, the loads must run further ahead still
  static constexpr int KROWS = 96;
  static constexpr int Q_BYTES = NCH * 16384;
  static constexpr int K_CHUNK = KROWS * 128;
  static constexpr int K_BYTES = NCH * K_CHUNK;
  static constexpr int V_CHUNK = D_PAD * 128;
  static constexpr int V_BYTES = 2 * V_CHUNK;
  static constexpr int OFF_K = NQ * Q_BYTES;
  static constexpr int OFF_V = OFF_K + 2 * K_BYTES;
  static constexpr int OFF_BAR = OFF_V + 2 * V_BYTES;
  static constexpr int TOTAL = OFF_BAR + 512;
  static constexpr bool ALIAS = D_PAD > 64;                      // O over the upper half of S (dead once P is written)
  static constexpr int TM_BUF = 160, TM_O = ALIAS ? 48 : 96;
  static_assert(TM_O + D_PAD <= TM_BUF && CROSS2_WGS * TM_BUF <= 512, "TMEM budget");
  static_assert(K_CHUNK % 1024 == 0 && V_CHUNK % 1024 == 0, "128B-swizzle atoms need 1 KB aligned chunks");
};

template <int D_PAD, int BF>
__global__ void __launch_bounds__(CROSS2_THREADS, 1)
attn_cross2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmVt, const AttnArgs a) {
  using C = Cross2Cfg<D_PAD>;
  constexpr int NQ = C::NQ;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if ((sbase & 1023u) != 0) { if (threadIdx.x == 0) printf("cid: cross-attn smem base not 1024-aligned\n"); __trap(); }
  const uint32_t bar0 = sbase + C::OFF_BAR;
  auto q_full = [&](int s) { return bar0 + 8u * s; };
  auto q_free = [&](int s) { return bar0 + 8u * (NQ + s); };
  auto kv_full = [&](int s) { return bar0 + 8u * (2 * NQ + s); };
  auto kv_free = [&](int s) { return bar0 + 8u * (2 * NQ + 2 + s); };
  constexpr int NB = CROSS2_WGS;
  auto s_full = [&](int j) { return bar0 + 8u * (2 * NQ + 4 + j); };            // S(u) complete (MMA commit)
  auto p_ready = [&](int j) { return bar0 + 8u * (2 * NQ + 4 + NB + j); };      // P(u) in TMEM (128 softmax threads)
  auto o_full = [&](int j) { return bar0 + 8u * (2 * NQ + 4 + 2 * NB + j); };   // P.V(u) retired (MMA commit)
  auto o_free = [&](int j) { return bar0 + 8u * (2 * NQ + 4 + 3 * NB + j); };   // O(u) pulled into registers (128 softmax threads)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + C::OFF_BAR + 8 * (2 * NQ + 4 + 4 * NB));
  static_assert(8 * (2 * NQ + 5 + 4 * NB) <= 512, "barrier block");

  const int warp = warp_id(), lane = lane_id();
  // this CTA's contiguous range of units; unit g = (sample * H + head) * tiles + query tile
  const int tiles = (a.Nq + 127) / 128;
  const long long U = (long long)a.B * a.H * tiles;
  const int g_beg = int(U * blockIdx.x / gridDim.x), g_end = int(U * (blockIdx.x + 1) / gridDim.x);

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmVt); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NQ; ++s) { mbar_init(q_full(s), 1); mbar_init(q_free(s), 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_free(s), 1); }
      for (int s = 0; s < NB; ++s) { mbar_init(s_full(s), 1); mbar_init(p_ready(s), 128); mbar_init(o_full(s), 1); mbar_init(o_free(s), 128); }
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
    // ============================================================ TMA producer
    int n_kv = -1;                 // index of the current K/V group (one per (sample, head) run inside the range)
    int uq = 0;
    for (int g = g_beg; g < g_end; ++g, ++uq) {
      const int bh = g / tiles, t = g - bh * tiles;
      const int b = bh / a.H, h = bh - b * a.H;
      if (g == g_beg || t == 0) {
        ++n_kv;
        const int s = n_kv & 1;
        mbar_wait(kv_free(s), uint32_t(((n_kv >> 1) & 1) ^ 1));
        const uint32_t kb = kv_full(s), kd = sbase + C::OFF_K + s * C::K_BYTES, vd = sbase + C::OFF_V + s * C::V_BYTES;
        if (elect_one()) {
          mbar_expect_tx(kb, C::K_BYTES + C::V_BYTES);
#pragma unroll
          for (int ch = 0; ch < C::NCH; ++ch) tma_load_4d(kd + ch * C::K_CHUNK, &tmK, kb, ch * 64, 0, h, b);
#pragma unroll
          for (int kc = 0; kc < 2; ++kc) tma_load_3d(vd + kc * C::V_CHUNK, &tmVt, kb, kc * 64, 0, bh);
        }
        __syncwarp();
      }
      const int slot = uq % NQ;
      mbar_wait(q_free(slot), uint32_t(((uq / NQ) & 1) ^ 1));
      const uint32_t qb = q_full(slot), qd = sbase + slot * C::Q_BYTES;
      if (elect_one()) {
        mbar_expect_tx(qb, C::Q_BYTES);
#pragma unroll
        for (int ch = 0; ch < C::NCH; ++ch) tma_load_4d(qd + ch * 16384, &tmQ, qb, ch * 64, t * 128, h, b);
      }
      __syncwarp();
    }
    if (elect_one()) griddep_launch_dependents();
    __syncwarp();
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    const uint32_t idesc_s = make_idesc(128, C::KROWS, BF);
    const uint32_t idesc_pv = make_idesc(128, D_PAD, BF);
    // P.V of unit x (its K/V slot kvs; `last`: the K/V group ends with it): six 16-key steps into the unit's single accumulator
    auto issue_PV = [&](int x, int kvs, bool last) {
      const int j = x % NB;
      const uint32_t par = uint32_t((x / NB) & 1);
      mbar_wait(p_ready(j), par);
      if (!C::ALIAS) mbar_wait(o_free(j), par ^ 1u);       // O(x - NB) drained (aliased layout: already waited for before S(x))
      tc_fence_after();
      const uint32_t tb = tmem + j * C::TM_BUF;
      const uint32_t sv = sbase + C::OFF_V + kvs * C::V_BYTES;
      const uint32_t ob = o_full(j), fb = kv_free(kvs);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 6; ++ks)
          umma_ts(tb + C::TM_O, tb + ks * 8, make_desc_sw128(sv + (ks >> 2) * C::V_CHUNK + (ks & 3) * 32), idesc_pv, ks ? 1u : 0u);
        umma_commit(ob);
        if (last) umma_commit(fb);
      }
      __syncwarp();
    };
#ifdef CID_ATTN_TRACE
    const bool trm = a.trace != nullptr && blockIdx.x < 10 && lane == 0;
    auto mstamp = [&](int u_, int e) { if (trm && u_ < 64) a.trace[((size_t)(32 + blockIdx.x) * 64 + u_) * 8 + e] = clock64(); };
#else
    auto mstamp = [&](int, int) {};
#endif
    // Two cursors over the unit range: the S stream runs NB units AHEAD of the P.V stream - S(x + NB) is issued right after P.V(x) (the
    // first moment its TMEM buffer is free), so a warpgroup finds its next scores ready when it has drained O(x).  (With S(u) issued only
    // after P.V(u-2) the warpgroups waited 1 100-2 500 cycles per unit for scores, profiles/r02_trace_cross2_s_late_*.txt.)
    // Cursor: g = unit, t = query tile inside its (sample, head) run (advanced without divisions), nkv = index of its K/V group.
    struct Cur { int g, t, nkv; };
    auto advance = [&](Cur& cu) { ++cu.g; if (++cu.t == tiles) { cu.t = 0; } if (cu.t == 0) ++cu.nkv; };
    auto issue_S = [&](const Cur& cu, int u) {
      if (cu.g == g_beg || cu.t == 0) mbar_wait(kv_full(cu.nkv & 1), uint32_t((cu.nkv >> 1) & 1));      // first unit of a K/V group
      const int kvs = cu.nkv & 1;
      const int slot = u % NQ;
      mbar_wait(q_full(slot), uint32_t((u / NQ) & 1));
      const int j = u % NB;
      if (C::ALIAS) mbar_wait(o_free(j), uint32_t(((u / NB) & 1) ^ 1));                                  // S(u) overwrites O(u - NB)
      tc_fence_after();
      const uint32_t d_tm = tmem + j * C::TM_BUF, sq = sbase + slot * C::Q_BYTES, sk = sbase + C::OFF_K + kvs * C::K_BYTES;
      const uint32_t sb = s_full(j), qb = q_free(slot);
      if (elect_one()) {
#pragma unroll
        for (int ch = 0; ch < C::NCH; ++ch) {
          const int ksteps = (D_PAD - ch * 64 >= 64) ? 4 : (D_PAD - ch * 64) / 16;
#pragma unroll
          for (int kk = 0; kk < ksteps; ++kk)
            umma_ss(d_tm, make_desc_sw128(sq + ch * 16384 + kk * 32), make_desc_sw128(sk + ch * C::K_CHUNK + kk * 32), idesc_s, (ch | kk) ? 1u : 0u);
        }
        umma_commit(sb);
        umma_commit(qb);
      }
      __syncwarp();
    };
    const int n_units = g_end - g_beg;
    Cur cs{g_beg, g_beg % tiles, 0}, cp = cs;
    int us = 0;
    for (; us < n_units && us < NB; ++us) { issue_S(cs, us); advance(cs); }
    for (int x = 0; x < n_units; ++x) {
      mstamp(x, 0);
      issue_PV(x, cp.nkv & 1, (cp.g + 1 == g_end) || (cp.t + 1 == tiles));
      advance(cp);
      mstamp(x, 1);
      if (us < n_units) { issue_S(cs, us); advance(cs); ++us; }
      mstamp(x, 2);
      mstamp(x, 3);
    }
  } else {
    // ============================================================ softmax warpgroups (warpgroup w: units w, w + 3, ...)
    const int wg = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    const uint32_t tS = tmem + wg * C::TM_BUF + lane_off;
    const float c = a.scale_log2;
    const int t_end = a.n_text, i_end = CROSS_IP_OFF + a.n_ip;
    const bool has_ip = a.n_ip > 0;
    int k = 0;                                            // uses of this warpgroup's buffer so far
    // Staggered start: the warpgroups would otherwise receive their first scores together and stay in lock-step - exponentials (MUFU)
    // contended, then all of them draining O with the MUFU idle (profiles/r02_trace_cross2_inphase_*.txt: 2 300-cycle exp phases at two
    // warpgroups).  A third of a period of head start each makes one warpgroup's exponentials run under the others' P.V wait and drain.
    if (CROSS2_STAGGER_CYCLES > 0 && wg > 0) {
      const long long t0 = clock64();
      while (clock64() - t0 < CROSS2_STAGGER_CYCLES * wg) { }
    }
#ifdef CID_ATTN_TRACE
    const bool tr = a.trace != nullptr && blockIdx.x < 10 && quarter == 0 && lane == 0;
    auto stamp = [&](int k_, int e) { if (tr && k_ < 64) a.trace[((size_t)(blockIdx.x * 3 + wg) * 64 + k_) * 8 + e] = clock64(); };
#else
    auto stamp = [&](int, int) {};
#endif
    for (int g = g_beg + wg; g < g_end; g += NB, ++k) {
      stamp(k, 0);
      const int bh = g / tiles, t = g - bh * tiles;
      const int b = bh / a.H, h = bh - b * a.H;
      const uint32_t par = uint32_t(k & 1);
      mbar_wait(s_full(wg), par);
      tc_fence_after();
      stamp(k, 1);
      uint32_t v[96];
      {
        uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
        uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
        uint32_t (&v2)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[64]);
        tmem_ld_x32(tS + 0, v0);
        tmem_ld_x32(tS + 32, v1);
        tmem_ld_x32(tS + 64, v2);
        tmem_ld_wait();
      }
      stamp(k, 2);
      // Both masked softmaxes BRANCH-FREE (selects on the key index): with `if (key < n_text) ...` inside the unrolled loops nvcc emits a
      // real branch + reconvergence pair per element - measured 35 / 54 cycles per element in the max / exp passes
      // (profiles/r02_trace_cross2_branchy_*.txt), 8 600 cycles per unit where the MUFU needs ~800.
      // mask ONCE (padding keys -> -inf: they drop out of the maxima and 2^-inf = 0), then plain max / exp loops
      const bool no_ip = !has_ip;
#pragma unroll
      for (int i = 0; i < 96; ++i) {
        const bool valid = (i < CROSS_IP_OFF) ? (i < t_end) : (i < t_end || i < i_end);
        v[i] = valid ? v[i] : 0xff800000u;
      }
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, mi = -INFINITY;
#pragma unroll
      for (int i = 0; i < CROSS_IP_OFF; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[i]));
#pragma unroll
      for (int i = CROSS_IP_OFF; i < 96; ++i) mi = fmaxf(mi, __uint_as_float(v[i]));
      // keys 80..95 are id keys, or - without id tokens - more text keys (then they share the text maximum and sum)
      const float mt = fmaxf(fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])), no_ip ? mi : -INFINITY);
      const float nmt = -mt * c, nmi = no_ip ? nmt : -mi * c;
      stamp(k, 3);
      float l4[4] = {0.f, 0.f, 0.f, 0.f}, li = 0.f;
#pragma unroll
      for (int key = 0; key < CROSS_IP_OFF; ++key) {
        const float pe = fast_exp2(fmaf(__uint_as_float(v[key]), c, nmt));
        l4[key & 3] += pe;
        v[key] = __float_as_uint(pe);
      }
#pragma unroll
      for (int key = CROSS_IP_OFF; key < 96; ++key) {
        const float pe = fast_exp2(fmaf(__uint_as_float(v[key]), c, nmi));
        li += pe;
        v[key] = __float_as_uint(pe);
      }
      if (no_ip) { l4[0] += li; li = 0.f; }
      const float lt = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      // normalised probabilities: text keys / l_text, id keys * ip_scale / l_ip -> one accumulator, no mixing epilogue
      const float wt = 1.f / lt, wi = has_ip ? a.ip_scale / li : wt;
      uint32_t pk[48];
#pragma unroll
      for (int i = 0; i < 96; i += 2) {
        const float w = (i < CROSS_IP_OFF) ? wt : wi;
        pk[i >> 1] = pack16(__uint_as_float(v[i]) * w, __uint_as_float(v[i + 1]) * w, BF);
      }
      stamp(k, 4);
#pragma unroll
      for (int cc = 0; cc < 48; cc += 16) {
        uint32_t (&p16)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[cc]);
        tmem_st_x16(tS + cc, p16);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready(wg));
      stamp(k, 5);
      // ---- drain O, convert, store
      const int row = t * 128 + r;
      const bool row_ok = row < a.Nq;
      uint16_t* dst = reinterpret_cast<uint16_t*>(a.O) + ((long long)b * a.Nq + row) * a.ldo + h * a.d;
      mbar_wait(o_full(wg), par);
      tc_fence_after();
      stamp(k, 6);
      uint32_t o[D_PAD];
#pragma unroll
      for (int cc = 0; cc < D_PAD; cc += 16) {
        uint32_t (&o16)[16] = *reinterpret_cast<uint32_t (*)[16]>(&o[cc]);
        tmem_ld_x16(tS + C::TM_O + cc, o16);
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_free(wg));                             // the accumulator is in registers: the buffer may be reused
      if (row_ok) {
#pragma unroll
        for (int cc = 0; cc < D_PAD; cc += 8) {
          if (cc < a.d) {
            uint4 q4;
            q4.x = pack16(__uint_as_float(o[cc + 0]), __uint_as_float(o[cc + 1]), BF); q4.y = pack16(__uint_as_float(o[cc + 2]), __uint_as_float(o[cc + 3]), BF);
            q4.z = pack16(__uint_as_float(o[cc + 4]), __uint_as_float(o[cc + 5]), BF); q4.w = pack16(__uint_as_float(o[cc + 6]), __uint_as_float(o[cc + 7]), BF);
            *reinterpret_cast<uint4*>(dst + cc) = q4;
          }
        }
      }
      stamp(k, 7);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

}  // namespace cid
