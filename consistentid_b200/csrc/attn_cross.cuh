// Decoupled cross-attention of Consistent_IPAttProcessor (attention.py:259-279); layouts: attn_common.cuh.
#pragma once
#include "attn_common.cuh"

namespace cid {

// ------------------------------------------------------------------------------------------------------------------
// Decoupled text + id cross-attention, one 96-row key/value tile.
template <int D_PAD>
struct CrossCfg {
  static constexpr int NCH = (D_PAD + 63) / 64;
  static constexpr int KROWS = 96;
  static constexpr int Q_BYTES = NCH * 16384;
  static constexpr int K_CHUNK = KROWS * 128;
  static constexpr int K_BYTES = NCH * K_CHUNK;
  static constexpr int V_CHUNK = D_PAD * 128;
  static constexpr int V_BYTES = 2 * V_CHUNK;
  static constexpr int OFF_K = Q_BYTES;
  static constexpr int OFF_V = OFF_K + K_BYTES;
  // P (A operand of the second MMAs) re-uses memory that is dead once S = Q K^T has been read: chunk 0 (keys 0-63) lives
  // in Q's first 16 KB, chunk 1 (keys 64-95) in Q's second chunk when the head dim has one, else in its own 16 KB.
  // O_text / O_ip accumulators likewise overwrite the S columns of TMEM.  -> 56-60 KB smem, 128 TMEM columns for
  // head dims <= 64: three CTAs per SM instead of two for this latency-bound kernel.
  static constexpr int OFF_P0 = 0;
  static constexpr int OFF_P1 = (NCH >= 2) ? 16384 : OFF_V + V_BYTES;
  static constexpr int OFF_BAR = (NCH >= 2) ? OFF_V + V_BYTES : OFF_V + V_BYTES + 16384;
  static constexpr int TOTAL = OFF_BAR + 64;
  static constexpr int TMEM_COLS = (2 * D_PAD <= 128) ? 128 : (2 * D_PAD <= 256) ? 256 : 512;
  static constexpr int MIN_CTAS = (TMEM_COLS == 128) ? 3 : (TMEM_COLS == 256) ? 2 : 1;
};

template <int D_PAD>
__global__ void __launch_bounds__(128, CrossCfg<D_PAD>::MIN_CTAS)
attn_cross_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmVt, const AttnArgs a) {
  using C = CrossCfg<D_PAD>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if ((sbase & 1023u) != 0) { if (threadIdx.x == 0) printf("cid: cross-attn smem base not 1024-aligned\n"); __trap(); }
  const uint32_t bar_ld = sbase + C::OFF_BAR, bar_s = bar_ld + 8, bar_o = bar_ld + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + C::OFF_BAR + 32);
  const int warp = warp_id(), lane = lane_id();
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;

  if (warp == 0) {
    if (lane == 0) {
      mbar_init(bar_ld, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<C::TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_S = tmem, tmem_Ot = tmem, tmem_Oi = tmem + D_PAD;      // O_* overwrite S after the softmax has read it
  griddep_wait();                  // PDL: the prologue above overlaps the predecessor's tail

  // warp 0 (converged; single-thread instructions under elect_one(), see gemm_tc2.cuh) loads Q, K_cat, V_cat^T and issues S = Q K_cat^T
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_ld, C::Q_BYTES + C::K_BYTES + C::V_BYTES);
#pragma unroll
      for (int ch = 0; ch < C::NCH; ++ch) {
        tma_load_4d(sbase + ch * 16384, &tmQ, bar_ld, ch * 64, q0, h, b);
        tma_load_4d(sbase + C::OFF_K + ch * C::K_CHUNK, &tmK, bar_ld, ch * 64, 0, h, b);
      }
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) tma_load_3d(sbase + C::OFF_V + kc * C::V_CHUNK, &tmVt, bar_ld, kc * 64, 0, b * a.H + h);
      griddep_launch_dependents();
    }
    __syncwarp();
    mbar_wait(bar_ld, 0);
    tc_fence_after();
    const uint32_t idesc_s = make_idesc(128, C::KROWS, a.is_bf16);
    if (elect_one()) {
#pragma unroll
      for (int ch = 0; ch < C::NCH; ++ch) {
        const int ksteps = (D_PAD - ch * 64 >= 64) ? 4 : (D_PAD - ch * 64) / 16;
#pragma unroll
        for (int kk = 0; kk < ksteps; ++kk)
          umma_ss(tmem_S, make_desc_sw128(sbase + ch * 16384 + kk * 32),
                  make_desc_sw128(sbase + C::OFF_K + ch * C::K_CHUNK + kk * 32), idesc_s, (ch | kk) ? 1u : 0u);
      }
      umma_commit(bar_s);
    }
    __syncwarp();
  }

  const int r = warp * 32 + lane;
  const uint32_t lane_off = uint32_t(warp * 32) << 16;
  const int bf = a.is_bf16;
  const float c = a.scale_log2;
  const int t_end = a.n_text, i_beg = a.ip_off, i_end = a.ip_off + a.n_ip;
  mbar_wait(bar_s, 0);
  tc_fence_after();
  float mt = -INFINITY, mi = -INFINITY;
#pragma unroll 1
  for (int cc = 0; cc < C::KROWS; cc += 32) {
    uint32_t v[32];
    tmem_ld_x32(tmem_S + lane_off + cc, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      // selects, not branches: a branch per element costs ~30 cycles here (profiles/r02_trace_cross2_branchy_*.txt)
      const int key = cc + i; const float s = __uint_as_float(v[i]);
      mt = fmaxf(mt, key < t_end ? s : -INFINITY);
      mi = fmaxf(mi, (key >= i_beg && key < i_end) ? s : -INFINITY);
    }
  }
  float lt = 0.f, li = 0.f;
  uint8_t* sP0 = smem_raw + C::OFF_P0;
  uint8_t* sP1 = smem_raw + C::OFF_P1;
#pragma unroll 1
  for (int cc = 0; cc < C::KROWS; cc += 32) {
    uint32_t v[32];
    tmem_ld_x32(tmem_S + lane_off + cc, v);
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float p[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = cc + i + e; const float s = __uint_as_float(v[i + e]);
        const bool is_t = key < t_end, is_i = !is_t && key >= i_beg && key < i_end;
        const float x = (s - (is_t ? mt : mi)) * c;
        const float pe = fast_exp2((is_t || is_i) ? x : -INFINITY);      // 2^-inf = 0 for the padding keys
        lt += is_t ? pe : 0.f;
        li += is_i ? pe : 0.f;
        p[e] = pe;
      }
      pk[i >> 1] = pack16(p[0], p[1], bf);
    }
    uint8_t* tile = (cc >> 6) ? sP1 : sP0;
    const int col = cc & 63;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      st_sw128(tile, r, col + q * 8, make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]));
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    const uint32_t idesc_pv = make_idesc(128, D_PAD, a.is_bf16);
    const uint32_t sp0 = sbase + C::OFF_P0, sp1 = sbase + C::OFF_P1, sv = sbase + C::OFF_V;
    // text range: keys [0, 80) = 5 k-steps (rows 77..79 of K_cat / V_cat are zero padding, P is 0 there); with no id tokens
    // (plain cross-attention, e.g. ControlNet's default processor over all 81 rows) the text range spans all 96 rows
    const bool has_ip = a.n_ip > 0;
    if (elect_one()) {
#pragma unroll
      for (int ks = 0; ks < 5; ++ks) {
        const int kc = ks >> 2, kk = ks & 3;
        umma_ss(tmem_Ot, make_desc_sw128((kc ? sp1 : sp0) + kk * 32), make_desc_sw128(sv + kc * C::V_CHUNK + kk * 32), idesc_pv, ks ? 1u : 0u);
      }
      // keys [80, 96) = k-step 5 (chunk 1, second 16-key slice): the id range (own accumulator) or, without id tokens, more text keys
      umma_ss(has_ip ? tmem_Oi : tmem_Ot, make_desc_sw128(sp1 + 32), make_desc_sw128(sv + C::V_CHUNK + 32), idesc_pv, has_ip ? 0u : 1u);
      umma_commit(bar_o);
    }
    __syncwarp();
  }
  mbar_wait(bar_o, 0);
  tc_fence_after();
  const float wt = 1.f / lt;
  const float wi = (a.n_ip > 0) ? a.ip_scale / li : 0.f;
  const bool row_ok = q0 + r < a.Nq;
  uint16_t* dst = reinterpret_cast<uint16_t*>(a.O) + ((long long)b * a.Nq + q0 + r) * a.ldo + h * a.d;
#pragma unroll 1
  for (int cc = 0; cc < D_PAD; cc += 16) {
    uint32_t vt[16], vi[16];
    tmem_ld_x16(tmem_Ot + lane_off + cc, vt);
    tmem_ld_x16(tmem_Oi + lane_off + cc, vi);
    tmem_ld_wait();
    if (row_ok && cc < a.d) {
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        // reference rounds each branch to 16-bit before the mix (attention.py:264,276-279)
        float2 tt = unpack16(pack16(__uint_as_float(vt[i]) * wt, 0.f, bf), bf);
        float2 ii = unpack16(pack16(__uint_as_float(vi[i]) * (1.f / li), 0.f, bf), bf);
        f[i] = tt.x + ((a.n_ip > 0) ? a.ip_scale * ii.x : 0.f);
      }
      (void)wi;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (cc + q * 8 < a.d) {
          uint4 u;
          u.x = pack16(f[q * 8 + 0], f[q * 8 + 1], bf); u.y = pack16(f[q * 8 + 2], f[q * 8 + 3], bf);
          u.z = pack16(f[q * 8 + 4], f[q * 8 + 5], bf); u.w = pack16(f[q * 8 + 6], f[q * 8 + 7], bf);
          *reinterpret_cast<uint4*>(dst + cc + q * 8) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<C::TMEM_COLS>(tmem); }
}

}  // namespace cid
