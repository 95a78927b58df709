// Microbenchmark: what bounds the tcgen05 GEMM main loop on B200?  (build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17
//  -I consistentid_b200/csrc -o gpurun_out/microbench tools/microbench.cu ; run on the GPU box)
// Per CTA (one per SM): a STAGES-deep ring of {A 128x64, B BNx64} bf16 tiles.
//   mode 0  MMA only   : issue 4 x tcgen05.mma (M128, N=BN, K16) per k-block on resident smem, no TMA
//   mode 1  TMA only   : stream k-blocks from global memory through the ring, no MMA
//   mode 2  both, independent (MMA never waits for data, TMA never waits for MMA)
//   mode 3  the real dependent producer/consumer pipeline
// Prints cycles per k-block (max over CTAs) and the derived tensor utilisation / ingest bytes per clock.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cudaTypedefs.h>
#include "common.cuh"
using namespace cid;

template <int BN, int STAGES, int KPI>
__global__ void __launch_bounds__(128, 1) mb_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                     int mode, int nkb, int k_blocks_total, long long* out_cycles) {
  constexpr int A_BYTES = 128 * 128 * KPI, B_BYTES = BN * 128 * KPI, STAGE = A_BYTES + B_BYTES;   // a stage holds KPI k-blocks
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + STAGES * STAGE;
  auto full = [&](int s) { return bar + 8u * s; };
  auto empty = [&](int s) { return bar + 8u * (STAGES + s); };
  const uint32_t done = bar + 8u * (2 * STAGES);
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem_raw + (base - smem_u32(smem_raw)) + STAGES * STAGE + 8 * (2 * STAGES + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) {
    if (lane == 0) { for (int s = 0; s < STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); } mbar_init(done, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc<256>(smem_u32(const_cast<uint32_t*>(slot)));
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  const long long t0 = clock64();
  if (warp == 0 && lane == 0 && mode != 0) {            // TMA producer
    int stage = 0; uint32_t phase = 0;
    const int m_row = blockIdx.x * 128;
    for (int kb = 0; kb < nkb; ++kb) {
      if (mode == 3) mbar_wait(empty(stage), phase ^ 1u);
      else if (kb >= STAGES) mbar_wait(full(stage), phase ^ 1u);      // previous load into this stage landed
      mbar_expect_tx(full(stage), STAGE);
      const int kchunk = (kb * KPI) % k_blocks_total;
      tma_load_3d(base + stage * STAGE, &tmA, full(stage), 0, m_row, kchunk);             // box {64, 128, KPI}
      tma_load_3d(base + stage * STAGE + A_BYTES, &tmB, full(stage), 0, 0, kchunk);       // box {64, BN, KPI}
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
    // drain
    for (int s = 0; s < STAGES && s < nkb; ++s) { /* last phases complete on their own; waited below by MMA or here */ }
    if (mode == 1) {
      int st = 0; uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) { if (kb >= nkb - STAGES) mbar_wait(full(st), ph); if (++st == STAGES) { st = 0; ph ^= 1u; } }
    }
  } else if (warp == 1 && lane == 0 && mode != 1) {     // MMA issuer
    const uint32_t idesc = make_idesc(128, BN, 1);
    int stage = 0; uint32_t phase = 0;
    const uint32_t a0 = desc_lo(base);
    for (int kb = 0; kb < nkb; ++kb) {
      if (mode == 3) { mbar_wait(full(stage), phase); tc_fence_after(); }
      const uint32_t a_lo = a0 + stage * (STAGE / 16), b_lo = a_lo + A_BYTES / 16;
#pragma unroll
      for (int q = 0; q < KPI; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem, desc_make(a_lo + q * 1024 + k * 2), desc_make(b_lo + q * (BN * 8) + k * 2), idesc, (kb | k | q) ? 1u : 0u);
      if (mode == 3) umma_commit(empty(stage));
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
    umma_commit(done);
    mbar_wait(done, 0);
  }
  __syncwarp();                                            // reconverge before the block barrier (diverged bar.sync is UB)
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  if (warp == 1) { tc_fence_after(); tmem_dealloc<256>(tmem); }
}

static PFN_cuTensorMapEncodeTiled_v12000 enc;
static void map3d(CUtensorMap* m, void* p, long long K, long long rows, int box_rows, int kpi) {
  // view [rows, K] as {64 (contiguous), rows (pitch K), K/64 chunks (pitch 64)}: one instruction fetches kpi k-blocks
  cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)(K / 64)}; cuuint64_t str[2] = {(cuuint64_t)K * 2, 128};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)kpi}, es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, p, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r) { printf("encode failed %d\n", (int)r); exit(1); }
}
template <int BN, int STAGES, int KPI>
void run(void* A, void* B, int K, int sms, long long* d_out, const char* tag) {
  printf("-- %s, %d k-block(s) per TMA instruction\n", tag, KPI);
  CUtensorMap ta, tb; map3d(&ta, A, K, (long long)sms * 128, 128, KPI); map3d(&tb, B, K, 256, BN, KPI);
  const int smem = STAGES * KPI * (128 * 128 + BN * 128) + 1024 + 256;
  cudaFuncSetAttribute(mb_kernel<BN, STAGES, KPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int nkb = 2048 / KPI;
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 2; ++rep) mb_kernel<BN, STAGES, KPI><<<sms, 128, smem>>>(ta, tb, mode, nkb, K / 64, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e) { printf("mode %d failed: %s\n", mode, cudaGetErrorString(e)); exit(1); }
    std::vector<long long> h(sms); cudaMemcpy(h.data(), d_out, sms * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
    const double cyc = double(mx) / (nkb * KPI);      // per 64-wide k-block
    printf("BN=%3d stages=%d mode=%d  %.1f cycles/k-block  tensor-util=%.0f%% (ideal %d)  ingest=%.1f B/clk/SM\n", BN, STAGES, mode, cyc,
           100.0 * (2.0 * BN) / cyc, 2 * BN, (128.0 * 128 + BN * 128) / cyc);
  }
}
int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q); enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int K = 8192;                                                     // A: [sms*128, 8192] bf16 = 310 MB (streams from HBM)
  void *A, *B; long long* d_out;
  cudaMalloc(&A, (size_t)sms * 128 * K * 2); cudaMalloc(&B, (size_t)256 * K * 2); cudaMalloc(&d_out, sms * 8);
  cudaMemset(A, 0, (size_t)sms * 128 * K * 2); cudaMemset(B, 0, (size_t)256 * K * 2);
  printf("SMs %d\n", sms);
  run<160, 5, 1>(A, B, 512, sms, d_out, "A L2-resident (K=512: 19 MB total)");
  run<160, 2, 2>(A, B, 512, sms, d_out, "A L2-resident");
  run<160, 3, 2>(A, B, 512, sms, d_out, "A L2-resident");
  run<256, 4, 1>(A, B, 512, sms, d_out, "A L2-resident");
  run<256, 2, 2>(A, B, 512, sms, d_out, "A L2-resident");
  run<64, 2, 4>(A, B, 512, sms, d_out, "A L2-resident");
  run<160, 3, 2>(A, B, K, sms, d_out, "A streams from HBM (310 MB)");
  run<256, 2, 2>(A, B, K, sms, d_out, "A streams from HBM (310 MB)");
  return 0;
}
