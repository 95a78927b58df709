// Microbenchmark: per-SM throughput of the pieces of a flash-attention softmax on B200 (what bounds attn_self*).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I consistentid_b200/csrc -o gpurun_out/mb_softmax tools/microbench_softmax.cu
//   run on the GPU box: gpurun_out/mb_softmax  -> one line per (test, warps/SM): elements per clock per SM
// Tests (each warp runs ITER iterations of 8 independent chains, so throughput - not latency - is measured):
//   ex2_f32     ex2.approx.ftz.f32                       (1 result / lane-op)
//   ex2_h2      ex2.approx.f16x2                         (2 results / lane-op)
//   ex2_bf2     ex2.approx.ftz.bf16x2
//   cvt_h2      cvt.rn.f16x2.f32                         (2 results / lane-op)
//   cvt_ex2_h2  fma + cvt.rn.f16x2.f32 + ex2.approx.f16x2   (the current kernel's per-pair sequence)
//   poly_f32    exp2 by Cody-Waite + degree-3 polynomial on the FMA/ALU pipes, fp32 (FA4-style MUFU offload), result packed with cvt
//   poly_h2     the same in packed half2 arithmetic (HFMA2), 2 results / lane-op
//   mix         half of the pairs through ex2_h2, half through poly_f32
//   ldtm_x32    tcgen05.ld 32x32b.x32 back to back (4 per wait), bytes per clock per SM
//   ldtm_x128   tcgen05.ld 32x32b.x128
//   ldtm_16x256 tcgen05.ld 16x256b.x8 (same 32 registers per thread as x32)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda.h>
#include "common.cuh"
using namespace cid;

constexpr int ITER = 2048;

__device__ __forceinline__ float ex2_f32(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_bf2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t cvt_h2(float lo, float hi) { uint32_t y; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }

// 2^x for x <= 0 (softmax argument), fp32: x = n + f, f in [0,1); 2^f by a degree-3 minimax polynomial; exponent added as an integer
__device__ __forceinline__ float poly_exp2_f32(float x) {
  x = fmaxf(x, -126.0f);
  const float n = floorf(x);                         // FRND (ALU) - alternatives measured below use the magic-number trick
  const float f = x - n;
  float p = fmaf(f, 0.0555041f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (int(n) << 23));
}
// magic-number variant: n = round-to-nearest via adding 1.5 * 2^23, f in [-0.5, 0.5]
__device__ __forceinline__ float poly_exp2_f32_magic(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;                   // 1.5 * 2^23: low mantissa bits of t = round(x)
  const float n = t - 12582912.0f;
  const float f = x - n;                             // [-0.5, 0.5]
  float p = fmaf(f, 0.0555041f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int TEST>
__global__ void __launch_bounds__(1024, 1) alu_kernel(float* sink, long long* cycles, float seed) {
  const int lane = threadIdx.x & 31;
  float a[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed * float(i + lane) - 3.0f; h[i] = 0xB800B400u + i + lane; }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (TEST == 0) a[i] = ex2_f32(a[i]) - 1.5f;
      if (TEST == 1) h[i] = ex2_h2(h[i]) ^ 0x80008000u;
      if (TEST == 2) h[i] = ex2_bf2(h[i]) ^ 0x80008000u;
      if (TEST == 3) { h[i] = cvt_h2(a[i], a[i] + 1.0f); a[i] += __uint_as_float(h[i] & 0x3f800000u); }
      if (TEST == 4) { h[i] = ex2_h2(cvt_h2(fmaf(a[i], seed, -2.0f), fmaf(a[i], seed, -3.0f))); a[i] = a[i] * 0.999f + __uint_as_float(h[i] & 0x007f0000u); }
      if (TEST == 5) { const float p0 = poly_exp2_f32(fmaf(a[i], seed, -2.0f)), p1 = poly_exp2_f32(fmaf(a[i], seed, -3.0f)); h[i] = cvt_h2(p0, p1); a[i] = a[i] * 0.999f + __uint_as_float(h[i] & 0x007f0000u); }
      if (TEST == 6) { const float p0 = poly_exp2_f32_magic(fmaf(a[i], seed, -2.0f)), p1 = poly_exp2_f32_magic(fmaf(a[i], seed, -3.0f)); h[i] = cvt_h2(p0, p1); a[i] = a[i] * 0.999f + __uint_as_float(h[i] & 0x007f0000u); }
      if (TEST == 7) {   // mix: even chains MUFU packed, odd chains polynomial
        if (i & 1) { const float p0 = poly_exp2_f32_magic(fmaf(a[i], seed, -2.0f)), p1 = poly_exp2_f32_magic(fmaf(a[i], seed, -3.0f)); h[i] = cvt_h2(p0, p1); }
        else h[i] = ex2_h2(cvt_h2(fmaf(a[i], seed, -2.0f), fmaf(a[i], seed, -3.0f)));
        a[i] = a[i] * 0.999f + __uint_as_float(h[i] & 0x007f0000u);
      }
      if (TEST == 8) {   // mix 3:1 (every 4th chain polynomial)
        if ((i & 3) == 3) { const float p0 = poly_exp2_f32_magic(fmaf(a[i], seed, -2.0f)), p1 = poly_exp2_f32_magic(fmaf(a[i], seed, -3.0f)); h[i] = cvt_h2(p0, p1); }
        else h[i] = ex2_h2(cvt_h2(fmaf(a[i], seed, -2.0f), fmaf(a[i], seed, -3.0f)));
        a[i] = a[i] * 0.999f + __uint_as_float(h[i] & 0x007f0000u);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  if (s == 12345.678f) sink[0] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ void tmem_ld_x128(uint32_t taddr, uint32_t (&r)[128]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x128.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63, "
      "%64, %65, %66, %67, %68, %69, %70, %71, %72, %73, %74, %75, %76, %77, %78, %79, %80, %81, %82, %83, %84, %85, %86, %87, %88, %89, %90, %91, %92, %93, %94, %95, "
      "%96, %97, %98, %99, %100, %101, %102, %103, %104, %105, %106, %107, %108, %109, %110, %111, %112, %113, %114, %115, %116, %117, %118, %119, %120, %121, %122, %123, %124, %125, %126, %127}, [%128];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
        "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
        "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63]),
        "=r"(r[64]), "=r"(r[65]), "=r"(r[66]), "=r"(r[67]), "=r"(r[68]), "=r"(r[69]), "=r"(r[70]), "=r"(r[71]), "=r"(r[72]), "=r"(r[73]), "=r"(r[74]), "=r"(r[75]), "=r"(r[76]), "=r"(r[77]), "=r"(r[78]), "=r"(r[79]),
        "=r"(r[80]), "=r"(r[81]), "=r"(r[82]), "=r"(r[83]), "=r"(r[84]), "=r"(r[85]), "=r"(r[86]), "=r"(r[87]), "=r"(r[88]), "=r"(r[89]), "=r"(r[90]), "=r"(r[91]), "=r"(r[92]), "=r"(r[93]), "=r"(r[94]), "=r"(r[95]),
        "=r"(r[96]), "=r"(r[97]), "=r"(r[98]), "=r"(r[99]), "=r"(r[100]), "=r"(r[101]), "=r"(r[102]), "=r"(r[103]), "=r"(r[104]), "=r"(r[105]), "=r"(r[106]), "=r"(r[107]), "=r"(r[108]), "=r"(r[109]), "=r"(r[110]), "=r"(r[111]),
        "=r"(r[112]), "=r"(r[113]), "=r"(r[114]), "=r"(r[115]), "=r"(r[116]), "=r"(r[117]), "=r"(r[118]), "=r"(r[119]), "=r"(r[120]), "=r"(r[121]), "=r"(r[122]), "=r"(r[123]), "=r"(r[124]), "=r"(r[125]), "=r"(r[126]), "=r"(r[127])
      : "r"(taddr)
      : "memory");
}
// 16 lanes x 256 bits per "row", x8: thread t of the warp gets 32 registers covering lanes {t/4 ... } (layout irrelevant for throughput)
__device__ __forceinline__ void tmem_ld_16x256_x8(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// TEST 0: 4 x x32 per wait; 1: x128 per wait; 2: 16x256b.x8 x4 per wait.  nwarps in {4, 8, 16}: warp w reads lane quarter w % 4.
template <int TEST>
__global__ void __launch_bounds__(256, 1) ldtm_kernel(float* sink, long long* cycles, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot + (uint32_t((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (TEST == 0) {
      uint32_t v[128];
      uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
      uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
      uint32_t (&v2)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[64]);
      uint32_t (&v3)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[96]);
      tmem_ld_x32(tmem + 0, v0); tmem_ld_x32(tmem + 32, v1); tmem_ld_x32(tmem + 64, v2); tmem_ld_x32(tmem + 96, v3);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 128; i += 16) acc ^= v[i];
    } else if (TEST == 1) {
      uint32_t v[128];
      tmem_ld_x128(tmem, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 128; i += 16) acc ^= v[i];
    } else {
      uint32_t v[128];
      uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
      uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
      uint32_t (&v2)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[64]);
      uint32_t (&v3)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[96]);
      const uint32_t t16 = slot + (uint32_t((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
      tmem_ld_16x256_x8(t16 + 0, v0); tmem_ld_16x256_x8(t16 + 64, v1);
      tmem_ld_16x256_x8(t16 + (16u << 16), v2); tmem_ld_16x256_x8(t16 + (16u << 16) + 64, v3);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 128; i += 16) acc ^= v[i];
    }
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) sink[0] = 1.f;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(slot); }
}

template <typename F>
static double run(F launch, long long* d_cyc, int nblk) {
  launch(); launch();
  cudaDeviceSynchronize();
  std::vector<long long> h(nblk);
  cudaMemcpy(h.data(), d_cyc, nblk * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (auto c : h) mx = c > mx ? c : mx;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  return double(mx);
}

int main() {
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  float* sink; long long* cyc;
  cudaMalloc(&sink, 64); cudaMalloc(&cyc, nsm * sizeof(long long));
  const char* names[] = {"ex2_f32", "ex2_h2", "ex2_bf2", "cvt_h2", "fma+cvt+ex2_h2", "poly_f32(floor)+cvt", "poly_f32(magic)+cvt", "mix 1:1 mufu:poly", "mix 3:1 mufu:poly"};
  const int per_op[] = {1, 2, 2, 2, 2, 2, 2, 2, 2};
  for (int t = 0; t < 9; ++t)
    for (int warps : {4, 8, 16}) {
      double c = 0;
      auto L = [&](auto k) { c = run([&] { k<<<nsm, warps * 32, 0>>>(sink, cyc, 1.0001f); }, cyc, nsm); };
      switch (t) {
        case 0: L(alu_kernel<0>); break; case 1: L(alu_kernel<1>); break; case 2: L(alu_kernel<2>); break; case 3: L(alu_kernel<3>); break;
        case 4: L(alu_kernel<4>); break; case 5: L(alu_kernel<5>); break; case 6: L(alu_kernel<6>); break; case 7: L(alu_kernel<7>); break;
        case 8: L(alu_kernel<8>); break;
      }
      const double elems = double(ITER) * 8 * per_op[t] * warps * 32;
      printf("%-22s warps/SM %2d : %8.0f cycles  -> %6.2f results/clk/SM\n", names[t], warps, c, elems / c);
    }
  const char* lnames[] = {"ldtm 4 x 32x32b.x32", "ldtm 32x32b.x128", "ldtm 4 x 16x256b.x8"};
  for (int t = 0; t < 3; ++t)
    for (int warps : {4, 8}) {
      double c = 0;
      const int iters = 4096;
      auto L = [&](auto k) { c = run([&] { k<<<nsm, warps * 32, 0>>>(sink, cyc, iters); }, cyc, nsm); };
      switch (t) { case 0: L(ldtm_kernel<0>); break; case 1: L(ldtm_kernel<1>); break; case 2: L(ldtm_kernel<2>); break; }
      const double bytes = double(iters) * 128 * 4 * warps * 32;
      printf("%-22s warps/SM %2d : %8.0f cycles  -> %6.1f B/clk/SM  (a 128x128 fp32 score tile = %5.0f cycles)\n", lnames[t], warps, c, bytes / c, 65536.0 / (bytes / c));
    }
  return 0;
}
