// Micro-benchmarks for the numbers the fused-attention design depends on (dev tool, B200):
//   1. ex2.approx.ftz.f32 vs ex2.approx.ftz.f16x2 issue rate per SM sub-partition (MUFU)
//   2. tcgen05.ld 32x32b.x32 read bandwidth per SM with 4 / 8 / 16 warps reading
//   3. FFMA / FADD / F2FP.PACK rate for reference
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/ubench tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../owl_vit_object_detection_b200/csrc/ptx.cuh"

using namespace owl;

constexpr int ITERS = 512;

template <int MODE>
__global__ void alu_kernel(float* out, long long* clk, float seed) {
  // 8 independent chains per thread so that latency never limits the rate
  float a[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + i); h[i] = __float_as_uint(a[i]) & 0x3bff3bffu; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(seed));
      if (MODE == 3) asm volatile("{.reg .b32 t; cvt.rn.f16x2.f32 t, %0, %0; mov.b32 %0, t;}" : "+f"(a[i]));
      if (MODE == 4) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 5) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

__global__ void tmem_ld_kernel(float* out, long long* clk, int reps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 128); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t r[32];
  float acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < reps; ++it) {
    tmem_ld32(base + (it & 3) * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 8) acc += __uint_as_float(r[i]);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 128); }
}

int main() {
  float* out;
  long long* clk;
  cudaMalloc(&out, 1 << 24);
  cudaMalloc(&clk, 1 << 16);
  long long h[1024];
  const char* names[] = {"ex2.approx.ftz.f32", "ex2.approx.ftz.bf16x2", "fma.rn.f32", "cvt.rn.f16x2.f32", "ex2.approx.f16x2", "tanh.approx.f32"};
  for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp *= 2) {
    const int threads = 128 * warps_per_smsp;   // one CTA per SM, warps spread over the 4 sub-partitions
    for (int mode = 0; mode < 6; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        switch (mode) {
          case 0: alu_kernel<0><<<148, threads>>>(out, clk, 0.001f); break;
          case 1: alu_kernel<1><<<148, threads>>>(out, clk, 0.001f); break;
          case 2: alu_kernel<2><<<148, threads>>>(out, clk, 0.001f); break;
          case 3: alu_kernel<3><<<148, threads>>>(out, clk, 0.001f); break;
          case 4: alu_kernel<4><<<148, threads>>>(out, clk, 0.001f); break;
          case 5: alu_kernel<5><<<148, threads>>>(out, clk, 0.001f); break;
        }
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, clk, 148 * 8, cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += h[i];
      avg /= 148;
      const double instr_per_smsp = 1.0 * ITERS * 8 * warps_per_smsp;   // warp-instructions issued on one sub-partition
      printf("%-22s %d warp(s)/SMSP: %.2f clk per warp-instruction per SMSP\n", names[mode], warps_per_smsp, avg / instr_per_smsp);
    }
  }
  for (int ctas = 1; ctas <= 4; ctas *= 2) {
    const int reps = 2048;
    for (int rep = 0; rep < 2; ++rep) { tmem_ld_kernel<<<148 * ctas, 128>>>(out, clk, reps); cudaDeviceSynchronize(); }
    cudaError_t e = cudaGetLastError();
    cudaMemcpy(h, clk, 148 * ctas * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148 * ctas; ++i) avg += h[i];
    avg /= 148 * ctas;
    // every warp reads 32 lanes x 32 columns x 4 B = 4 KB per iteration; ctas * 4 warps per SM
    printf("tcgen05.ld.32x32b.x32, %d CTA(s) of 4 warps per SM: %.1f clk per load per warp -> %.1f B/clk/SM (%s)\n", ctas,
           avg / reps, 4096.0 * 4 * ctas / (avg / reps), cudaGetErrorString(e));
  }
  return 0;
}
