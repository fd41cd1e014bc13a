// Micro-measurements that size the attention kernels' softmax warps on B200 (one CTA on one SM, clock64 around the loop):
//   MUFU.EX2 issue rate per scheduler, a softmax-like instruction mix, tcgen05.ld (32x32b.x32) bandwidth per scheduler / SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I transformer4sed_b200/csrc -I include scripts/probe/sm_rates.cu -o scripts/probe/sm_rates
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "ptx.cuh"
using namespace t4s;

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// mode 0: 32 independent ex2 per iteration; mode 1: softmax mix per pair: FFMA2, 2 MUFU, FADD2, F2FP;  mode 2: mix without MUFU
template <int kMode>
__global__ void mufu_kernel(float* out, long long* cyc, int iters) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = -0.001f * (threadIdx.x + i);
  uint64_t acc = ptx::pack2(0.f, 0.f);
  uint32_t pk = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (kMode == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = ex2(v[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float a, b;
        ptx::unpack2(ptx::fma2(ptx::pack2(v[2 * i], v[2 * i + 1]), ptx::pack2(0.999f, 0.999f), ptx::pack2(-0.01f, -0.01f)), a, b);
        if (kMode == 1) { a = ex2(a); b = ex2(b); }
        acc = ptx::add2(acc, ptx::pack2(a, b));
        __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
        pk ^= *reinterpret_cast<uint32_t*>(&t);
        v[2 * i] = a; v[2 * i + 1] = b;
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f, a, b;
  ptx::unpack2(acc, a, b);
#pragma unroll
  for (int i = 0; i < 32; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + a + b + __uint_as_float(pk);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void ldtm_kernel(float* out, long long* cyc, int iters, int cols_per_iter) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  for (int c = 0; c < 512; c += 32) ptx::tmem_st_32x32(t_lane + c, v);
  ptx::tmem_st_wait();
  __syncthreads();
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int c = 0; c < cols_per_iter; c += 128) {
      uint32_t a[32], b[32], d[32], e[32];
      ptx::tmem_ld_32x32(t_lane + ((c) & 511), a);
      ptx::tmem_ld_32x32(t_lane + ((c + 32) & 511), b);
      ptx::tmem_ld_32x32(t_lane + ((c + 64) & 511), d);
      ptx::tmem_ld_32x32(t_lane + ((c + 96) & 511), e);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= a[i] ^ b[i] ^ d[i] ^ e[i];
    }
  }
  const long long t1 = clock64();
  out[threadIdx.x] = __uint_as_float(acc);
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {4, 8, 16, 32}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) mufu_kernel<0><<<1, warps * 32>>>(out, cyc, iters);
        if (mode == 1) mufu_kernel<1><<<1, warps * 32>>>(out, cyc, iters);
        if (mode == 2) mufu_kernel<2><<<1, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double per = (double)h / ((double)iters * 32.0 * (warps / 4.0));   // cycles per warp-level element-instruction slot per scheduler
      printf("mode %d (%s) warps/SM %2d: %lld cycles, %.2f cycles per warp-wide element per scheduler (32 elements/iter/warp)\n", mode,
             mode == 0 ? "ex2 only" : mode == 1 ? "softmax mix" : "mix without ex2", warps, h, per);
    }
  for (int warps : {1, 2, 4, 8, 16}) {
    for (int rep = 0; rep < 2; ++rep) { ldtm_kernel<<<1, warps * 32>>>(out, cyc, 500, 512); cudaDeviceSynchronize(); }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double bytes = 500.0 * 512 * 32 * 4 * warps;
    printf("tcgen05.ld 32x32b.x32, %2d warps: %lld cycles, %.1f B/clk per SM, %.1f B/clk per warp\n", warps, h, bytes / h, bytes / h / warps);
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
