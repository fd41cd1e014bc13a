// Shared helpers for libt4s (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "t4s.h"

namespace t4s {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int sm_count();
// Bind the primary context to the calling thread (autograd worker threads may not have one yet): driver entry points
// such as cuTensorMapEncodeTiled fail with CUDA_ERROR_INVALID_CONTEXT otherwise.
void ensure_context();

#define T4S_CUDA(expr)                                                        \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) return ::t4s::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

extern long long g_launches;  // kernels launched by this library (bench.py reports it as gpu_launches)
#define T4S_LAUNCH_CHECK()          \
  do {                              \
    ++::t4s::g_launches;            \
    T4S_CUDA(cudaGetLastError());   \
  } while (0)

#define T4S_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::t4s::set_error(__VA_ARGS__);    \
      return T4S_ERR_ARG;               \
    }                                   \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

}  // namespace t4s
