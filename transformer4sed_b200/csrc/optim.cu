// K9 and the parameter-side kernels of a training step: fused AdamW over a flat fp32 parameter arena (optionally refreshing
// the bf16 GEMM-operand shadow in the same pass), multi-tensor gradient pack into the flat all-reduce buffer, EMA teacher.
// Replaces torch.optim.AdamW (reference recipes/desed/setting.py:254-258), update_ema (src/utils/scheduler.py:125-130) and
// nn.DataParallel's reduce_add_coalesced: one process per GPU, one NCCL all-reduce over the packed buffer.
#include <algorithm>

#include "common.cuh"

namespace t4s {
namespace optim {

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             __nv_bfloat16* __restrict__ shadow, size_t n, float lr, float b1, float b2, float eps, float wd, float bc1,
                             float bc2_sqrt, float grad_scale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float pi = p[i];
    pi *= 1.0f - lr * wd;  // decoupled weight decay (torch.optim.AdamW)
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * mi / denom;
    p[i] = pi;
    if (shadow) shadow[i] = __float2bfloat16_rn(pi);
  }
}

// teacher = alpha * teacher + (1 - alpha) * student
__global__ void ema_kernel(float* __restrict__ teacher, const float* __restrict__ student, __nv_bfloat16* __restrict__ shadow, size_t n,
                           float alpha) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float t = alpha * teacher[i] + (1.0f - alpha) * student[i];
    teacher[i] = t;
    if (shadow) shadow[i] = __float2bfloat16_rn(t);
  }
}

struct PackEntry {
  const float* src;  // NULL -> zero fill
  long long offset;  // element offset in the flat buffer
  long long n;
};

// one block-row per tensor chunk: table lives in global memory
__global__ void pack_kernel(const PackEntry* __restrict__ table, int n_entries, float* __restrict__ flat, int chunks_per_entry) {
  const int e = blockIdx.x / chunks_per_entry, c = blockIdx.x % chunks_per_entry;
  if (e >= n_entries) return;
  const PackEntry ent = table[e];
  const long long per = (ent.n + chunks_per_entry - 1) / chunks_per_entry;
  const long long lo = c * per, hi = min(ent.n, lo + per);
  float* dst = flat + ent.offset;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = ent.src ? ent.src[i] : 0.f;
}

}  // namespace optim
}  // namespace t4s

extern "C" {

int t4s_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* bf16_shadow, size_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
  T4S_REQUIRE(params && grads && exp_avg && exp_avg_sq && step >= 1, "t4s_adamw_step: bad arguments");
  if (n == 0) return T4S_OK;
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)t4s::sm_count() * 16);
  t4s::optim::adamw_kernel<<<grid, 256, 0, t4s::as_stream(stream)>>>(params, grads, exp_avg, exp_avg_sq, static_cast<__nv_bfloat16*>(bf16_shadow), n, lr,
                                                                    beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), grad_scale);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_ema_update(float* teacher, const float* student, void* bf16_shadow, size_t n, float alpha, void* stream) {
  T4S_REQUIRE(teacher && student, "t4s_ema_update: null pointer");
  if (n == 0) return T4S_OK;
  const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)t4s::sm_count() * 16);
  t4s::optim::ema_kernel<<<grid, 256, 0, t4s::as_stream(stream)>>>(teacher, student, static_cast<__nv_bfloat16*>(bf16_shadow), n, alpha);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

/* table: n_entries x {const float* src, int64 offset, int64 n} in DEVICE memory */
int t4s_grad_pack(const void* table, int n_entries, float* flat, void* stream) {
  T4S_REQUIRE(table && flat && n_entries > 0, "t4s_grad_pack: bad arguments");
  const int chunks = 8;
  t4s::optim::pack_kernel<<<n_entries * chunks, 256, 0, t4s::as_stream(stream)>>>(static_cast<const t4s::optim::PackEntry*>(table), n_entries, flat, chunks);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

}  // extern "C"
