// Probe: which TMA row-store configurations are legal on sm_100a (run each variant in its own process).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, int x0, int smem_off, int dims) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned short* s = reinterpret_cast<unsigned short*>(smem + smem_off);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s[i] = (unsigned short)(i + 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(s);
    if (dims == 4)
      asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)&tm), "r"(sa), "r"(x0), "r"(1), "r"(0), "r"(0) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)&tm), "r"(sa), "r"(x0), "r"(1) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
int main(int argc, char** argv) {
  int box = atoi(argv[1]), x0 = atoi(argv[2]), smem_off = atoi(argv[3]), dims = atoi(argv[4]), sw = atoi(argv[5]);
  cudaFree(0);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  unsigned short* d; cudaMalloc(&d, 256 * 8 * 2); cudaMemset(d, 0, 256 * 8 * 2);
  CUtensorMap tm;
  cuuint64_t gd[4] = {256, 8, 1, 1}; cuuint64_t gs[3] = {512, 4096, 4096}; cuuint32_t bx[4] = {(cuuint32_t)box, 1, 1, 1}; cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dims, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("box=%d x0=%d off=%d dims=%d sw=%d encode rc=%d ", box, x0, smem_off, dims, sw, (int)rc);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
  k<<<1, 64, 8192>>>(tm, x0, smem_off, dims);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned short h[256 * 2]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("sync=%s row1[%d..]= %d %d %d ... row1[%d]=%d\n", cudaGetErrorString(e), x0, h[256 + x0], h[256 + x0 + 1], h[256 + x0 + 2], x0 + box - 1, h[256 + x0 + box - 1]);
  return 0;
}
