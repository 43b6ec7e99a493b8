// minimal 2D TMA probe with variants (argv[1]): 0 plain, 1 + fence.proxy.async after barrier init, 2 elect.sync issue
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, int variant, float* out) {
  __shared__ __align__(1024) float s[8 * 32];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
    if (variant >= 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(8u * 32u * 4u) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_u32(s)), "l"(&tmap), "r"(0), "r"(0), "r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  } while (!ok);
  for (int i = threadIdx.x; i < 8 * 32; i += blockDim.x) out[i] = s[i];
}
typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                       const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int H = 64, W = 64;
  std::vector<float> h((size_t)H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o;
  CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMalloc(&o, 8 * 32 * 4));
  CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  alignas(64) CUtensorMap tm;
  const cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)H};
  const cuuint64_t gstr[1] = {(cuuint64_t)W * 4};
  const cuuint32_t box[2] = {32, 8}, es[2] = {1, 1};
  CUresult r = ((Fn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d encode rc=%d q=%d d=%p\n", variant, (int)r, (int)q, (void*)d);
  const unsigned long long* w = reinterpret_cast<const unsigned long long*>(&tm);
  for (int i = 0; i < 16; ++i) printf("%016llx%s", w[i], i % 4 == 3 ? "\n" : " ");
  k<<<1, 128>>>(tm, variant, o);
  CK(cudaDeviceSynchronize());
  std::vector<float> r2(8 * 32);
  CK(cudaMemcpy(r2.data(), o, r2.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int y = 0; y < 8; ++y) for (int x = 0; x < 32; ++x) if (r2[y * 32 + x] != h[y * W + x]) ++bad;
  printf("%d mismatches\n", bad);
  return 0;
}
