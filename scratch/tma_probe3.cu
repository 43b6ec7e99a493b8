// 3D TMA probe, variants by argv[1]: 0: box {32,8,1} coords 0 | 1: box {20,10,32} coords (16,8,0) | 2: same, coords (-1,-1,0)
// | 3: as 1 with L2 promotion 128B | 4: as 1, dynamic smem manually aligned to 128
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, int dyn, int x, int y, int z, uint32_t bytes, float* out) {
  __shared__ __align__(1024) float sbuf[32 * 10 * 24];
  extern __shared__ unsigned char dynraw[];
  __shared__ __align__(8) uint64_t bar;
  float* s = dyn ? reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(dynraw) + 127) & ~(uintptr_t)127) : sbuf;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(s)), "l"(&tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  } while (!ok);
  for (int i = threadIdx.x; i < (int)(bytes / 4); i += blockDim.x) out[i] = s[i];
}
typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                       const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int B = 3, C = 32, H = 64, W = 64;
  std::vector<float> h((size_t)B * C * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003) + 1.f;
  float *d, *o;
  CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMalloc(&o, 32 * 10 * 24 * 4));
  CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  alignas(64) CUtensorMap tm;
  const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * C};
  const cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  cuuint32_t box[3] = {32, 8, 1}; const cuuint32_t es[3] = {1, 1, 1};
  int x = 0, y = 0, z = 0;
  if (variant >= 1) { box[0] = 20; box[1] = 10; box[2] = 32; x = 16; y = 8; }
  if (variant == 2) { x = -1; y = -1; }
  if (variant >= 5) { box[0] = 24; x = -4; y = -1; }
  if (variant == 6) { x = 60; y = 57; z = 64; }
  CUresult r = ((Fn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, variant == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const uint32_t bytes = box[0] * box[1] * box[2] * 4;
  printf("variant %d encode rc=%d bytes %u\n", variant, (int)r, bytes);
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 10 * 20 * 4 + 256));
  k<<<1, 128, variant == 4 ? 32 * 10 * 20 * 4 + 256 : 0>>>(tm, variant == 4, x, y, z, bytes, o);
  CK(cudaDeviceSynchronize());
  std::vector<float> r2(bytes / 4);
  CK(cudaMemcpy(r2.data(), o, bytes, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int c = 0; c < (int)box[2]; ++c) for (int yy = 0; yy < (int)box[1]; ++yy) for (int xx = 0; xx < (int)box[0]; ++xx) {
    const int gx = x + xx, gy = y + yy, gz = z + c;
    const float want = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0.f : h[((size_t)gz * H + gy) * W + gx];
    if (r2[(c * box[1] + yy) * box[0] + xx] != want) ++bad;
  }
  printf("%d mismatches\n", bad);
  return 0;
}
