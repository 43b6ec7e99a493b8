// TMA probe: one cp.async.bulk.tensor.3d of a [C][10][20] box (halo of an 8x16 tile) out of a [B*C][H][W] fp32 tensor, with
// negative start coordinates (zero fill = the perception's zero halo).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
constexpr int C = 32, BH = 10;
#ifndef BW
#define BW 20
#endif
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, const CUtensorMap* gmap, int x, int y, int z, float* out) {
  extern __shared__ __align__(128) unsigned char raw[];
  float* s = reinterpret_cast<float*>(raw);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"((uint32_t)(C * BH * BW * 4)) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(s)), "l"(gmap ? gmap : &tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  } while (!ok);
  for (int i = threadIdx.x; i < C * BH * BW; i += blockDim.x) out[i] = s[i];
}
typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                       const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int B = 3, H = 64, W = 64;
  std::vector<float> h((size_t)B * C * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003) * 0.001f + 1.f;
  float *d, *o;
  CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMalloc(&o, C * BH * BW * 4));
  CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  CUtensorMap tm;
  const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * C};
  const cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  const cuuint32_t box[3] = {BW, BH, C}, es[3] = {1, 1, 1};
  CUresult r = ((Fn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d BW %d encode rc=%d\n", variant, BW, (int)r);
  CUtensorMap* gm; CK(cudaMalloc(&gm, sizeof(CUtensorMap))); CK(cudaMemcpy(gm, &tm, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C * BH * BW * 4));
  int bad_total = 0;
  const int cases[4][3] = {{-1, -1, 0}, {15, 7, 32}, {47, 55, 64}, {31, -1, 32}};
  for (auto& cs : cases) {
    k<<<1, 128, C * BH * BW * 4>>>(tm, variant == 1 ? gm : nullptr, cs[0], cs[1], cs[2], o);
    CK(cudaDeviceSynchronize());
    std::vector<float> r2(C * BH * BW);
    CK(cudaMemcpy(r2.data(), o, r2.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int c = 0; c < C; ++c) for (int yy = 0; yy < BH; ++yy) for (int xx = 0; xx < BW; ++xx) {
      const int gx = cs[0] + xx, gy = cs[1] + yy, gz = cs[2] + c;
      const float want = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0.f : h[((size_t)gz * H + gy) * W + gx];
      if (r2[(c * BH + yy) * BW + xx] != want) ++bad;
    }
    printf("box at (x=%d, y=%d, z=%d): %d mismatches of %d\n", cs[0], cs[1], cs[2], bad, C * BH * BW);
    bad_total += bad;
  }
  printf(bad_total ? "FAIL\n" : "TMA probe OK\n");
  return 0;
}
