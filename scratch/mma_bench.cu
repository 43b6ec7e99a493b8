// mma.sync m16n8k8 TF32 throughput / latency per SM on sm_100a (legacy tensor path), vs FFMA.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int ILP>
__global__ void k_mma(float* out, long long* cyc, int iters) {
  unsigned a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x * 3, threadIdx.x * 5};
  float d[ILP][4];
  for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) mma_tf32(d[i], a, b);
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void k_ffma(float* out, long long* cyc, int iters) {
  float a = threadIdx.x * 1e-3f, b = 1.0001f;
  float d[ILP];
  for (int i = 0; i < ILP; ++i) d[i] = i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) d[i] = fmaf(d[i], b, a);
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < ILP; ++i) s += d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 4096);
  const int iters = 2000;
  for (int warps : {1, 2, 4, 8, 16}) {
    long long h;
    k_mma<1><<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("mma tf32 m16n8k8  warps=%2d ILP=1: %.2f cyc/mma/warp (latency-bound), %.0f MAC/clk/SM\n", warps, (double)h / iters, warps * 1024.0 * iters / h);
    k_mma<8><<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("mma tf32 m16n8k8  warps=%2d ILP=8: %.2f cyc/mma/warp, %.0f MAC/clk/SM\n", warps, (double)h / iters / 8, warps * 8 * 1024.0 * iters / h);
    k_ffma<16><<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ffma              warps=%2d ILP=16: %.0f MAC/clk/SM\n", warps, warps * 16 * 32.0 * iters / h);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
