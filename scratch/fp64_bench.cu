#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dadd(double* out, int iters, double x) {
  double a = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { a = a + x; a = a * x; }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) out[blockDim.x] = (double)(t1 - t0) / (2.0 * iters);
}
__global__ void k_fadd(float* out, int iters, float x) {
  float a = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { a = a + x; a = a * x; }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) out[blockDim.x] = (float)(t1 - t0) / (2.0f * iters);
}
__global__ void k_shfl_dadd(double* out, int iters) {
  double a = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) out[blockDim.x] = (double)(t1 - t0) / (5.0 * iters);
}
int main() {
  double* d; cudaMalloc(&d, 4096 * 8); float* f; cudaMalloc(&f, 4096 * 4);
  for (int th : {32, 128, 512}) {
    k_dadd<<<1, th>>>(d, 1000, 1.0000001); cudaDeviceSynchronize();
    double r; cudaMemcpy(&r, d + th, 8, cudaMemcpyDeviceToHost);
    k_fadd<<<1, th>>>(f, 1000, 1.0000001f); cudaDeviceSynchronize();
    float rf; cudaMemcpy(&rf, f + th, 4, cudaMemcpyDeviceToHost);
    k_shfl_dadd<<<1, th>>>(d, 200); cudaDeviceSynchronize();
    double rs; cudaMemcpy(&rs, d + th, 8, cudaMemcpyDeviceToHost);
    printf("threads=%d  dependent DADD/DMUL: %.1f cyc/op   FADD/FMUL: %.1f cyc/op   SHFL(64b)+DADD level: %.1f cyc\n", th, r, rf, rs);
  }
  return 0;
}
