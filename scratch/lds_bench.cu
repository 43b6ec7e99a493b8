// microbench: shared-memory load throughput for broadcast / distinct LDS.32 / LDS.128, and FFMA:LDS mixes
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  __shared__ __align__(16) float s[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = i * 0.001f;
  __syncthreads();
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  int lane = threadIdx.x & 31;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 32; ++u) {
      int base = ((it * 32 + u) * 4) & 4095;
      if (MODE == 0) {  // broadcast LDS.128
        float4 v = *reinterpret_cast<float4*>(&s[base]);
        a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
      } else if (MODE == 1) {  // distinct conflict-free LDS.128
        float4 v = *reinterpret_cast<float4*>(&s[(base + lane * 4) & 8191]);
        a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
      } else if (MODE == 2) {  // broadcast LDS.32
        a0 += s[base];
      } else if (MODE == 3) {  // distinct LDS.32
        a0 += s[(base + lane) & 8191];
      } else if (MODE == 4) {  // half-warp broadcast LDS.128 (2 distinct addrs)
        float4 v = *reinterpret_cast<float4*>(&s[(base + (lane >> 4) * 4) & 8191]);
        a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
      } else if (MODE == 5) {  // 8 distinct addrs (each 4 lanes share)
        float4 v = *reinterpret_cast<float4*>(&s[(base + (lane >> 2) * 4) & 8191]);
        a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
// FFMA : broadcast LDS.128 mix, R ffma per load
template <int R>
__global__ void kmix(float* out, int iters) {
  __shared__ __align__(16) float s[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = i * 0.001f;
  __syncthreads();
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.1f + i;
  float y0 = threadIdx.x * 0.5f, y1 = y0 + 1.f, y2 = y0 + 2.f, y3 = y0 + 3.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 32; ++u) {
      int base = ((it * 32 + u) * 4) & 4095;
      float4 v = *reinterpret_cast<float4*>(&s[base]);
      float w[4] = {v.x, v.y, v.z, v.w};
      float yy[4] = {y0, y1, y2, y3};
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r % 16] = fmaf(w[r & 3], yy[(r >> 2) & 3], acc[r % 16]);
    }
  }
  long long t1 = clock64();
  float t = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) t += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
template <typename F>
void run(const char* name, F f, int threads, float per_iter_instr) {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  int iters = 2000;
  f<<<148, threads>>>(d, iters);
  f<<<148, threads>>>(d, iters);
  cudaDeviceSynchronize();
  float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
  int warps = threads / 32;
  printf("%-40s threads=%4d cycles=%10.0f  cyc per warp-instr (per SM) = %.3f\n", name, threads, cyc,
         cyc / (iters * 32.0 * warps * per_iter_instr));
  cudaFree(d);
}
int main() {
  for (int th : {128, 256, 512, 1024}) {
    run("broadcast LDS.128", k<0>, th, 1);
    run("distinct LDS.128", k<1>, th, 1);
    run("broadcast LDS.32", k<2>, th, 1);
    run("distinct LDS.32", k<3>, th, 1);
    run("2-addr LDS.128", k<4>, th, 1);
    run("8-addr LDS.128", k<5>, th, 1);
    run("mix 4 FFMA per bcast LDS.128 (per ffma)", kmix<4>, th, 4);
    run("mix 8 FFMA per bcast LDS.128 (per ffma)", kmix<8>, th, 8);
    run("mix 16 FFMA per bcast LDS.128 (per ffma)", kmix<16>, th, 16);
  }
  return 0;
}
