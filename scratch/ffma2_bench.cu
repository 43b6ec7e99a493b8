// Does packed fp32 FMA (fma.rn.f32x2, __ffma2_rn, sm_100) shorten the register-tiled layer-1 loop of the resident kernels?
// Shape of k_rep_fwd's 2-cell tile: per k-step one LDS.128 of weights (4 hidden units of the lane), the tile's two y values,
// 8 dependent FFMA chains.  A: 8 FFMA + LDS.128 + LDS.64.  B: 4 FFMA2 + LDS.128 + LDS.128 (y stored duplicated so that the
// broadcast operand is a register pair straight from the load).  16 warps per SM (4 per scheduler), 128 SMs busy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/ffma2_bench scratch/ffma2_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
constexpr int K = 48, HID = 128, REP = 400;
__global__ void __launch_bounds__(512) k_a(const float* w, const float* y, float* out, long long* cyc) {
  __shared__ __align__(16) float sW[K * HID];
  __shared__ __align__(16) float sY[16][K * 2];
  for (int i = threadIdx.x; i < K * HID; i += 512) sW[i] = w[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < K * 2; i += 32) sY[warp][i] = y[i];
  __syncthreads();
  float acc[2][4] = {};
  const long long t0 = clock64();
  for (int r = 0; r < REP; ++r) {
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
      const float4 ww = *reinterpret_cast<const float4*>(sW + k * HID + 4 * lane);
      const float2 yy = *reinterpret_cast<const float2*>(&sY[warp][k * 2]);
      acc[0][0] = fmaf(yy.x, ww.x, acc[0][0]); acc[0][1] = fmaf(yy.x, ww.y, acc[0][1]);
      acc[0][2] = fmaf(yy.x, ww.z, acc[0][2]); acc[0][3] = fmaf(yy.x, ww.w, acc[0][3]);
      acc[1][0] = fmaf(yy.y, ww.x, acc[1][0]); acc[1][1] = fmaf(yy.y, ww.y, acc[1][1]);
      acc[1][2] = fmaf(yy.y, ww.z, acc[1][2]); acc[1][3] = fmaf(yy.y, ww.w, acc[1][3]);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * 512 + threadIdx.x] = acc[0][0] + acc[0][1] + acc[0][2] + acc[0][3] + acc[1][0] + acc[1][1] + acc[1][2] + acc[1][3];
}
__global__ void __launch_bounds__(512) k_b(const float* w, const float* y, float* out, long long* cyc) {
  __shared__ __align__(16) float sW[K * HID];
  __shared__ __align__(16) float sY[16][K * 4];                 // (y0, y0, y1, y1) per k
  for (int i = threadIdx.x; i < K * HID; i += 512) sW[i] = w[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < K * 4; i += 32) sY[warp][i] = y[(i >> 2) * 2 + ((i >> 1) & 1)];
  __syncthreads();
  float2 acc[2][2] = {};
  const long long t0 = clock64();
  for (int r = 0; r < REP; ++r) {
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
      const float4 ww = *reinterpret_cast<const float4*>(sW + k * HID + 4 * lane);
      const float4 yy = *reinterpret_cast<const float4*>(&sY[warp][k * 4]);
      const float2 w01 = make_float2(ww.x, ww.y), w23 = make_float2(ww.z, ww.w);
      const float2 y0 = make_float2(yy.x, yy.y), y1 = make_float2(yy.z, yy.w);
      acc[0][0] = __ffma2_rn(y0, w01, acc[0][0]); acc[0][1] = __ffma2_rn(y0, w23, acc[0][1]);
      acc[1][0] = __ffma2_rn(y1, w01, acc[1][0]); acc[1][1] = __ffma2_rn(y1, w23, acc[1][1]);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * 512 + threadIdx.x] = acc[0][0].x + acc[0][0].y + acc[0][1].x + acc[0][1].y + acc[1][0].x + acc[1][0].y + acc[1][1].x + acc[1][1].y;
}
int main() {
  float *w, *y, *o; long long* c;
  cudaMalloc(&w, K * HID * 4); cudaMalloc(&y, K * 2 * 4); cudaMalloc(&o, 148 * 512 * 4); cudaMalloc(&c, 148 * 8);
  cudaMemset(w, 0, K * HID * 4); cudaMemset(y, 0, K * 2 * 4);
  long long h[2];
  for (int it = 0; it < 2; ++it) {
    k_a<<<128, 512>>>(w, y, o, c); cudaDeviceSynchronize(); cudaMemcpy(&h[0], c, 8, cudaMemcpyDeviceToHost);
    k_b<<<128, 512>>>(w, y, o, c); cudaDeviceSynchronize(); cudaMemcpy(&h[1], c, 8, cudaMemcpyDeviceToHost);
  }
  printf("layer-1 loop of a 2-cell tile, 16 warps per SM: FFMA %.1f cycles per k-step-tile, FFMA2 %.1f  (ratio %.2f)  err %s\n",
         (double)h[0] / (REP * K), (double)h[1] / (REP * K), (double)h[0] / h[1], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
