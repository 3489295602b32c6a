// Cost of *_sync collectives when ptxas cannot prove the warp converged (it guards each one with BRA.DIV).
// nvcc -arch=sm_100a -O3 tests/micro/shfl_div.cu -o /tmp/shfl_div && /tmp/shfl_div
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_uniform(long long* out, float seed) {
  float f = seed + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) f = __shfl_xor_sync(0xffffffffu, f, 1) + 1.0f;
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)f; }
}
// warp-uniform at run time, but the condition depends on threadIdx.x
__global__ void k_guarded(long long* out, float seed, int which) {
  if ((int)(threadIdx.x >> 5) != which) return;
  float f = seed + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) f = __shfl_xor_sync(0xffffffffu, f, 1) + 1.0f;
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = (long long)f; }
}
// same, after lane 0 has run alone and the block met at a non-aligned barrier (the control CTA's pattern)
__global__ void k_split(long long* out, float seed, int which, int spin, int resync) {
  if ((int)(threadIdx.x >> 5) != which) return;
  float f = seed + threadIdx.x;
  if ((threadIdx.x & 31) == 0) { for (int i = 0; i < spin; ++i) f = f * 1.0001f + 0.5f; }
  asm volatile("barrier.sync 1, 32;" ::: "memory");
  if (resync) __syncwarp();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) f = __shfl_xor_sync(0xffffffffu, f, 1) + 1.0f;
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = (long long)f; }
}
int main() {
  long long* out; cudaMalloc(&out, 16); long long h[2];
  k_uniform<<<1, 32>>>(out, 3.f); k_uniform<<<1, 32>>>(out, 3.f); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("uniform            %.1f cycles/shuffle\n", h[0] / 256.0);
  k_guarded<<<1, 64>>>(out, 3.f, 0); k_guarded<<<1, 64>>>(out, 3.f, 0); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("threadIdx-guarded  %.1f cycles/shuffle\n", h[0] / 256.0);
  for (int resync = 0; resync < 2; ++resync) {
    k_split<<<1, 64>>>(out, 3.f, 0, 1000, resync); k_split<<<1, 64>>>(out, 3.f, 0, 1000, resync); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("after split+barrier (syncwarp=%d) %.1f cycles/shuffle\n", resync, h[0] / 256.0);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
