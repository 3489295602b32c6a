// Cost of the block barriers the control CTA uses (all warps arriving together): cycles per barrier.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/barrier tests/micro/barrier.cu && /tmp/barrier
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, int iters) {
  __shared__ int sink;
  long long t0, t1;
  // (a) non-aligned named barrier, 512 threads
  if (threadIdx.x < 512) {
    asm volatile("barrier.sync 1, 512;" ::: "memory");
    t0 = clock64();
    for (int i = 0; i < iters; ++i) asm volatile("barrier.sync 1, 512;" ::: "memory");
    t1 = clock64();
    if (threadIdx.x == 32) out[0] = t1 - t0;
  }
  __syncthreads();
  // (b) aligned __syncthreads, whole block (1024)
  t0 = clock64();
  for (int i = 0; i < iters; ++i) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 32) out[1] = t1 - t0;
  // (c) named barrier 128 threads (warps 1..4)
  if (threadIdx.x >= 32 && threadIdx.x < 160) {
    t0 = clock64();
    for (int i = 0; i < iters; ++i) asm volatile("barrier.sync 13, 128;" ::: "memory");
    t1 = clock64();
    if (threadIdx.x == 32) out[2] = t1 - t0;
  }
  __syncthreads();
  // (d) 512-thread barrier where warp 0 arrives with lane 0 only late (split warp): lanes 1..31 arrive first
  if (threadIdx.x < 512) {
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (threadIdx.x == 0) { sink = i; }
      asm volatile("barrier.sync 1, 512;" ::: "memory");
    }
    t1 = clock64();
    if (threadIdx.x == 32) out[3] = t1 - t0;
  }
  // (e) shared-memory atomicAdd same address from 40 threads + barrier
  __syncthreads();
  if (threadIdx.x < 512) {
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (threadIdx.x >= 32 && threadIdx.x < 72) atomicAdd(&sink, 1);
      asm volatile("barrier.sync 1, 512;" ::: "memory");
    }
    t1 = clock64();
    if (threadIdx.x == 32) out[4] = t1 - t0;
  }
  // (f) globaltimer read latency (dependent)
  if (threadIdx.x == 0) {
    unsigned long long a = 0, acc = 0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(a)); acc += a; }
    t1 = clock64();
    out[5] = t1 - t0; out[7] = (long long)acc;
  }
  // (g) fence.acq_rel.gpu with nothing outstanding
  if (threadIdx.x == 0) {
    t0 = clock64();
    for (int i = 0; i < iters; ++i) asm volatile("fence.acq_rel.gpu;" ::: "memory");
    t1 = clock64();
    out[6] = t1 - t0;
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  const int iters = 2000;
  k<<<1, 1024>>>(d, iters); cudaDeviceSynchronize();
  k<<<1, 1024>>>(d, iters); cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  const char* names[] = {"barrier.sync 1,512 (non-aligned)", "__syncthreads (1024)", "barrier.sync 13,128", "512 with thread-0 store before", "512 + 40 smem atomics same addr", "globaltimer read", "fence.acq_rel.gpu idle"};
  for (int i = 0; i < 7; ++i) printf("%-40s %.1f cycles\n", names[i], (double)h[i] / iters);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
