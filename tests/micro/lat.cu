// Micro-benchmarks behind the dataflow design (run on the GPU box: nvcc -arch=sm_100a -O3 tests/micro/lat.cu -o /tmp/lat && /tmp/lat)
//  (1) dependent ld.cg chain in L2 (cycles per hop), idle GPU and with 147 CTAs polling one line
//  (2) producer -> consumers flag latency (ns), release store vs polling loads
//  (3) cost of fence.acq_rel.gpu / red.release in the consumer (cycles)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__global__ void chase(const unsigned* next, int hops, long long* out, unsigned* flag, int pollers) {
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      unsigned i = 0;
      long long t0 = clock64();
      for (int h = 0; h < hops; ++h) i = __ldcg(next + i);
      long long t1 = clock64();
      out[0] = t1 - t0; out[1] = i;
      st_release(flag, 1u);
    }
  } else if ((int)blockIdx.x <= pollers) {
    if (threadIdx.x == 0) while (ld_relaxed(flag) == 0u) { }
  }
}

// ping: CTA 0 publishes epochs; every other CTA polls, then (optionally fences), reads a payload line, bumps `done`
__global__ void ping(unsigned* flag, unsigned* done, unsigned* payload, int rounds, int fence_mode, unsigned long long* stats) {
  const int W = gridDim.x - 1;
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      unsigned long long tot = 0, tpub = 0;
      for (int r = 1; r <= rounds; ++r) {
        payload[threadIdx.x] = r;
        unsigned long long t0 = gt();
        st_release(flag, (unsigned)r);
        unsigned long long t1 = gt();
        while (ld_relaxed(done) < (unsigned)(W * r)) { }
        unsigned long long t2 = gt();
        tpub += t1 - t0; tot += t2 - t0;
      }
      stats[0] = tot / rounds; stats[1] = tpub / rounds;
    }
  } else {
    unsigned long long cyc_f = 0, cyc_l = 0, cyc_r = 0;
    for (int r = 1; r <= rounds; ++r) {
      __shared__ unsigned s_v;
      if (threadIdx.x == 0) {
        while (ld_relaxed(flag) < (unsigned)r) { }
        long long a = clock64();
        if (fence_mode) asm volatile("fence.acq_rel.gpu;" ::: "memory");
        long long b = clock64();
        unsigned v = __ldcg(payload);
        s_v = v;
        long long c = clock64();
        cyc_f += b - a; cyc_l += c - b + (v == 0xFFFFFFFFu);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        long long a = clock64();
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(done), "r"(1u) : "memory");
        long long b = clock64();
        cyc_r += b - a;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 1) { stats[2] = cyc_f / rounds; stats[3] = cyc_l / rounds; stats[4] = cyc_r / rounds; }
  }
}

int main() {
  const int n = 1 << 22;  // 16 MB chase table (L2 resident)
  unsigned* h = (unsigned*)malloc(n * 4);
  unsigned stride = 4099 * 32;   // jump by many lines
  for (int i = 0; i < n; ++i) h[i] = (unsigned)(((unsigned long long)i + stride) % n);
  unsigned *d, *flag; long long* out;
  cudaMalloc(&d, n * 4); cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&flag, 4096); cudaMalloc(&out, 64);
  for (int pollers : {0, 147}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(flag, 0, 4096);
      chase<<<148, 32>>>(d, 2000, out, flag, pollers);
      long long o[2]; cudaMemcpy(o, out, 16, cudaMemcpyDeviceToHost);
      printf("ld.cg chase (L2), %3d pollers: %.0f cycles/hop\n", pollers, (double)o[0] / 2000);
    }
  }
  unsigned long long* st; cudaMalloc(&st, 64);
  unsigned* payload; cudaMalloc(&payload, 4096);
  for (int fm : {0, 1}) {
    for (int threads : {32, 1024}) {
      cudaMemset(flag, 0, 4096); cudaMemset(flag + 256, 0, 4);
      ping<<<148, threads>>>(flag, flag + 256, payload, 2000, fm, st);
      unsigned long long s[5]; cudaMemcpy(s, st, 40, cudaMemcpyDeviceToHost);
      printf("ping 147 consumers x %4d thr, consumer fence=%d: round trip %llu ns (publish %llu ns); consumer cycles: fence %llu, payload ld.cg %llu, red.release %llu\n",
             threads, fm, s[0], s[1], s[2], s[3], s[4]);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
