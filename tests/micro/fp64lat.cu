// Dependent-chain latency (cycles/op) of the scalar ops the control phase uses, one warp on one SM.
// nvcc -arch=sm_100a -O3 tests/micro/fp64lat.cu -o /tmp/fp64lat && /tmp/fp64lat
#include <cstdio>
#include <cuda_runtime.h>
#define CHAIN(name, init, body)                                                     \
  __global__ void k_##name(long long* out, double seed) {                           \
    double x = seed + threadIdx.x; float f = (float)seed + threadIdx.x; unsigned long long u = (unsigned long long)seed + threadIdx.x; \
    init;                                                                           \
    long long t0 = clock64();                                                       \
    _Pragma("unroll 1") for (int i = 0; i < 256; ++i) { body; }                     \
    long long t1 = clock64();                                                       \
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)x + (long long)f + (long long)u; } \
  }
CHAIN(dadd, , x = __dadd_rn(x, 1.25))
CHAIN(dmul, , x = __dmul_rn(x, 1.0000001))
CHAIN(dfma, , x = __fma_rn(x, 1.0000001, 0.5))
CHAIN(ddiv, , x = __ddiv_rn(x, 1.0000001))
CHAIN(dsqrt, , x = __dsqrt_rn(x) + 3.0)
CHAIN(d2f2d, , f = (float)x; x = (double)f + 1.0)
CHAIN(f2ull, , u = __float2ull_rn(f) + 1ull; f = (float)(u & 1023ull) + 0.5f)
CHAIN(d2ll, , u = (unsigned long long)__double2ll_rn(x); x = (double)(long long)(u & 1023ull) + 0.5)
CHAIN(drint, , x = rint(x * 0.999) + 0.25)
CHAIN(ffma, , f = __fmaf_rn(f, 1.0001f, 0.5f))
CHAIN(fdiv, , f = __fdiv_rn(f, 1.0001f))
CHAIN(imad, , u = u * 3ull + 1ull)
CHAIN(mulhi64, , u = __umul64hi(u, 0x9E3779B97F4A7C15ull) + 1ull)
CHAIN(shfl, , f = __shfl_xor_sync(0xffffffffu, f, 1) + 1.0f)
CHAIN(redux, unsigned r = threadIdx.x, r = __reduce_add_sync(0xffffffffu, r) + 1u; f = (float)r)
CHAIN(lds, __shared__ unsigned s[64]; s[threadIdx.x] = (threadIdx.x + 1) & 31; s[threadIdx.x + 32] = threadIdx.x & 31; __syncwarp(); unsigned r = threadIdx.x, r = s[r]; f = (float)r)
int main() {
  long long* out; cudaMalloc(&out, 16);
  long long h[2];
#define RUN(name) k_##name<<<1, 32>>>(out, 3.0); k_##name<<<1, 32>>>(out, 3.0); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); printf("%-8s %7.1f cycles/iter\n", #name, h[0] / 256.0);
  RUN(dadd) RUN(dmul) RUN(dfma) RUN(ddiv) RUN(dsqrt) RUN(d2f2d) RUN(f2ull) RUN(d2ll) RUN(drint) RUN(ffma) RUN(fdiv) RUN(imad) RUN(mulhi64) RUN(shfl) RUN(redux) RUN(lds)
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
