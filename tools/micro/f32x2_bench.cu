// Microbenchmark: issue cost of the packed FP32 instructions of sm_100a (add/mul.rn.f32x2 -> FADD2 / FMUL2) against the
// scalar FADD / FMUL doing the same work.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 f32x2_bench.cu && ./a.out
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float addf(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float mulf(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

template <int PACKED>
__global__ void k(float* out, int iters, float s) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  u64 p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = ((u64)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
  u64 s2 = ((u64)__float_as_uint(s) << 32) | __float_as_uint(s);
  for (int it = 0; it < iters; ++it) {
    if (PACKED) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = mul2(add2(p[i], s2), s2);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = mulf(addf(a[i], s), s);
    }
  }
  float r = 0;
  if (PACKED) {
#pragma unroll
    for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int packed = 0; packed < 2; ++packed)
    for (int warps = 1; warps <= 8; warps *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (packed) k<1><<<148 * 4, 32 * warps>>>(out, iters, 1.0000001f);
        else k<0><<<148 * 4, 32 * warps>>>(out, iters, 1.0000001f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      double flops = 148.0 * 4 * 32 * warps * (double)iters * 32;   // 16 adds + 16 muls per iteration per thread
      printf("%s warps/CTA %d (4 CTAs/SM): %.3f ms  %.1f GFLOP/s  (%.2f scalar-op-equivalents per clk per SM at 1.965 GHz)\n",
             packed ? "FADD2/FMUL2" : "FADD /FMUL ", warps, ms, flops / ms / 1e6, flops / (ms * 1e-3) / 148 / 1.965e9);
    }
  return 0;
}
