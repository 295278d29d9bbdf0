// Micro-benchmark: issue rate of MUFU.EX2 in fp32 vs the f16 / f16x2 / bf16x2 forms on sm_100a (does a packed ex2 halve the
// MUFU work of the softmax?).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mufu_bench tools/mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a0 = seed + threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  unsigned h0 = __float_as_uint(a0) & 0x3bff3bffu, h1 = h0 ^ 0x01000100u, h2 = h0 ^ 0x02000200u, h3 = h0 ^ 0x03000300u;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 1) {
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h0)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h2)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h3));
    } else if (MODE == 2) {
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h0)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h2)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h3));
    } else {
      unsigned short s0 = (unsigned short)h0, s1 = (unsigned short)h1, s2 = (unsigned short)h2, s3 = (unsigned short)h3;
      asm volatile("ex2.approx.f16 %0, %0;" : "+h"(s0)); asm volatile("ex2.approx.f16 %0, %0;" : "+h"(s1));
      asm volatile("ex2.approx.f16 %0, %0;" : "+h"(s2)); asm volatile("ex2.approx.f16 %0, %0;" : "+h"(s3));
      h0 = s0; h1 = s1; h2 = s2; h3 = s3;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(h0 ^ h1 ^ h2 ^ h3);
}
template <int MODE>
void run(const char* name, int per_instr) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000;
  k<MODE><<<148 * 8, 256>>>(out, 10, 0.5f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(out, iters, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double instr = 148.0 * 8 * 256 * iters * 4;
  printf("%-22s %8.3f ms  %7.2f G ptx-instr/s  %7.2f G exps/s (per SM per clk @1.9GHz: %.2f exps)\n", name, ms, instr / ms / 1e6,
         instr * per_instr / ms / 1e6, instr * per_instr / ms / 1e6 / 148 / 1.9);
  cudaFree(out);
}
int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("ex2.approx.f16", 1);
  return 0;
}
