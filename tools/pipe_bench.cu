// Micro-benchmark of the SM pipes the fused edge kernel leans on (MUFU.TANH / EX2, FFMA, HFMA2.BF16):
// gives the measured denominators for the "sfu" roofline in bench.py / DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench tools/pipe_bench.cu && ./pipe_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, int iters, float seed) {
  float v[8];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = seed + threadIdx.x * 1e-3f + i; u[i] = __float_as_uint(v[i]); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u[i]));
      if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(v[i]) : "f"(seed));
      if (OP == 5) asm volatile("fma.rn.bf16x2 %0, %0, %1, %0;" : "+r"(u[i]) : "r"(0x3f803f80u));
      if (OP == 7) { unsigned long long r = ((unsigned long long)u[i] << 32) | u[(i + 1) & 7];
                     asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(r) : "l"(0x3f8000003f800000ull));
                     u[i] = (uint32_t)(r >> 32); }
      if (OP == 6) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));            // mixed: 1 MUFU + 3 FMA
                     asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(v[i]) : "f"(seed));
                     asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(v[i]) : "f"(seed));
                     asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(v[i]) : "f"(seed)); }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i] + __uint_as_float(u[i]);
  if (s == 12345.678f) out[0] = s;
}

template <int OP>
void run(const char* name, double ops_per_inner, int sms) {
  float* out; cudaMalloc(&out, 4);
  const int iters = 4096, blocks = sms * 2;
  k<OP><<<blocks, 1024>>>(out, 16, 1.0f);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<OP><<<blocks, 1024>>>(out, iters, 1.0f);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double total = (double)blocks * 1024 * iters * 8 * ops_per_inner;
  double per_s = total / (ms * 1e-3);
  printf("%-28s %8.3f ms  %10.2f Gop/s  = %6.2f op/clk/SM @1965MHz\n", name, ms, per_s / 1e9, per_s / sms / 1.965e9);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, sms);
  run<0>("tanh.approx.f32", 1, sms);
  run<1>("ex2.approx.f32", 1, sms);
  run<2>("rcp.approx.f32", 1, sms);
  run<3>("tanh.approx.bf16x2 (x2)", 2, sms);
  run<4>("fma.f32", 1, sms);
  run<5>("fma.bf16x2 (x2)", 2, sms);
  run<6>("tanh.f32 + 3 fma (per tanh)", 1, sms);
  run<7>("fma.f32x2 (x2, +2 int ops)", 2, sms);
  return 0;
}
