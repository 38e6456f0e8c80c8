// ubench.cu — per-SM instruction throughput probes on B200 that decide the softmax design of the
// item-attention kernel: MUFU ex2 (f32 / f16x2 / bf16x2), packed f32x2 FMA, 3-input max, and a
// degree-3 polynomial exp2 on the FMA pipe.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o tools/ubench tools/ubench.cu ; run on the GPU box.  Prints results per clock per SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int MODE>
__global__ void __launch_bounds__(1024) probe(float* out, float seed) {
  float a[ILP];
  uint32_t u[ILP];
  unsigned long long w[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    a[i] = seed * (threadIdx.x + i) * 1e-3f - 1.0f;
    u[i] = __float_as_uint(a[i]) & 0x3c003c00u;
    w[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f);
  }
  const unsigned long long c2 = ((unsigned long long)__float_as_uint(0.999f) << 32) | __float_as_uint(0.998f);
  const unsigned long long c3 = ((unsigned long long)__float_as_uint(-1e-4f) << 32) | __float_as_uint(-2e-4f);
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) {          // ex2.f32
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      } else if (MODE == 1) {   // ex2.f16x2
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      } else if (MODE == 2) {   // ex2.bf16x2
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
      } else if (MODE == 3) {   // fma.f32x2
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[i]) : "l"(c2), "l"(c3));
      } else if (MODE == 4) {   // scalar ffma
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(0.999f), "f"(-1e-4f));
      } else if (MODE == 5) {   // 3-input max
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) % ILP]), "f"(seed));
      } else if (MODE == 6) {   // 2-input max
        asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed));
      } else if (MODE == 7) {   // cvt pack f32x2 -> bf16x2
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(__uint_as_float(u[i])));
      } else if (MODE == 8) {   // cvt pack f32x2 -> f16x2
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(__uint_as_float(u[i])));
      } else if (MODE == 9) {   // add.f32x2
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(c3));
      } else if (MODE == 10) {  // scalar polynomial exp2 (deg 3): floor via magic add, Horner, exponent splice
        float x = fmaxf(a[i], -125.f);
        float r = __fadd_rd(x, 12582912.f);           // 1.5 * 2^23: integer part in the low mantissa bits
        float fl = r - 12582912.f;
        float f = x - fl;
        float p = fmaf(f, 0.0555054f, 0.2402265f);
        p = fmaf(p, f, 0.6931472f);
        p = fmaf(p, f, 1.0f);
        a[i] = __uint_as_float(__float_as_uint(p) + (__float_as_uint(r) << 23)) - 1.5f;
      } else if (MODE == 11) {  // f16x2 fma (HFMA2)
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(0x3c003c00u), "r"(0x00010001u));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i] + __uint_as_float(u[i]) + __uint_as_float((uint32_t)w[i]) + __uint_as_float((uint32_t)(w[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int elems_per_op, float* out) {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int blocks = sms * 2, threads = 1024;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<MODE><<<blocks, threads>>>(out, 0.37f);
  cudaDeviceSynchronize();
  float best = 1e9f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    probe<MODE><<<blocks, threads>>>(out, 0.37f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double ops = (double)blocks * threads * ITERS * ILP;   // thread-level instructions
  const double per_clk_sm = ops / (best * 1e-3) / ((double)khz * 1e3) / sms;
  printf("%-28s %8.3f ms  %7.1f lane-ops/clk/SM  %7.1f results/clk/SM (at %d MHz nominal)\n", name, best, per_clk_sm,
         per_clk_sm * elems_per_op, khz / 1000);
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 2 * 1024 * 4 * 2);
  run<0>("ex2.approx.ftz.f32", 1, out);
  run<1>("ex2.approx.f16x2", 2, out);
  run<2>("ex2.approx.ftz.bf16x2", 2, out);
  run<3>("fma.rn.f32x2", 2, out);
  run<4>("fma.rn.f32", 1, out);
  run<5>("max.f32 (3 in)", 2, out);
  run<6>("max.f32 (2 in)", 1, out);
  run<7>("cvt.rn.bf16x2.f32", 2, out);
  run<8>("cvt.rn.f16x2.f32", 2, out);
  run<9>("add.rn.f32x2", 2, out);
  run<10>("poly3 exp2 (scalar)", 1, out);
  run<11>("fma.rn.f16x2", 2, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
