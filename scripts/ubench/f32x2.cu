// Microbenchmark: issue/throughput of FFMA vs FFMA2 (packed f32x2, sm_100) and mixes with
// integer/XU work. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float2 a[8], b, c;
  b = make_float2(seed, seed * 1.0001f);
  c = make_float2(0.5f, 0.25f);
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  int acc = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {  // scalar FFMA x2 per pair
        a[i].x = __fmaf_rn(a[i].x, b.x, c.x);
        a[i].y = __fmaf_rn(a[i].y, b.y, c.y);
      } else if (MODE == 1) {  // packed
        a[i] = __ffma2_rn(a[i], b, c);
      } else if (MODE == 2) {  // packed + one integer op per pair
        a[i] = __ffma2_rn(a[i], b, c);
        acc = acc * 3 + i;
      } else if (MODE == 3) {  // scalar + one integer op per pair
        a[i].x = __fmaf_rn(a[i].x, b.x, c.x);
        a[i].y = __fmaf_rn(a[i].y, b.y, c.y);
        acc = acc * 3 + i;
      } else if (MODE == 4) {  // packed add/mul
        a[i] = __fadd2_rn(__fmul2_rn(a[i], b), c);
      } else if (MODE == 5) {  // scalar add/mul
        a[i].x = __fadd_rn(__fmul_rn(a[i].x, b.x), c.x);
        a[i].y = __fadd_rn(__fmul_rn(a[i].y, b.y), c.y);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
}

template <int MODE>
void run(const char* name, float* d) {
  const int iters = 20000, blocks = 148 * 4, threads = 256;
  k<MODE><<<blocks, threads>>>(d, 100, 1.0f);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(d, iters, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double pairs = (double)iters * 8 * blocks * threads;
  printf("%-28s %8.3f ms  %7.2f G pair-ops/s  (%.2f TFLOP/s if FMA)\n", name, ms, pairs / ms / 1e6, pairs * 4 / ms / 1e9);
}

int main() {
  float* d;
  cudaMalloc(&d, 148 * 4 * 256 * 4);
  run<0>("scalar FFMA x2", d);
  run<1>("FFMA2", d);
  run<3>("scalar FFMA x2 + IMAD", d);
  run<2>("FFMA2 + IMAD", d);
  run<5>("scalar FMUL,FADD x2", d);
  run<4>("FMUL2,FADD2", d);
  return 0;
}
