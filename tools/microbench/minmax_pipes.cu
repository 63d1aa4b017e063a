// Which pipe do integer and float min/max run on (sm_100a)?  Throughput of independent chains per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o minmax_pipes minmax_pipes.cu && ./minmax_pipes
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096, CH = 8;
template <int MODE>
__global__ void __launch_bounds__(256) k(int* out, int seed) {
  int a[CH];
  float f[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    a[i] = seed * (threadIdx.x + i + 1);
    f[i] = (float)(seed + i) * 0.37f + threadIdx.x;
  }
  int kk = seed * 7 + threadIdx.x;
  float kf = seed * 0.11f + threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0) a[i] = max(a[i], min(a[(i + 1) % CH], kk + i));           // 2 VIMNMX
      if (MODE == 1) f[i] = fmaxf(f[i], fminf(f[(i + 1) % CH], kf + (float)i)); // 2 FMNMX
      if (MODE == 2) {                                                         // 1 + 1
        a[i] = max(a[i], min(a[(i + 1) % CH], kk + i));
        f[i] = fmaxf(f[i], fminf(f[(i + 1) % CH], kf));
      }
      if (MODE == 3) f[i] = fmaf(f[i], 1.0001f, kf);                           // FFMA reference
      if (MODE == 4) {                                                         // VIMNMX + FFMA
        a[i] = max(a[i], min(a[(i + 1) % CH], kk + i));
        f[i] = fmaf(f[i], 1.0001f, kf);
      }
    }
    kk += it;
    kf += 1.f;
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += a[i] + __float_as_int(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, double ops_per_iter) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  int* out; cudaMalloc(&out, sms * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms * 8, 256>>>(out, 3);
  cudaEventRecord(e0);
  k<MODE><<<sms * 8, 256>>>(out, 3);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double thread_ops = (double)sms * 8 * 256 * ITERS * CH * ops_per_iter;
  printf("%-18s %.3f ms  %.1f thread-ops/clk/SM (at %d MHz nominal)\n", name, ms,
         thread_ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  cudaFree(out);
}
int main() {
  run<0>("VIMNMX x2", 2);
  run<1>("FMNMX x2", 2);
  run<2>("VIMNMX+FMNMX x2", 4);
  run<3>("FFMA", 1);
  run<4>("2 VIMNMX + FFMA", 3);
  return 0;
}
