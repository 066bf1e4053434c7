// Micro-probe: FP64 FMA / DIV / SQRT throughput and FP32 FMA throughput on the current GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(double* out, int iters) {
  double a = threadIdx.x * 1e-3 + 1.0, b = 1.000001, c = 0.5, d = a + 1;
  float fa = a, fb = 1.000001f, fc = 0.5f, fd = fa + 1;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) { a = fma(a, b, c); d = fma(d, b, c); }
    if (MODE == 1) { a = c / (a + 1.0); d = c / (d + 1.0); }
    if (MODE == 2) { a = sqrt(a + 1.0); d = sqrt(d + 1.0); }
    if (MODE == 3) { fa = fmaf(fa, fb, fc); fd = fmaf(fd, fb, fc); }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + d + fa + fd;
}
template <int MODE> void run(const char* name, int iters) {
  double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(out, 16);
  cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 2.0 * iters * 148 * 8 * 256;
  printf("%s: %.3f ms, %.2f Gop/s (ops = fma/div/sqrt instructions)\n", name, ms, ops / ms / 1e6);
  cudaFree(out);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s sm_%d%d SMs %d clock %d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
  run<0>("dfma", 20000); run<1>("ddiv", 2000); run<2>("dsqrt", 2000); run<3>("ffma", 20000);
  return 0;
}
