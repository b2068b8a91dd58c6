// Latency microbenchmarks on B200: dependent chains of fp64/fp32 ops (cycles per op, one thread).
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
template <int MODE> __global__ void k(double* out, double a, double b, long long* cyc) {
  double x = a; float xf = (float)a; double y = b;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
      if (MODE == 0) x = fma(x, y, a);                 // DFMA
      if (MODE == 1) x = x * y;                        // DMUL
      if (MODE == 2) xf = fmaf(xf, (float)b, (float)a);  // FFMA
      if (MODE == 3) x = (double)((float)x) + y;       // F2F down + up + DADD
      if (MODE == 4) xf = rsqrtf(xf) + 1.0f;           // MUFU.RSQ + FADD
      if (MODE == 5) x = rsqrt(x) + y;                 // double rsqrt
      if (MODE == 6) x = 1.0 / sqrt(x) + y;            // double sqrt + div
      if (MODE == 7) { double r = (double)rsqrtf((float)x); r = r * (1.5 - 0.5 * x * r * r); x = r + y; }  // seed+1 newton
      if (MODE == 8) x = 1.0 / x + y;                  // double div
      if (MODE == 9) { double r = (double)__frcp_rn((float)x); r = fma(r, fma(-x, r, 1.0), r); x = r + y; } // rcp seed + newton
      if (MODE == 10) x = x + y;                       // DADD
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x + xf;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name, int extra) {
  double* out; long long* cyc; cudaMalloc(&out, 8 * 64); cudaMalloc(&cyc, 8);
  k<MODE><<<1, 32>>>(out, 1.0000001, 1.0000002, cyc); k<MODE><<<1, 32>>>(out, 1.0000001, 1.0000002, cyc);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %.1f cycles/iter (minus %d for extras)\n", name, (double)h / N, extra);
}
int main() {
  run<0>("DFMA", 0); run<1>("DMUL", 0); run<10>("DADD", 0); run<2>("FFMA", 0); run<3>("F2F dn+up+DADD", 0);
  run<4>("MUFU.RSQ+FADD", 0); run<5>("rsqrt(double)+DADD", 0); run<6>("1/sqrt(double)+DADD", 0);
  run<7>("seed+1newton+DADD", 0); run<8>("1/x double + DADD", 0); run<9>("rcp seed+newton+DADD", 0);
  // smem / barrier latencies
  return 0;
}
