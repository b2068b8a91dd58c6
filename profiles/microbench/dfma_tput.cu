// FP64 / FP32 FMA throughput of one SM (one CTA of 256 threads, 8 independent accumulators per thread).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_tput dfma_tput.cu
#include <cstdio>
template <typename T>
__global__ void k(T* out, long long* cyc, int iters) {
  T a[8];
  for (int i = 0; i < 8; ++i) a[i] = (T)(threadIdx.x + i);
  const T m = (T)1.0000001, c = (T)0.5;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a[i] * m + c;
  __syncthreads();
  const long long t1 = clock64();
  T s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <typename T> void run(const char* name, int threads) {
  T* out; long long* cyc; cudaMalloc(&out, 1024 * sizeof(T)); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<T><<<1, threads>>>(out, cyc, iters); k<T><<<1, threads>>>(out, cyc, iters);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%s threads=%d: %.2f FMA/clk/SM (%.2f cycles per warp-instruction per SMSP)\n", name, threads,
         (double)threads * 8 * iters / h, (double)h / (8.0 * iters * (threads / 32) / 4.0));
}
int main() {
  run<double>("fp64", 256); run<double>("fp64", 1024); run<double>("fp64", 32);
  run<float>("fp32", 256); run<float>("fp32", 1024);
  return 0;
}
