// Stand-alone timing harness for solve_small_kernel (includes the kernel source directly).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -DPGBA_SOLVE_TIMING -o solve_bench solve_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../cdv-slam_b200/csrc/ba_numeric.cu"
namespace pgba { void count_launch() {} long long launch_count() { return 0; } int chunk_grid(const Problem&, int64_t) { return 1; } bool pdl_enabled() { return false; } cudaError_t launch_big_solve(const Problem&, int64_t, cudaStream_t) { return cudaSuccess; } }
using namespace pgba;
int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 10, n = 6 * N, F = N + 12, batch = argc > 2 ? atoi(argv[2]) : 1;
  Layout L = make_layout(40000, F, F * 96, N, batch, 8);
  char* ws; cudaMalloc(&ws, total_bytes(L, batch)); cudaMemset(ws, 0, total_bytes(L, batch));
  std::vector<float> S(n * n), y(n), poses(F * 7, 0.f);
  std::vector<double> M(n * n);
  srand(1);
  for (auto& v : M) v = rand() / (double)RAND_MAX - 0.5;
  for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) { double a = 0; for (int k = 0; k < n; ++k) a += M[r * n + k] * M[c * n + k]; S[r * n + c] = (float)(a * 100 + (r == c ? n : 0)); }
  for (auto& v : y) v = rand() / (float)RAND_MAX;
  for (int f = 0; f < F; ++f) poses[f * 7 + 6] = 1.f;
  float *dS, *dy, *dposes; cudaMalloc(&dS, 4 * n * n); cudaMalloc(&dy, 4 * n); cudaMalloc(&dposes, 4 * F * 7 * batch);
  cudaMemcpy(dS, S.data(), 4 * n * n, cudaMemcpyHostToDevice); cudaMemcpy(dy, y.data(), 4 * n, cudaMemcpyHostToDevice);
  for (int b = 0; b < batch; ++b) cudaMemcpy(dposes + b * F * 7, poses.data(), 4 * F * 7, cudaMemcpyHostToDevice);
  Problem pb{}; pb.poses = dposes; pb.st.poses = F * 7; pb.F = F; pb.K = F * 96; pb.P = 3; pb.t0 = 12; pb.t1 = 12 + N; pb.apply = 1; pb.with_schur = 1; pb.ws = ws; pb.L = L; pb.E = 40000;
  const size_t smem = solve_small_smem_bytes(n);
  cudaFuncSetAttribute(solve_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float tot = 0; const int reps = 50;
  for (int it = 0; it < reps + 5; ++it) {
    for (int w = 0; w < batch; ++w) { WinPtrs wp = win_ptrs(ws, L, w); cudaMemcpyAsync(wp.S, dS, 4 * n * n, cudaMemcpyDeviceToDevice); cudaMemcpyAsync(wp.y, dy, 4 * n, cudaMemcpyDeviceToDevice); }
    cudaEventRecord(a); solve_small_kernel<<<batch, 256, smem>>>(pb); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (it >= 5) tot += ms;
  }
  printf("N=%d n=%d batch=%d: %.2f us per launch (event-timed), err=%s\n", N, n, batch, 1e3 * tot / reps, cudaGetErrorString(cudaGetLastError()));
  // residual check in double on the host
  std::vector<float> dX(n); WinPtrs wp = win_ptrs(ws, L, 0); cudaMemcpy(dX.data(), wp.dX, 4 * n, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int r = 0; r < n; ++r) { double acc = 0; for (int c = 0; c < n; ++c) { double s = (double)S[(r >= c ? r * n + c : c * n + r)]; if (r == c) s += 1e-4 * s + 1.0; acc += s * dX[c]; } worst = fmax(worst, fabs(acc - y[r])); }
  printf("max |S x - y| = %.3e\n", worst);
#ifdef PGBA_SOLVE_TIMING
  long long ts[64]; cudaMemcpyFromSymbol(ts, g_solve_ts, sizeof(ts));
  printf("phase clocks (cycles since start; 1 loaded, 2 factored, 3 back-substituted, 4 end; 10+/30+ worker thread and 20+/40+ look-ahead\n"
         " warp in steps kb = 0 / 24, see ba_chol32.cuh):"); for (int i = 1; i < 64; ++i) if (ts[i]) printf(" [%d]%lld", i, ts[i] - ts[0]); printf("\n");
#endif
  return 0;
}
