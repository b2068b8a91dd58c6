// Blocked (6 wide) right-looking Cholesky of a small dense SPD matrix held in shared memory as fp64, for one CTA of
// 256 threads.  Shared by solve_small_kernel (whole 6N x 6N system, right-hand side riding along as an extra row)
// and by the diagonal-panel factorisation of the large (global BA) solver.
#pragma once

namespace pgba {

#ifndef LA_WARP
#define LA_WARP 7      // the look-ahead (critical path) warp
#endif

// Cholesky of the 6x6 diagonal block at kb (lower triangle, in place) by ONE thread; rd = 1 / diag(L).
// rsqrt of a non-positive pivot gives NaN/inf, which propagates like the reference's unchecked potrf (info ignored).
__device__ __forceinline__ void factor_diag6(double* A, double* rd, int ld, int kb) {
  double Lk[6][6];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) Lk[r][c] = A[(kb + r) * ld + kb + c];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    double d = Lk[c][c];
#pragma unroll
    for (int e = 0; e < c; ++e) d -= Lk[c][e] * Lk[c][e];
    const double ri = rsqrt(d);
    Lk[c][c] = d * ri;
    rd[kb + c] = ri;
#pragma unroll
    for (int r = c + 1; r < 6; ++r) {
      double v = Lk[r][c];
#pragma unroll
      for (int e = 0; e < c; ++e) v -= Lk[r][e] * Lk[c][e];
      Lk[r][c] = v * ri;
    }
  }
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) A[(kb + r) * ld + kb + c] = Lk[r][c];
}


// In-place Cholesky of the lower triangle of A [.. x ld] (n x n, n a multiple of 6); rows n..last_row (if any) are
// carried along as extra right-hand-side rows: on exit they hold (L^-1 b)^T.  rd[n] = 1 / diag(L).
// All 256 threads of the CTA must call; A must be fully written and visible (a __syncthreads() is done first).
__device__ __forceinline__ void chol6_smem(double* A, double* rd, int n, int last_row, int ld) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __syncthreads();
  if (tid == 0) factor_diag6(A, rd, ld, 0);
  __syncthreads();
  for (int kb = 0; kb < n; kb += 6) {
    // panel: rows below: x L11^T = a
    for (int r = kb + 6 + tid; r <= last_row; r += 256) {
      double x[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double v = A[r * ld + kb + c];
#pragma unroll
        for (int e = 0; e < c; ++e) v -= x[e] * A[(kb + c) * ld + kb + e];
        x[c] = v * rd[kb + c];
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) A[r * ld + kb + c] = x[c];
    }
    __syncthreads();
    const int nb = kb + 6;
    if (nb >= n) break;
    if (warp == LA_WARP) {
      // look-ahead: update and factor the next diagonal block while the other warps update the rest
      if (lane < 21) {
        int r = 0, c = lane;
        while (c > r) { c -= r + 1; ++r; }          // lane -> (r, c), c <= r < 6
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) acc += A[(nb + r) * ld + kb + a] * A[(nb + c) * ld + kb + a];
        A[(nb + r) * ld + nb + c] -= acc;
      }
      __syncwarp();
      if (lane == 0) factor_diag6(A, rd, ld, nb);
    } else {
      // trailing update of rows >= nb + 6: item = (row r, column block cb <= r)
      const int rb0 = nb + 6;
      const int nrows = last_row + 1 - rb0;
      const int nblk = (n - nb) / 6;                // column blocks nb, nb+6, ..., n-6
      for (int it = (warp < LA_WARP ? tid : tid - 32); it < nrows * nblk; it += 224) {
        const int r = rb0 + it / nblk, cb = nb + 6 * (it % nblk);
        if (cb > r) continue;
        double lr[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) lr[a] = A[r * ld + kb + a];
#pragma unroll
        for (int b = 0; b < 6; ++b) {
          const double* lc = A + (cb + b) * ld + kb;
          double acc = 0.0;
#pragma unroll
          for (int a = 0; a < 6; ++a) acc += lr[a] * lc[a];
          A[r * ld + cb + b] -= acc;
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace pgba
