// Blocked (6 wide) right-looking Cholesky of a small dense SPD matrix held in shared memory as fp64, for one CTA of
// 256 threads.  Shared by solve_small_kernel (whole 6N x 6N system, right-hand side riding along as an extra row)
// and by the diagonal-panel factorisation of the large (global BA) solver.
#pragma once

namespace pgba {

#ifdef PGBA_SOLVE_TIMING
__device__ long long g_solve_ts[64];
#define SOLVE_TS(i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && (i) < 64) g_solve_ts[i] = clock64(); } while (0)
#define SOLVE_TS_LA(i) do { if (threadIdx.x == 32 * LA_WARP && blockIdx.x == 0 && (i) < 64) g_solve_ts[i] = clock64(); } while (0)
#else
#define SOLVE_TS(i) do { } while (0)
#define SOLVE_TS_LA(i) do { } while (0)
#endif

#ifndef LA_WARP
#define LA_WARP 7      // the look-ahead (critical path) warp
#endif

// Cholesky of the 6x6 diagonal block at kb (lower triangle, in place) by ONE thread; rd = 1 / diag(L).
// rsqrt of a non-positive pivot gives NaN/inf, which propagates like the reference's unchecked potrf (info ignored).
__device__ __forceinline__ void factor_diag6(double* A, double* rd, int ld, int kb) {
  // Right-looking inside the block: as soon as a column is scaled, every trailing entry is updated (independent FMAs),
  // so the dependent chain per column is rsqrt -> scale -> one FMA into the next pivot instead of a growing dot product.
  // The subtractions reach every entry in the same order as in the left-looking form (bitwise the same result).
  double Lk[6][6];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) Lk[r][c] = A[(kb + r) * ld + kb + c];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const double d = Lk[c][c];
    const double ri = rsqrt(d);
    Lk[c][c] = d * ri;
    rd[kb + c] = ri;
#pragma unroll
    for (int r = c + 1; r < 6; ++r) Lk[r][c] *= ri;
#pragma unroll
    for (int c2 = c + 1; c2 < 6; ++c2)
#pragma unroll
      for (int r = c2; r < 6; ++r) Lk[r][c2] -= Lk[r][c] * Lk[c2][c];
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
  int la_r = 0, la_c = lane;                       // look-ahead warp: lane -> entry (r, c), c <= r < 6, of the next block
  if (warp == LA_WARP) {
    while (la_c > la_r) { la_c -= la_r + 1; ++la_r; }
  }
  __syncthreads();
  for (int kb = 0; kb < n; kb += 6) {
    const int tsb = (kb == 0) ? 10 : (kb == 24 ? 20 : 100);
    (void)tsb;                                        // only used by the PGBA_SOLVE_TIMING build
    SOLVE_TS(tsb);
    // panel: rows below: x L11^T = a (right-looking over the 6 columns: chain = one multiply + one FMA per column)
    if (kb + 6 + tid <= last_row) {
      double l[6][6], ri[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        ri[c] = rd[kb + c];
#pragma unroll
        for (int e = 0; e < c; ++e) l[c][e] = A[(kb + c) * ld + kb + e];
      }
      for (int r = kb + 6 + tid; r <= last_row; r += 256) {
        double a[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) a[c] = A[r * ld + kb + c];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const double x = a[c] * ri[c];
          a[c] = x;
#pragma unroll
          for (int c2 = c + 1; c2 < 6; ++c2) a[c2] -= x * l[c2][c];
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) A[r * ld + kb + c] = a[c];
      }
    }
    SOLVE_TS(tsb + 1);
    __syncthreads();
    SOLVE_TS(tsb + 2);
    const int nb = kb + 6;
    if (nb >= n) break;
    if (warp == LA_WARP) {
      SOLVE_TS_LA(tsb + 5);
      // look-ahead: update and factor the next diagonal block while the other warps update the rest
      if (lane < 21) {
        const double* pr = A + (nb + la_r) * ld + kb;
        const double* pc = A + (nb + la_c) * ld + kb;
        double acc0 = 0.0, acc1 = 0.0;              // two chains of three
#pragma unroll
        for (int a = 0; a < 6; a += 2) {
          acc0 += pr[a] * pc[a];
          acc1 += pr[a + 1] * pc[a + 1];
        }
        A[(nb + la_r) * ld + nb + la_c] -= acc0 + acc1;
      }
      __syncwarp();
      SOLVE_TS_LA(tsb + 6);
      if (lane == 0) factor_diag6(A, rd, ld, nb);
      SOLVE_TS_LA(tsb + 7);
    } else {
      // trailing update of rows >= nb + 6: item = (row pair r0, r0 + 1; column block cb): the 6 x 6 factor block of the
      // column rows is loaded once for both rows.  (Measured on B200, n = 60: the update is bound by the 64-bit
      // shared-memory loads; 1 x 6, 3 x 6 and 6 x 6 items and hoisting all loads were slower, see profiles/README.md.)
      const int rb0 = nb + 6;
      const int nrows = last_row + 1 - rb0;
      const int npair = (nrows + 1) >> 1;
      const int nblk = (n - nb) / 6;                // column blocks nb, nb+6, ..., n-6
      // only the items of the lower block triangle are enumerated: column block cbi (cb = nb + 6 cbi) pairs with the row
      // pairs rp >= max(0, 3 cbi - 3); at n = 60 the first step has 141 such items for the 224 threads (one round; the
      // rectangular enumeration gave thread 0 two items and made the early steps trailing-bound)
      const int mb = nblk - 1;
      const int total = npair * nblk - 3 * ((mb * (mb - 1)) >> 1);    // npair >= 3 (nblk - 1): no column is empty
      for (int it = (warp < LA_WARP ? tid : tid - 32); it < total; it += 224) {
        int cbi = 0, rem = it, cnt = npair;
        if (rem >= cnt) {
          rem -= cnt; cbi = 1;
          while (rem >= cnt && cbi < mb) { rem -= cnt; cnt -= 3; ++cbi; }
        }
        const int rp = max(0, 3 * cbi - 3) + rem;
        const int r0 = rb0 + 2 * rp, r1 = r0 + 1, cb = nb + 6 * cbi;
        const bool p0 = cb <= r0, p1 = r1 <= last_row && cb <= r1;
        if (!p0 && !p1) continue;
        double l0[6], l1[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          l0[a] = A[r0 * ld + kb + a];
          l1[a] = p1 ? A[r1 * ld + kb + a] : 0.0;
        }
#pragma unroll
        for (int b = 0; b < 6; ++b) {
          const double* lc = A + (cb + b) * ld + kb;
          double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            const double c = lc[a];
            acc0 += l0[a] * c;
            acc1 += l1[a] * c;
          }
          if (p0) A[r0 * ld + cb + b] -= acc0;
          if (p1) A[r1 * ld + cb + b] -= acc1;
        }
      }
      SOLVE_TS(tsb + 3);
    }
    __syncthreads();
    SOLVE_TS(tsb + 4);
  }
}

}  // namespace pgba
