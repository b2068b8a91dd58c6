// Small dense SPD solve in fp32 for one CTA of 256 threads (6N <= 156): blocked (6 wide) right-looking Cholesky in shared
// memory with the right-hand side riding along as an extra row, inverse-based backward substitution.
//
// The reference factors the damped Schur complement in fp32 (torch::linalg_cholesky_ex / cholesky_solve on a float
// tensor, cdvslam/fastba/ba_cuda.cu:576-577, 590-591), so this is its own arithmetic; round 1 did the same steps in fp64
// and was bound by the latency of the double-precision chain (rsqrt(double) 53 cycles, DFMA 8.8: 18 k cycles per
// factorisation at n = 60).  Measured accuracy of the fp32 factorisation on the c1 / c2 / c5 systems: dX within 1e-5 of
// the float64 solve (profiles/README.md, round 2).
//
// Critical path of a step kb: [6 rows of the next diagonal block: panel solve -> update -> 6 pivots] runs on ONE warp (the
// look-ahead warp) with warp-level synchronisation only; the other seven warps solve the remaining panel rows, wait for
// the look-ahead rows on a named barrier the look-ahead warp only ARRIVES at, and apply the trailing update.  One
// CTA-wide barrier per step.  The 6x6 inverses of the diagonal blocks of L are computed by six otherwise idle threads
// during the panel phase; the backward substitution then is, per block, six independent dot products plus one
// rank-6 update over the lanes of a warp instead of a 6-step dependent solve.
#pragma once

namespace pgba {

constexpr int C32_LA = 7;                      // look-ahead warp
constexpr int C32_WORKERS = 224;               // threads of warps 0..6
constexpr int C32_INV0 = 192;                  // threads C32_INV0 .. +5 invert the diagonal block (never hold a panel row)

__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("barrier.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// Cholesky of the 6x6 diagonal block at kb (lower triangle, in place) by ONE thread; rd = 1 / diag(L).  Right-looking
// inside the block: the dependent chain per pivot is rsqrt -> scale -> one FMA into the next pivot.  A non-positive pivot
// gives NaN / inf, which propagates like the reference's unchecked potrf (info ignored, ba_cuda.cu:576).
__device__ __forceinline__ void factor_diag6_f32(float* A, float* rd, int ld, int kb) {
  float Lk[6][6];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) Lk[r][c] = A[(kb + r) * ld + kb + c];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const float d = Lk[c][c];
    const float ri = rsqrtf(d);
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

// Row r below the diagonal block kb: x L11^T = a, right-looking over the 6 columns.
__device__ __forceinline__ void panel_row_f32(float* A, const float* rd, int ld, int kb, int r) {
  float l[6][6], ri[6], a[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    ri[c] = rd[kb + c];
    a[c] = A[r * ld + kb + c];
#pragma unroll
    for (int e = 0; e < c; ++e) l[c][e] = A[(kb + c) * ld + kb + e];
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const float x = a[c] * ri[c];
    a[c] = x;
#pragma unroll
    for (int c2 = c + 1; c2 < 6; ++c2) a[c2] -= x * l[c2][c];
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) A[r * ld + kb + c] = a[c];
}

// Column c of the inverse of the 6x6 lower-triangular diagonal block at kb -> inv[r * 6 + c], r >= c.
__device__ __forceinline__ void invert_diag_col_f32(const float* A, const float* rd, float* inv, int ld, int kb, int c) {
  float x[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) x[r] = 0.f;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    if (r < c) continue;
    if (r == c) { x[r] = rd[kb + r]; continue; }
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 6; ++e)
      if (e >= c && e < r) s += A[(kb + r) * ld + kb + e] * x[e];
    x[r] = -s * rd[kb + r];
  }
#pragma unroll
  for (int r = 0; r < 6; ++r)
    if (r >= c) inv[r * 6 + c] = x[r];
}

// In-place Cholesky of the lower triangle of A [(n + 1) x ld] (n a multiple of 6, ld odd); row n is the right-hand side
// and holds (L^-1 b)^T on exit.  rd [n] = 1 / diag(L); dinv [n / 6][36] = inverses of the diagonal blocks of L.  All 256
// threads must call; a __syncthreads() is done first and last.
__device__ __forceinline__ void chol6_f32(float* A, float* rd, float* dinv, int n, int ld) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __syncthreads();
  if (tid == 0) factor_diag6_f32(A, rd, ld, 0);
  int la_r = 0, la_c = lane;                       // look-ahead warp: lane -> entry (r, c), c <= r < 6, of the next block
  if (warp == C32_LA) {
    while (la_c > la_r) { la_c -= la_r + 1; ++la_r; }
  }
  __syncthreads();
  for (int kb = 0; kb < n; kb += 6) {
    const int nb = kb + 6;
    if (warp == C32_LA) {
      if (nb < n) {
        if (lane < 6) panel_row_f32(A, rd, ld, kb, nb + lane);
        __syncwarp();
        __threadfence_block();
        named_bar_arrive(1, 256);                 // rows nb .. nb+5 of the panel are written: release the trailing update
        if (lane < 21) {                          // update of the next diagonal block from those rows
          const float* pr = A + (nb + la_r) * ld + kb;
          const float* pc = A + (nb + la_c) * ld + kb;
          float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
          for (int a = 0; a < 6; a += 2) {
            acc0 += pr[a] * pc[a];
            acc1 += pr[a + 1] * pc[a + 1];
          }
          A[(nb + la_r) * ld + nb + la_c] -= acc0 + acc1;
        }
        __syncwarp();
        if (lane == 0) factor_diag6_f32(A, rd, ld, nb);
      } else {
        named_bar_arrive(1, 256);
      }
    } else {
      // panel rows below the look-ahead rows (the last step has only the right-hand-side row n)
      const int r0 = (nb < n) ? nb + 6 : nb;
      for (int r = r0 + tid; r <= n; r += C32_WORKERS) panel_row_f32(A, rd, ld, kb, r);
      if (tid >= C32_INV0 && tid < C32_INV0 + 6) invert_diag_col_f32(A, rd, dinv + (kb / 6) * 36, ld, kb, tid - C32_INV0);
      named_bar_sync(1, 256);
      if (nb < n) {
        // trailing update of rows >= nb + 6 (incl. the right-hand-side row): item = (row pair r0, r0 + 1; column block cb);
        // only items of the lower block triangle are enumerated: column block cbi (cb = nb + 6 cbi) pairs with the row
        // pairs rp >= max(0, 3 cbi - 3)
        const int rb0 = nb + 6;
        const int nrows = n + 1 - rb0;
        const int npair = (nrows + 1) >> 1;
        const int nblk = (n - nb) / 6;
        const int mb = nblk - 1;
        const int total = npair * nblk - 3 * ((mb * (mb - 1)) >> 1);
        for (int it = tid; it < total; it += C32_WORKERS) {
          int cbi = 0, rem = it, cnt = npair;
          if (rem >= cnt) {
            rem -= cnt; cbi = 1;
            while (rem >= cnt && cbi < mb) { rem -= cnt; cnt -= 3; ++cbi; }
          }
          const int rp = max(0, 3 * cbi - 3) + rem;
          const int r0i = rb0 + 2 * rp, r1i = r0i + 1, cb = nb + 6 * cbi;
          const bool p0 = cb <= r0i, p1 = r1i <= n && cb <= r1i;
          if (!p0 && !p1) continue;
          float l0[6], l1[6];
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            l0[a] = A[r0i * ld + kb + a];
            l1[a] = p1 ? A[r1i * ld + kb + a] : 0.f;
          }
#pragma unroll
          for (int b = 0; b < 6; ++b) {
            const float* lc = A + (cb + b) * ld + kb;
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
              const float c = lc[a];
              acc0 += l0[a] * c;
              acc1 += l1[a] * c;
            }
            if (p0) A[r0i * ld + cb + b] -= acc0;
            if (p1) A[r1i * ld + cb + b] -= acc1;
          }
        }
      }
    }
    __syncthreads();
  }
}

// Backward substitution L^T x = z by ONE warp: z = row n of A on entry, x on exit.  Per block (from the bottom): x_b =
// L_bb^-T z_b as six independent dot products with the precomputed inverse, then z_c -= L[b, c]^T x_b for the rows above.
__device__ __forceinline__ void backsub6_f32(float* A, const float* dinv, int n, int ld, int lane) {
  float* xv = A + n * ld;
  for (int kb = n - 6; kb >= 0; kb -= 6) {
    float x = 0.f;
    if (lane < 6) {
      const float* inv = dinv + (kb / 6) * 36;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int e = 0; e < 6; e += 2) {             // x_c = sum_{e >= c} inv[e][c] z[e]   (entries above the diagonal are never read)
        if (e >= lane) a0 += inv[e * 6 + lane] * xv[kb + e];
        if (e + 1 >= lane) a1 += inv[(e + 1) * 6 + lane] * xv[kb + e + 1];
      }
      x = a0 + a1;
    }
    __syncwarp();
    if (lane < 6) xv[kb + lane] = x;
    __syncwarp();
    for (int c = lane; c < kb; c += 32) {
      float v0 = xv[c], v1 = 0.f;
#pragma unroll
      for (int a = 0; a < 6; a += 2) {
        v0 -= A[(kb + a) * ld + c] * xv[kb + a];
        v1 -= A[(kb + a + 1) * ld + c] * xv[kb + a + 1];
      }
      xv[c] = v0 + v1;
    }
    __syncwarp();
  }
}

}  // namespace pgba
