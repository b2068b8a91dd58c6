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
// CTA-wide barrier per step.  The 6x6 inverses of the diagonal blocks of L are computed after the factorisation, one
// thread per column; the backward substitution then is, per block, six independent dot products plus one rank-6 update
// over the lanes of a warp instead of a 6-step dependent solve.
#pragma once

namespace pgba {

#ifdef PGBA_SOLVE_TIMING      // per-phase clocks of steps kb = 0 and kb = 24 (profiles/microbench/solve_bench.cu)
#define C32_TS_W(i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && (kb == 0 || kb == 24)) g_solve_ts[(kb ? 30 : 10) + (i)] = clock64(); } while (0)
#define C32_TS_LA(i) do { if (threadIdx.x == 224 && blockIdx.x == 0 && (kb == 0 || kb == 24)) g_solve_ts[(kb ? 40 : 20) + (i)] = clock64(); } while (0)
#else
#define C32_TS_W(i) do { } while (0)
#define C32_TS_LA(i) do { } while (0)
#endif

constexpr int C32_LA = 7;                      // look-ahead warp
constexpr int C32_WORKERS = 224;               // threads of warps 0..6

__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("barrier.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ float rsqrt_fast(float x) {          // one MUFU.RSQ (2 ulp); pivots are >= 1 after the damping
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Cholesky of a 6x6 block held in REGISTERS (lower triangle a[r][c], c <= r), in place: a <- L, ri[c] = 1 / L[c][c].
// Square-root-free elimination with the unscaled column (a[r][c2] -= a[r][c] a[c2][c] / d_c, products formed while the
// reciprocal is in flight), so the dependent chain from pivot to pivot is one MUFU.RCP + one FFMA; the scaling by
// rsqrt(d_c) that turns the column into L runs off that chain.  A non-positive pivot gives NaN / inf, which propagates like
// the reference's unchecked potrf (info ignored, ba_cuda.cu:576).
__device__ __forceinline__ void factor6_regs(float a[6][6], float ri[6]) {
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const float d = a[c][c];
    const float inv = rcp_fast(d);
    const float rs = rsqrt_fast(d);
#pragma unroll
    for (int c2 = c + 1; c2 < 6; ++c2)
#pragma unroll
      for (int r = c2; r < 6; ++r) a[r][c2] = fmaf(-(a[r][c] * a[c2][c]), inv, a[r][c2]);
    a[c][c] = d * rs;
    ri[c] = rs;
#pragma unroll
    for (int r = c + 1; r < 6; ++r) a[r][c] *= rs;
  }
}

// Look-ahead warp: diagonal block at nb.  With `update`, first subtracts P P^T of the 6 panel rows just solved against block
// kb (21 lanes, one entry each); the 21 entries are gathered into lane 0 by shuffles (no shared-memory round trip), factored
// in registers and written back with rd.  la_r, la_c: this lane's entry (lanes >= 21: unused).
__device__ __forceinline__ void la_diag_block(float* A, float* rd, int ld, int kb, int nb, bool update, int lane, int la_r,
                                              int la_c) {
  float dv = 0.f;
  if (lane < 21) {
    dv = A[(nb + la_r) * ld + nb + la_c];
    if (update) {
      const float* pr = A + (nb + la_r) * ld + kb;
      const float* pc = A + (nb + la_c) * ld + kb;
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
      for (int a = 0; a < 6; a += 2) {
        acc0 += pr[a] * pc[a];
        acc1 += pr[a + 1] * pc[a + 1];
      }
      dv -= acc0 + acc1;
    }
  }
  float a6[6][6], ri[6];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) a6[r][c] = __shfl_sync(0xffffffffu, dv, (r * (r + 1)) / 2 + c);
  if (lane == 0) {
    factor6_regs(a6, ri);
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      rd[nb + r] = ri[r];
#pragma unroll
      for (int c = 0; c <= r; ++c) A[(nb + r) * ld + nb + c] = a6[r][c];
    }
  }
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

// Column c of the inverse of the 6x6 lower-triangular diagonal block at kb -> inv[r * 6 + c]: forward substitution
// L x = e_c, x[r] = (delta_rc - sum_{e < r} L[r][e] x[e]) / L[r][r], written without a branch on c (x[e] == 0 for e < c), so
// the threads of a warp that invert different columns do not diverge.
__device__ __forceinline__ void invert_diag_col_f32(const float* A, const float* rd, float* inv, int ld, int kb, int c) {
  float l[6][6], ri[6], x[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    ri[r] = rd[kb + r];
#pragma unroll
    for (int e = 0; e < r; ++e) l[r][e] = A[(kb + r) * ld + kb + e];
  }
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    float s0 = (r == c) ? 1.f : 0.f, s1 = 0.f;
#pragma unroll
    for (int e = 0; e < r; ++e) {
      if (e & 1) s1 = fmaf(-l[r][e], x[e], s1); else s0 = fmaf(-l[r][e], x[e], s0);
    }
    x[r] = (s0 + s1) * ri[r];
  }
#pragma unroll
  for (int r = 0; r < 6; ++r) inv[r * 6 + c] = x[r];
}

// In-place Cholesky of the lower triangle of A [(n + 1) x ld] (n a multiple of 6, ld odd); row n is the right-hand side
// and holds (L^-1 b)^T on exit.  rd [n] = 1 / diag(L); dinv [n / 6][36] = inverses of the diagonal blocks of L.  All 256
// threads must call; a __syncthreads() is done first and last.
__device__ __forceinline__ void chol6_f32(float* A, float* rd, float* dinv, int n, int ld) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int la_r = 0, la_c = lane;                       // look-ahead warp: lane -> entry (r, c), c <= r < 6, of a diagonal block
  if (warp == C32_LA) {
    while (la_c > la_r) { la_c -= la_r + 1; ++la_r; }
  }
  __syncthreads();
  if (warp == C32_LA) la_diag_block(A, rd, ld, 0, 0, false, lane, la_r, la_c);
  __syncthreads();
  for (int kb = 0; kb < n; kb += 6) {
    const int nb = kb + 6;
    if (warp == C32_LA) {
      if (nb < n) {
        C32_TS_LA(0);
        if (lane < 6) panel_row_f32(A, rd, ld, kb, nb + lane);
        __syncwarp();
        C32_TS_LA(1);
        // rows nb .. nb+5 of the panel are written: release the trailing update.  barrier.arrive is a release and the
        // consumers' barrier.sync an acquire at CTA scope (PTX memory model), so no separate fence (MEMBAR.SC costs ~100
        // cycles per step on this chain; measured identical results with and without, profiles/README.md)
        named_bar_arrive(1, 256);
        C32_TS_LA(2);
        la_diag_block(A, rd, ld, kb, nb, true, lane, la_r, la_c);
        C32_TS_LA(4);
      } else {
        named_bar_arrive(1, 256);
      }
    } else {
      // panel rows below the look-ahead rows (the last step has only the right-hand-side row n)
      const int r0 = (nb < n) ? nb + 6 : nb;
      C32_TS_W(0);
      for (int r = r0 + tid; r <= n; r += C32_WORKERS) panel_row_f32(A, rd, ld, kb, r);
      C32_TS_W(1);
      named_bar_sync(1, 256);
      C32_TS_W(2);
      if (nb < n) {
        // trailing update of rows >= nb + 6 (incl. the right-hand-side row): item = (row pair r0, r0 + 1; column block cb);
        // only items of the lower block triangle are enumerated: column block cbi (cb = nb + 6 cbi) pairs with the row
        // pairs rp >= max(0, 3 cbi - 3).  Shared-memory latency (~30 cycles) dominates an item, so ALL its operands -- two
        // panel rows, the 6 x 6 factor block of the column rows and the 12 targets -- are loaded before the first FMA
        // (measured: 1085 -> see profiles/README.md cycles per item against loads interleaved with read-modify-writes).
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
          const int r0i = rb0 + 2 * rp, cb = nb + 6 * cbi;
          const bool p0 = cb <= r0i, p1 = r0i + 1 <= n && cb <= r0i + 1;
          if (!p0 && !p1) continue;
          const int r1i = p1 ? r0i + 1 : r0i;      // a missing second row re-reads the first (never stored)
          float l0[6], l1[6], lc[6][6], t0[6], t1[6];
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            l0[a] = A[r0i * ld + kb + a];
            l1[a] = A[r1i * ld + kb + a];
            t0[a] = A[r0i * ld + cb + a];
            t1[a] = A[r1i * ld + cb + a];
          }
#pragma unroll
          for (int b = 0; b < 6; ++b)
#pragma unroll
            for (int a = 0; a < 6; ++a) lc[b][a] = A[(cb + b) * ld + kb + a];
#pragma unroll
          for (int b = 0; b < 6; ++b) {
#pragma unroll
            for (int a = 0; a < 6; ++a) {
              t0[b] = fmaf(-l0[a], lc[b][a], t0[b]);
              t1[b] = fmaf(-l1[a], lc[b][a], t1[b]);
            }
          }
#pragma unroll
          for (int b = 0; b < 6; ++b) {
            if (p0) A[r0i * ld + cb + b] = t0[b];
            if (p1) A[r1i * ld + cb + b] = t1[b];
          }
        }
      }
      C32_TS_W(3);
    }
    __syncthreads();
    C32_TS_W(4);
  }
  // inverses of the diagonal blocks of L, one thread per column (n <= 156 < 256 columns), for the backward substitution
  if (tid < n) invert_diag_col_f32(A, rd, dinv + (tid / 6) * 36, ld, 6 * (tid / 6), tid % 6);
  __syncthreads();
}

// Backward substitution L^T x = z by ONE warp: z = row n of A on entry, x on exit.  Every lane keeps its rows (lane,
// lane + 32, ...) of the running right-hand side in registers; per block (from the bottom) the six entries of the block
// are read from shared memory, every lane forms x_b = L_bb^-T z_b redundantly with the precomputed inverse (six
// independent dot products), updates its own rows with the 6 x NSLOT factor entries it fetched one step ahead, and the
// owners of the next block's rows publish them.  One shared-memory round trip and one __syncwarp per block.
template <int NSLOT>
__device__ __forceinline__ void backsub6_f32(float* A, const float* dinv, int n, int ld, int lane) {
  float* xv = A + n * ld;
  float v[NSLOT];
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) v[s] = (lane + 32 * s < n) ? xv[lane + 32 * s] : 0.f;
  float Ln[6][NSLOT], inv[21];
  auto fetch = [&](int kb) {                       // operands of block kb that do not depend on x
    const float* ib = dinv + (kb / 6) * 36;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) inv[(r * (r + 1)) / 2 + c] = ib[r * 6 + c];
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int row = lane + 32 * s;
        Ln[a][s] = (row < kb) ? A[(kb + a) * ld + row] : 0.f;
      }
  };
  fetch(n - 6);
  for (int kb = n - 6; kb >= 0; kb -= 6) {
    float z[6], x[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) z[a] = xv[kb + a];
    float Lc[6][NSLOT];
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) Lc[a][s] = Ln[a][s];
    float ic[21];
#pragma unroll
    for (int i = 0; i < 21; ++i) ic[i] = inv[i];
    if (kb >= 6) fetch(kb - 6);                    // next block's operands: in flight during this block's arithmetic
#pragma unroll
    for (int c = 0; c < 6; ++c) {                  // x_c = sum_{e >= c} inv[e][c] z[e]
      float acc = 0.f;
#pragma unroll
      for (int e = c; e < 6; ++e) acc = fmaf(ic[(e * (e + 1)) / 2 + c], z[e], acc);
      x[c] = acc;
    }
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      const int row = lane + 32 * s;
      float acc0 = v[s], acc1 = 0.f;
#pragma unroll
      for (int a = 0; a < 6; a += 2) {
        acc0 = fmaf(-Lc[a][s], x[a], acc0);
        acc1 = fmaf(-Lc[a + 1][s], x[a + 1], acc1);
      }
      float nv = acc0 + acc1;                       // rows < kb; rows of this block take x; rows above kb + 5 are final
#pragma unroll
      for (int a = 0; a < 6; ++a) nv = (row == kb + a) ? x[a] : nv;
      if (row < kb + 6) v[s] = nv;
      if (row >= kb - 6 && row < kb) xv[row] = nv;  // the next block's entries
    }
    __syncwarp();
  }
#pragma unroll
  for (int s = 0; s < NSLOT; ++s)
    if (lane + 32 * s < n) xv[lane + 32 * s] = v[s];
  __syncwarp();
}

__device__ __forceinline__ void backsub6_f32_any(float* A, const float* dinv, int n, int ld, int lane) {
  const int nslot = (n + 31) >> 5;
  if (nslot <= 2) backsub6_f32<2>(A, dinv, n, ld, lane);
  else if (nslot == 3) backsub6_f32<3>(A, dinv, n, ld, lane);
  else if (nslot == 4) backsub6_f32<4>(A, dinv, n, ld, lane);
  else backsub6_f32<5>(A, dinv, n, ld, lane);
}

}  // namespace pgba
