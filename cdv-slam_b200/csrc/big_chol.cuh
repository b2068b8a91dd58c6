// Right-looking blocked Cholesky solve of a large dense SPD system whose lower triangle is mostly zero tiles
// (block-banded pose graphs plus a few loop-closure rows), in place in global memory, panel width NB = 48:
//   big_potf2_kernel   diagonal tile: fp64 shared-memory Cholesky (ba_chol.cuh) + explicit inverse W = L11^-T
//   big_trsm_kernel    row tiles below: X = A21 * W (a small GEMM, no serial chain); records which row tiles are
//                      non-zero -- all-zero tiles stay zero and are skipped by every later stage (identical results)
//   big_syrk_kernel    trailing update A22 -= X X^T over PAIRS OF ACTIVE row tiles only (persistent grid)
//   big_back_kernel    backward substitution, one panel per launch, partial dot products + last-block finish
// The right-hand side rides along as an extra row; on exit y holds the solution.
//
// Templated on a "system provider" Sys: `typename Sys::T` is the storage type (float for the bundle adjustment, double
// for the pose-graph solve) and `sys.get(w)` returns the pointers of system w (blockIdx.y).  Used by ba_bigsolve.cu
// (global BA, reference: cdvslam/fastba/ba_cuda.cu:575-578, 589-591) and pgo.cu (cuda_ba.solve_system, ba.cpp:99-180).
#pragma once
#include <stdlib.h>

#include "ba_common.cuh"
#include "ba_chol.cuh"

namespace pgba {

constexpr int NB = BIG_NB;          // 48

template <typename T> struct BigSys {
  T* S;          // [n x ld] row-major, lower triangle used
  T* y;          // [n] right-hand side -> solution
  int n, ld;
  T* rdiag;      // [n]
  T* winv;       // [steps][NB][NB]
  int* active;   // [steps][big_tiles]
  int* nact;     // [steps]   (zero-initialised not required: potf2 resets its entry)
  T* tbuf;       // [NB] zero-initialised scratch of the backward substitution
  int* ticket;   // zero-initialised
  int big_tiles;
  int* chol_info;  // 0, or 1 + index of the first non-positive pivot
};

// grid = (1, batch), block = 256, dynamic smem: (2*NB) x (NB|1) doubles + NB
template <typename T>
__device__ __forceinline__ void big_potf2_dev(const BigSys<T>& bp, int step, double* sd) {
  const int tid = threadIdx.x;
  const int n6 = bp.n, kb = step * NB;
  const int nh = min(NB, n6 - kb), ld = NB | 1;
  double* A = sd;                       // rows 0..nh-1: diagonal tile; rows nh..2nh-1: identity (-> L^-T, see below)
  double* rd = sd + 2 * NB * ld;
  {
    constexpr int NI = NB * NB / 256;              // all loads of the tile in flight before the first shared-memory store
    T v[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int x = tid + 256 * i, r = x / nh, c = x - r * nh;
      v[i] = (x < nh * nh && c <= r) ? bp.S[(size_t)(kb + r) * bp.ld + kb + c] : (T)0;
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int x = tid + 256 * i, r = x / nh, c = x - r * nh;
      if (x < nh * nh) {
        A[r * ld + c] = (double)v[i];
        A[(nh + r) * ld + c] = (r == c) ? 1.0 : 0.0;
      }
    }
  }
  chol6_smem(A, rd, nh, 2 * nh - 1, ld);
  // rows nh + i now hold (L^-1 e_i)^T, i.e. W[i][c] = Linv[c][i]  (so X = A21 * W solves X L^T = A21)
  for (int x = tid; x < nh * nh; x += 256) {
    const int r = x / nh, c = x - r * nh;
    if (c <= r) bp.S[(size_t)(kb + r) * bp.ld + kb + c] = (T)A[r * ld + c];
    bp.winv[(size_t)step * NB * NB + r * NB + c] = (T)A[(nh + r) * ld + c];
  }
  for (int x = tid; x < nh; x += 256) {
    bp.rdiag[kb + x] = (T)rd[x];
    if (!(rd[x] > 0.0) || !isfinite(rd[x])) atomicCAS(bp.chol_info, 0, kb + x + 1);
  }
  if (tid == 0) bp.nact[step] = 0;
}

template <typename Sys>
__global__ void __launch_bounds__(256, 1) big_potf2_kernel(Sys sys, int step) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ double sd[];
  big_potf2_dev(sys.get(blockIdx.y), step, sd);
}

// Row "tile" t of the rows below panel `step`: t < ntb -> rows [kb+nh + 48 t, +48) of S; t == ntb -> the rhs row y.
// grid = (ntb + 1, batch), block = 256
template <typename T>
__device__ __forceinline__ void big_trsm_dev(const BigSys<T>& bp, int step, int t, T (*sA)[NB + 1], T (*sW)[NB + 1]) {
  const int tid = threadIdx.x;
  const int n6 = bp.n, kb = step * NB;
  const int nh = min(NB, n6 - kb), r0 = kb + nh;
  const int ntb = (n6 - r0 + NB - 1) / NB;
  __syncthreads();                               // the tiles may still be in use by the caller's previous item
  const bool rhs = (t == ntb);
  const int rows = rhs ? 1 : min(NB, n6 - (r0 + t * NB));
  T* src = rhs ? (bp.y + kb) : (bp.S + (size_t)(r0 + t * NB) * bp.ld + kb);
  const size_t rstride = rhs ? 0 : (size_t)bp.ld;
  int nz = 0;
  {
    // the row tile and W are requested together, all loads ahead of the first shared-memory store (one round trip instead
    // of one per store; the all-zero test of the row tile only decides whether the product runs)
    constexpr int NI = NB * NB / 256;
    T va[NI], vw[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int x = tid + 256 * i, r = x / nh, c = x - r * nh;
      va[i] = (x < rows * nh) ? src[r * rstride + c] : (T)0;
      vw[i] = (x < nh * nh) ? bp.winv[(size_t)step * NB * NB + r * NB + c] : (T)0;
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int x = tid + 256 * i, r = x / nh, c = x - r * nh;
      if (x < rows * nh) { sA[r][c] = va[i]; nz |= (va[i] != (T)0); }
      if (x < nh * nh) sW[r][c] = vw[i];
    }
  }
  nz = __syncthreads_or(nz);
  if (!nz && !rhs) return;                       // an all-zero tile stays zero: inactive for this panel
  for (int x = tid; x < rows * nh; x += 256) {
    const int r = x / nh, c = x - r * nh;
    T acc = 0;
    for (int e = 0; e <= c; ++e) acc += sA[r][e] * sW[e][c];      // W[e][c] = Linv[c][e] is zero for e > c
    src[r * rstride + c] = acc;
  }
  if (tid == 0) {
    const int slot = atomicAdd(&bp.nact[step], 1);
    bp.active[(size_t)step * bp.big_tiles + slot] = rhs ? -1 : t;   // -1 marks the rhs row
  }
}

template <typename Sys>
__global__ void __launch_bounds__(256) big_trsm_kernel(Sys sys, int step) {
  pdl_wait();
  pdl_trigger();
  using T = typename Sys::T;
  __shared__ T sA[NB][NB + 1];
  __shared__ T sW[NB][NB + 1];
  big_trsm_dev(sys.get(blockIdx.y), step, (int)blockIdx.x, sA, sW);
}

// Trailing update over pairs (a >= b) of active row tiles of this panel: S[tile a][tile b] -= X_a X_b^T.
// Persistent grid: grid = (gx, batch), block = 256 (16 x 16 threads, 3 x 3 outputs each).
template <typename T>
__device__ __forceinline__ void big_syrk_pair_dev(const BigSys<T>& bp, int step, int ta, int tb, T (*sXa)[NB + 1],
                                                  T (*sXb)[NB + 1]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n6 = bp.n, kb = step * NB;
  const int nh = min(NB, n6 - kb), r0 = kb + nh;
  if (ta == -1 && tb == -1) return;              // rhs x rhs: nothing to update
  // order so that "a" is the lower tile (larger row index); the rhs row is below everything
  if (tb == -1 || (ta != -1 && tb > ta)) { const int s = ta; ta = tb; tb = s; }
  const bool rhs = (ta == -1);
  const int ra = rhs ? 0 : r0 + ta * NB, rb = r0 + tb * NB;
  const int rows_a = rhs ? 1 : min(NB, n6 - ra), rows_b = min(NB, n6 - rb);
  const T* xa = rhs ? (bp.y + kb) : (bp.S + (size_t)ra * bp.ld + kb);
  const size_t sa = rhs ? 0 : (size_t)bp.ld;
  const T* xb = bp.S + (size_t)rb * bp.ld + kb;
  __syncthreads();
  {
    constexpr int NI = NB * NB / 256;              // both tiles in flight before the first shared-memory store
    T va[NI], vb[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int x = tid + 256 * i, r = x / NB, k = x - r * NB;
      va[i] = (r < rows_a && k < nh) ? xa[r * sa + k] : (T)0;
      vb[i] = (r < rows_b && k < nh) ? xb[(size_t)r * bp.ld + k] : (T)0;
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int x = tid + 256 * i, r = x / NB, k = x - r * NB;
      sXa[k][r] = va[i];
      sXb[k][r] = vb[i];
    }
  }
  __syncthreads();
  T acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int k = 0; k < nh; ++k) {
    T a[3], b[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { a[i] = sXa[k][ty + 16 * i]; b[i] = sXb[k][tx + 16 * i]; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) acc[i][j] += a[i] * b[j];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int r = ty + 16 * i, c = tx + 16 * j;
      if (r >= rows_a || c >= rows_b) continue;
      if (rhs) {
        bp.y[rb + c] -= acc[i][j];
      } else if (rb + c <= ra + r) {             // lower triangle only
        bp.S[(size_t)(ra + r) * bp.ld + rb + c] -= acc[i][j];
      }
    }
}

__device__ __forceinline__ void pair_of(int pr, int& ia, int& ib) {
  ia = (int)((sqrtf(8.f * pr + 1.f) - 1.f) * 0.5f);
  while (ia * (ia + 1) / 2 > pr) --ia;
  while ((ia + 1) * (ia + 2) / 2 <= pr) ++ia;
  ib = pr - ia * (ia + 1) / 2;
}

template <typename Sys>
__global__ void __launch_bounds__(256) big_syrk_kernel(Sys sys, int step) {
  pdl_wait();
  pdl_trigger();
  using T = typename Sys::T;
  __shared__ T sXa[NB][NB + 1];     // [k][row]
  __shared__ T sXb[NB][NB + 1];
  const BigSys<T> bp = sys.get(blockIdx.y);
  const int na = bp.nact[step];
  const int* act = bp.active + (size_t)step * bp.big_tiles;
  const int npairs = na * (na + 1) / 2;
  for (int pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
    int ia, ib;
    pair_of(pr, ia, ib);
    big_syrk_pair_dev(bp, step, act[ia], act[ib], sXa, sXb);
  }
}

// Backward substitution, panel `step` (launched for step = last .. 0): x_k = L11^-T (y_k - sum_below L[r][k]^T x[r]).
// grid = (gx, batch), block = 256.  The active row tiles of this panel are the only rows with non-zero L[r][k].
template <typename T>
__device__ __forceinline__ void big_back_partial_dev(const BigSys<T>& bp, int step, int first, int stride, T (*spart)[NB]) {
  const int tid = threadIdx.x;
  const int n6 = bp.n, kb = step * NB;
  const int nh = min(NB, n6 - kb), r0 = kb + nh;
  const int na = bp.nact[step];
  const int* act = bp.active + (size_t)step * bp.big_tiles;
  const int c = tid % NB, grp = tid / NB;          // 5 row groups x 48 columns (threads 240..255 idle)
  for (int ia = first; ia < na; ia += stride) {
    const int t = act[ia];
    if (t == -1) continue;
    const int ra = r0 + t * NB, rows = min(NB, n6 - ra);
    T acc = 0;
    if (grp < 5 && c < nh)
      for (int r = grp; r < rows; r += 5) acc += bp.S[(size_t)(ra + r) * bp.ld + kb + c] * bp.y[ra + r];
    __syncthreads();
    if (grp < 5) spart[grp][c] = acc;
    __syncthreads();
    if (tid < nh) atomicAdd(&bp.tbuf[tid], spart[0][tid] + spart[1][tid] + spart[2][tid] + spart[3][tid] + spart[4][tid]);
  }
}

template <typename T>
__device__ __forceinline__ void big_back_finish_dev(const BigSys<T>& bp, int step, T* sz) {
  const int tid = threadIdx.x;
  const int n6 = bp.n, kb = step * NB;
  const int nh = min(NB, n6 - kb);
  if (tid < nh) {
    sz[tid] = bp.y[kb + tid] - __ldcg(&bp.tbuf[tid]);
    bp.tbuf[tid] = 0;
  }
  __syncthreads();
  if (tid < nh) {                                   // x[c] = sum_{e >= c} W[c][e] z[e]
    const T* W = bp.winv + (size_t)step * NB * NB + tid * NB;
    T acc = 0;
    for (int e = tid; e < nh; ++e) acc += W[e] * sz[e];
    bp.y[kb + tid] = acc;
  }
}

template <typename Sys>
__global__ void __launch_bounds__(256) big_back_kernel(Sys sys, int step) {
  pdl_wait();
  pdl_trigger();
  using T = typename Sys::T;
  __shared__ T spart[5][NB];
  __shared__ T sz[NB];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const BigSys<T> bp = sys.get(blockIdx.y);
  big_back_partial_dev(bp, step, (int)blockIdx.x, (int)gridDim.x, spart);
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(bp.ticket, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid == 0) *bp.ticket = 0;
  big_back_finish_dev(bp, step, sz);
}

// The whole backward substitution L^T x = z in ONE launch: one CTA of 1024 threads per system walks the panels from the last
// to the first (the 125 dependent launches of big_back_kernel cost ~9 us each on the global BA: launch + cold round trips
// for ~50 KB of work).  The running solution lives in shared memory (n floats); per panel the 32 warps split its ACTIVE row
// tiles (the only rows with non-zero L[r][panel]), each warp accumulating its 48 partial dot products with coalesced row
// reads, partials are folded through shared memory, then 48 threads apply the explicit inverse of the diagonal tile.
// The tile list and W of the next panel are independent of x, so their first touch overlaps the current panel's tail.
// dynamic smem: (n + 32 * NB + NB) * sizeof(T)
template <typename Sys>
__global__ void __launch_bounds__(1024, 1) big_backsolve_kernel(Sys sys, int nsteps) {
  pdl_wait();
  pdl_trigger();
  using T = typename Sys::T;
  extern __shared__ __align__(16) unsigned char bsm[];
  const BigSys<T> bp = sys.get(blockIdx.y);
  const int n6 = bp.n;
  T* sx = reinterpret_cast<T*>(bsm);                // [n]      z on entry, x as the panels complete
  T* spart = sx + n6;                               // [32][NB] per-warp partial sums
  T* sz = spart + 32 * NB;                          // [NB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int x = tid; x < n6; x += 1024) sx[x] = bp.y[x];
  __syncthreads();
  for (int step = nsteps - 1; step >= 0; --step) {
    const int kb = step * NB;
    const int nh = min(NB, n6 - kb), r0 = kb + nh;
    const int na = bp.nact[step];
    const int* act = bp.active + (size_t)step * bp.big_tiles;
    T acc0 = 0, acc1 = 0;                           // columns lane and lane + 32 of the panel
    for (int ia = warp; ia < na; ia += 32) {
      const int t = act[ia];
      if (t == -1) continue;                        // the right-hand-side row of the factorisation
      const int ra = r0 + t * NB, rows = min(NB, n6 - ra);
      const T* Lp = bp.S + (size_t)ra * bp.ld + kb;
      // 12 rows (24 loads) in flight per lane: the tile is cold in L2 / DRAM and a dependent row-by-row walk would pay a
      // full round trip per row
      constexpr int RB = 12;
      for (int rb = 0; rb < rows; rb += RB) {
        T v0[RB], v1[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const bool in = rb + r < rows;
          const T* row = Lp + (size_t)(rb + r) * bp.ld;
          v0[r] = (in && lane < nh) ? row[lane] : (T)0;
          v1[r] = (in && lane + 32 < nh) ? row[lane + 32] : (T)0;
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const T xr = (rb + r < rows) ? sx[ra + rb + r] : (T)0;
          acc0 += v0[r] * xr;
          acc1 += v1[r] * xr;
        }
      }
    }
    spart[warp * NB + lane] = acc0;
    if (lane + 32 < NB) spart[warp * NB + lane + 32] = acc1;
    __syncthreads();
    if (tid < nh) {
      T s = 0;
#pragma unroll
      for (int w = 0; w < 32; ++w) s += spart[w * NB + tid];
      sz[tid] = sx[kb + tid] - s;
    }
    __syncthreads();
    if (tid < nh) {                                 // x[c] = sum_{e >= c} W[c][e] z[e]
      const T* W = bp.winv + (size_t)step * NB * NB + tid * NB;
      T a = 0;
      for (int e = tid; e < nh; ++e) a += W[e] * sz[e];
      sx[kb + tid] = a;
      bp.y[kb + tid] = a;
    }
    __syncthreads();
  }
}

// Host: factorisation + both substitutions of `batch` systems of order n (the damping has been applied by the caller).
template <typename Sys>
cudaError_t launch_big_chol(const Sys& sys, int n, int64_t batch, cudaStream_t stream) {
  const int nsteps = (n + NB - 1) / NB;
  const unsigned B = (unsigned)batch;
  const size_t psm = sizeof(double) * ((size_t)2 * NB * (NB | 1) + NB);
  cudaFuncSetAttribute(big_potf2_kernel<Sys>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm);
  const int gx = batch > 1 ? 64 : 148 * 2;
  for (int s = 0; s < nsteps; ++s) {
    launch_k(big_potf2_kernel<Sys>, dim3(1, B), dim3(256), psm, stream, sys, s);
    count_launch();
    const int r0 = s * NB + (n - s * NB < NB ? n - s * NB : NB);
    const int ntb = (n - r0 + NB - 1) / NB;
    launch_k(big_trsm_kernel<Sys>, dim3(ntb + 1, B), dim3(256), 0, stream, sys, s);
    count_launch();
    if (ntb > 0) {
      launch_k(big_syrk_kernel<Sys>, dim3(gx, B), dim3(256), 0, stream, sys, s);
      count_launch();
    }
  }
  // backward substitution: one launch (x in shared memory) when the system fits, else one launch per panel
  using T = typename Sys::T;
  const size_t bsmem = sizeof(T) * ((size_t)n + 32 * NB + NB);
  static int one_launch = -1;                       // PGBA_BIG_BACK_SPLIT=1: the per-panel kernels (A/B runs)
  if (one_launch < 0) { const char* e = getenv("PGBA_BIG_BACK_SPLIT"); one_launch = (e && e[0] == '1') ? 0 : 1; }
  if (one_launch && bsmem <= 200 * 1024) {
    cudaFuncSetAttribute(big_backsolve_kernel<Sys>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem);
    launch_k(big_backsolve_kernel<Sys>, dim3(1, B), dim3(1024), bsmem, stream, sys, nsteps);
    count_launch();
  } else {
    for (int s = nsteps - 1; s >= 0; --s) {
      launch_k(big_back_kernel<Sys>, dim3(16, B), dim3(256), 0, stream, sys, s);
      count_launch();
    }
  }
  return cudaGetLastError();
}

}  // namespace pgba
