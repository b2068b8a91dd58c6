// Large dense solve for the global (loop-closure) bundle adjustment: 6N > 156, e.g. BASELINE config c4 with
// N = 999 free poses (S is 5994 x 5994).  Replaces at::linalg_cholesky_ex + torch::cholesky_solve on the dense S of
// the reference (cdvslam/fastba/ba_cuda.cu:575-578 for eff_impl, :589-591 dense).
//
// Right-looking blocked Cholesky on the lower triangle of S (fp32 storage in the workspace, in place), panel width
// NB = 48 (8 pose blocks), with the right-hand side y riding along as an extra row:
//   big_potf2_kernel   diagonal tile: fp64 shared-memory Cholesky (ba_chol.cuh) + explicit inverse W = L11^-T
//   big_trsm_kernel    row tiles below: X = A21 * W (a small GEMM, no serial chain); records which row tiles are
//                      non-zero -- S of a pose graph is block-banded plus a few loop-closure rows, so most tiles
//                      stay exactly zero and are skipped by every later stage (identical results, ~10x less work)
//   big_syrk_kernel    trailing update A22 -= X X^T over PAIRS OF ACTIVE row tiles only (persistent grid)
//   big_back_kernel    backward substitution, one panel per launch, partial dot products + last-block finish
//   big_finish_kernel  dX, SE3 retraction of the free poses (ba_cuda.cu:88-206)
#include "ba_common.cuh"
#include "ba_chol.cuh"

namespace pgba {

constexpr int NB = BIG_NB;          // 48

struct BigPtrs {
  float* S; float* y; float* rdiag; float* winv; int* active; int* nact; float* tbuf; int* ticket;
};

__device__ __forceinline__ BigPtrs big_ptrs(const Problem& pb, const WinPtrs& wp, int w) {
  char* base = (char*)pb.ws + pb.L.body0 + (size_t)w * pb.L.body_bytes;
  char* z = (char*)pb.ws + (size_t)w * pb.L.zero_bytes;
  BigPtrs b;
  b.S = wp.S; b.y = wp.y;
  b.rdiag = (float*)(base + pb.L.o_rdiag);
  b.winv = (float*)(base + pb.L.o_winv);
  b.active = (int*)(base + pb.L.o_active);
  b.nact = (int*)(z + pb.L.z_nact);
  b.tbuf = (float*)(z + pb.L.z_bs);
  b.ticket = (int*)(z + pb.L.z_bs) + 64;
  return b;
}

// S += I * (1e-4 * S + 1)   (ba_cuda.cu:575/589)
__global__ void big_damp_kernel(Problem pb) {
  pdl_wait();
  pdl_trigger();
  const int w = blockIdx.y;
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  const int n6 = 6 * (pb.t1 - pb.t0);
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n6) {
    float* d = wp.S + (size_t)r * n6 + r;
    *d = *d + (1e-4f * *d + 1.0f);
  }
}

// grid = (1, batch), block = 256, dynamic smem: (2*NB) x (NB|1) doubles + NB
__global__ void __launch_bounds__(256, 1) big_potf2_kernel(Problem pb, int step) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ double sd[];
  const int w = blockIdx.y, tid = threadIdx.x;
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  const BigPtrs bp = big_ptrs(pb, wp, w);
  const int n6 = 6 * (pb.t1 - pb.t0), kb = step * NB;
  const int nh = min(NB, n6 - kb), ld = NB | 1;
  double* A = sd;                       // rows 0..nh-1: diagonal tile; rows nh..2nh-1: identity (-> L^-T ... see below)
  double* rd = sd + 2 * NB * ld;
  for (int x = tid; x < nh * nh; x += 256) {
    const int r = x / nh, c = x - r * nh;
    A[r * ld + c] = (c <= r) ? (double)bp.S[(size_t)(kb + r) * n6 + kb + c] : 0.0;
    A[(nh + r) * ld + c] = (r == c) ? 1.0 : 0.0;
  }
  chol6_smem(A, rd, nh, 2 * nh - 1, ld);
  // rows nh + i now hold (L^-1 e_i)^T, i.e. W[i][c] = Linv[c][i]  (so X = A21 * W solves X L^T = A21)
  for (int x = tid; x < nh * nh; x += 256) {
    const int r = x / nh, c = x - r * nh;
    if (c <= r) bp.S[(size_t)(kb + r) * n6 + kb + c] = (float)A[r * ld + c];
    bp.winv[(size_t)step * NB * NB + r * NB + c] = (float)A[(nh + r) * ld + c];
  }
  for (int x = tid; x < nh; x += 256) {
    bp.rdiag[kb + x] = (float)rd[x];
    if (!(rd[x] > 0.0) || !isfinite(rd[x])) atomicCAS(&wp.hdr->chol_info, 0, kb + x + 1);
  }
  if (tid == 0) bp.nact[step] = 0;
}

// Row "tile" t of the rows below panel `step`: t < ntb -> rows [kb+nh + 48 t, +48) of S; t == ntb -> the rhs row y.
// grid = (ntb + 1, batch), block = 256
__global__ void __launch_bounds__(256) big_trsm_kernel(Problem pb, int step) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sA[NB][NB + 1];
  __shared__ float sW[NB][NB + 1];
  const int w = blockIdx.y, tid = threadIdx.x;
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  const BigPtrs bp = big_ptrs(pb, wp, w);
  const int n6 = 6 * (pb.t1 - pb.t0), kb = step * NB;
  const int nh = min(NB, n6 - kb), r0 = kb + nh;
  const int ntb = (n6 - r0 + NB - 1) / NB;
  const int t = blockIdx.x;
  const bool rhs = (t == ntb);
  const int rows = rhs ? 1 : min(NB, n6 - (r0 + t * NB));
  float* src = rhs ? (bp.y + kb) : (bp.S + (size_t)(r0 + t * NB) * n6 + kb);
  const size_t rstride = rhs ? 0 : (size_t)n6;
  int nz = 0;
  for (int x = tid; x < rows * nh; x += 256) {
    const int r = x / nh, c = x - r * nh;
    const float v = src[r * rstride + c];
    sA[r][c] = v;
    nz |= (v != 0.0f);
  }
  nz = __syncthreads_or(nz);
  if (!nz && !rhs) return;                       // an all-zero tile stays zero: inactive for this panel
  for (int x = tid; x < nh * nh; x += 256) {
    const int r = x / nh, c = x - r * nh;
    sW[r][c] = bp.winv[(size_t)step * NB * NB + r * NB + c];
  }
  __syncthreads();
  for (int x = tid; x < rows * nh; x += 256) {
    const int r = x / nh, c = x - r * nh;
    float acc = 0.f;
    for (int e = 0; e <= c; ++e) acc += sA[r][e] * sW[e][c];      // W[e][c] = Linv[c][e] is zero for e > c
    src[r * rstride + c] = acc;
  }
  if (tid == 0) {
    const int slot = atomicAdd(&bp.nact[step], 1);
    bp.active[(size_t)step * pb.L.big_tiles + slot] = rhs ? -1 : t;   // -1 marks the rhs row
  }
}

// Trailing update over pairs (a >= b) of active row tiles of this panel: S[tile a][tile b] -= X_a X_b^T.
// Persistent grid: grid = (gx, batch), block = 256 (16 x 16 threads, 3 x 3 outputs each).
__global__ void __launch_bounds__(256) big_syrk_kernel(Problem pb, int step) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sXa[NB][NB + 1];     // [k][row]
  __shared__ float sXb[NB][NB + 1];
  const int w = blockIdx.y, tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  const BigPtrs bp = big_ptrs(pb, wp, w);
  const int n6 = 6 * (pb.t1 - pb.t0), kb = step * NB;
  const int nh = min(NB, n6 - kb), r0 = kb + nh;
  const int na = bp.nact[step];
  const int* act = bp.active + (size_t)step * pb.L.big_tiles;
  const int npairs = na * (na + 1) / 2;
  for (int pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
    int ia = (int)((sqrtf(8.f * pr + 1.f) - 1.f) * 0.5f);
    while (ia * (ia + 1) / 2 > pr) --ia;
    while ((ia + 1) * (ia + 2) / 2 <= pr) ++ia;
    const int ib = pr - ia * (ia + 1) / 2;
    int ta = act[ia], tb = act[ib];
    if (ta == -1 && tb == -1) continue;            // rhs x rhs: nothing to update
    // order so that "a" is the lower tile (larger row index); the rhs row is below everything
    if (tb == -1 || (ta != -1 && tb > ta)) { const int s = ta; ta = tb; tb = s; }
    const bool rhs = (ta == -1);
    const int ra = rhs ? 0 : r0 + ta * NB, rb = r0 + tb * NB;
    const int rows_a = rhs ? 1 : min(NB, n6 - ra), rows_b = min(NB, n6 - rb);
    const float* xa = rhs ? (bp.y + kb) : (bp.S + (size_t)ra * n6 + kb);
    const size_t sa = rhs ? 0 : (size_t)n6;
    const float* xb = bp.S + (size_t)rb * n6 + kb;
    __syncthreads();
    for (int x = tid; x < NB * NB; x += 256) {
      const int r = x / NB, k = x - r * NB;
      sXa[k][r] = (r < rows_a && k < nh) ? xa[r * sa + k] : 0.f;
      sXb[k][r] = (r < rows_b && k < nh) ? xb[(size_t)r * n6 + k] : 0.f;
    }
    __syncthreads();
    float acc[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    for (int k = 0; k < nh; ++k) {
      float a[3], b[3];
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
          bp.S[(size_t)(ra + r) * n6 + rb + c] -= acc[i][j];
        }
      }
  }
}

// Backward substitution, panel `step` (launched for step = last .. 0): x_k = L11^-T (y_k - sum_below L[r][k]^T x[r]).
// grid = (gx, batch), block = 256.  The active row tiles of this panel are the only rows with non-zero L[r][k].
__global__ void __launch_bounds__(256) big_back_kernel(Problem pb, int step) {
  pdl_wait();
  pdl_trigger();
  __shared__ float spart[5][NB];
  __shared__ float sz[NB];
  __shared__ int s_last;
  const int w = blockIdx.y, tid = threadIdx.x;
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  const BigPtrs bp = big_ptrs(pb, wp, w);
  const int n6 = 6 * (pb.t1 - pb.t0), kb = step * NB;
  const int nh = min(NB, n6 - kb), r0 = kb + nh;
  const int na = bp.nact[step];
  const int* act = bp.active + (size_t)step * pb.L.big_tiles;
  const int c = tid % NB, grp = tid / NB;          // 5 row groups x 48 columns (threads 240..255 idle)
  for (int ia = blockIdx.x; ia < na; ia += gridDim.x) {
    const int t = act[ia];
    if (t == -1) continue;
    const int ra = r0 + t * NB, rows = min(NB, n6 - ra);
    float acc = 0.f;
    if (grp < 5 && c < nh)
      for (int r = grp; r < rows; r += 5) acc += bp.S[(size_t)(ra + r) * n6 + kb + c] * bp.y[ra + r];
    __syncthreads();
    if (grp < 5) spart[grp][c] = acc;
    __syncthreads();
    if (tid < nh) atomicAdd(&bp.tbuf[tid], spart[0][tid] + spart[1][tid] + spart[2][tid] + spart[3][tid] + spart[4][tid]);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(bp.ticket, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < nh) {
    sz[tid] = bp.y[kb + tid] - __ldcg(&bp.tbuf[tid]);
    bp.tbuf[tid] = 0.f;
  }
  if (tid == 0) *bp.ticket = 0;
  __syncthreads();
  if (tid < nh) {                                   // x[c] = sum_{e >= c} W[c][e] z[e]
    const float* W = bp.winv + (size_t)step * NB * NB + tid * NB;
    float acc = 0.f;
    for (int e = tid; e < nh; ++e) acc += W[e] * sz[e];
    bp.y[kb + tid] = acc;
  }
}

__global__ void big_finish_kernel(Problem pb) {
  pdl_wait();
  pdl_trigger();
  const int w = blockIdx.y;
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  const int N = pb.t1 - pb.t0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float xi[6];
#pragma unroll
  for (int a = 0; a < 6; ++a) { xi[a] = wp.y[6 * i + a]; wp.dX[6 * i + a] = xi[a]; }
  if (pb.apply) retract_pose(pb.poses + (int64_t)w * pb.st.poses + 7 * (int64_t)(pb.t0 + i), xi);
}

// Host: the whole dense solve of one Gauss-Newton iteration (S, y -> dX, poses).
cudaError_t launch_big_solve(const Problem& pb, int64_t batch, cudaStream_t stream) {
  const int N = pb.t1 - pb.t0, n6 = 6 * N;
  const int nsteps = (n6 + NB - 1) / NB;
  const unsigned B = (unsigned)batch;
  launch_k(big_damp_kernel, dim3((n6 + 255) / 256, B), dim3(256), 0, stream, pb);
  count_launch();
  const size_t psm = sizeof(double) * ((size_t)2 * NB * (NB | 1) + NB);
  cudaFuncSetAttribute(big_potf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm);
  const int gx = batch > 1 ? 64 : 148 * 2;
  for (int s = 0; s < nsteps; ++s) {
    launch_k(big_potf2_kernel, dim3(1, B), dim3(256), psm, stream, pb, s);
    count_launch();
    const int r0 = s * NB + (n6 - s * NB < NB ? n6 - s * NB : NB);
    const int ntb = (n6 - r0 + NB - 1) / NB;
    launch_k(big_trsm_kernel, dim3(ntb + 1, B), dim3(256), 0, stream, pb, s);
    count_launch();
    if (ntb > 0) {
      launch_k(big_syrk_kernel, dim3(gx, B), dim3(256), 0, stream, pb, s);
      count_launch();
    }
  }
  for (int s = nsteps - 1; s >= 0; --s) {
    launch_k(big_back_kernel, dim3(16, B), dim3(256), 0, stream, pb, s);
    count_launch();
  }
  launch_k(big_finish_kernel, dim3((N + 127) / 128, B), dim3(128), 0, stream, pb);
  count_launch();
  return cudaGetLastError();
}

}  // namespace pgba
