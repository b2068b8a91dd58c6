// Pose-graph normal equations + solve on the device: replaces cuda_ba.solve_system (reference:
// cdvslam/fastba/ba.cpp:99-180, which moves everything to the CPU and uses Eigen's sparse SimplicialCholesky in double;
// called by the classical loop closure, cdvslam/loop_closure/optim_utils.py:212-244).
//
//   J is [7r x 7n] with two 7x7 blocks per residual row-block x: J_Ginv_i[x] at column block ii[x], J_Ginv_j[x] at
//   jj[x] (ba.cpp:140-156).  A = J^T J, b = -J^T res (:160-162), A.diag += A.diag * lm + ep (:164-165), then
//   A[:m,:m] delta[:m] = b[:m] with m = 7 * freen (all of A when freen < 0), delta[m:] = 0 (:101-118, :166).
//
// Everything is double like the reference.  A is held dense in the caller's workspace; its tiles outside the pose
// graph's band + loop-closure rows stay exactly zero and are skipped by the tile-sparse blocked Cholesky of
// big_chol.cuh, the same code the global bundle adjustment uses in fp32.
#include "big_chol.cuh"

namespace pgba {

struct PgoSys {
  using T = double;
  BigSys<double> s;
  __device__ BigSys<double> get(int) const { return s; }
};

struct PgoLayout { size_t A, b, rdiag, winv, active, nact, tbuf, ticket, info, total; int steps, tiles; };

// leading dimension of A: 7n rounded up to a multiple of 6 (the diagonal-tile factorisation works on 6-wide blocks; the
// solved block is padded with identity rows up to the next multiple of 6)
static size_t pgo_ld(int64_t n_poses) { return ((7 * (size_t)n_poses + 5) / 6) * 6; }

static PgoLayout pgo_layout(int64_t n_poses) {
  PgoLayout L{};
  const size_t n7 = pgo_ld(n_poses);
  L.steps = (int)((n7 + NB - 1) / NB);
  L.tiles = L.steps + 1;
  size_t o = 0;
  L.A = o;      o = align256(o + 8 * n7 * n7);
  L.b = o;      o = align256(o + 8 * n7);
  L.tbuf = o;   o = align256(o + 8 * NB);
  L.ticket = o; o = align256(o + 16);
  L.info = o;   o = align256(o + 16);
  L.nact = o;   o = align256(o + 4 * (size_t)(L.steps + 1));
  const size_t zeroed = o;                       // everything up to here is cleared at the start of a call
  (void)zeroed;
  L.rdiag = o;  o = align256(o + 8 * n7);
  L.winv = o;   o = align256(o + 8 * (size_t)L.steps * NB * NB);
  L.active = o; o = align256(o + 4 * (size_t)L.steps * L.tiles);
  L.total = o;
  return L;
}

// grid = r (one CTA per residual block), block = 256: one output per thread
__global__ void __launch_bounds__(256) pgo_assemble_kernel(const float* __restrict__ Ji, const float* __restrict__ Jj,
                                                           const int64_t* __restrict__ ii, const int64_t* __restrict__ jj,
                                                           const float* __restrict__ res, int n_poses, size_t ld, double* A, double* b) {
  __shared__ double sJi[49], sJj[49], sv[7];
  const int x = blockIdx.x, tid = threadIdx.x;
  const int64_t i = ii[x], j = jj[x];
  if (tid < 49) { sJi[tid] = (double)Ji[(size_t)x * 49 + tid]; sJj[tid] = (double)Jj[(size_t)x * 49 + tid]; }
  if (tid < 7) sv[tid] = (double)res[(size_t)x * 7 + tid];
  __syncthreads();
  if (i < 0 || j < 0 || i >= n_poses || j >= n_poses) return;
  const size_t n7 = ld;
  if (tid < 147) {
    const int blk = tid / 49, e = tid - blk * 49, a = e / 7, c = e - a * 7;
    // blk 0: (i,i) += Ji^T Ji ; blk 1: (j,j) += Jj^T Jj ; blk 2: the off-diagonal block in the LOWER triangle
    const double* L = (blk == 1) ? sJj : sJi;
    const double* Rm = (blk == 0) ? sJi : sJj;
    int64_t rb = (blk == 1) ? j : i, cb = (blk == 0) ? i : j;
    const double* Lm = L;
    if (blk == 2 && j > i) { rb = j; cb = i; Lm = sJj; Rm = sJi; }      // (j,i) += Jj^T Ji
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) acc += Lm[k * 7 + a] * Rm[k * 7 + c];
    atomicAdd(&A[(size_t)(rb * 7 + a) * n7 + cb * 7 + c], acc);
  } else if (tid < 161) {
    const int t = tid - 147, side = t / 7, a = t - side * 7;
    const double* Jm = side ? sJj : sJi;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) acc += Jm[k * 7 + a] * sv[k];
    atomicAdd(&b[(side ? j : i) * 7 + a], -acc);
  }
}

__global__ void pgo_damp_kernel(double* A, size_t n7, size_t ld, double lm, double ep) {
  const size_t d = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d < n7) {
    double* p = A + d * ld + d;
    const double v = *p;
    *p = (v + v * lm) + ep;              // A.diagonal() += A.diagonal() * lm; A.diagonal().array() += ep  (ba.cpp:164-165)
  }
}

// rows [m, mp) of the solved block become identity rows with a zero right-hand side (lower triangle only is read)
__global__ void pgo_pad_kernel(double* A, double* b, size_t ld, size_t m, size_t mp) {
  const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t rows = mp - m;
  if (x >= rows * mp) return;
  const size_t d = m + x / mp, c = x % mp;
  if (c <= d) A[d * ld + c] = (c == d) ? 1.0 : 0.0;
  if (c == 0) b[d] = 0.0;
}

__global__ void pgo_finish_kernel(const double* y, size_t m, size_t n7, float* delta) {
  const size_t d = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d < n7) delta[d] = d < m ? (float)y[d] : 0.0f;
}

}  // namespace pgba

using namespace pgba;

extern "C" {

int pgba_pgo_workspace_bytes(int64_t n_poses, size_t* bytes) {
  if (!bytes) return PGBA_ERR_NULL;
  if (n_poses <= 0 || n_poses > 100000) return PGBA_ERR_SHAPE;
  *bytes = pgo_layout(n_poses).total;
  return PGBA_OK;
}

int pgba_pgo_solve(const float* J_Ginv_i, const float* J_Ginv_j, const int64_t* ii, const int64_t* jj, const float* res,
                   int64_t n_res, int64_t n_poses, float ep, float lm, int freen, float* delta, int32_t* info,
                   void* workspace, size_t workspace_bytes, pgba_stream_t stream) {
  if (!J_Ginv_i || !J_Ginv_j || !ii || !jj || !res || !delta) return PGBA_ERR_NULL;
  if (n_res < 0 || n_poses <= 0 || n_poses > 100000 || n_res >= ((int64_t)1 << 31)) return PGBA_ERR_SHAPE;
  if (!workspace || ((uintptr_t)workspace & 255)) return PGBA_ERR_WORKSPACE;
  const PgoLayout L = pgo_layout(n_poses);
  if (L.total > workspace_bytes) return PGBA_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  const size_t n7 = 7 * (size_t)n_poses, ld = pgo_ld(n_poses);
  const int64_t m64 = (int64_t)freen * 7;
  const size_t m = (m64 < 0) ? n7 : ((size_t)m64 < n7 ? (size_t)m64 : n7);
  const size_t mp = ((m + 5) / 6) * 6;                                // padded order of the solved block (<= ld)
  double* A = (double*)(ws + L.A);
  double* b = (double*)(ws + L.b);
  cudaError_t e = cudaMemsetAsync(ws, 0, L.rdiag, s);                 // A, b, tbuf, ticket, info, nact
  if (e != cudaSuccess) return (int)e;
  if (n_res > 0) {
    pgo_assemble_kernel<<<(unsigned)n_res, 256, 0, s>>>(J_Ginv_i, J_Ginv_j, ii, jj, res, (int)n_poses, ld, A, b);
    count_launch();
  }
  pgo_damp_kernel<<<(unsigned)((n7 + 255) / 256), 256, 0, s>>>(A, n7, ld, (double)lm, (double)ep);
  count_launch();
  if (m > 0) {
    if (mp > m) {
      pgo_pad_kernel<<<(unsigned)(((mp - m) * mp + 255) / 256), 256, 0, s>>>(A, b, ld, m, mp);
      count_launch();
    }
    PgoSys sys;
    sys.s.S = A; sys.s.y = b; sys.s.n = (int)mp; sys.s.ld = (int)ld;
    sys.s.rdiag = (double*)(ws + L.rdiag);
    sys.s.winv = (double*)(ws + L.winv);
    sys.s.active = (int*)(ws + L.active);
    sys.s.nact = (int*)(ws + L.nact);
    sys.s.tbuf = (double*)(ws + L.tbuf);
    sys.s.ticket = (int*)(ws + L.ticket);
    sys.s.big_tiles = L.tiles;
    sys.s.chol_info = (int*)(ws + L.info);
    e = launch_big_chol(sys, (int)mp, 1, s);
    if (e != cudaSuccess) return (int)e;
  }
  pgo_finish_kernel<<<(unsigned)((n7 + 255) / 256), 256, 0, s>>>(b, m, n7, delta);
  count_launch();
  if (info) {
    e = cudaMemcpyAsync(info, ws + L.info, sizeof(int32_t), cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return (int)e;
  }
  return (int)cudaGetLastError();
}

}  // extern "C"
