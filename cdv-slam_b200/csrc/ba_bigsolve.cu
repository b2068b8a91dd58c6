// Large dense solve for the global (loop-closure) bundle adjustment: 6N > 156, e.g. BASELINE config c4 with
// N = 999 free poses (S is 5994 x 5994).  Replaces at::linalg_cholesky_ex + torch::cholesky_solve on the dense S of
// the reference (cdvslam/fastba/ba_cuda.cu:575-578 for eff_impl, :589-591 dense).  The factorisation itself is the
// tile-sparse blocked Cholesky of big_chol.cuh on the fp32 S held in the workspace, in place; this file adds the
// damping (ba_cuda.cu:575/589) and the pose retraction (ba_cuda.cu:88-206).
#include "big_chol.cuh"

namespace pgba {

// System provider of the BA workspace: window w -> pointers (see ba_common.cuh: Layout)
struct BaBigSys {
  using T = float;
  Problem pb;
  __device__ BigSys<float> get(int w) const {
    const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
    char* base = (char*)pb.ws + pb.L.body0 + (size_t)w * pb.L.body_bytes;
    char* z = (char*)pb.ws + (size_t)w * pb.L.zero_bytes;
    BigSys<float> b;
    b.S = wp.S; b.y = wp.y;
    b.n = 6 * (pb.t1 - pb.t0); b.ld = b.n;
    b.rdiag = (float*)(base + pb.L.o_rdiag);
    b.winv = (float*)(base + pb.L.o_winv);
    b.active = (int*)(base + pb.L.o_active);
    b.nact = (int*)(z + pb.L.z_nact);
    b.tbuf = (float*)(z + pb.L.z_bs);
    b.ticket = (int*)(z + pb.L.z_bs) + 64;
    b.big_tiles = pb.L.big_tiles;
    b.chol_info = &wp.hdr->chol_info;
    return b;
  }
};

// S += I * (1e-4 * S + 1)   (ba_cuda.cu:575/589)
__global__ void big_damp_kernel(Problem pb) {
  pdl_wait();
  pdl_trigger();
  const int w = blockIdx.y + pb.w0;
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  const int n6 = 6 * (pb.t1 - pb.t0);
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n6) {
    float* d = wp.S + (size_t)r * n6 + r;
    *d = *d + (1e-4f * *d + 1.0f);
  }
}

__global__ void big_finish_kernel(Problem pb) {
  pdl_wait();
  pdl_trigger();
  const int w = blockIdx.y + pb.w0;
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
  const unsigned B = (unsigned)batch;
  launch_k(big_damp_kernel, dim3((n6 + 255) / 256, B), dim3(256), 0, stream, pb);
  count_launch();
  BaBigSys sys{pb};
  cudaError_t e = launch_big_chol(sys, n6, batch, stream);
  if (e != cudaSuccess) return e;
  launch_k(big_finish_kernel, dim3((N + 127) / 128, B), dim3(128), 0, stream, pb);
  count_launch();
  return cudaGetLastError();
}

}  // namespace pgba
