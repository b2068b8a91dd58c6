// Patch correlation lookup and patch gather for sm_100a (see include/pcorr.h).
//
//   corr_forward_kernel       one CTA per (edge, batch): the patch feature fmap1[ii] is staged in shared memory as
//                             fp32, every thread owns window taps (pixel, y-off, x-off) and walks the channels with
//                             x-contiguous loads of fmap2[jj]; the (2R+2)^2 volume stays in shared memory, the 4-tap
//                             bilinear blend of the reference's host code (correlation_kernel.cu:221-230) and its
//                             final permute (:232) are fused, and the result is written once, contiguously, in the
//                             final [x-off][y-off][P][P] order (optionally interleaving two pyramid levels, which is
//                             what slam.py:321-323 builds with two calls + torch.stack).
//   corr_backward_kernel      scatter of the blended gradient into fmap1 / fmap2 gradients (training path)
//   patchify_forward/backward gather / scatter of (2R+2)^2 windows (correlation_kernel.cu:17-80)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pcorr.h"

namespace pgba { void count_launch(); }

namespace pcorr {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

struct CorrLevel {
  const void* fmap2;
  int H2, W2;
  float inv_scale;    // coords are multiplied by this (1 for level 0, 0.25 for level 1: slam.py:321-322)
};

// smem: f1 [C*PP] floats | vol [PP*D*D] floats | cx, cy [PP] floats
template <typename T, int NLEV>
__global__ void __launch_bounds__(256) corr_forward_kernel(const T* __restrict__ fmap1, CorrLevel lv0, CorrLevel lv1,
                                                          const float* __restrict__ coords,
                                                          const int64_t* __restrict__ us,
                                                          const int64_t* __restrict__ vs, int64_t E, int64_t K,
                                                          int64_t F, int C, int P, int R, T* __restrict__ out) {
  extern __shared__ float sm[];
  const int PP = P * P, D = 2 * R + 2, DD = D * D, Do = D - 1;
  float* f1 = sm;
  float* vol = f1 + C * PP;
  float* sx = vol + PP * DD;
  float* sy = sx + PP;
  const int64_t m = blockIdx.x;
  const int b = blockIdx.y;
  const int tid = threadIdx.x, T_ = blockDim.x;
  const int64_t ix = us[m], jx = vs[m];

  const T* f1g = fmap1 + ((int64_t)b * K + ix) * C * PP;
  for (int x = tid; x < C * PP; x += T_) f1[x] = to_f<T>(f1g[x]);
  const float* cg = coords + ((int64_t)b * E + m) * 2 * PP;
  __syncthreads();

#pragma unroll
  for (int lev = 0; lev < NLEV; ++lev) {
    const CorrLevel lv = lev == 0 ? lv0 : lv1;
    const int H2 = lv.H2, W2 = lv.W2;
    const int64_t plane = (int64_t)H2 * W2;
    const T* f2g = (const T*)lv.fmap2 + ((int64_t)b * F + jx) * C * plane;
    if (tid < PP) {
      // level > 0 uses coords / 4 computed in fp32 exactly like the caller's `coords / 4` (slam.py:322)
      sx[tid] = lev == 0 ? cg[tid] : cg[tid] * lv.inv_scale;
      sy[tid] = lev == 0 ? cg[PP + tid] : cg[PP + tid] * lv.inv_scale;
    }
    __syncthreads();
    for (int o = tid; o < PP * DD; o += T_) {
      const int p = o / DD, pos = o - p * DD;
      const int io = pos / D, jo = pos - io * D;
      const int i1 = (int)floorf(sy[p]) + (io - R);
      const int j1 = (int)floorf(sx[p]) + (jo - R);
      float acc = 0.f;
      if (i1 >= 0 && i1 < H2 && j1 >= 0 && j1 < W2) {
        const T* src = f2g + (int64_t)i1 * W2 + j1;
        const float* a = f1 + p;
#pragma unroll 8
        for (int c = 0; c < C; ++c) acc += a[c * PP] * to_f<T>(src[c * plane]);
      }
      vol[o] = acc;
    }
    __syncthreads();
    // bilinear blend + permute: out[m][xo][yo][i0][j0] (correlation_kernel.cu:221-232)
    T* og = out + ((int64_t)b * E + m) * (int64_t)(Do * Do * PP) * NLEV;
    for (int o = tid; o < Do * Do * PP; o += T_) {
      const int p = o % PP;
      const int yo = (o / PP) % Do;
      const int xo = o / (PP * Do);
      const float dx = sx[p] - floorf(sx[p]), dy = sy[p] - floorf(sy[p]);
      const float* v = vol + p * DD + yo * D + xo;
      const float r = (1.f - dx) * (1.f - dy) * v[0] + dx * (1.f - dy) * v[1] + (1.f - dx) * dy * v[D] +
                      dx * dy * v[D + 1];
      og[(int64_t)o * NLEV + lev] = from_f<T>(r);
    }
    __syncthreads();
  }
}

// Gradient scatter: one CTA per (edge, batch).  grad [B,E,Do(x),Do(y),P,P] f32.
template <typename T>
__global__ void __launch_bounds__(256) corr_backward_kernel(const T* __restrict__ fmap1, const T* __restrict__ fmap2,
                                                           const float* __restrict__ coords,
                                                           const int64_t* __restrict__ us,
                                                           const int64_t* __restrict__ vs,
                                                           const float* __restrict__ grad, int64_t E, int64_t K,
                                                           int64_t F, int C, int H2, int W2, int P, int R,
                                                           T* __restrict__ g1, T* __restrict__ g2) {
  extern __shared__ float sm[];
  const int PP = P * P, D = 2 * R + 2, DD = D * D, Do = D - 1;
  float* gv = sm;            // [PP*DD] gradient w.r.t. the raw volume
  float* sx = gv + PP * DD;
  float* sy = sx + PP;
  const int64_t m = blockIdx.x;
  const int b = blockIdx.y, tid = threadIdx.x, T_ = blockDim.x;
  const int64_t ix = us[m], jx = vs[m];
  const float* cg = coords + ((int64_t)b * E + m) * 2 * PP;
  if (tid < PP) { sx[tid] = cg[tid]; sy[tid] = cg[PP + tid]; }
  __syncthreads();
  const float* gg = grad + ((int64_t)b * E + m) * (int64_t)(Do * Do * PP);
  // transpose of the blend (correlation_kernel.cu:252-269)
  for (int o = tid; o < PP * DD; o += T_) {
    const int p = o / DD, pos = o - p * DD;
    const int io = pos / D, jo = pos - io * D;
    const float dx = sx[p] - floorf(sx[p]), dy = sy[p] - floorf(sy[p]);
    float acc = 0.f;
    auto G = [&](int yo, int xo) { return gg[((int64_t)xo * Do + yo) * PP + p]; };
    if (io < Do && jo < Do) acc += (1.f - dx) * (1.f - dy) * G(io, jo);
    if (io < Do && jo >= 1) acc += dx * (1.f - dy) * G(io, jo - 1);
    if (io >= 1 && jo < Do) acc += (1.f - dx) * dy * G(io - 1, jo);
    if (io >= 1 && jo >= 1) acc += dx * dy * G(io - 1, jo - 1);
    gv[o] = acc;
  }
  __syncthreads();
  const int64_t plane = (int64_t)H2 * W2;
  const T* f1g = fmap1 + ((int64_t)b * K + ix) * C * PP;
  const T* f2g = fmap2 + ((int64_t)b * F + jx) * C * plane;
  T* g1g = g1 + ((int64_t)b * K + ix) * C * PP;
  T* g2g = g2 + ((int64_t)b * F + jx) * C * plane;
  for (int o = tid; o < PP * DD; o += T_) {
    const int p = o / DD, pos = o - p * DD;
    const int io = pos / D, jo = pos - io * D;
    const int i1 = (int)floorf(sy[p]) + (io - R);
    const int j1 = (int)floorf(sx[p]) + (jo - R);
    if (i1 < 0 || i1 >= H2 || j1 < 0 || j1 >= W2) continue;
    const float g = gv[o];
    const int64_t off2 = (int64_t)i1 * W2 + j1;
    for (int c = 0; c < C; ++c) {
      atomicAdd(&g1g[c * PP + p], from_f<T>(g * to_f<T>(f2g[c * plane + off2])));
      atomicAdd(&g2g[c * plane + off2], from_f<T>(g * to_f<T>(f1g[c * PP + p])));
    }
  }
}

// one CTA per (patch m, batch b); threads over (c, i, j)
template <typename T>
__global__ void __launch_bounds__(256) patchify_forward_kernel(const T* __restrict__ net,
                                                              const float* __restrict__ coords, int64_t M, int C,
                                                              int H, int W, int R, T* __restrict__ patches) {
  const int D = 2 * R + 2, DD = D * D;
  const int64_t m = blockIdx.x;
  const int b = blockIdx.y;
  const float x = coords[((int64_t)b * M + m) * 2], y = coords[((int64_t)b * M + m) * 2 + 1];
  const int fy = (int)floorf(y), fx = (int)floorf(x);
  const T* ng = net + (int64_t)b * C * H * W;
  T* pg = patches + ((int64_t)b * M + m) * C * DD;
  for (int o = threadIdx.x; o < C * DD; o += blockDim.x) {
    const int c = o / DD, pos = o - c * DD;
    const int i = fy + (pos / D - R), j = fx + (pos % D - R);
    T v = from_f<T>(0.f);
    if (i >= 0 && i < H && j >= 0 && j < W) v = ng[((int64_t)c * H + i) * W + j];
    pg[o] = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) patchify_backward_kernel(const T* __restrict__ pgrad,
                                                               const float* __restrict__ coords, int64_t M, int C,
                                                               int H, int W, int R, T* __restrict__ ngrad) {
  const int D = 2 * R + 2, DD = D * D;
  const int64_t m = blockIdx.x;
  const int b = blockIdx.y;
  const float x = coords[((int64_t)b * M + m) * 2], y = coords[((int64_t)b * M + m) * 2 + 1];
  const int fy = (int)floorf(y), fx = (int)floorf(x);
  T* ng = ngrad + (int64_t)b * C * H * W;
  const T* pg = pgrad + ((int64_t)b * M + m) * C * DD;
  for (int o = threadIdx.x; o < C * DD; o += blockDim.x) {
    const int c = o / DD, pos = o - c * DD;
    const int i = fy + (pos / D - R), j = fx + (pos % D - R);
    if (i >= 0 && i < H && j >= 0 && j < W) atomicAdd(&ng[((int64_t)c * H + i) * W + j], pg[o]);
  }
}

template <typename T, int NLEV>
static int launch_corr(const void* fmap1, CorrLevel l0, CorrLevel l1, const float* coords, const int64_t* ii,
                       const int64_t* jj, int B, int64_t E, int64_t K, int64_t F, int C, int P, int R, void* out,
                       cudaStream_t s) {
  const int PP = P * P, D = 2 * R + 2;
  const size_t smem = sizeof(float) * ((size_t)C * PP + (size_t)PP * D * D + 2 * PP);
  if (smem > 200 * 1024) return PCORR_ERR_UNSUPPORTED;
  auto kern = corr_forward_kernel<T, NLEV>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<dim3((unsigned)E, (unsigned)B), 256, smem, s>>>((const T*)fmap1, l0, l1, coords, ii, jj, E, K, F, C, P, R,
                                                         (T*)out);
  pgba::count_launch();
  return (int)cudaGetLastError();
}

}  // namespace pcorr

using namespace pcorr;

extern "C" {

int pcorr_forward(const void* fmap1, const void* fmap2, const float* coords, const int64_t* ii, const int64_t* jj,
                  int B, int64_t E, int64_t K, int64_t F, int C, int H2, int W2, int P, int radius, int dtype,
                  void* out, pcorr_stream_t stream) {
  if (E == 0 || B == 0) return PCORR_OK;
  if (!fmap1 || !fmap2 || !coords || !ii || !jj || !out) return PCORR_ERR_NULL;
  if (B < 0 || E < 0 || K <= 0 || F <= 0 || C <= 0 || H2 <= 0 || W2 <= 0 || P <= 0 || radius < 0) return PCORR_ERR_SHAPE;
  if (E >= (int64_t)1 << 31 || B > 65535) return PCORR_ERR_UNSUPPORTED;
  CorrLevel l0{fmap2, H2, W2, 1.f}, l1{nullptr, 0, 0, 1.f};
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == PCORR_F32) return launch_corr<float, 1>(fmap1, l0, l1, coords, ii, jj, B, E, K, F, C, P, radius, out, s);
  if (dtype == PCORR_F16) return launch_corr<__half, 1>(fmap1, l0, l1, coords, ii, jj, B, E, K, F, C, P, radius, out, s);
  return PCORR_ERR_DTYPE;
}

int pcorr_forward_pyramid2(const void* fmap1, const void* fmap2_l0, const void* fmap2_l1, const float* coords,
                           const int64_t* ii, const int64_t* jj, int B, int64_t E, int64_t K, int64_t F, int C,
                           int H0, int W0, int H1, int W1, int P, int radius, int dtype, void* out,
                           pcorr_stream_t stream) {
  if (E == 0 || B == 0) return PCORR_OK;
  if (!fmap1 || !fmap2_l0 || !fmap2_l1 || !coords || !ii || !jj || !out) return PCORR_ERR_NULL;
  if (B < 0 || E < 0 || K <= 0 || F <= 0 || C <= 0 || H0 <= 0 || W0 <= 0 || H1 <= 0 || W1 <= 0 || P <= 0 || radius < 0)
    return PCORR_ERR_SHAPE;
  if (E >= (int64_t)1 << 31 || B > 65535) return PCORR_ERR_UNSUPPORTED;
  CorrLevel l0{fmap2_l0, H0, W0, 1.f}, l1{fmap2_l1, H1, W1, 0.25f};
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == PCORR_F32) return launch_corr<float, 2>(fmap1, l0, l1, coords, ii, jj, B, E, K, F, C, P, radius, out, s);
  if (dtype == PCORR_F16) return launch_corr<__half, 2>(fmap1, l0, l1, coords, ii, jj, B, E, K, F, C, P, radius, out, s);
  return PCORR_ERR_DTYPE;
}

int pcorr_backward(const void* fmap1, const void* fmap2, const float* coords, const int64_t* ii, const int64_t* jj,
                   const float* grad, int B, int64_t E, int64_t K, int64_t F, int C, int H2, int W2, int P,
                   int radius, int dtype, void* fmap1_grad, void* fmap2_grad, pcorr_stream_t stream) {
  if (E == 0 || B == 0) return PCORR_OK;
  if (!fmap1 || !fmap2 || !coords || !ii || !jj || !grad || !fmap1_grad || !fmap2_grad) return PCORR_ERR_NULL;
  if (B < 0 || E < 0 || K <= 0 || F <= 0 || C <= 0 || H2 <= 0 || W2 <= 0 || P <= 0 || radius < 0) return PCORR_ERR_SHAPE;
  if (E >= (int64_t)1 << 31 || B > 65535) return PCORR_ERR_UNSUPPORTED;
  const int PP = P * P, D = 2 * radius + 2;
  const size_t smem = sizeof(float) * ((size_t)PP * D * D + 2 * PP);
  if (smem > 48 * 1024) return PCORR_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid((unsigned)E, (unsigned)B);
  if (dtype == PCORR_F32)
    corr_backward_kernel<float><<<grid, 256, smem, s>>>((const float*)fmap1, (const float*)fmap2, coords, ii, jj, grad,
                                                        E, K, F, C, H2, W2, P, radius, (float*)fmap1_grad,
                                                        (float*)fmap2_grad);
  else if (dtype == PCORR_F16)
    corr_backward_kernel<__half><<<grid, 256, smem, s>>>((const __half*)fmap1, (const __half*)fmap2, coords, ii, jj,
                                                         grad, E, K, F, C, H2, W2, P, radius, (__half*)fmap1_grad,
                                                         (__half*)fmap2_grad);
  else
    return PCORR_ERR_DTYPE;
  pgba::count_launch();
  return (int)cudaGetLastError();
}

int pcorr_patchify_forward(const void* net, const float* coords, int B, int64_t M, int C, int H, int W, int radius,
                           int dtype, void* patches, pcorr_stream_t stream) {
  if (M == 0 || B == 0) return PCORR_OK;
  if (!net || !coords || !patches) return PCORR_ERR_NULL;
  if (B < 0 || M < 0 || C <= 0 || H <= 0 || W <= 0 || radius < 0) return PCORR_ERR_SHAPE;
  if (M >= (int64_t)1 << 31 || B > 65535) return PCORR_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid((unsigned)M, (unsigned)B);
  const int D = 2 * radius + 2;
  const int threads = (C * D * D >= 256) ? 256 : ((C * D * D + 31) / 32) * 32;
  if (dtype == PCORR_F32)
    patchify_forward_kernel<float><<<grid, threads, 0, s>>>((const float*)net, coords, M, C, H, W, radius, (float*)patches);
  else if (dtype == PCORR_F16)
    patchify_forward_kernel<__half><<<grid, threads, 0, s>>>((const __half*)net, coords, M, C, H, W, radius, (__half*)patches);
  else
    return PCORR_ERR_DTYPE;
  pgba::count_launch();
  return (int)cudaGetLastError();
}

int pcorr_patchify_backward(const void* patch_grad, const float* coords, int B, int64_t M, int C, int H, int W,
                            int radius, int dtype, void* net_grad, pcorr_stream_t stream) {
  if (M == 0 || B == 0) return PCORR_OK;
  if (!patch_grad || !coords || !net_grad) return PCORR_ERR_NULL;
  if (B < 0 || M < 0 || C <= 0 || H <= 0 || W <= 0 || radius < 0) return PCORR_ERR_SHAPE;
  if (M >= (int64_t)1 << 31 || B > 65535) return PCORR_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid((unsigned)M, (unsigned)B);
  if (dtype == PCORR_F32)
    patchify_backward_kernel<float><<<grid, 256, 0, s>>>((const float*)patch_grad, coords, M, C, H, W, radius, (float*)net_grad);
  else if (dtype == PCORR_F16)
    patchify_backward_kernel<__half><<<grid, 256, 0, s>>>((const __half*)patch_grad, coords, M, C, H, W, radius, (__half*)net_grad);
  else
    return PCORR_ERR_DTYPE;
  pgba::count_launch();
  return (int)cudaGetLastError();
}

}  // extern "C"
