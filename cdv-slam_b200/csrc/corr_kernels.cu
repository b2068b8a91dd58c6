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

// ---------------------------------------------------------------------------------------------------------------
// Staged lookup for the shape slam.py uses (P = 3, R = 3, C % 8 == 0): one WARP per (edge, level) task, persistent
// CTAs of 4 warps.  The 9 patch pixels land within a few pixels of each other, so their nine 8x8 windows live in one
// small region of fmap2[jj] (<= 12 x 12 pixels): the region is staged through shared memory 8 channels at a time
// (coalesced row segments, zero-filled outside the map), every lane owns <= 5 region pixels and accumulates the full
// [9 patch pixels x its region pixels] block of products with packed fp32x2 FMAs (FFMA2), i.e. each fmap2 value is
// fetched once per edge instead of once per (patch pixel, tap).  The per-pixel window selection, the bilinear blend
// (correlation_kernel.cu:221-230) and the final permute (:232) are done from the region volume in shared memory.
// Edges whose windows do not fit the 12 x 12 region (strong zoom) take the per-tap path inside the same kernel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int RG = 12;                 // region is RG x RG pixels
constexpr int RPX = RG * RG;           // 144
constexpr int CK = 8;                  // channels per staging chunk
constexpr int SW_WARPS = 4;
struct __align__(16) WarpSmem {
  float reg[CK][RPX];                  // staged region chunk (fp32)
  float f1[CK][12];                    // patch features of the chunk: [c][p] (9 used, padded to 12 for 16-byte rows)
  float vol[9][RPX + 4];               // region volume: vol[p][region pixel]; slow path: vol[p][64 taps]
  float4 wgt[9];                       // bilinear weights of pixel p: (1-dx)(1-dy), dx(1-dy), (1-dx)dy, dx dy
  float sx[9], sy[9];                  // coords of the 9 patch pixels at this level
  int wox[9], woy[9];                  // window origin of pixel p inside the region
  int vbase[9];                        // p * (RPX + 4) + woy * RG + wox
};

template <typename T, int NLEV>
__global__ void __launch_bounds__(32 * SW_WARPS) corr_staged_kernel(const T* __restrict__ fmap1, CorrLevel lv0, CorrLevel lv1,
                                                                   const float* __restrict__ coords,
                                                                   const int64_t* __restrict__ us,
                                                                   const int64_t* __restrict__ vs, int B, int64_t E,
                                                                   int64_t K, int64_t F, int C, T* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int R = 3, D = 8, Do = 7, PP = 9;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpSmem& S = reinterpret_cast<WarpSmem*>(smraw)[warp];
  const int64_t ntask = (int64_t)B * E * NLEV;
  // output o = lane + 32 t  ->  (p, yo, xo), the same for every task: packed p | (yo*RG + xo) << 8 | (yo*D + xo) << 16
  constexpr int NOUT = (7 * 7 * 9 + 31) / 32;                      // 14
  int odec[NOUT];
#pragma unroll
  for (int t = 0; t < NOUT; ++t) {
    const int o = lane + 32 * t;
    const int p = o % 9, yo = (o / 9) % 7, xo = o / 63;
    odec[t] = p | ((yo * RG + xo) << 8) | ((yo * 8 + xo) << 16);
  }
  for (int64_t task = (int64_t)blockIdx.x * SW_WARPS + warp; task < ntask; task += (int64_t)gridDim.x * SW_WARPS) {
    const int lev = (int)(task % NLEV);
    const int64_t be = task / NLEV;
    const int b = (int)(be / E);
    const int64_t m = be - (int64_t)b * E;
    const CorrLevel lv = (NLEV == 1 || lev == 0) ? lv0 : lv1;
    const int H2 = lv.H2, W2 = lv.W2;
    const int64_t plane = (int64_t)H2 * W2;
    const int64_t ix = us[m], jx = vs[m];
    const T* f1g = fmap1 + ((int64_t)b * K + ix) * C * PP;
    const T* f2g = (const T*)lv.fmap2 + ((int64_t)b * F + jx) * C * plane;
    const float* cg = coords + ((int64_t)b * E + m) * 2 * PP;
    __syncwarp();
    // ---- per-pixel coordinates, window origins, region bounding box
    int fxp = 0, fyp = 0;
    if (lane < PP) {
      // level > 0 uses coords / 4 computed in fp32 exactly like the caller's `coords / 4` (slam.py:322)
      const float x = (lev == 0) ? cg[lane] : cg[lane] * lv.inv_scale;
      const float y = (lev == 0) ? cg[PP + lane] : cg[PP + lane] * lv.inv_scale;
      S.sx[lane] = x; S.sy[lane] = y;
      // clamp far-out / non-finite coordinates to "entirely outside the map" (the window is then all zeros)
      const float xf = floorf(x), yf = floorf(y);
      fxp = (xf > -1e6f && xf < 1e6f) ? (int)xf : -1000000;
      fyp = (yf > -1e6f && yf < 1e6f) ? (int)yf : -1000000;
    }
    int xmin = (lane < PP) ? fxp : 0x7fffffff, xmax = (lane < PP) ? fxp : -0x7fffffff;
    int ymin = (lane < PP) ? fyp : 0x7fffffff, ymax = (lane < PP) ? fyp : -0x7fffffff;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
      ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    xmin = __shfl_sync(0xffffffffu, xmin, 0); xmax = __shfl_sync(0xffffffffu, xmax, 0);
    ymin = __shfl_sync(0xffffffffu, ymin, 0); ymax = __shfl_sync(0xffffffffu, ymax, 0);
    const int x0 = xmin - R, y0 = ymin - R;                      // region origin in the map
    const bool fits = (xmax - xmin + D <= RG) && (ymax - ymin + D <= RG);
    if (lane < PP) {
      S.wox[lane] = fxp - R - x0; S.woy[lane] = fyp - R - y0;
      S.vbase[lane] = lane * (RPX + 4) + (fyp - R - y0) * RG + (fxp - R - x0);
      const float xs = S.sx[lane], ys = S.sy[lane];
      const float dx = xs - floorf(xs), dy = ys - floorf(ys);
      S.wgt[lane] = make_float4((1.f - dx) * (1.f - dy), dx * (1.f - dy), (1.f - dx) * dy, dx * dy);
    }
    T* og = out + ((int64_t)b * E + m) * (int64_t)(Do * Do * PP) * NLEV;
    __syncwarp();

    if (fits) {
      // ---- which region pixels does this lane own: px = lane + 32 m
      int goff[5];
      bool gok[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const int px = lane + 32 * q;
        const int ry = px / RG, rx = px - ry * RG;
        const int gy = y0 + ry, gx = x0 + rx;
        gok[q] = (px < RPX) && gy >= 0 && gy < H2 && gx >= 0 && gx < W2;
        goff[q] = gok[q] ? gy * W2 + gx : 0;
      }
      float2 acc[5][5];                                          // [owned pixel][patch-pixel pair]; pair 4 = (p8, -)
#pragma unroll
      for (int q = 0; q < 5; ++q)
#pragma unroll
        for (int pp = 0; pp < 5; ++pp) acc[q][pp] = make_float2(0.f, 0.f);
      for (int c0 = 0; c0 < C; c0 += CK) {
        // stage: region chunk (each lane loads its own pixels: coalesced row segments) and the patch features
        {
          T tmp[CK][5];
#pragma unroll
          for (int c = 0; c < CK; ++c) {
            const T* src = f2g + (int64_t)(c0 + c) * plane;
#pragma unroll
            for (int q = 0; q < 5; ++q) tmp[c][q] = src[goff[q]];       // goff = 0 (always mapped) when masked
          }
#pragma unroll
          for (int c = 0; c < CK; ++c)
#pragma unroll
            for (int q = 0; q < 5; ++q) {
              const int px = lane + 32 * q;
              if (q < 4 || px < RPX) S.reg[c][px] = gok[q] ? to_f<T>(tmp[c][q]) : 0.f;
            }
        }
        for (int x = lane; x < CK * PP; x += 32) {
          const int c = x / PP, p = x - c * PP;
          S.f1[c][p] = to_f<T>(f1g[(c0 + c) * PP + p]);
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < CK; ++c) {
          const float4 fa = *reinterpret_cast<const float4*>(&S.f1[c][0]);
          const float4 fb = *reinterpret_cast<const float4*>(&S.f1[c][4]);
          const float f8 = S.f1[c][8];
          const float2 fp[5] = {make_float2(fa.x, fa.y), make_float2(fa.z, fa.w), make_float2(fb.x, fb.y),
                                make_float2(fb.z, fb.w), make_float2(f8, 0.f)};
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            const int px = lane + 32 * q;
            const float r = (px < RPX) ? S.reg[c][px] : 0.f;
            const float2 rr = make_float2(r, r);
#pragma unroll
            for (int pp = 0; pp < 5; ++pp) acc[q][pp] = __ffma2_rn(rr, fp[pp], acc[q][pp]);
          }
        }
        __syncwarp();
      }
      // ---- region volume to shared memory
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const int px = lane + 32 * q;
        if (px < RPX) {
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) { S.vol[2 * pp][px] = acc[q][pp].x; S.vol[2 * pp + 1][px] = acc[q][pp].y; }
          S.vol[8][px] = acc[q][4].x;
        }
      }
      __syncwarp();
      // ---- window selection + bilinear blend + permute: out[xo][yo][p] (correlation_kernel.cu:221-232)
      {
        const float* vol0 = &S.vol[0][0];
#pragma unroll
        for (int t = 0; t < NOUT; ++t) {
          const int o = lane + 32 * t;
          if (o >= Do * Do * PP) break;
          const int p = odec[t] & 0xff;
          const float4 wg = S.wgt[p];
          const float* v = vol0 + S.vbase[p] + ((odec[t] >> 8) & 0xff);
          const float r = wg.x * v[0] + wg.y * v[1] + wg.z * v[RG] + wg.w * v[RG + 1];
          og[(int64_t)o * NLEV + lev] = from_f<T>(r);
        }
      }
    } else {
      // ---- per-tap path (windows too far apart for one region): same arithmetic as corr_forward_kernel
      for (int o = lane; o < PP * D * D; o += 32) {
        const int p = o / (D * D), pos = o - p * (D * D);
        const int io = pos / D, jo = pos - io * D;
        const float xf = floorf(S.sx[p]), yf = floorf(S.sy[p]);
        float acc = 0.f;
        if (xf > -1e6f && xf < 1e6f && yf > -1e6f && yf < 1e6f) {
          const int i1 = (int)yf + (io - R), j1 = (int)xf + (jo - R);
          if (i1 >= 0 && i1 < H2 && j1 >= 0 && j1 < W2) {
            const T* src = f2g + (int64_t)i1 * W2 + j1;
            for (int c = 0; c < C; ++c) acc += to_f<T>(f1g[c * PP + p]) * to_f<T>(src[c * plane]);
          }
        }
        S.vol[p][pos] = acc;
      }
      __syncwarp();
#pragma unroll
      for (int t = 0; t < NOUT; ++t) {
        const int o = lane + 32 * t;
        if (o >= Do * Do * PP) break;
        const int p = odec[t] & 0xff;
        const float xs = S.sx[p], ys = S.sy[p];
        const float dx = xs - floorf(xs), dy = ys - floorf(ys);
        const float* v = &S.vol[p][(odec[t] >> 16) & 0xff];
        const float r = (1.f - dx) * (1.f - dy) * v[0] + dx * (1.f - dy) * v[1] + (1.f - dx) * dy * v[D] +
                        dx * dy * v[D + 1];
        og[(int64_t)o * NLEV + lev] = from_f<T>(r);
      }
    }
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

// The modes the reference applies in python on the gathered (2R+2)^2 window (cdvslam/altcorr/correlation.py:51-71), fused
// into the gather.  'bilinear': the four (2R+1)^2 sub-windows blended with weights from coords - floor(coords), evaluated
// the way torch evaluates the reference's expression -- float32 weights (the coordinates are float32), the half window
// promoted to float32, separate (unfused) multiplies and adds in the same order -- so the result is float32 and equals
// the reference's bit for bit.  'upperleft': element [0][0] of the window, in the map's dtype.
template <typename T>
__global__ void __launch_bounds__(256) patchify_bilinear_kernel(const T* __restrict__ net,
                                                               const float* __restrict__ coords, int64_t M, int C,
                                                               int H, int W, int R, float* __restrict__ out) {
  const int d = 2 * R + 1, dd = d * d;
  const int64_t m = blockIdx.x;
  const int b = blockIdx.y;
  const float x = coords[((int64_t)b * M + m) * 2], y = coords[((int64_t)b * M + m) * 2 + 1];
  const float flx = floorf(x), fly = floorf(y);
  const int fy = (int)fly, fx = (int)flx;
  const float dx = __fsub_rn(x, flx), dy = __fsub_rn(y, fly);
  const float w00 = __fmul_rn(__fsub_rn(1.f, dy), __fsub_rn(1.f, dx)), w01 = __fmul_rn(__fsub_rn(1.f, dy), dx);
  const float w10 = __fmul_rn(dy, __fsub_rn(1.f, dx)), w11 = __fmul_rn(dy, dx);
  const T* ng = net + (int64_t)b * C * H * W;
  float* og = out + ((int64_t)b * M + m) * C * dd;
  for (int o = threadIdx.x; o < C * dd; o += blockDim.x) {
    const int c = o / dd, pos = o - c * dd;
    const int i = fy + (pos / d - R), j = fx + (pos % d - R);
    const T* plane = ng + (int64_t)c * H * W;
    auto at = [&](int ii, int jj) -> float {
      return (ii >= 0 && ii < H && jj >= 0 && jj < W) ? to_f<T>(plane[(int64_t)ii * W + jj]) : 0.f;
    };
    float v = __fmul_rn(w00, at(i, j));
    v = __fadd_rn(v, __fmul_rn(w01, at(i, j + 1)));
    v = __fadd_rn(v, __fmul_rn(w10, at(i + 1, j)));
    v = __fadd_rn(v, __fmul_rn(w11, at(i + 1, j + 1)));
    og[o] = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) patchify_upperleft_kernel(const T* __restrict__ net,
                                                                const float* __restrict__ coords, int64_t M, int C,
                                                                int H, int W, int R, T* __restrict__ out) {
  const int64_t m = blockIdx.x;
  const int b = blockIdx.y;
  const float x = coords[((int64_t)b * M + m) * 2], y = coords[((int64_t)b * M + m) * 2 + 1];
  const int i = (int)floorf(y) - R, j = (int)floorf(x) - R;
  const bool ok = i >= 0 && i < H && j >= 0 && j < W;
  const T* ng = net + (int64_t)b * C * H * W;
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    out[((int64_t)b * M + m) * C + c] = ok ? ng[((int64_t)c * H + i) * W + j] : from_f<T>(0.f);
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

// Adjoint of patchify_bilinear_kernel / patchify_upperleft_kernel w.r.t. the map: what autograd derives for the
// reference's python blend followed by patchify_backward_kernel (the four weighted copies of the gradient, cast to the
// map's dtype, scattered with atomics).  grad: float32 [B, M, C, d, d] (bilinear) or T [B, M, C] (upperleft).
template <typename T>
__global__ void __launch_bounds__(256) patchify_bilinear_backward_kernel(const float* __restrict__ grad,
                                                                        const float* __restrict__ coords, int64_t M,
                                                                        int C, int H, int W, int R,
                                                                        T* __restrict__ ngrad) {
  const int d = 2 * R + 1, dd = d * d;
  const int64_t m = blockIdx.x;
  const int b = blockIdx.y;
  const float x = coords[((int64_t)b * M + m) * 2], y = coords[((int64_t)b * M + m) * 2 + 1];
  const float flx = floorf(x), fly = floorf(y);
  const int fy = (int)fly, fx = (int)flx;
  const float dx = x - flx, dy = y - fly;
  const float w[4] = {(1.f - dy) * (1.f - dx), (1.f - dy) * dx, dy * (1.f - dx), dy * dx};
  T* ng = ngrad + (int64_t)b * C * H * W;
  const float* gg = grad + ((int64_t)b * M + m) * C * dd;
  for (int o = threadIdx.x; o < C * dd; o += blockDim.x) {
    const int c = o / dd, pos = o - c * dd;
    const int i = fy + (pos / d - R), j = fx + (pos % d - R);
    const float g = gg[o];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int ii = i + (t >> 1), jj = j + (t & 1);
      if (ii >= 0 && ii < H && jj >= 0 && jj < W) atomicAdd(&ng[((int64_t)c * H + ii) * W + jj], from_f<T>(w[t] * g));
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) patchify_upperleft_backward_kernel(const T* __restrict__ grad,
                                                                         const float* __restrict__ coords, int64_t M,
                                                                         int C, int H, int W, int R,
                                                                         T* __restrict__ ngrad) {
  const int64_t m = blockIdx.x;
  const int b = blockIdx.y;
  const float x = coords[((int64_t)b * M + m) * 2], y = coords[((int64_t)b * M + m) * 2 + 1];
  const int i = (int)floorf(y) - R, j = (int)floorf(x) - R;
  if (i < 0 || i >= H || j < 0 || j >= W) return;
  T* ng = ngrad + (int64_t)b * C * H * W;
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    atomicAdd(&ng[((int64_t)c * H + i) * W + j], grad[((int64_t)b * M + m) * C + c]);
}

template <typename T, int NLEV>
static int launch_corr(const void* fmap1, CorrLevel l0, CorrLevel l1, const float* coords, const int64_t* ii,
                       const int64_t* jj, int B, int64_t E, int64_t K, int64_t F, int C, int P, int R, void* out,
                       cudaStream_t s) {
  if (P == 3 && R == 3 && C % CK == 0) {
    const size_t smem = sizeof(WarpSmem) * SW_WARPS;
    auto kern = corr_staged_kernel<T, NLEV>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t tasks = (int64_t)B * E * NLEV;
    int64_t grid = (tasks + SW_WARPS - 1) / SW_WARPS;
    if (grid > 148 * 5) grid = 148 * 5;
    kern<<<(unsigned)grid, 32 * SW_WARPS, smem, s>>>((const T*)fmap1, l0, l1, coords, ii, jj, B, E, K, F, C, (T*)out);
    pgba::count_launch();
    return (int)cudaGetLastError();
  }
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

int pcorr_patchify_mode_forward(const void* net, const float* coords, int B, int64_t M, int C, int H, int W, int radius,
                                int mode, int dtype, void* out, pcorr_stream_t stream) {
  if (mode == PCORR_PATCH_RAW) return pcorr_patchify_forward(net, coords, B, M, C, H, W, radius, dtype, out, stream);
  if (mode != PCORR_PATCH_BILINEAR && mode != PCORR_PATCH_UPPERLEFT) return PCORR_ERR_SHAPE;
  if (M == 0 || B == 0) return PCORR_OK;
  if (!net || !coords || !out) return PCORR_ERR_NULL;
  if (B < 0 || M < 0 || C <= 0 || H <= 0 || W <= 0 || radius < 0) return PCORR_ERR_SHAPE;
  if (M >= (int64_t)1 << 31 || B > 65535) return PCORR_ERR_UNSUPPORTED;
  if (dtype != PCORR_F32 && dtype != PCORR_F16) return PCORR_ERR_DTYPE;
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid((unsigned)M, (unsigned)B);
  if (mode == PCORR_PATCH_BILINEAR) {
    const int d = 2 * radius + 1;
    const int threads = (C * d * d >= 256) ? 256 : ((C * d * d + 31) / 32) * 32;
    if (dtype == PCORR_F32)
      patchify_bilinear_kernel<float><<<grid, threads, 0, s>>>((const float*)net, coords, M, C, H, W, radius, (float*)out);
    else
      patchify_bilinear_kernel<__half><<<grid, threads, 0, s>>>((const __half*)net, coords, M, C, H, W, radius, (float*)out);
  } else {
    const int threads = (C >= 256) ? 256 : ((C + 31) / 32) * 32;
    if (dtype == PCORR_F32)
      patchify_upperleft_kernel<float><<<grid, threads, 0, s>>>((const float*)net, coords, M, C, H, W, radius, (float*)out);
    else
      patchify_upperleft_kernel<__half><<<grid, threads, 0, s>>>((const __half*)net, coords, M, C, H, W, radius, (__half*)out);
  }
  pgba::count_launch();
  return (int)cudaGetLastError();
}

int pcorr_patchify_mode_backward(const void* out_grad, const float* coords, int B, int64_t M, int C, int H, int W,
                                 int radius, int mode, int dtype, void* net_grad, pcorr_stream_t stream) {
  if (mode == PCORR_PATCH_RAW) return pcorr_patchify_backward(out_grad, coords, B, M, C, H, W, radius, dtype, net_grad, stream);
  if (mode != PCORR_PATCH_BILINEAR && mode != PCORR_PATCH_UPPERLEFT) return PCORR_ERR_SHAPE;
  if (M == 0 || B == 0) return PCORR_OK;
  if (!out_grad || !coords || !net_grad) return PCORR_ERR_NULL;
  if (B < 0 || M < 0 || C <= 0 || H <= 0 || W <= 0 || radius < 0) return PCORR_ERR_SHAPE;
  if (M >= (int64_t)1 << 31 || B > 65535) return PCORR_ERR_UNSUPPORTED;
  if (dtype != PCORR_F32 && dtype != PCORR_F16) return PCORR_ERR_DTYPE;
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid((unsigned)M, (unsigned)B);
  if (mode == PCORR_PATCH_BILINEAR) {
    if (dtype == PCORR_F32)
      patchify_bilinear_backward_kernel<float><<<grid, 256, 0, s>>>((const float*)out_grad, coords, M, C, H, W, radius, (float*)net_grad);
    else
      patchify_bilinear_backward_kernel<__half><<<grid, 256, 0, s>>>((const float*)out_grad, coords, M, C, H, W, radius, (__half*)net_grad);
  } else {
    if (dtype == PCORR_F32)
      patchify_upperleft_backward_kernel<float><<<grid, 128, 0, s>>>((const float*)out_grad, coords, M, C, H, W, radius, (float*)net_grad);
    else
      patchify_upperleft_backward_kernel<__half><<<grid, 128, 0, s>>>((const __half*)out_grad, coords, M, C, H, W, radius, (__half*)net_grad);
  }
  pgba::count_launch();
  return (int)cudaGetLastError();
}

}  // extern "C"
