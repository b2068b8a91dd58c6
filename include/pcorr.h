/*
 * pcorr.h -- C ABI of the B200-native patch correlation lookup / patch gather (libpgba.so).
 *
 * Drop-in boundary for the `cuda_corr` extension module of FrankYard/CDV-SLAM
 * (reference pybind table: cdvslam/altcorr/correlation.cpp:57-63; python: cdvslam/altcorr/correlation.py).
 * Plain device pointers and sizes; return value 0 = ok, negative = PCORR_ERR_*, positive = cudaError_t.
 * No host synchronisation, no allocation.
 */
#ifndef PCORR_H_
#define PCORR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* pcorr_stream_t; /* == cudaStream_t */

enum { PCORR_OK = 0, PCORR_ERR_NULL = -1, PCORR_ERR_SHAPE = -2, PCORR_ERR_DTYPE = -3, PCORR_ERR_UNSUPPORTED = -4 };
enum { PCORR_F32 = 0, PCORR_F16 = 1 };
enum { PCORR_PATCH_RAW = 0, PCORR_PATCH_BILINEAR = 1, PCORR_PATCH_UPPERLEFT = 2 };

/* Correlation lookup.  Replaces cuda_corr.forward == corr_cuda_forward() (reference:
 * cdvslam/altcorr/correlation_kernel.cu:83-136 kernel, :193-233 host incl. the bilinear blend and the final
 * permute(0,1,3,2,4,5)).
 *   fmap1  [B, K, C, P, P]      patch features (dtype f32 or f16)
 *   fmap2  [B, F, C, H2, W2]    frame feature maps, same dtype
 *   coords f32 [B, E, 2, P, P]  (ch0 = x, ch1 = y) reprojected pixel positions in fmap2's resolution
 *   ii i64 [E] (index into K), jj i64 [E] (index into F)
 *   out    [B, E, 2R+1 (x offset), 2R+1 (y offset), P, P], dtype of fmap1, CONTIGUOUS (the reference returns the
 *          same logical tensor as a permuted view).
 * Window taps outside the map contribute 0.  Accumulation is fp32 for both dtypes. */
int pcorr_forward(const void* fmap1, const void* fmap2, const float* coords, const int64_t* ii, const int64_t* jj,
                  int B, int64_t E, int64_t K, int64_t F, int C, int H2, int W2, int P, int radius, int dtype,
                  void* out, pcorr_stream_t stream);

/* Fused two-level lookup: what cdvslam/slam.py:316-323 does with two corr() calls + torch.stack(..., -1):
 * level 0 uses coords, level 1 uses coords / 4 on fmap2_l1; out [B, E, 2R+1, 2R+1, P, P, 2] contiguous. */
int pcorr_forward_pyramid2(const void* fmap1, const void* fmap2_l0, const void* fmap2_l1, const float* coords,
                           const int64_t* ii, const int64_t* jj, int B, int64_t E, int64_t K, int64_t F, int C,
                           int H0, int W0, int H1, int W1, int P, int radius, int dtype, void* out,
                           pcorr_stream_t stream);

/* TMA + tensor-core lookup for the production shapes (fp16 features, C in {24, 32, 128}, P = 3, radius = 3) -- the default
 * path of cuda_corr.forward for that shape.  The frame maps are first copied to a channel-last layout in the
 * caller-owned, 256-byte aligned workspace (pcorr_tma_workspace_bytes()), then one warp per edge fetches its region
 * of fmap2[jj] with one TMA tile load per pyramid level (out-of-map pixels are zero-filled by the hardware), contracts
 * it with the patch features on the tensor cores (fp32 accumulate) and writes the blended, permuted result once.
 * Same results layout as pcorr_forward (nlev = 1; fmap2_l1 may be NULL) / pcorr_forward_pyramid2 (nlev = 2).
 * Maps smaller than 12 x 12 pixels are not supported (PCORR_ERR_UNSUPPORTED; use pcorr_forward).
 * The host builds two CUtensorMap descriptors per call (no allocation, no synchronisation; graph-capturable). */
int pcorr_tma_supported(int C, int P, int radius, int dtype);
int pcorr_tma_workspace_bytes(int nlev, int B, int64_t F, int C, int H0, int W0, int H1, int W1,
                              size_t* bytes /* host, out */);
/* the same with the feature dtype stated (fp32 maps, C = 128: the 3xTF32 tile kernel needs a 4-byte channel-last copy) */
int pcorr_tma_workspace_bytes_dt(int nlev, int B, int64_t F, int C, int dtype, int H0, int W0, int H1, int W1,
                                 size_t* bytes /* host, out */);
int pcorr_forward_tma(const void* fmap1, const void* fmap2_l0, const void* fmap2_l1, const float* coords,
                      const int64_t* ii, const int64_t* jj, int nlev, int B, int64_t E, int64_t K, int64_t F, int C,
                      int H0, int W0, int H1, int W1, int P, int radius, int dtype, void* out, void* workspace,
                      size_t workspace_bytes, pcorr_stream_t stream);

/* Persistent channel-last mirror ("ring") of the frame-map pyramid.  pcorr_forward_tma copies ALL frame maps to the
 * channel-last layout on every call, although the SLAM front end rewrites exactly one ring slot per new frame
 * (cdvslam/slam.py:681-682: pyramid[0][:, n % mem] = fmap, pyramid[1][:, n % mem] = avg_pool2d(fmap, 4, 4)).  A caller
 * that owns a ring buffer (pcorr_tma_workspace_bytes() bytes, 256-byte aligned, contents kept between calls) instead
 *   - calls pcorr_ring_update(...) for the slots it has just written (frames [first_frame, first_frame + n_frames) of
 *     every batch entry; all frames once at start-up), and
 *   - looks up with pcorr_forward_ring(...), which is pcorr_forward_tma without the per-call copy (the NCHW maps are not
 *     read at all).  Same results bit for bit. */
int pcorr_ring_update(const void* fmap2_l0, const void* fmap2_l1, int nlev, int B, int64_t F, int C, int H0, int W0, int H1,
                      int W1, int64_t first_frame, int64_t n_frames, void* ring, size_t ring_bytes, pcorr_stream_t stream);
int pcorr_forward_ring(const void* fmap1, const float* coords, const int64_t* ii, const int64_t* jj, int nlev, int B,
                       int64_t E, int64_t K, int64_t F, int C, int H0, int W0, int H1, int W1, int P, int radius, int dtype,
                       void* out, const void* ring, size_t ring_bytes, pcorr_stream_t stream);

/* Gradient of pcorr_forward w.r.t. fmap1 and fmap2.  Replaces cuda_corr.backward == corr_cuda_backward()
 * (reference: correlation_kernel.cu:140-190, 236-286).  grad is the gradient of `out` in out's layout, f32;
 * fmap1_grad / fmap2_grad have the shapes and dtype of fmap1 / fmap2 and must be zero-filled by the caller. */
int pcorr_backward(const void* fmap1, const void* fmap2, const float* coords, const int64_t* ii, const int64_t* jj,
                   const float* grad, int B, int64_t E, int64_t K, int64_t F, int C, int H2, int W2, int P,
                   int radius, int dtype, void* fmap1_grad, void* fmap2_grad, pcorr_stream_t stream);

/* Patch gather.  Replaces cuda_corr.patchify_forward == patchify_cuda_forward() (reference:
 * correlation_kernel.cu:17-47, 288-307).  net [B, C, H, W]; coords f32 [B, M, 2] (x, y);
 * patches [B, M, C, D, D] with D = 2R+2, patches[b,m,c,i,j] = net[b,c,floor(y)+i-R,floor(x)+j-R] or 0 outside. */
int pcorr_patchify_forward(const void* net, const float* coords, int B, int64_t M, int C, int H, int W, int radius,
                           int dtype, void* patches, pcorr_stream_t stream);

/* Replaces cuda_corr.patchify_backward (reference: correlation_kernel.cu:50-80, 310-333): scatter-add of the
 * patch gradient [B, M, C, D, D] into net_grad [B, C, H, W] (zero-filled by the caller). */
int pcorr_patchify_backward(const void* patch_grad, const float* coords, int B, int64_t M, int C, int H, int W,
                            int radius, int dtype, void* net_grad, pcorr_stream_t stream);

/* altcorr.patchify(net, coords, radius, mode) in one kernel (reference: cdvslam/altcorr/correlation.py:51-71, which runs
 * patchify_cuda_forward and then blends / crops the (2R+2)^2 window with torch ops):
 *   PCORR_PATCH_RAW        out [B, M, C, 2R+2, 2R+2], dtype of net  (== pcorr_patchify_forward)
 *   PCORR_PATCH_BILINEAR   out f32 [B, M, C, 2R+1, 2R+1] (float32 also for f16 maps: torch promotes the half window against
 *                          the float32 weights); (1-dy)(1-dx) w[:d,:d] + (1-dy)dx w[:d,1:] + dy(1-dx) w[1:,:d] + dy dx w[1:,1:]
 *                          with unfused float32 multiplies / adds in that order, i.e. bit-identical to the reference
 *   PCORR_PATCH_UPPERLEFT  out [B, M, C, 1, 1], dtype of net: window element [0][0]
 * pcorr_patchify_mode_backward is the adjoint w.r.t. net (out_grad in out's layout and dtype; net_grad zero-filled by the
 * caller, dtype of net). */
int pcorr_patchify_mode_forward(const void* net, const float* coords, int B, int64_t M, int C, int H, int W, int radius,
                                int mode, int dtype, void* out, pcorr_stream_t stream);
int pcorr_patchify_mode_backward(const void* out_grad, const float* coords, int B, int64_t M, int C, int H, int W,
                                 int radius, int mode, int dtype, void* net_grad, pcorr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PCORR_H_ */
