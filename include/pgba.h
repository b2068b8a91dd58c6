/*
 * pgba.h -- C ABI of the B200-native patch-graph bundle adjustment (libpgba.so).
 *
 * Drop-in boundary for the `cuda_ba` extension module of FrankYard/CDV-SLAM
 * (reference pybind table: cdvslam/fastba/ba.cpp:183-188; python shim: cdvslam/fastba/ba.py:4-8).
 * Plain pointers and sizes only -- no torch types.  All data pointers are DEVICE pointers unless marked host.
 * Every entry point returns 0 on success, a negative PGBA_ERR_* for an invalid argument and a positive value
 * (a cudaError_t) for a CUDA failure; nothing throws, exits or synchronises the host, and nothing allocates:
 * temporaries live in a caller-owned workspace, so a whole call can be captured in a CUDA graph.
 *
 * Tensor layouts are the reference's:
 *   poses      f32 [n_pose_rows, 7]     (tx ty tz qx qy qz qw), world->camera, updated IN PLACE for rows t0..t1-1
 *   patches    f32 [n_patch_rows, 3, P, P]  ch0 = x, ch1 = y, ch2 = inverse depth, ch2 updated IN PLACE
 *   intrinsics f32 [>=1, 4]             (fx fy cx cy); only row 0 is used (reference: ba_cuda.cu:253-259)
 *   target     f32 [n_edges, 2]   weight f32 [n_edges, 2]   lmbda f32 [1]
 *   ii, jj, kk i64 [n_edges]            source frame, target frame, global patch id
 */
#ifndef PGBA_H_
#define PGBA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* pgba_stream_t; /* == cudaStream_t */

enum {
  PGBA_OK = 0,
  PGBA_ERR_NULL = -1,        /* a required pointer is NULL */
  PGBA_ERR_SHAPE = -2,       /* negative / inconsistent size, P < 2, t1 < t0, ... */
  PGBA_ERR_WORKSPACE = -3,   /* workspace too small or misaligned (needs 256-byte alignment) */
  PGBA_ERR_UNSUPPORTED = -4  /* configuration outside the implemented range (see pgba_ba_limits) */
};

/* Bits of the device-side status word (pgba_ba_status): set by kernels when the edge list cannot be processed. */
enum {
  PGBA_ST_INDEX_RANGE = 1,   /* an ii/jj/kk value is outside [0, n_pose_rows) / [0, n_patch_rows) */
  PGBA_ST_TOO_MANY_SLOTS = 2,/* a source frame sees more distinct target frames than PGBA_MAX_SLOTS */
  PGBA_ST_CAPACITY = 4,      /* internal table capacity exceeded (extremely sparse patch x frame incidence) */
  PGBA_ST_MIXED_SOURCE = 8   /* (informational) a patch id appears with two different source frames */
};

#define PGBA_MAX_SLOTS 128   /* distinct target frames per (source frame, 128-patch chunk) */
#define PGBA_MAX_POSE_ROWS 8192

const char* pgba_error_string(int code);
int pgba_version(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Bundle adjustment.  Replaces cuda_ba.forward == cuda_ba() (reference: cdvslam/fastba/ba_cuda.cu:462-611; the
 * block-sparse variant cdvslam/fastba/block_e.cu:43-300 selected there by eff_impl).  `iterations` Gauss-Newton
 * steps; per step: per-edge SE3 reprojection residual + Jacobians (ba_cuda.cu:265-343), assembly of the normal
 * equations (:346-403), Q = 1/(C+lmbda), Schur complement S = B - E Q E^T, damping S += I*(1e-4*S + 1), Cholesky
 * solve, dZ = Q (u - E^T dX) (:548-592), SE3 / inverse-depth retraction (:178-229).  t1 == t0 is the reference's
 * structure-only branch (:550-560).  `ppf` (the reference's PPF/M) and `eff_impl` are accepted for signature
 * parity; the implementation always uses its own block-sparse layout, which is valid for both settings.
 *
 * Batched form: `batch` independent windows with identical shapes, tensor b at base + b*stride (strides in
 * ELEMENTS of the tensor's dtype; a stride of 0 shares the tensor, e.g. intrinsics or lmbda).  n_edges_dev, when
 * not NULL, is an i32 [batch] device array of per-window edge counts (<= n_edges) for ragged batches.
 * ------------------------------------------------------------------------------------------------------------- */
int pgba_ba_workspace_bytes(int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows, int t0, int t1,
                            int64_t batch, size_t* bytes /* host, out */);

int pgba_ba_solve(float* poses, float* patches, const float* intrinsics, const float* target, const float* weight,
                  const float* lmbda, const int64_t* ii, const int64_t* jj, const int64_t* kk, int64_t n_edges,
                  int64_t n_pose_rows, int64_t n_patch_rows, int P, int ppf, int t0, int t1, int iterations,
                  int eff_impl, void* workspace, size_t workspace_bytes, pgba_stream_t stream);

typedef struct {
  int64_t poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk; /* per-window strides, in elements */
} pgba_strides;

int pgba_ba_solve_batched(float* poses, float* patches, const float* intrinsics, const float* target,
                          const float* weight, const float* lmbda, const int64_t* ii, const int64_t* jj,
                          const int64_t* kk, const int32_t* n_edges_dev, const pgba_strides* strides /* host */,
                          int64_t batch, int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows, int P, int ppf,
                          int t0, int t1, int iterations, int eff_impl, void* workspace, size_t workspace_bytes,
                          pgba_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * The same operator for a caller whose tensors live in HOST memory (pinned for asynchronous copies): all pointers
 * marked _h are host pointers in the layouts of pgba_ba_solve; `lmbda_h` points to one float.  The call enqueues, and
 * returns without synchronising:
 *     stream      : H2D ii, jj, kk -> graph analysis (plan)          | wait for aux | iterations -> D2H poses, patches
 *     aux_stream  : H2D poses, patches, intrinsics[0], target, weight, lmbda |
 * i.e. the upload of everything the plan does not need overlaps the plan; `poses_h` rows t0..t1-1 and channel 2 of
 * `patches_h` hold the result once `stream` has been synchronised (in-place semantics of cuda_ba.forward, on the host
 * tensors).  `staging` is a caller-owned DEVICE buffer of pgba_ba_host_staging_bytes() (256-byte aligned), `workspace`
 * as for pgba_ba_solve.  aux_stream may equal stream (no overlap).  The two cudaEvents used for the fork / join are
 * created once per device and cached by the library; everything is capturable into a CUDA graph from `stream`.
 * ------------------------------------------------------------------------------------------------------------- */
int pgba_ba_host_staging_bytes(int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows, int P,
                               size_t* bytes /* host, out */);

/* Arena mode: when the nine host tensors are views of ONE allocation at the byte offsets returned here (order: poses,
 * patches, intrinsics [n_pose_rows, 4], target, weight, lmbda, ii, jj, kk; `bytes` = size of the allocation),
 * pgba_ba_solve_host uploads it with two copies (indices | everything else) and downloads poses + patches with one. */
int pgba_ba_host_arena_offsets(int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows, int P,
                               size_t* offsets9 /* host, out */, size_t* bytes /* host, out */);

int pgba_ba_solve_host(float* poses_h, float* patches_h, const float* intrinsics_h, const float* target_h,
                       const float* weight_h, const float* lmbda_h, const int64_t* ii_h, const int64_t* jj_h,
                       const int64_t* kk_h, int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows, int P, int ppf,
                       int t0, int t1, int iterations, int eff_impl, void* staging, size_t staging_bytes,
                       void* workspace, size_t workspace_bytes, pgba_stream_t stream, pgba_stream_t aux_stream);

/* The same with 32-bit indices (the reference API's index tensors are int64; a caller that builds its edge list on the
 * host can write int32 into the arena and halve the index upload -- 0.45 of 1.74 MB on the default window).  The values
 * are range-checked on the device like the 64-bit ones (PGBA_ST_INDEX_RANGE). */
int pgba_ba_host_arena_offsets_i32(int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows, int P,
                                   size_t* offsets9 /* host, out */, size_t* bytes /* host, out */);

int pgba_ba_solve_host_i32(float* poses_h, float* patches_h, const float* intrinsics_h, const float* target_h,
                           const float* weight_h, const float* lmbda_h, const int32_t* ii_h, const int32_t* jj_h,
                           const int32_t* kk_h, int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows, int P,
                           int ppf, int t0, int t1, int iterations, int eff_impl, void* staging, size_t staging_bytes,
                           void* workspace, size_t workspace_bytes, pgba_stream_t stream, pgba_stream_t aux_stream);

/* Measurement hooks (used by bench.py only).  pgba_ba_solve_profiled runs exactly the launch sequence of
 * pgba_ba_solve_batched with cudaEvents between the stages, SYNCHRONISES the stream and fills the host array
 * stage_ms [1 + 3*iterations]: workspace clear + plan, then per iteration {linearize+Schur, solve + pose
 * retraction, back-substitution + depth retraction}.  pgba_launch_count returns the number of kernels this library
 * has launched. */
int pgba_ba_solve_profiled(float* poses, float* patches, const float* intrinsics, const float* target,
                           const float* weight, const float* lmbda, const int64_t* ii, const int64_t* jj,
                           const int64_t* kk, const pgba_strides* strides /* host */, int64_t batch, int64_t n_edges,
                           int64_t n_pose_rows, int64_t n_patch_rows, int P, int t0, int t1, int iterations,
                           void* workspace, size_t workspace_bytes, pgba_stream_t stream, float* stage_ms /* host */);
long long pgba_launch_count(void);

/* Debug / parity export (the reference API never exposes the Hessian or gradient the tolerance is stated on).
 * Runs ONE linearisation (no retraction, nothing is modified) of a single window and copies out, when the pointer
 * is not NULL:  S f32 [6N,6N] (full symmetric; before damping), y f32 [6N], dX f32 [6N] (solution of the damped
 * system), patch_ids i64 [n_unique] (internal order), C, u, Q, dZ f32 [n_unique] in the same order, n_unique
 * i32 [1], status i32 [1].  with_schur = 0 skips the E Q E^T terms, so S == B and y == v of ba_cuda.cu:364-398. */
int pgba_ba_linearize_debug(const float* poses, const float* patches, const float* intrinsics, const float* target,
                            const float* weight, const float* lmbda, const int64_t* ii, const int64_t* jj,
                            const int64_t* kk, int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows, int P,
                            int t0, int t1, int with_schur, float* S, float* y, float* dX, int64_t* patch_ids,
                            float* C, float* u, float* Q, float* dZ, int32_t* n_unique, int32_t* status,
                            void* workspace, size_t workspace_bytes, pgba_stream_t stream);

/* Status word of window `b` of the last call that used `workspace` with the same sizes (device pointer to an i32;
 * read it after the stream has been synchronised).  0 means every edge was processed. */
const int32_t* pgba_ba_status_ptr(const void* workspace, int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows,
                                  int t0, int t1, int64_t batch, int64_t b);

/* Plan cache.  The graph analysis of a call ("plan": chunk tables, edge permutation, cell tables -- what replaces
 * at::_unique(kk) and the EfficentE constructor of the reference) depends only on ii / jj / kk and the call's sizes.  Windows
 * handled by the single-launch plans (everything except the global BA) keep a 128-bit
 * fingerprint of their edge list in the workspace; a call that finds the workspace untouched since a call with the same
 * sizes and an unchanged edge list (the 12 x initialisation loop of slam.py:715-716, repeated BA calls between two
 * frames) skips the analysis: the index arrays are still read once (for the fingerprint), everything else is reused.
 * Requirements on the caller: none beyond the existing one that the workspace is private to these calls (any write into it
 * by other code must also clobber its first 256 bytes).  PGBA_PLAN_CACHE=0 in the environment disables the reuse.
 * pgba_ba_plan_hit_ptr: device pointer to an i32 that is 1 when window b of the last call reused its tables. */
const int32_t* pgba_ba_plan_hit_ptr(const void* workspace, int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows,
                                    int t0, int t1, int64_t batch, int64_t b);

/* Frame ordering of the large solve.  Beyond 26 free poses (6N > 156) the dense solve of the reference (ba_cuda.cu:575-578, 589-591)
 * runs on a symmetric permutation of the pose system, computed on the device from the edge list by the call itself:
 * [chain segment 0 | ... | chain segment P-1 | border], segments independent of one another (csrc/ba_bignd.cu).
 * pgba_ba_order_ptr (diagnostics, tests): device pointer to the i32 [n_pose_rows] array `pos` of window b of the last call --
 * pos[f] = position (in frames; 8 frames per 48 x 48 tile) of free frame f in the permuted, tile-padded system -- or NULL
 * when the sizes do not take the reordered solve.  *segments = P, *tile_capacity = tiles the permuted system may use,
 * *header = device pointer to 4 + 2 * 32 i32: tiles in use, first border tile, border tiles, border frames, tiles of segment
 * p [32], first tile of segment p [32].  Read after the stream has been synchronised. */
const int32_t* pgba_ba_order_ptr(const void* workspace, int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows,
                                 int t0, int t1, int64_t batch, int64_t b, int32_t* segments /* host, out */,
                                 int32_t* tile_capacity /* host, out */, const int32_t** header /* host, out */);

/* ---------------------------------------------------------------------------------------------------------------
 * Reprojection of all PxP pixels of each edge's patch from frame ii to frame jj.  Replaces cuda_ba.reproject ==
 * cuda_reproject() (reference: cdvslam/fastba/ba_cuda.cu:408-458, 614-645).  coords f32 [n_edges, 2, P, P].
 * clamp_depth = 0 reproduces the kernel (intrinsics row 0, unguarded X/Z); clamp_depth = 1 is pops.transform as
 * slam.py's reproject() calls it (slam.py:328; cdvslam/projective_ops.py:19-68): back-projection with intrinsics[ii],
 * projection with intrinsics[jj] (intrinsics must then hold a row per referenced frame) and d = 1 / Z.clamp(min=0.1).
 * ------------------------------------------------------------------------------------------------------------- */
int pgba_reproject(const float* poses, const float* patches, const float* intrinsics, const int64_t* ii,
                   const int64_t* jj, const int64_t* kk, int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows,
                   int P, int clamp_depth, float* coords, pgba_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Temporal neighbours of every edge.  Replaces cuda_ba.neighbors == neighbors() (reference: cdvslam/fastba/ba.cpp:59-97;
 * called as fastba.neighbors(kk, jj) by every network update, net_cdv.py:102-107).  Edges are grouped by ii; inside a
 * group they are ordered by jj, ties by edge index (the reference's std::stable_sort of a group listed in input order);
 * ix[e] / jx[e] = index of the previous / next edge of e's group in that order, -1 at the ends.  ii, jj, ix, jx i64
 * [n_edges]; any int64 key values.  Two launches (cluster binning kernel + warp-per-bin link kernel), no host round trip (the reference synchronises, sorts on the
 * CPU and copies back).  Workspace: pgba_neighbors_workspace_bytes(), 256-byte aligned.
 * ------------------------------------------------------------------------------------------------------------- */
int pgba_neighbors_workspace_bytes(int64_t n_edges, size_t* bytes /* host, out */);
int pgba_neighbors(const int64_t* ii, const int64_t* jj, int64_t n_edges, int64_t* ix, int64_t* jx, void* workspace,
                   size_t workspace_bytes, pgba_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Pose-graph normal equations + solve.  Replaces cuda_ba.solve_system == solve_system() (reference:
 * cdvslam/fastba/ba.cpp:99-180; caller cdvslam/loop_closure/optim_utils.py:230).  J_Ginv_i, J_Ginv_j f32 [r,7,7];
 * ii, jj i64 [r] (ii[x] != jj[x]); res f32 [r,7]; delta f32 [n_poses,7] (out).  A = J^T J, b = -J^T res,
 * A.diag += A.diag*lm + ep, solve the leading 7*freen x 7*freen block (all of A when freen < 0), rest of delta = 0.
 * Arithmetic in double like the reference (Eigen SimplicialCholesky); `info` (device i32, may be NULL) receives 0 or
 * 1 + the index of the first non-positive pivot.  Workspace: pgba_pgo_workspace_bytes() (dense 7n x 7n doubles).
 * ------------------------------------------------------------------------------------------------------------- */
int pgba_pgo_workspace_bytes(int64_t n_poses, size_t* bytes /* host, out */);
int pgba_pgo_solve(const float* J_Ginv_i, const float* J_Ginv_j, const int64_t* ii, const int64_t* jj, const float* res,
                   int64_t n_res, int64_t n_poses, float ep, float lm, int freen, float* delta, int32_t* info,
                   void* workspace, size_t workspace_bytes, pgba_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PGBA_H_ */
