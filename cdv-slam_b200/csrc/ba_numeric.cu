// Numeric phase of the patch-graph BA (sm_100a): linearisation + Schur complement, small dense solve + pose
// retraction, back-substitution + inverse-depth retraction.
//
//   linearize_kernel   per chunk (one source frame i, <= pc patches, <= 128 target-frame slots):
//                      lanes <-> target-frame slots, loop over patches.  Per edge (reference arithmetic:
//                      cdvslam/fastba/ba_cuda.cu:265-343): reprojection residual, Jj rows, Jz.  Because
//                      Ji = Ad^T(Gij) Jj is the same linear map for every edge of a frame pair (ba_cuda.cu:353,
//                      57-72), only H_ij = sum w Jj Jj^T and g_ij = sum w r Jj are accumulated per edge (registers,
//                      no atomics); B_ii += A H A^T, B_ij -= A H, B_jj += H, v_i -= A g, v_j += g are formed once
//                      per frame pair (ba_cuda.cu:364-398).  Per patch: C, u, E columns (ba_cuda.cu:380-402) by
//                      warp shuffles; Q = 1/(C+lambda); the chunk's Schur update S -= E Q E^T, y -= E Q u
//                      (ba_cuda.cu:583-587, block_e.cu:147-234) is applied from shared memory; E is kept for the
//                      back-substitution in a compact [patch][column][6] layout.  Only the lower block triangle
//                      of S (full diagonal blocks) is accumulated, with 8-byte vector reductions.
//   solve_small_kernel S += I*(1e-4*S+1) (ba_cuda.cu:575/589), blocked Cholesky + solves (ba_cuda.cu:576-577/
//                      590-591) in fp64 shared memory, one CTA per window, 6N <= 156; then the SE3 retraction of
//                      the free poses (ba_cuda.cu:88-206) and re-zeroing of S, y for the next iteration.
//   update_kernel      dZ = Q (u - E^T dX) (ba_cuda.cu:592, block_e.cu:253-283) + inverse-depth retraction
//                      (ba_cuda.cu:209-229)
#include <stdlib.h>

#include "ba_common.cuh"
#include "ba_chol.cuh"
#include "ba_chol32.cuh"
#include "ba_umma.cuh"

namespace pgba {

int chunk_grid(const Problem& pb, int64_t batch);

// Per-edge terms.  xi0, xi1: normalised patch centre ((px-cx)/fx, (py-cy)/fy), pd: inverse depth; R, t: relative
// pose Gij.  Same formulas as ba_cuda.cu:282-343; the Jacobian rows have the structural zeros Jx[1] == Jy[0] == 0,
// so H[0][1] is identically zero and rows 0 / 1 of H, g, e only see one of the two residual rows.
__device__ __forceinline__ void edge_terms(float xi0, float xi1, float pd, float fx, float fy, float cx, float cy,
                                           const float R[9], const float t[3], float2 tg, float2 wt, float H[21],
                                           float g[6], float e[6], float& c_out, float& u_out) {
  const float X = R[0] * xi0 + R[1] * xi1 + R[2] + pd * t[0];
  const float Y = R[3] * xi0 + R[4] * xi1 + R[5] + pd * t[1];
  const float Z = R[6] * xi0 + R[7] * xi1 + R[8] + pd * t[2];
  const float W = pd;
  const float iz = 1.0f / Z;                                   // unguarded, like the reference's X / Z (:299-300)
  const float d = (Z >= 0.2f) ? iz : 0.0f;                     // ba_cuda.cu:296
  const float d2 = d * d;
  const float x1 = fx * (X * iz) + cx;
  const float y1 = fy * (Y * iz) + cy;
  const float rx = tg.x - x1, ry = tg.y - y1;
  const bool in_bounds = (rx * rx + ry * ry < 128.f * 128.f) && (Z > 0.2f) && (x1 > -64.f) && (y1 > -64.f) &&
                         (x1 < 2.f * cx + 64.f) && (y1 < 2.f * cy + 64.f);
  const float wx = in_bounds ? wt.x : 0.0f, wy = in_bounds ? wt.y : 0.0f;
  // Jj rows (ba_cuda.cu:323-341)
  const float fxd = fx * d, fyd = fy * d, Xd2 = X * d2, Yd2 = Y * d2;
  const float Jx0 = fxd * W, Jx2 = -fx * Xd2 * W, Jx3 = -fx * Xd2 * Y, Jx4 = fx * (1.0f + X * Xd2), Jx5 = -fxd * Y;
  const float Jy1 = fyd * W, Jy2 = -fy * Yd2 * W, Jy3 = -fy * (1.0f + Y * Yd2), Jy4 = fy * Xd2 * Y, Jy5 = fyd * X;
  const float Jzx = fx * (t[0] * d - t[2] * Xd2);
  const float Jzy = fy * (t[1] * d - t[2] * Yd2);
  const float Jx[6] = {Jx0, 0.f, Jx2, Jx3, Jx4, Jx5};
  const float Jy[6] = {0.f, Jy1, Jy2, Jy3, Jy4, Jy5};
  const float wrx = wx * rx, wry = wy * ry;
  const float wzx = wx * Jzx, wzy = wy * Jzy;
  {                                                            // row 0: only Jx, row 1: only Jy
    const float wa = wx * Jx0, wb = wy * Jy1;
    H[sym6(0, 0)] += wa * Jx0;
    H[sym6(1, 1)] += wb * Jy1;
#pragma unroll
    for (int b = 2; b < 6; ++b) {
      H[sym6(0, b)] += wa * Jx[b];
      H[sym6(1, b)] += wb * Jy[b];
    }
    g[0] += wrx * Jx0;
    g[1] += wry * Jy1;
    e[0] = wzx * Jx0;
    e[1] = wzy * Jy1;
  }
#pragma unroll
  for (int a = 2; a < 6; ++a) {
    const float wa = wx * Jx[a], wb = wy * Jy[a];
#pragma unroll
    for (int b = a; b < 6; ++b) H[sym6(a, b)] += wa * Jx[b] + wb * Jy[b];
    g[a] += wrx * Jx[a] + wry * Jy[a];
    e[a] = wzx * Jx[a] + wzy * Jy[a];
  }
  c_out = wzx * Jzx + wzy * Jzy;
  u_out = wrx * Jzx + wry * Jzy;
}

// Shared-memory carve-up of linearize_kernel (floats unless noted)
#ifndef LIN_MIN_CTAS
#define LIN_MIN_CTAS 2
#endif
constexpr int PQS = 9;                              // stride of the per-patch records sPQ
constexpr int HW_STRIDE = 29;                       // odd stride: conflict-free per-lane rows
constexpr int LIN_FIXED_FLOATS = SMAX * 12 + SMAX * 28 + 8 * 32 * HW_STRIDE + 48 + SMAX;

struct LinSmem {
  float* sRt;      // [SMAX][12]  relative pose per slot (R row-major, t)
  float* sH;       // [SMAX][28]  H (21) + g (6) per slot, reduced over warps and patch batches
  float* sHw;      // [8][32][29] per-warp partials of the current slot block; reused as sAH [SMAX][36]
  float* sBii;     // [36 + 6]    B_ii and v_i of the chunk
  int* sFrame;     // [SMAX]
  float* sPatch;   // [pc][4]     px, py, pd, -
  float* sPQ;      // [pc][PQS]   per patch: C, u, E_i[6] (source-frame column accumulators); odd stride: lanes <-> patches is conflict-free
  float* sQ;       // [pc]
  float* sE;       // [ebudget]   E tile [patch][col][6]
  float* sAH;      // [SMAX][36]  A_s H_s of the pose-block phase: aliases sHw, or (tcgen05 Schur: sHw holds an operand array
                   //             while that phase runs) its own region behind the E tile
  float* sSq;      // [pc]        sqrt(Q) (tcgen05 Schur only)
};

size_t lin_smem_bytes(int pc, int ebudget, bool umma_regions) {
  return sizeof(float) * ((size_t)LIN_FIXED_FLOATS + (size_t)pc * (4 + PQS + 1) + (size_t)ebudget +
                          (umma_regions ? (size_t)SMAX * 36 + (size_t)pc : 0));
}

__device__ __forceinline__ int pow2_ceil(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

// Width of the next block of target-frame slots in the edge loop.  Lanes <-> slots wastes (pow2 - width) / pow2 of every
// warp step, so with many patches per chunk (split == true: >= 64, i.e. several steps per block) 17..27 remaining slots
// are taken as 16 + the rest and 9 as 8 + 1 (the extra block costs about one step: fold + two barriers); small chunks
// (single window, one step per block) always take min(32, rem).
__device__ __forceinline__ int slot_block_width(int rem, bool split) {
  if (rem >= 32) return 32;
  if (!split) return rem;
  if (rem > 16) return rem >= 28 ? rem : 16;
  if (rem == 9) return 8;
  return rem;
}

__device__ __forceinline__ void red_add2(float* p, float a, float b) {       // p 8-byte aligned
  atomicAdd(reinterpret_cast<float2*>(p), make_float2(a, b));
}

// Everything the back-substitution of a chunk needs that does NOT depend on dX, requested ahead of time (small-chunk
// path: the first patch of every warp): E row, Q, u, patch id -> old depth, and the index of the thread's dX entry.
constexpr int UPD_PRE = 2;     // patches per warp fetched ahead (16 patches per chunk in the single-window regime = 2 per warp)
struct UpdPre {
  float2 e0[UPD_PRE], e1[UPD_PRE];
  float q[UPD_PRE], u[UPD_PRE], d[UPD_PRE];
  int kx[UPD_PRE], dxi;
};
__device__ __forceinline__ UpdPre upd_prefetch(const Problem& pb, const WinPtrs& wp, const Chunk& ch, const float* patches,
                                               bool apply);
__device__ __forceinline__ void chunk_depth_update(const Problem& pb, const WinPtrs& wp, const Chunk& ch, float* patches,
                                                   float* sdx, bool apply, const UpdPre& pre, float* s_depth);

#ifdef PGBA_LIN_TIMING
// per-CTA wall-clock trace (globaltimer, ns): [kernel slot][entry / after pdl_wait / exit][flattened CTA index < 512];
// slots: 0 first linearize, 1 solve, 2 update, 3 linearize with the fused update (profiles/cta_trace.py)
__device__ unsigned long long g_cta_ts[4][3][512];
__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define CTA_TS(k, f) do { if (threadIdx.x == 0) { const int b_ = blockIdx.x + gridDim.x * blockIdx.y; \
                          if (b_ < 512) g_cta_ts[k][f][b_] = gtime_ns(); } } while (0)
__device__ long long g_lin_ts[32];
#ifndef PGBA_LIN_TS_BLOCK
#define PGBA_LIN_TS_BLOCK 0
#endif
#define LIN_TS(i) do { if (threadIdx.x == 0 && blockIdx.y == PGBA_LIN_TS_BLOCK && blockIdx.x == 0) g_lin_ts[i] = clock64(); } while (0)
#else
#define LIN_TS(i) do { } while (0)
#define CTA_TS(k, f) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------------------------
// Schur update of one patch batch on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulator in TMEM); see
// ba_umma.cuh.  Large chunks only (batched windows, global BA): nq <= 96 patches, ncols <= 11 columns.  Two halves, so
// that the pose-block phase of the chunk runs on warps 0..6 while warp 7 feeds the tensor core:
//   schur_umma_issue
//     1. (ncols == 11 only) rows 64, 65 of the product do not fit the M = 64 instruction: their 2 x 2 corner and gradient
//        entries are five dot products on one warp; their other entries come from the symmetry D[64 + x][n] = D[n][64 + x]
//     2. staging, in place: X = sqrt(Q) E (gradient row sqrt(Q) u) split into TF32 hi / lo, written in the canonical
//        K-major core-matrix layout: hi over the E tile (its fp32 contents have just gone to global memory), lo over the
//        per-warp partials region (idle from here on).  A warp writes one 128-byte core matrix per step: conflict-free
//     3. ONE thread (the first of warp 7) issues 3 MMAs per 8 patches (hi hi^T, lo hi^T, hi lo^T) and commits to an
//        mbarrier; everybody else returns at once
//   schur_umma_finish
//     4. wait for the accumulator; thread = one row of D (tcgen05.ld 32x32b, 24 columns at a time; warps 0..3 take the even
//        column blocks, warps 4..7 -- same TMEM lanes -- the odd ones), 8-byte vector reductions into S, y
// Not inlined: their register demand (27 staged values, 24 accumulator columns) stays out of the edge loop's allocation.
// All arguments are scalars / pointers passed by value -- handing the kernel's LinSmem / Chunk / WinPtrs structs over by
// reference forces them into local memory for the WHOLE kernel (measured: c5 linearisation 113 -> 199 us).
// ---------------------------------------------------------------------------------------------------------------
#ifndef SCHUR_UMMA_ATTR
#define SCHUR_UMMA_ATTR __noinline__
#endif
constexpr int UMMA_ISSUE_THREAD = 224;         // first thread of warp 7

__device__ SCHUR_UMMA_ATTR void schur_umma_issue(float* sE, float* sHw, const float* sQ, const float* sSq, const float* sPQ,
                                                 float* gS, float* gy, int f10, int nq, int ncols, int n6, uint32_t tmem,
                                                 uint64_t* mbar) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int estride = ncols * 6;
  // ---- 1. corner entries of rows 64, 65 (from the fp32 tile, before it is overwritten)
  if (ncols == 11 && warp == 6) {
    float corner[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int q = lane; q < nq; q += 32) {
      const float Q = sQ[q], u = sPQ[q * PQS + 1];
      const float e4 = sE[q * estride + 64], e5 = sE[q * estride + 65];
      corner[0] = fmaf(Q * e4, e4, corner[0]);
      corner[1] = fmaf(Q * e5, e4, corner[1]);
      corner[2] = fmaf(Q * e5, e5, corner[2]);
      corner[3] = fmaf(Q * u, e4, corner[3]);
      corner[4] = fmaf(Q * u, e5, corner[4]);
    }
#pragma unroll
    for (int x = 0; x < 5; ++x)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) corner[x] += __shfl_xor_sync(0xffffffffu, corner[x], o);
    if (lane == 0) {
      float* d = gS + (size_t)(6 * f10) * n6 + 6 * f10;
      atomicAdd(d + (size_t)4 * n6 + 4, -corner[0]);
      atomicAdd(d + (size_t)5 * n6 + 4, -corner[1]);
      atomicAdd(d + (size_t)4 * n6 + 5, -corner[1]);
      atomicAdd(d + (size_t)5 * n6 + 5, -corner[2]);
      atomicAdd(&gy[6 * f10 + 4], -corner[3]);
      atomicAdd(&gy[6 * f10 + 5], -corner[4]);
    }
  }
  LIN_TS(12);
  // ---- 2. staging: read everything, barrier, write in place.  Core matrix cm = warp + 8 it  ->  (8-row group g, K group kq)
  constexpr int KG = umma::KMAX / 4;                               // 24 core matrices per 8-row group
  constexpr int PER_WARP = (umma::NROWS / 8) * KG / 8;             // 27
  float val[PER_WARP];
  {
    int g = 0, kq = warp;
#pragma unroll
    for (int it = 0; it < PER_WARP; ++it) {
      const int n = 8 * g + (lane & 7), q = 4 * kq + (lane >> 3);
      // unconditional loads from clamped addresses (all 27 in flight together), selection afterwards
      const int qc = min(q, nq - 1);
      const float e = sE[qc * estride + min(n, estride - 1)], u = sPQ[qc * PQS + 1], sq = sSq[qc];
      val[it] = (q < nq && n <= estride) ? sq * ((n < estride) ? e : u) : 0.f;
      kq += 8;
      if (kq >= KG) { kq -= KG; ++g; }
    }
  }
  LIN_TS(13);
  __syncthreads();
  {
    char* xhi = reinterpret_cast<char*>(sE) + (lane & 7) * 16 + (lane >> 3) * 4;
    char* xlo = reinterpret_cast<char*>(sHw) + (lane & 7) * 16 + (lane >> 3) * 4;
#pragma unroll
    for (int it = 0; it < PER_WARP; ++it) {
      const int cm = warp + 8 * it;
      float hi, lo;
      umma::split_tf32(val[it], hi, lo);
      *reinterpret_cast<float*>(xhi + cm * 128) = hi;
      *reinterpret_cast<float*>(xlo + cm * 128) = lo;
    }
  }
  umma::fence_smem_to_async();
  __syncthreads();
  LIN_TS(14);
  // ---- 3. MMAs
  if (tid == UMMA_ISSUE_THREAD) {
    const int nk = (nq + 7) >> 3;
    const int N = ((estride + 1 + 7) >> 3) << 3;                    // accumulator columns actually needed
    umma::fence_after_sync();
    const uint32_t ahi = umma::smem_u32(sE), alo = umma::smem_u32(sHw);
    const uint32_t idesc = umma::instr_desc_tf32_m64(N);
    for (int ks = 0; ks < nk; ++ks) {
      const uint64_t dhi = umma::smem_desc(ahi + ks * 2 * umma::LBO), dlo = umma::smem_desc(alo + ks * 2 * umma::LBO);
      umma::mma_tf32(tmem, dhi, dhi, idesc, ks > 0 ? 1u : 0u);
      umma::mma_tf32(tmem, dlo, dhi, idesc, 1u);
      umma::mma_tf32(tmem, dhi, dlo, idesc, 1u);
    }
    umma::mma_commit(mbar);
  }
  __syncwarp();                 // the other lanes of the issuing warp do not run ahead into the mbarrier wait
  LIN_TS(15);
}

// All 256 threads.  Ends with a __syncthreads(); returns the flipped mbarrier parity.
__device__ SCHUR_UMMA_ATTR uint32_t schur_umma_finish(const int* sFrame, int first_free, int n_free, float* gS, float* gy,
                                                     int ncols, int fi, int t0, int n6, uint32_t tmem, uint64_t* mbar,
                                                     uint32_t mbar_parity) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int estride = ncols * 6;
  const int N = ((estride + 1 + 7) >> 3) << 3;
  auto col_frame = [&](int cbk) { return ((cbk < n_free) ? sFrame[first_free + cbk] : fi) - t0; };
  umma::mbar_wait(mbar, mbar_parity);
  umma::fence_after_sync();
  LIN_TS(16);
  // ---- 4. epilogue: the M = 64 accumulator rows live in TMEM lanes 32 * (m / 16) + m % 16; a warp reaches the lanes of
  //      its quarter (warp % 4), so warps w and w + 4 share rows and split the column blocks by parity
  {
    const int wq = warp & 3, half = warp >> 2;
    const int m = 16 * wq + lane;                                  // valid for lane < 16
    const bool row_ok = lane < 16 && m < estride;
    const int ca = m / 6, r = m - 6 * ca;
    const int fa = row_ok ? col_frame(ca) : 0;
    const uint32_t trow = tmem + ((uint32_t)(32 * wq) << 16);
#pragma unroll
    for (int part = 0; part < 3; ++part) {                         // 24 columns = 4 column blocks at a time
      if (24 * part >= N) break;                                   // warp-uniform
      float v[24];
      umma::tmem_ld16(trow + 24 * part, v);
      umma::tmem_ld8(trow + 24 * part + 16, v + 16);
      umma::tmem_ld_wait();
      if (!row_ok) continue;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int cb = 4 * part + b;
        if ((cb & 1) != half) continue;
        if (cb <= ca && cb < ncols) {                              // lower block triangle, full diagonal blocks
          const int fb = col_frame(cb);
          if (fa >= fb) {
            float* dst = gS + (size_t)(6 * fa + r) * n6 + 6 * fb;
            red_add2(dst, -v[6 * b], -v[6 * b + 1]);
            red_add2(dst + 2, -v[6 * b + 2], -v[6 * b + 3]);
            red_add2(dst + 4, -v[6 * b + 4], -v[6 * b + 5]);
          } else {                                                 // transposed into block (fb, fa)
            float* dst = gS + (size_t)(6 * fb) * n6 + 6 * fa + r;
#pragma unroll
            for (int c = 0; c < 6; ++c) atomicAdd(dst + (size_t)c * n6, -v[6 * b + c]);
          }
        }
        if (cb == ncols) atomicAdd(&gy[6 * fa + r], -v[6 * b]);    // gradient column (index estride)
      }
      if (part == 2 && ncols == 11 && half == 0) {                 // rows 64, 65 by symmetry: D[64 + x][m] = D[m][64 + x]
        const int f10 = col_frame(10);
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          const float d = -v[64 - 48 + x];
          if (f10 >= fa) atomicAdd(gS + (size_t)(6 * f10 + 4 + x) * n6 + 6 * fa + r, d);   // incl. the diagonal block's mirror
          else atomicAdd(gS + (size_t)(6 * fa + r) * n6 + 6 * f10 + 4 + x, d);
        }
      }
    }
  }
  LIN_TS(17);
  umma::fence_before_sync();
  __syncthreads();
  return mbar_parity ^ 1u;
}

// grid = (batch, gx), block = 256, dynamic smem = lin_smem_bytes(pc, ebudget).  fuse_update: first apply the previous
// iteration's back-substitution + depth retraction to the chunk's patches (saves the separate update launch).
template <bool FUSE, bool UMMA>
__global__ void __launch_bounds__(256, LIN_MIN_CTAS) linearize_kernel(Problem pb, int ebudget, int flags) {
  CTA_TS((flags & 1) ? 3 : 0, 0);
  // flags & 8 (second and later linearisations of a call, small solve in between): the previous kernel is the solve,
  // which lets this kernel start only after its own pdl_wait(), i.e. after the previous linearisation and the plan have
  // completed.  Only dX and the poses are still being produced; the chunk tables, cells, targets / weights, patch
  // coordinates and the inputs of the fused back-substitution are final and are requested BEFORE pdl_wait().
  // flags & 64 (not the first linearisation of the call): the plan completed kernels ago, so the chunk count is final and a
  // CTA in a chunk slot beyond it (the grid is sized for the worst case) leaves at once, without waiting
  if ((flags & 64) && (int)blockIdx.y >= win_ptrs(pb.ws, pb.L, blockIdx.x + pb.w0).hdr->n_chunks) return;
  bool waited = (flags & 8) == 0;
  if (waited) {
    pdl_wait();
    pdl_trigger();
    CTA_TS((flags & 1) ? 3 : 0, 1);
    LIN_TS(0);
  }
  if ((flags & 32) && waited && blockIdx.x == 0 && blockIdx.y == 0 && pb.w0 == 0 && threadIdx.x == 0) {
    // first linearisation after a plan: record the descriptor of the call whose tables this workspace now holds (plan cache,
    // see WinHeader; written here, after every window's plan has completed, because the windows' plan clusters read it)
    int* d = win_ptrs(pb.ws, pb.L, 0).hdr->desc;
    d[0] = PLAN_DESC_MAGIC; d[1] = (int)pb.E; d[2] = pb.F; d[3] = pb.K; d[4] = pb.t0; d[5] = pb.t1; d[6] = pb.L.pc; d[7] = pb.batch;
  }
  extern __shared__ __align__(16) float smem[];
  // tcgen05 Schur product (flags & 16): per-CTA tensor-memory allocation + one mbarrier, released at the end of the kernel
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_mbar;
  constexpr bool use_umma = UMMA;
  uint32_t mbar_parity = 0;
  const int pc = pb.L.pc;
  LinSmem s;
  s.sRt = smem;
  s.sH = s.sRt + SMAX * 12;
  s.sHw = s.sH + SMAX * 28;
  s.sBii = s.sHw + 8 * 32 * HW_STRIDE;
  s.sFrame = (int*)(s.sBii + 48);
  s.sPatch = (float*)(s.sFrame + SMAX);
  s.sPQ = s.sPatch + pc * 4;
  s.sQ = s.sPQ + pc * PQS;
  s.sE = s.sQ + pc;
  s.sAH = use_umma ? s.sE + ebudget : s.sHw;
  s.sSq = s.sAH + SMAX * 36;                           // only touched by the tcgen05 instance
  float* sAH = s.sAH;

  // grid = (windows, chunk slots): blockIdx.x is the window, so CTAs are dispatched chunk slot by chunk slot over all
  // windows and the slots beyond a window's chunk count (the grid is sized for the worst case) come LAST instead of taking
  // resident-CTA slots between the real ones
  const int w = blockIdx.x + pb.w0, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  const float* poses = pb.poses + (int64_t)w * pb.st.poses;
  const float* patches = pb.patches + (int64_t)w * pb.st.patches;
  const float* intr = pb.intrinsics + (int64_t)w * pb.st.intrinsics;
  const float2* target = (const float2*)(pb.target + (int64_t)w * pb.st.target);
  const float2* weight = (const float2*)(pb.weight + (int64_t)w * pb.st.weight);
  const float lmbda = pb.lmbda[(int64_t)w * pb.st.lmbda];
  const float fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  const int t0 = pb.t0, N = pb.t1 - pb.t0, n6 = 6 * N;
  const int PP = pb.P * pb.P, pstride = 3 * PP, cidx = pb.P + 1;   // centre = [1][1] (ba_cuda.cu:282-285)
  const int n_chunks = wp.hdr->n_chunks;
  const int n_dups = wp.hdr->n_dups;
  const bool schur = pb.with_schur != 0;
  // the grid is sized for the worst-case chunk count: only CTAs that own a chunk allocate tensor memory
  const bool umma_cta = use_umma && (int)blockIdx.y < n_chunks;
  if (umma_cta) {
    if (threadIdx.x < 32) umma::tmem_alloc(&s_tmem);
    if (threadIdx.x == 32) umma::mbar_init(&s_mbar, 1);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }
  constexpr bool fuse_update = FUSE;

  for (int c = blockIdx.y; c < n_chunks; c += gridDim.y) {
    const Chunk ch = wp.chunks[c];
    if (ch.n_patches == 0) continue;
    __syncthreads();
    LIN_TS(1);
    const int ns = ch.n_slots, ncols = ch.ncols, fi = ch.frame, np = ch.n_patches;
    const bool i_free = ch.icol >= 0;
    const int* cells = wp.cells + ch.cell_base;
    const int* kx = wp.kx + ch.patch_base;
    const int estride = ncols * 6;
    int PB = (estride > 0) ? min(np, max(ebudget / estride, 1)) : np;
    // ---- early loads.  The chunk's tables are chains of dependent global loads (slot -> pose, patch id -> patch,
    //      cell -> target / weight); in the single-window regime their round trips are the kernel's critical path, so
    //      the first link of every chain is issued here, before the fused depth update, and the second links together
    //      right after it (3 + 1 round trips instead of 6 in sequence).
    const int e_fj = (tid < ns) ? wp.slots[ch.slot_base + tid] : 0;                       // ns <= SMAX <= 256
    const int e_np = min(PB, np);                                                         // first patch batch [0, e_np)
    const int e_kx = (tid < e_np) ? kx[tid] : 0;                                          // e_np <= PMAX <= 256
    const bool split_slots = pc >= 64;
    const int e_w = slot_block_width(ns, split_slots);                                    // width of the first slot block
    const int e_DW = pow2_ceil(e_w), e_PW = 32 / e_DW;
    const bool e_ok = (lane & (e_DW - 1)) < e_w;
    const int e_p = warp * e_PW + lane / e_DW, e_sl = lane & (e_DW - 1);
    const int e_n0 = (e_ok && e_p < e_np) ? cells[e_p * ns + e_sl] : -1;
    const int e_n1 = (e_ok && e_p + 8 * e_PW < e_np) ? cells[(e_p + 8 * e_PW) * ns + e_sl] : -1;
    UpdPre upre;
    if (fuse_update) upre = upd_prefetch(pb, wp, ch, patches, true);
    float e_px = 0.f, e_py = 0.f, e_pd = 0.f;                  // x, y never change; the depth is rewritten by the update
    if (FUSE && tid < e_np) {
      const float* pr = patches + (int64_t)e_kx * pstride;
      e_px = pr[cidx]; e_py = pr[PP + cidx];
    }
    if (!waited) {
      pdl_wait();
      pdl_trigger();
      waited = true;
      CTA_TS((flags & 1) ? 3 : 0, 1);
      LIN_TS(0);
    }
    float Pi[7], Pj[7];                                         // pose rows: requested before the depth update
    if (FUSE && tid < ns) {
#pragma unroll
      for (int x = 0; x < 7; ++x) { Pi[x] = poses[7 * (int64_t)fi + x]; Pj[x] = poses[7 * (int64_t)e_fj + x]; }
    }
    if (fuse_update) {
      chunk_depth_update(pb, wp, ch, pb.patches + (int64_t)w * pb.st.patches, s.sHw, true, upre, s.sQ);
      if (tid < e_np) e_pd = s.sQ[tid];                         // the retracted depth, straight from the update
    }
    float2 e_tg = make_float2(0.f, 0.f), e_wt = make_float2(0.f, 0.f);
    if (e_n0 >= 0) { e_tg = __ldg(target + e_n0); e_wt = __ldg(weight + e_n0); }
    if (!FUSE && tid < e_np) {
      const float* pr = patches + (int64_t)e_kx * pstride;
      e_px = pr[cidx]; e_py = pr[PP + cidx]; e_pd = pr[2 * PP + cidx];
    }
    // ---- per-slot relative poses, zero the H accumulators
    if (tid < ns) {
      const int sl = tid, fj = e_fj;
      s.sFrame[sl] = fj;
      float R[9], t[3];
      if (!FUSE) {
#pragma unroll
        for (int x = 0; x < 7; ++x) { Pi[x] = poses[7 * (int64_t)fi + x]; Pj[x] = poses[7 * (int64_t)fj + x]; }
      }
      rel_pose(Pi, Pj, R, t);
#pragma unroll
      for (int x = 0; x < 9; ++x) s.sRt[sl * 12 + x] = R[x];
#pragma unroll
      for (int x = 0; x < 3; ++x) s.sRt[sl * 12 + 9 + x] = t[x];
    }
    for (int x = tid; x < ns * 28; x += 256) s.sH[x] = 0.f;

    bool umma_pending = false;
    for (int b0 = 0; b0 < np; b0 += PB) {
      const int b1 = min(b0 + PB, np);
      __syncthreads();
    LIN_TS(2);
      // ---- stage patch centres, zero the E tile
      if (b0 == 0) {
        if (tid < b1) {
          s.sPatch[tid * 4 + 0] = (e_px - cx) / fx;              // ba_cuda.cu:282-283
          s.sPatch[tid * 4 + 1] = (e_py - cy) / fy;
          s.sPatch[tid * 4 + 2] = e_pd;
        }
      } else {
        for (int p = b0 + tid; p < b1; p += 256) {
          const float* pr = patches + (int64_t)kx[p] * pstride;
          const int q = p - b0;
          s.sPatch[q * 4 + 0] = (pr[cidx] - cx) / fx;
          s.sPatch[q * 4 + 1] = (pr[PP + cidx] - cy) / fy;
          s.sPatch[q * 4 + 2] = pr[2 * PP + cidx];
        }
      }
      for (int x = tid; x < (b1 - b0) * estride; x += 256) s.sE[x] = 0.f;
      __syncthreads();
    LIN_TS(3);

      // ---- tile loop: lanes <-> slots (DW wide), PW patches per warp step
      for (int sb = 0, ns_here = 0; sb < ns; sb += ns_here) {
        ns_here = slot_block_width(ns - sb, split_slots);
        const int DW = pow2_ceil(ns_here), PW = 32 / DW;
        const int sl = sb + (lane & (DW - 1));
        const int pl = lane / DW;
        const bool slot_ok = (lane & (DW - 1)) < ns_here;
        float R[9], t[3];
        if (slot_ok) {
#pragma unroll
          for (int x = 0; x < 9; ++x) R[x] = s.sRt[sl * 12 + x];
#pragma unroll
          for (int x = 0; x < 3; ++x) t[x] = s.sRt[sl * 12 + 9 + x];
        } else {
#pragma unroll
          for (int x = 0; x < 9; ++x) R[x] = 0.f;
          t[0] = t[1] = t[2] = 0.f;
        }
        const int col = slot_ok ? (sl - ch.first_free) : -1;            // E column of this slot if it is free
        const bool col_ok = slot_ok && col >= 0 && col < ch.n_free;
        float H[21], g[6];
#pragma unroll
        for (int x = 0; x < 21; ++x) H[x] = 0.f;
#pragma unroll
        for (int x = 0; x < 6; ++x) g[x] = 0.f;

        // software pipeline over the patch steps: the cell index is fetched two steps ahead, target / weight one
        // step ahead, so the two dependent global loads of an edge are off the critical path
        const int pstep = 8 * PW;
        const int p_first = b0 + warp * PW + pl;
        int n_cur, n_nxt;
        float2 tg_cur = make_float2(0.f, 0.f), wt_cur = make_float2(0.f, 0.f);
        if (b0 == 0 && sb == 0) {                    // loaded early (same DW / PW / lane mapping)
          n_cur = e_n0; n_nxt = e_n1; tg_cur = e_tg; wt_cur = e_wt;
        } else {
          n_cur = (slot_ok && p_first < b1) ? cells[p_first * ns + sl] : -1;
          n_nxt = (slot_ok && p_first + pstep < b1) ? cells[(p_first + pstep) * ns + sl] : -1;
          if (n_cur >= 0) { tg_cur = __ldg(target + n_cur); wt_cur = __ldg(weight + n_cur); }
        }
        for (int p0 = b0 + warp * PW; p0 < b1; p0 += pstep) {
          const int p = p0 + pl;
          const int n = n_cur;
          const float2 tg = tg_cur, wt = wt_cur;
          {
            const int p2 = p + 2 * pstep;
            const int n_nn = (slot_ok && p2 < b1) ? cells[p2 * ns + sl] : -1;
            n_cur = n_nxt;
            if (n_cur >= 0) { tg_cur = __ldg(target + n_cur); wt_cur = __ldg(weight + n_cur); }
            n_nxt = n_nn;
          }
          float e[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, ck = 0.f, uk = 0.f;
          float ei[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (n >= 0) {
            const int q = p - b0;
            edge_terms(s.sPatch[q * 4], s.sPatch[q * 4 + 1], s.sPatch[q * 4 + 2], fx, fy, cx, cy, R, t, tg, wt, H, g,
                       e, ck, uk);
            if (col_ok && schur) {
#pragma unroll
              for (int a = 0; a < 6; ++a) s.sE[q * estride + col * 6 + a] = e[a];
            }
            if (i_free) adj_map(R, t, e, ei);
          }
          // reduce the per-patch quantities over the DW lanes of this patch (plain butterfly: measured faster than a
          // transposed, select-heavy 8-value reduction on this kernel)
          for (int o = DW >> 1; o > 0; o >>= 1) {
            ck += __shfl_xor_sync(0xffffffffu, ck, o);
            uk += __shfl_xor_sync(0xffffffffu, uk, o);
            if (i_free) {
#pragma unroll
              for (int a = 0; a < 6; ++a) ei[a] += __shfl_xor_sync(0xffffffffu, ei[a], o);
            }
          }
          if ((lane & (DW - 1)) == 0 && p < b1) {       // exactly one lane group owns patch p in this slot block
            float* dst = s.sPQ + (p - b0) * PQS;          // [0]: C, [1]: u, [2..7]: E_i -= w Jz Ji
            if (sb == 0) {
              dst[0] = ck; dst[1] = uk;
#pragma unroll
              for (int a = 0; a < 6; ++a) dst[2 + a] = -ei[a];
            } else {
              dst[0] += ck; dst[1] += uk;
#pragma unroll
              for (int a = 0; a < 6; ++a) dst[2 + a] -= ei[a];
            }
          }
        }
        // fold the lane groups that hold the same slot, then reduce over warps through shared memory
        for (int o = DW; o < 32; o <<= 1) {
#pragma unroll
          for (int x = 0; x < 21; ++x) H[x] += __shfl_xor_sync(0xffffffffu, H[x], o);
#pragma unroll
          for (int x = 0; x < 6; ++x) g[x] += __shfl_xor_sync(0xffffffffu, g[x], o);
        }
        {
          float* dst = s.sHw + (warp * 32 + lane) * HW_STRIDE;
#pragma unroll
          for (int x = 0; x < 21; ++x) dst[x] = H[x];
#pragma unroll
          for (int x = 0; x < 6; ++x) dst[21 + x] = g[x];
        }
        __syncthreads();
    LIN_TS(4);
        for (int x = tid; x < ns_here * 27; x += 256) {
          const int sx = x / 27, v = x - sx * 27;
          float acc = 0.f;
#pragma unroll
          for (int ww = 0; ww < 8; ++ww) acc += s.sHw[(ww * 32 + sx) * HW_STRIDE + v];
          s.sH[(sb + sx) * 28 + v] += acc;
        }
        __syncthreads();
    LIN_TS(5);
      }

      // ---- duplicated (patch, slot) edges: rare slow path, one thread (deterministic, no shared atomics), from the
      //      plan's global list
      if (n_dups > 0) {
        if (tid == 0) {
          auto process_dup = [&](int q, int sl, int n) {
            float R[9], t[3], H[21], g[6], e[6], ei[6], ck, uk;
            for (int x = 0; x < 9; ++x) R[x] = s.sRt[sl * 12 + x];
            for (int x = 0; x < 3; ++x) t[x] = s.sRt[sl * 12 + 9 + x];
            for (int x = 0; x < 21; ++x) H[x] = 0.f;
            for (int x = 0; x < 6; ++x) g[x] = 0.f;
            edge_terms(s.sPatch[q * 4], s.sPatch[q * 4 + 1], s.sPatch[q * 4 + 2], fx, fy, cx, cy, R, t,
                       __ldg(target + n), __ldg(weight + n), H, g, e, ck, uk);
            const int col = sl - ch.first_free;
            if (col >= 0 && col < ch.n_free && schur)
              for (int a = 0; a < 6; ++a) s.sE[q * estride + col * 6 + a] += e[a];
            if (i_free) {
              adj_map(R, t, e, ei);
              for (int a = 0; a < 6; ++a) s.sPQ[q * PQS + 2 + a] -= ei[a];
            }
            s.sPQ[q * PQS] += ck;
            s.sPQ[q * PQS + 1] += uk;
            for (int x = 0; x < 21; ++x) s.sH[sl * 28 + x] += H[x];
            for (int x = 0; x < 6; ++x) s.sH[sl * 28 + 21 + x] += g[x];
          };
          for (int dix = 0; dix < n_dups; ++dix) {
            const DupEdge de = wp.dups[dix];
            if (de.chunk != c || de.p < b0 || de.p >= b1) continue;
            process_dup(de.p - b0, de.s, de.n);
          }
        }
        __syncthreads();
      }

      // ---- per patch: fold the source-frame column, Q = 1/(C + lambda), export Q, u
      for (int p = b0 + tid; p < b1; p += 256) {
        const int q = p - b0;
        const float Q = 1.0f / (s.sPQ[q * PQS] + lmbda);
        s.sQ[q] = Q;
        if (use_umma) s.sSq[q] = sqrtf(Q);
        wp.Q[ch.patch_base + p] = Q;
        wp.u[ch.patch_base + p] = s.sPQ[q * PQS + 1];
        if (i_free && schur) {
#pragma unroll
          for (int a = 0; a < 6; ++a) s.sE[q * estride + ch.icol * 6 + a] += s.sPQ[q * PQS + 2 + a];
        }
      }
      __syncthreads();
    LIN_TS(6);
      if (N > 0 && schur && ncols > 0) {
        float* eg = wp.ecells + 6 * ((int64_t)ch.ecell_base + (int64_t)b0 * ncols);
        for (int x = tid; x < (b1 - b0) * estride; x += 256) eg[x] = s.sE[x];

        if (use_umma && split_slots && (b1 - b0) <= umma::KMAX && ncols <= 11) {
          const int f10 = (ncols == 11) ? ((10 < ch.n_free) ? s.sFrame[ch.first_free + 10] : fi) - t0 : 0;
          schur_umma_issue(s.sE, s.sHw, s.sQ, s.sSq, s.sPQ, wp.S, wp.y, f10, b1 - b0, ncols, n6, s_tmem, &s_mbar);
          if (b1 < np) {                                  // more patch batches follow: the operand arrays are needed again
            mbar_parity = schur_umma_finish(s.sFrame, ch.first_free, ch.n_free, wp.S, wp.y, ncols, fi, t0, n6, s_tmem,
                                            &s_mbar, mbar_parity);
          } else {
            umma_pending = true;                          // finished after the pose-block phase, which overlaps the MMAs
          }
          continue;
        }
        if (split_slots) {
          // ---- large chunks: the phase below is bound by shared-memory wavefronts (5 loads per 12 FMAs).  Here a QUAD of
          //      lanes owns a whole 6 x 6 block (ca >= cb), each lane takes every fourth patch (7 loads per 36 FMAs, 18
          //      FFMA2), the quad is summed with two butterfly steps and its lanes share the 18 vector reductions.  The
          //      gradient term y runs the same way on the quads at the top of the block.
          const int npairs = ncols * (ncols + 1) / 2;
          const int nq = b1 - b0;
          const int qs = tid & 3, quad = tid >> 2;
          const unsigned qmask = 0xFu << (lane & 28);        // the quads of a warp have different trip counts
          for (int pr = quad; pr < npairs; pr += 64) {
            int ca = (int)((sqrtf(8.f * pr + 1.f) - 1.f) * 0.5f);
            while (ca * (ca + 1) / 2 > pr) --ca;
            while ((ca + 1) * (ca + 2) / 2 <= pr) ++ca;
            const int cb = pr - ca * (ca + 1) / 2;
            float2 acc[6][3];
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
              for (int k = 0; k < 3; ++k) acc[a][k] = make_float2(0.f, 0.f);
            const float* ea = s.sE + ca * 6;
            const float* eb = s.sE + cb * 6;
            for (int q = qs; q < nq; q += 4) {
              const float Q = s.sQ[q];
              const float2* pa = reinterpret_cast<const float2*>(ea + q * estride);
              const float2* pb2 = reinterpret_cast<const float2*>(eb + q * estride);
              const float2 a01 = pa[0], a23 = pa[1], a45 = pa[2];
              const float2 bk[3] = {pb2[0], pb2[1], pb2[2]};
              const float va[6] = {Q * a01.x, Q * a01.y, Q * a23.x, Q * a23.y, Q * a45.x, Q * a45.y};
#pragma unroll
              for (int a = 0; a < 6; ++a) {
                const float2 v = make_float2(va[a], va[a]);
#pragma unroll
                for (int k = 0; k < 3; ++k) acc[a][k] = __ffma2_rn(v, bk[k], acc[a][k]);
              }
            }
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                float x = acc[a][k].x, y2 = acc[a][k].y;
                x += __shfl_xor_sync(qmask, x, 1); y2 += __shfl_xor_sync(qmask, y2, 1);
                x += __shfl_xor_sync(qmask, x, 2); y2 += __shfl_xor_sync(qmask, y2, 2);
                acc[a][k] = make_float2(x, y2);
              }
            const int fa = ((ca < ch.n_free) ? s.sFrame[ch.first_free + ca] : fi) - t0;
            const int fb = ((cb < ch.n_free) ? s.sFrame[ch.first_free + cb] : fi) - t0;
            if (fa >= fb) {                       // block (fa, fb): row a, column pair k
              float* dst = wp.S + (size_t)(6 * fa) * n6 + 6 * fb;
#pragma unroll
              for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int k = 0; k < 3; ++k)
                  if (((a * 3 + k) & 3) == qs) red_add2(dst + (size_t)a * n6 + 2 * k, -acc[a][k].x, -acc[a][k].y);
            } else {                              // transposed into block (fb, fa): row b, column pair (a, a + 1)
              float* dst = wp.S + (size_t)(6 * fb) * n6 + 6 * fa;
#pragma unroll
              for (int a2 = 0; a2 < 3; ++a2)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                  if (((a2 * 6 + 2 * k) & 3) == qs)
                    red_add2(dst + (size_t)(2 * k) * n6 + 2 * a2, -acc[2 * a2][k].x, -acc[2 * a2 + 1][k].x);
                  if (((a2 * 6 + 2 * k + 1) & 3) == qs)
                    red_add2(dst + (size_t)(2 * k + 1) * n6 + 2 * a2, -acc[2 * a2][k].y, -acc[2 * a2 + 1][k].y);
                }
            }
          }
          for (int ca = 63 - quad; ca < ncols; ca += 64) {          // y[ca] -= sum_p Q_p u_p E_p[ca]
            float2 g01 = make_float2(0.f, 0.f), g23 = g01, g45 = g01;
            for (int q = qs; q < nq; q += 4) {
              const float w = s.sQ[q] * s.sPQ[q * PQS + 1];
              const float2* pa = reinterpret_cast<const float2*>(s.sE + q * estride + ca * 6);
              const float2 v = make_float2(w, w);
              g01 = __ffma2_rn(v, pa[0], g01); g23 = __ffma2_rn(v, pa[1], g23); g45 = __ffma2_rn(v, pa[2], g45);
            }
            float gv[6] = {g01.x, g01.y, g23.x, g23.y, g45.x, g45.y};
#pragma unroll
            for (int a = 0; a < 6; ++a) {
              gv[a] += __shfl_xor_sync(qmask, gv[a], 1);
              gv[a] += __shfl_xor_sync(qmask, gv[a], 2);
            }
            const int fa = ((ca < ch.n_free) ? s.sFrame[ch.first_free + ca] : fi) - t0;
#pragma unroll
            for (int a = 0; a < 6; ++a)
              if ((a & 3) == qs) atomicAdd(&wp.y[6 * fa + a], -gv[a]);
          }
          continue;
        }
        // ---- Schur update of this batch.  Item = (column pair ca >= cb, row pair 2*a2, 2*a2+1): twelve sums over
        //      the patches, as packed fp32x2 FMAs (FFMA2).  M = sum_p Q_p E_p[ca] E_p[cb]^T is block (ca, cb); it
        //      lands at (frame(ca), frame(cb)) or transposed.
        const int npairs = ncols * (ncols + 1) / 2;
        const int nq = b1 - b0;
        for (int it = tid; it < npairs * 3; it += 256) {
          const int pr = it / 3, a2 = it - pr * 3;
          int ca = (int)((sqrtf(8.f * pr + 1.f) - 1.f) * 0.5f);
          while (ca * (ca + 1) / 2 > pr) --ca;
          while ((ca + 1) * (ca + 2) / 2 <= pr) ++ca;
          const int cb = pr - ca * (ca + 1) / 2;
          float2 acc0[3], acc1[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) { acc0[k] = make_float2(0.f, 0.f); acc1[k] = make_float2(0.f, 0.f); }
          const float* ea = s.sE + ca * 6 + 2 * a2;
          const float* eb = s.sE + cb * 6;
          for (int q = 0; q < nq; ++q) {
            const float Q = s.sQ[q];
            const float2 va = *reinterpret_cast<const float2*>(ea + q * estride);
            const float2 v0 = make_float2(Q * va.x, Q * va.x), v1 = make_float2(Q * va.y, Q * va.y);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const float2 bk = *reinterpret_cast<const float2*>(eb + q * estride + 2 * k);
              acc0[k] = __ffma2_rn(v0, bk, acc0[k]);
              acc1[k] = __ffma2_rn(v1, bk, acc1[k]);
            }
          }
          const int fa = ((ca < ch.n_free) ? s.sFrame[ch.first_free + ca] : fi) - t0;
          const int fb = ((cb < ch.n_free) ? s.sFrame[ch.first_free + cb] : fi) - t0;
          if (fa >= fb) {                       // block (fa, fb), rows 2*a2, 2*a2 + 1
            float* dst = wp.S + (size_t)(6 * fa + 2 * a2) * n6 + 6 * fb;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              red_add2(dst + 2 * k, -acc0[k].x, -acc0[k].y);
              red_add2(dst + n6 + 2 * k, -acc1[k].x, -acc1[k].y);
            }
          } else {                              // transposed into block (fb, fa): columns 2*a2, 2*a2 + 1
            float* dst = wp.S + (size_t)(6 * fb) * n6 + 6 * fa + 2 * a2;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              red_add2(dst + (size_t)(2 * k) * n6, -acc0[k].x, -acc1[k].x);
              red_add2(dst + (size_t)(2 * k + 1) * n6, -acc0[k].y, -acc1[k].y);
            }
          }
        }
        // y[ca] -= sum_p Q_p u_p E_p[ca]
        // (taken from the top of the block: the threads the Schur items above leave idle)
        for (int x = 255 - tid; x < ncols * 6; x += 256) {
          const int ca = x / 6, a = x - ca * 6;
          float acc = 0.f;
          for (int q = 0; q < nq; ++q) acc += s.sQ[q] * s.sPQ[q * PQS + 1] * s.sE[q * estride + x];
          const int fa = ((ca < ch.n_free) ? s.sFrame[ch.first_free + ca] : fi) - t0;
          atomicAdd(&wp.y[6 * fa + a], -acc);
        }
      }
    }
    LIN_TS(7);
    // With MMAs in flight (umma_pending) the pose-block phase below runs on warps 0..6 only, synchronised by a named barrier
    // of 224 threads, while the first thread of warp 7 is still issuing; everything the phase reads was final before the
    // Schur phase, and schur_umma_issue ended with a CTA-wide barrier.
    if (!umma_pending) __syncthreads();
    LIN_TS(8);
    const bool b_active = !(umma_pending && warp == 7);
    const int b_stride = umma_pending ? 224 : 256;
    auto b_sync = [&]() {
      if (!umma_pending) __syncthreads();
      else if (warp < 7) named_bar_sync(3, 224);
    };

    // ---- pose blocks of this chunk (ba_cuda.cu:364-398)
    if (N > 0) {
      const int io = 6 * (fi - t0);
      if (i_free) {
        // B1: AH_s = A_s H_s, column b per warp (warps 0..5), slots on lanes; warp 6: v_i -= sum_s A_s g_s
        float vi[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int sb = 0; sb < ns; sb += 32) {
          const int sl = sb + lane;
          if (warp < 7 && sl < ns) {
            float R[9], t[3], colv[6], outv[6];
#pragma unroll
            for (int x = 0; x < 9; ++x) R[x] = s.sRt[sl * 12 + x];
#pragma unroll
            for (int x = 0; x < 3; ++x) t[x] = s.sRt[sl * 12 + 9 + x];
            if (warp < 6) {
#pragma unroll
              for (int a = 0; a < 6; ++a) colv[a] = s.sH[sl * 28 + (a <= warp ? sym6(a, warp) : sym6(warp, a))];
              adj_map(R, t, colv, outv);
#pragma unroll
              for (int a = 0; a < 6; ++a) sAH[sl * 36 + a * 6 + warp] = outv[a];
            } else {
#pragma unroll
              for (int a = 0; a < 6; ++a) colv[a] = s.sH[sl * 28 + 21 + a];
              adj_map(R, t, colv, outv);
#pragma unroll
              for (int a = 0; a < 6; ++a) vi[a] += outv[a];
            }
          }
        }
        if (warp == 6) {
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            float v = vi[a];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s.sBii[36 + a] = -v;
          }
        }
        b_sync();
    LIN_TS(9);
        // B2: B_ii = sum_s AH_s A_s^T ; row a per warp: (AH A^T)[a][:] = A * (AH[a][:])^T
        if (warp < 6) {
          float bi[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int sb = 0; sb < ns; sb += 32) {
            const int sl = sb + lane;
            if (sl < ns) {
              float R[9], t[3], rowv[6], outv[6];
#pragma unroll
              for (int x = 0; x < 9; ++x) R[x] = s.sRt[sl * 12 + x];
#pragma unroll
              for (int x = 0; x < 3; ++x) t[x] = s.sRt[sl * 12 + 9 + x];
#pragma unroll
              for (int b = 0; b < 6; ++b) rowv[b] = sAH[sl * 36 + warp * 6 + b];
              adj_map(R, t, rowv, outv);
#pragma unroll
              for (int b = 0; b < 6; ++b) bi[b] += outv[b];
            }
          }
#pragma unroll
          for (int b = 0; b < 6; ++b) {
            float v = bi[b];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s.sBii[warp * 6 + b] = v;
          }
        }
        b_sync();
    LIN_TS(10);
      }
      // B3: scatter.  Item = (slot, row a).
      for (int it = b_active ? tid : ns * 6; it < ns * 6; it += b_stride) {
        const int sl = it / 6, a = it - sl * 6;
        const int fj = s.sFrame[sl];
        if (fj < t0 || fj >= pb.t1) continue;
        const int jo = 6 * (fj - t0);
        {                                                               // B_jj += H ; v_j += g
          float h[6];
#pragma unroll
          for (int b = 0; b < 6; ++b) h[b] = s.sH[sl * 28 + (a <= b ? sym6(a, b) : sym6(b, a))];
          float* dst = wp.S + (size_t)(jo + a) * n6 + jo;
          red_add2(dst, h[0], h[1]); red_add2(dst + 2, h[2], h[3]); red_add2(dst + 4, h[4], h[5]);
          atomicAdd(&wp.y[jo + a], s.sH[sl * 28 + 21 + a]);
        }
        if (i_free) {                                                   // B_ij -= AH, B_ji -= AH^T
          if (fi > fj) {
            float* dst = wp.S + (size_t)(io + a) * n6 + jo;
            const float* r = sAH + sl * 36 + a * 6;
            red_add2(dst, -r[0], -r[1]); red_add2(dst + 2, -r[2], -r[3]); red_add2(dst + 4, -r[4], -r[5]);
          } else if (fi < fj) {
            float* dst = wp.S + (size_t)(jo + a) * n6 + io;              // row a of AH^T = column a of AH
            const float* r = sAH + sl * 36 + a;
            red_add2(dst, -r[0], -r[6]); red_add2(dst + 2, -r[12], -r[18]); red_add2(dst + 4, -r[24], -r[30]);
          } else {
            float* dst = wp.S + (size_t)(io + a) * n6 + io;
            const float* r = sAH + sl * 36;
            red_add2(dst, -(r[a * 6 + 0] + r[0 * 6 + a]), -(r[a * 6 + 1] + r[1 * 6 + a]));
            red_add2(dst + 2, -(r[a * 6 + 2] + r[2 * 6 + a]), -(r[a * 6 + 3] + r[3 * 6 + a]));
            red_add2(dst + 4, -(r[a * 6 + 4] + r[4 * 6 + a]), -(r[a * 6 + 5] + r[5 * 6 + a]));
          }
        }
      }
      if (i_free && tid < 6) {                                          // B_ii, v_i
        float* dst = wp.S + (size_t)(io + tid) * n6 + io;
        const float* r = s.sBii + tid * 6;
        red_add2(dst, r[0], r[1]); red_add2(dst + 2, r[2], r[3]); red_add2(dst + 4, r[4], r[5]);
        atomicAdd(&wp.y[io + tid], s.sBii[36 + tid]);
      }
    }
    if (use_umma && umma_pending)
      mbar_parity = schur_umma_finish(s.sFrame, ch.first_free, ch.n_free, wp.S, wp.y, ncols, fi, t0, n6, s_tmem, &s_mbar,
                                      mbar_parity);
    LIN_TS(11);
  }
  if (!waited) {                                                // a CTA without a chunk still has to release its dependents
    pdl_wait();
    pdl_trigger();
    CTA_TS((flags & 1) ? 3 : 0, 1);
  }
  if (umma_cta) {
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_dealloc(s_tmem);
  }
  CTA_TS((flags & 1) ? 3 : 0, 2);
}

#ifdef PGBA_LIN_TIMING
void lin_timestamps(long long* out) { cudaMemcpyFromSymbol(out, g_lin_ts, sizeof(long long) * 32); }
void cta_timestamps(unsigned long long* out) { cudaMemcpyFromSymbol(out, g_cta_ts, sizeof(unsigned long long) * 4 * 3 * 512); }
#endif

// ---------------------------------------------------------------------------------------------------------------
// Small dense solve: one CTA per window, fp32 in shared memory (the reference's own precision: ba_cuda.cu:576-577),
// blocked (6 wide) right-looking Cholesky on the matrix augmented with the right-hand side as an extra row (so the
// forward substitution comes for free), inverse-based warp-level backward substitution (ba_chol32.cuh), then the SE3
// retraction of the free poses.
// ---------------------------------------------------------------------------------------------------------------
size_t solve_small_smem_bytes(int n) {
  const int ld = n | 1;
  return sizeof(float) * ((size_t)(n + 1) * ld + n + 6 * (size_t)n + 8);
}

__global__ void __launch_bounds__(256, 1) solve_small_kernel(Problem pb) {
  CTA_TS(1, 0);
  pdl_wait();
  pdl_trigger();
  CTA_TS(1, 1);
  extern __shared__ float sf[];
  SOLVE_TS(0);
  const int w = blockIdx.x + pb.w0, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  const int N = pb.t1 - pb.t0, n = 6 * N, ld = n | 1;
  float* A = sf;                   // [n + 1][ld]; row n = right-hand side
  float* rd = sf + (n + 1) * ld;   // [n] reciprocal diagonal of L
  float* dinv = rd + n;            // [n / 6][36] inverses of the diagonal blocks of L
  const bool rezero = pb.apply != 0;
  // prefetch the pose rows that are retracted at the end
  float* prow = pb.poses + (int64_t)w * pb.st.poses + 7 * (int64_t)(pb.t0 + tid);
  float pose[7];
  if (tid < N) {
#pragma unroll
    for (int x = 0; x < 7; ++x) pose[x] = prow[x];
  }
  // ---- load S (only the lower block triangle is meaningful) with the damping S += I * (1e-4 * S + 1)
  //      (ba_cuda.cu:575/589); all loads are independent.  n*n is a multiple of 4.
  {
    const float4* S4 = reinterpret_cast<const float4*>(wp.S);
    const int n4 = (n * n) >> 2;
    // four 16-byte loads per thread in flight before the first shared-memory store (n = 60: the whole matrix in ONE round
    // trip; issued one per store they were four dependent trips)
    const float yv = (tid < n) ? wp.y[tid] : 0.f;
    for (int x0 = tid; x0 < n4; x0 += 4 * 256) {
      float4 v4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v4[u] = (x0 + 256 * u < n4) ? S4[x0 + 256 * u] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int x4 = x0 + 256 * u;
        if (x4 >= n4) break;
        const float vv[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w};
        int r = (4 * x4) / n, c = 4 * x4 - r * n;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float v = vv[k];
          if (r == c) v += 1e-4f * v + 1.0f;
          A[r * ld + c] = v;
          if (++c == n) { c = 0; ++r; }
        }
      }
    }
    if (tid < n) A[n * ld + tid] = yv;            // n <= 156 < 256
  }
  SOLVE_TS(1);
  chol6_f32(A, rd, dinv, n, ld);        // rows 0..n: the rhs row n rides along (forward substitution)
  SOLVE_TS(2);
  if (warp == 0) {
    backsub6_f32_any(A, dinv, n, ld, lane);
  } else if (rezero) {             // S, y are consumed: the other warps clear them for the next iteration's accumulation
    float4* S4 = reinterpret_cast<float4*>(wp.S);
    const int n4 = (n * n) >> 2;
    for (int x4 = tid - 32; x4 < n4; x4 += 224) S4[x4] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int x = tid - 32; x < n; x += 224) wp.y[x] = 0.f;
  }
  __syncthreads();
  SOLVE_TS(3);
  const float* xv = A + n * ld;
  for (int x = tid; x < n; x += 256) wp.dX[x] = xv[x];
  // ---- SE3 retraction of the free poses (ba_cuda.cu:178-206)
  if (pb.apply && tid < N) {
    float xi[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) xi[a] = xv[6 * tid + a];
    retract_pose(pose, xi);
#pragma unroll
    for (int x = 0; x < 7; ++x) prow[x] = pose[x];
  }
  SOLVE_TS(4);
  CTA_TS(1, 2);
}

// ---------------------------------------------------------------------------------------------------------------
// Back-substitution dZ = Q (u - E^T dX) and inverse-depth retraction (ba_cuda.cu:592, 209-229; block_e.cu:253-283)
// grid = (gx, batch), block = 256: one warp per patch, lanes over the E row.  apply = 0 only computes dZ.
// ---------------------------------------------------------------------------------------------------------------
// Per-chunk back-substitution + depth retraction by the whole CTA (256 threads).  sdx: >= (SMAX + 1) * 6 floats of
// shared memory, 8-byte aligned.  Ends with a __syncthreads().
__device__ __forceinline__ UpdPre upd_prefetch(const Problem& pb, const WinPtrs& wp, const Chunk& ch, const float* patches,
                                               bool apply) {
  const int tid = threadIdx.x;
  const int N = pb.t1 - pb.t0;
  const int PP = pb.P * pb.P, pstride = 3 * PP;
  const int ncols = (N > 0 && pb.with_schur != 0) ? ch.ncols : 0;
  const int len2 = ncols * 3;
  UpdPre pre;
  pre.dxi = -1;
  if (tid < ncols * 6) {
    const int col = tid / 6, a = tid - col * 6;
    const int f = (col < ch.n_free) ? wp.slots[ch.slot_base + ch.first_free + col] : ch.frame;
    pre.dxi = 6 * (f - pb.t0) + a;
  }
#pragma unroll
  for (int r = 0; r < UPD_PRE; ++r) {
    pre.e0[r] = make_float2(0.f, 0.f); pre.e1[r] = make_float2(0.f, 0.f);
    pre.q[r] = 0.f; pre.u[r] = 0.f; pre.d[r] = 0.f; pre.kx[r] = 0;
    const int p = (tid >> 5) + 8 * r, lane = tid & 31;
    if (pb.L.pc <= 32 && p < ch.n_patches) {
      const float2* eg = reinterpret_cast<const float2*>(wp.ecells + 6 * ((int64_t)ch.ecell_base + (int64_t)p * ch.ncols));
      if (lane < len2) pre.e0[r] = eg[lane];
      if (lane + 32 < len2) pre.e1[r] = eg[lane + 32];
      pre.q[r] = wp.Q[ch.patch_base + p];
      pre.u[r] = wp.u[ch.patch_base + p];
      pre.kx[r] = wp.kx[ch.patch_base + p];
      if (apply) pre.d[r] = patches[(int64_t)pre.kx[r] * pstride + 2 * PP];       // reads [2][0][0] (ba_cuda.cu:218)
    }
  }
  return pre;
}

// s_depth (optional, shared memory, >= n_patches floats): receives the retracted inverse depth of every patch of the chunk.
__device__ __forceinline__ void chunk_depth_update(const Problem& pb, const WinPtrs& wp, const Chunk& ch, float* patches,
                                                   float* sdx, bool apply, const UpdPre& pre, float* s_depth) {
  const int tid = threadIdx.x;
  const int N = pb.t1 - pb.t0, t0 = pb.t0;
  const int PP = pb.P * pb.P, pstride = 3 * PP;
  const int ncols = (N > 0 && pb.with_schur != 0) ? ch.ncols : 0;
  if (pre.dxi >= 0) sdx[tid] = wp.dX[pre.dxi];
  for (int x = tid + 256; x < ncols * 6; x += 256) {
    const int col = x / 6, a = x - col * 6;
    const int f = (col < ch.n_free) ? wp.slots[ch.slot_base + ch.first_free + col] : ch.frame;
    sdx[x] = wp.dX[6 * (f - t0) + a];
  }
  const int len2 = ncols * 3;
  __syncthreads();
  const float2* dx2 = reinterpret_cast<const float2*>(sdx);
  if (pb.L.pc > 32) {
    // large chunks: one thread per patch, its E row (ncols * 6 floats, 8-byte aligned) is read with independent
    // 8-byte loads (no shuffles; consecutive threads read consecutive rows)
    for (int p = tid; p < ch.n_patches; p += 256) {
      const float2* eg = reinterpret_cast<const float2*>(wp.ecells + 6 * ((int64_t)ch.ecell_base + (int64_t)p * ch.ncols));
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 4
      for (int x = 0; x < len2; ++x) {
        const float2 e = eg[x], d = dx2[x];
        acc0 += e.x * d.x;
        acc1 += e.y * d.y;
      }
      const float dz = wp.Q[ch.patch_base + p] * (wp.u[ch.patch_base + p] - (acc0 + acc1));
      wp.dZ[ch.patch_base + p] = dz;
      if (apply) {
        float* pr = patches + (int64_t)wp.kx[ch.patch_base + p] * pstride + 2 * PP;
        float d = pr[0] + dz;                   // reads [2][0][0] (ba_cuda.cu:218)
        d = (d > 20.f) ? 1.0f : d;
        d = fmaxf(d, 1e-4f);
        for (int x = 0; x < PP; ++x) pr[x] = d;
        if (s_depth) s_depth[p] = d;
      }
    }
  } else {
    // small chunks (single window): one warp per patch, lanes over the E row.  Latency-bound: everything that does not
    // depend on dX (E row, Q, u, patch id -> old depth) was requested before the barrier above (pre[] below).
    const int lane = tid & 31, warp = tid >> 5;
    int r = 0;
    for (int p = warp; p < ch.n_patches; p += 8, ++r) {
      const float2* eg = reinterpret_cast<const float2*>(wp.ecells + 6 * ((int64_t)ch.ecell_base + (int64_t)p * ch.ncols));
      const bool first = r < UPD_PRE;                  // fetched ahead (upd_prefetch)
      float2 pe0 = make_float2(0.f, 0.f), pe1 = pe0;
      float pq = 0.f, pu = 0.f, pd = 0.f;
      int pkx = 0;
#pragma unroll
      for (int rr = 0; rr < UPD_PRE; ++rr)
        if (rr == r) { pe0 = pre.e0[rr]; pe1 = pre.e1[rr]; pq = pre.q[rr]; pu = pre.u[rr]; pd = pre.d[rr]; pkx = pre.kx[rr]; }
      float acc = 0.f;
      for (int x = lane, k = 0; x < len2; x += 32, ++k) {
        const float2 e = (first && k < 2) ? (k == 0 ? pe0 : pe1) : eg[x];
        const float2 d = dx2[x];
        acc += e.x * d.x + e.y * d.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      const float q = first ? pq : wp.Q[ch.patch_base + p], u = first ? pu : wp.u[ch.patch_base + p];
      const float dz = q * (u - acc);
      if (lane == 0) wp.dZ[ch.patch_base + p] = dz;
      if (apply) {
        float* pr = patches + (int64_t)(first ? pkx : wp.kx[ch.patch_base + p]) * pstride + 2 * PP;
        float d = (first ? pd : pr[0]) + dz;
        d = (d > 20.f) ? 1.0f : d;
        d = fmaxf(d, 1e-4f);
        __syncwarp();
        for (int x = lane; x < PP; x += 32) pr[x] = d;
        if (s_depth && lane == 0) s_depth[p] = d;
      }
    }
  }
  __syncthreads();
}

// early != 0: the kernel before this one is the small solve, which passes its own pdl_wait() -- i.e. the linearisation
// that wrote E, Q, u has completed -- before it lets this kernel start.  Everything except dX (chunk table, E rows, Q, u,
// old depths) is then final when the CTA begins and is requested BEFORE pdl_wait(), while the solve is still running.
__global__ void __launch_bounds__(256) update_kernel(Problem pb, int early) {
  CTA_TS(2, 0);
  bool waited = early == 0;
  if (waited) { pdl_wait(); pdl_trigger(); CTA_TS(2, 1); }
  __shared__ __align__(8) float sdx[(SMAX + 1) * 6];
  const int w = blockIdx.x + pb.w0;                  // grid = (windows, chunk slots), see linearize_kernel
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  float* patches = pb.patches + (int64_t)w * pb.st.patches;
  const int n_chunks = wp.hdr->n_chunks;
  for (int c = blockIdx.y; c < n_chunks; c += gridDim.y) {
    const Chunk ch = wp.chunks[c];
    if (ch.n_patches == 0) continue;
    __syncthreads();
    const UpdPre pre = upd_prefetch(pb, wp, ch, patches, pb.apply != 0);
    if (!waited) { pdl_wait(); pdl_trigger(); waited = true; CTA_TS(2, 1); }
    chunk_depth_update(pb, wp, ch, patches, sdx, pb.apply != 0, pre, nullptr);
  }
  if (!waited) { pdl_wait(); pdl_trigger(); CTA_TS(2, 1); }
  CTA_TS(2, 2);
}

// Back-substitution + depth retraction for large chunks (batched windows): one thread per patch.  The chunk's E tile
// (n_patches x ncols x 6 floats, contiguous) is staged in shared memory with coalesced loads, and Q, u, the patch id and
// the old depth are fetched, BEFORE pdl_wait() when the previous kernel is the small solve (early != 0, see
// update_kernel): on the batched shape the solve occupies one SM per window for ~20 us, during which the other SMs
// pull the whole E matrix on chip; after the wait only dX is loaded.
// dynamic smem: tile_floats (E tile) + (SMAX + 1) * 6 (dX of the chunk's columns)
__global__ void __launch_bounds__(256) update_large_kernel(Problem pb, int early, int tile_floats) {
  CTA_TS(2, 0);
  extern __shared__ __align__(16) float usm[];
  float* sE = usm;
  float* sdx = usm + tile_floats;
  bool waited = early == 0;
  if (waited) { pdl_wait(); pdl_trigger(); CTA_TS(2, 1); }
  const int w = blockIdx.x + pb.w0, tid = threadIdx.x;   // grid = (windows, chunk slots), see linearize_kernel
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  float* patches = pb.patches + (int64_t)w * pb.st.patches;
  const int N = pb.t1 - pb.t0, t0 = pb.t0;
  const int PP = pb.P * pb.P, pstride = 3 * PP;
  const bool apply = pb.apply != 0;
  const int n_chunks = wp.hdr->n_chunks;
  // the grid is sized for the worst-case chunk count: a CTA without a chunk leaves at once (an exiting CTA needs no wait,
  // and while it sat at pdl_wait() it would hold a slot that a CTA with a tile to prefetch could use)
  if ((int)blockIdx.y >= n_chunks) return;
  for (int c = blockIdx.y; c < n_chunks; c += gridDim.y) {
    const Chunk ch = wp.chunks[c];
    if (ch.n_patches == 0) continue;
    __syncthreads();
    const int np = ch.n_patches;
    const int ncols = (N > 0 && pb.with_schur != 0) ? ch.ncols : 0;
    const int len2 = ncols * 3;                                      // float2 per E row
    {
      const float2* eg = reinterpret_cast<const float2*>(wp.ecells + 6 * (int64_t)ch.ecell_base);
      float2* dst = reinterpret_cast<float2*>(sE);
      for (int x = tid; x < np * len2; x += 256) dst[x] = eg[x];
    }
    float q = 0.f, u = 0.f, d_old = 0.f;
    int kx = 0, dxi = -1;
    if (tid < np) {
      q = wp.Q[ch.patch_base + tid];
      u = wp.u[ch.patch_base + tid];
      kx = wp.kx[ch.patch_base + tid];
      if (apply) d_old = patches[(int64_t)kx * pstride + 2 * PP];    // reads [2][0][0] (ba_cuda.cu:218)
    }
    if (tid < ncols * 6) {
      const int col = tid / 6, a = tid - col * 6;
      const int f = (col < ch.n_free) ? wp.slots[ch.slot_base + ch.first_free + col] : ch.frame;
      dxi = 6 * (f - t0) + a;
    }
    if (!waited) { pdl_wait(); pdl_trigger(); waited = true; CTA_TS(2, 1); }
    if (dxi >= 0) sdx[tid] = wp.dX[dxi];
    for (int x = tid + 256; x < ncols * 6; x += 256) {
      const int col = x / 6, a = x - col * 6;
      const int f = (col < ch.n_free) ? wp.slots[ch.slot_base + ch.first_free + col] : ch.frame;
      sdx[x] = wp.dX[6 * (f - t0) + a];
    }
    __syncthreads();
    for (int p = tid; p < np; p += 256) {                            // np <= PMAX <= 256: one trip
      const float2* er = reinterpret_cast<const float2*>(sE) + p * len2;
      const float2* dx2 = reinterpret_cast<const float2*>(sdx);
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 4
      for (int x = 0; x < len2; ++x) {
        const float2 e = er[x], d = dx2[x];
        acc0 += e.x * d.x;
        acc1 += e.y * d.y;
      }
      const float dz = q * (u - (acc0 + acc1));
      wp.dZ[ch.patch_base + p] = dz;
      if (apply) {
        float* pr = patches + (int64_t)kx * pstride + 2 * PP;
        float d = d_old + dz;
        d = (d > 20.f) ? 1.0f : d;
        d = fmaxf(d, 1e-4f);
        for (int x = 0; x < PP; ++x) pr[x] = d;
      }
    }
  }
  if (!waited) { pdl_wait(); pdl_trigger(); CTA_TS(2, 1); }
  CTA_TS(2, 2);
}

// ---------------------------------------------------------------------------------------------------------------
// Host-side launch sequence of one Gauss-Newton iteration
// ---------------------------------------------------------------------------------------------------------------
int lin_ebudget(const Problem& pb) {
  const int N = pb.t1 - pb.t0;
  const int64_t cols = (N < SMAX ? N : SMAX) + 1;
  int64_t need = (int64_t)pb.L.pc * 6 * cols;
  if (need > EBUDGET) need = EBUDGET;
  if (need < 64) need = 64;
  return (int)need;
}

cudaError_t launch_big_solve(const Problem& pb, int64_t batch, cudaStream_t stream);
bool nd_active(const Problem& pb);
cudaError_t launch_nd_solve(const Problem& pb, int64_t batch, cudaStream_t stream, bool more);

// Chunk Schur product on the tcgen05 tensor cores (schur_umma_issue / _finish) for chunks of >= 64 patches (batched windows,
// global BA): opt-in, PGBA_SCHUR_UMMA=1.  Measured on the B200 (c5, 64 windows, chunks of 96 patches x 10 columns, same box,
// profiles/README.md round 2): linearisation 133 us against 112 us for the FFMA2 quad product -- the 36 small dependent
// MMAs (M = 64, N = 64, K = 8, ~100 cycles each) and the tensor core itself are cheap, but the in-place TF32 hi / lo
// staging of the operand arrays (4.7 k cycles) and the 64-row epilogue (2.6 k) cost as much as the 9.2 k cycles of packed
// FMAs they replace, and busy CTAs run 22.3 instead of 20.0 us with two CTAs per SM.  Accuracy: S to 2.8e-7 .. 6.3e-7 of the
// float64 oracle (FFMA2: 1.2e-7).  Kept, parity-tested (tests/test_ba_gpu.py::test_forced_chunk_size), for the A/B record.
static bool schur_on_tcgen05(const Problem& pb) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PGBA_SCHUR_UMMA");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  // the operand arrays alias the E tile (>= 27 648 bytes) and the per-warp partials region
  return v != 0 && pb.L.pc >= 64 && pb.t1 > pb.t0 && (size_t)lin_ebudget(pb) * sizeof(float) >= (size_t)umma::X_BYTES;
}

// Loads ahead of pdl_wait() in the second linearisation / the update kernel (PGBA_EARLY=0 disables, for A/B runs).
static bool early_loads_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PGBA_EARLY");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

void launch_linearize(const Problem& pb, int64_t batch, cudaStream_t stream, bool fuse_update, bool first) {
  const int gx = chunk_grid(pb, batch);
  const int ebudget = lin_ebudget(pb);
  const bool umma_on = schur_on_tcgen05(pb);
  static int extra_smem = -1;                         // PGBA_LIN_EXTRA_SMEM=1: size the FFMA2 instance like the tcgen05 one (A/B probe)
  if (extra_smem < 0) { const char* e = getenv("PGBA_LIN_EXTRA_SMEM"); extra_smem = (e && e[0] == '1') ? 1 : 0; }
  const size_t lsm = lin_smem_bytes(pb.L.pc, ebudget, umma_on || extra_smem);
  const bool early = fuse_update && pb.t1 > pb.t0 && !pb.L.big && early_loads_enabled();
  const int flags = (fuse_update ? 1 : 0) | (early ? 8 : 0) | (umma_on ? 16 : 0) | (first ? 32 : 64);
  auto go = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm);
    launch_k(kern, dim3((unsigned)batch, (unsigned)gx), dim3(256), lsm, stream, pb, ebudget, flags);
  };
  if (flags & 16) {
    if (fuse_update) go(linearize_kernel<true, true>); else go(linearize_kernel<false, true>);
  } else {
    if (fuse_update) go(linearize_kernel<true, false>); else go(linearize_kernel<false, false>);
  }
  count_launch();
}

// Solve S dX = y (damped), retract the poses (if pb.apply).  Small systems: one CTA per window, S and y are re-zeroed
// by the kernel.  Large systems: blocked Cholesky in global memory; S holds the factor afterwards.
void launch_solve(const Problem& pb, int64_t batch, cudaStream_t stream, bool more) {
  const int N = pb.t1 - pb.t0;
  if (N <= 0) return;
  if (pb.L.big) {
    if (nd_active(pb)) launch_nd_solve(pb, batch, stream, more);
    else launch_big_solve(pb, batch, stream);
    return;
  }
  const size_t smem = solve_small_smem_bytes(6 * N);
  cudaFuncSetAttribute(solve_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  launch_k(solve_small_kernel, dim3((unsigned)batch), dim3(256), smem, stream, pb);
  count_launch();
}

void launch_update(const Problem& pb, int64_t batch, cudaStream_t stream) {
  const int gx = chunk_grid(pb, batch);
  const int N = pb.t1 - pb.t0;
  const int early = (N > 0 && !pb.L.big && early_loads_enabled()) ? 1 : 0;
  const int64_t cols = (N < SMAX ? N : SMAX) + 1;
  const int64_t tile_floats = (int64_t)pb.L.pc * cols * 6;
  static int large_on = -1;                        // PGBA_UPDATE_LARGE=0: the generic kernel for every chunk size (A/B runs)
  if (large_on < 0) {
    const char* e = getenv("PGBA_UPDATE_LARGE");
    large_on = (e && e[0] == '0') ? 0 : 1;
  }
  if (large_on && pb.L.pc > 32 && tile_floats * 4 <= 48 * 1024) {
    const size_t smem = sizeof(float) * ((size_t)tile_floats + (SMAX + 1) * 6);
    cudaFuncSetAttribute(update_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(update_large_kernel, dim3((unsigned)batch, (unsigned)gx), dim3(256), smem, stream, pb, early, (int)tile_floats);
  } else {
    launch_k(update_kernel, dim3((unsigned)batch, (unsigned)gx), dim3(256), 0, stream, pb, early);
  }
  count_launch();
}

// After a large solve S holds the Cholesky factor: clear S, y of every window before the next linearisation.
cudaError_t clear_big_system(const Problem& pb, int64_t batch, cudaStream_t stream) {
  cudaError_t e = cudaSuccess;
  for (int64_t w = 0; w < batch && e == cudaSuccess; ++w)
    e = cudaMemsetAsync((char*)pb.ws + (size_t)w * pb.L.zero_bytes + pb.L.z_y, 0, pb.L.zero_bytes - pb.L.z_y, stream);
  return e;
}

// One Gauss-Newton iteration.  The back-substitution + depth retraction of iteration k is fused into the linearisation
// of iteration k + 1 (`first` = false); only the last iteration (`more` = false) launches update_kernel.
// ev (optional): 4 events recorded before linearize / solve / update and after update
cudaError_t launch_iteration(const Problem& pb, int64_t batch, cudaStream_t stream, cudaEvent_t* ev, bool first, bool more) {
  // fusing pays in the latency-bound single-window regime (small chunks); with large chunks (batched windows) the
  // separate, fully parallel update kernel measured faster
  const bool fuse = pb.L.pc <= 32;
  if (ev) cudaEventRecord(ev[0], stream);
  launch_linearize(pb, batch, stream, fuse && !first, first);
  if (ev) cudaEventRecord(ev[1], stream);
  launch_solve(pb, batch, stream, more);
  if (ev) cudaEventRecord(ev[2], stream);
  if (!fuse || !more) launch_update(pb, batch, stream);
  if (pb.L.big && more && !nd_active(pb)) {        // (the reordered solve re-zeroes what it touched itself)
    cudaError_t e = clear_big_system(pb, batch, stream);
    if (e != cudaSuccess) return e;
  }
  if (ev) cudaEventRecord(ev[3], stream);
  return cudaGetLastError();
}

bool solve_supported(int N) { return N <= PGBA_MAX_POSE_ROWS; }

}  // namespace pgba
