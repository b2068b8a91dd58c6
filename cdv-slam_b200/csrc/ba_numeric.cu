// Numeric phase of the patch-graph BA (sm_100a): linearisation + Schur complement, small dense solve,
// back-substitution + retraction.
//
//   linearize_kernel   per chunk (one source frame i, <= 128 patches, <= 128 target-frame slots):
//                      lanes <-> target-frame slots, loop over patches.  Per edge (reference arithmetic:
//                      cdvslam/fastba/ba_cuda.cu:265-343): reprojection residual, Jj rows, Jz.  Because
//                      Ji = Ad^T(Gij) Jj is the same linear map for every edge of a frame pair (ba_cuda.cu:353,
//                      57-72), only H_ij = sum w Jj Jj^T and g_ij = sum w r Jj are accumulated per edge (registers,
//                      no atomics); B_ii += A H A^T, B_ij -= A H, B_jj += H, v_i -= A g, v_j += g are formed once
//                      per frame pair (ba_cuda.cu:364-398).  Per patch: C, u, E columns (ba_cuda.cu:380-402) by
//                      warp shuffles; Q = 1/(C+lambda); the chunk's Schur update S -= E Q E^T, y -= E Q u
//                      (ba_cuda.cu:583-587, block_e.cu:147-234) is applied from shared memory; E is kept for the
//                      back-substitution in a compact [patch][column][6] layout.
//   solve_small_kernel S += I*(1e-4*S+1) (ba_cuda.cu:575/589), Cholesky + solves (ba_cuda.cu:576-577/590-591) in
//                      fp64 shared memory, one CTA per window, 6N <= 156.
//   update_kernel      dZ = Q (u - E^T dX) (ba_cuda.cu:592, block_e.cu:253-283) + inverse-depth retraction
//                      (ba_cuda.cu:209-229)
//   pose_retr_kernel   SE3 retraction (ba_cuda.cu:88-206)
#include "ba_common.cuh"

namespace pgba {

struct EdgeOut {
  float H[21];   // sum_rows w * Jj Jj^T (upper triangle)
  float g[6];    // sum_rows w * r * Jj
  float e[6];    // sum_rows w * Jz * Jj
  float c, u;    // sum_rows w * Jz^2, sum_rows w * r * Jz
};

// Per-edge terms.  px, py, pd: patch centre and inverse depth; R, t: relative pose Gij.
__device__ __forceinline__ void edge_terms(float px, float py, float pd, float fx, float fy, float cx, float cy,
                                           const float R[9], const float t[3], float2 tg, float2 wt, float H[21],
                                           float g[6], float e[6], float& c_out, float& u_out) {
  const float xi0 = (px - cx) / fx, xi1 = (py - cy) / fy;
  const float X = R[0] * xi0 + R[1] * xi1 + R[2] + pd * t[0];
  const float Y = R[3] * xi0 + R[4] * xi1 + R[5] + pd * t[1];
  const float Z = R[6] * xi0 + R[7] * xi1 + R[8] + pd * t[2];
  const float W = pd;
  const float d = (Z >= 0.2f) ? 1.0f / Z : 0.0f;
  const float d2 = d * d;
  const float x1 = fx * (X / Z) + cx;
  const float y1 = fy * (Y / Z) + cy;
  const float rx = tg.x - x1, ry = tg.y - y1;
  const bool in_bounds = (sqrtf(rx * rx + ry * ry) < 128.f) && (Z > 0.2f) && (x1 > -64.f) && (y1 > -64.f) &&
                         (x1 < 2.f * cx + 64.f) && (y1 < 2.f * cy + 64.f);
  const float mask = in_bounds ? 1.0f : 0.0f;
  const float wx = mask * wt.x, wy = mask * wt.y;
  // Jj rows (ba_cuda.cu:323-341); Jx[1] == 0 and Jy[0] == 0
  const float Jx0 = fx * W * d, Jx2 = -fx * X * W * d2, Jx3 = -fx * X * Y * d2, Jx4 = fx * (1.0f + X * X * d2),
              Jx5 = -fx * Y * d;
  const float Jy1 = fy * W * d, Jy2 = -fy * Y * W * d2, Jy3 = -fy * (1.0f + Y * Y * d2), Jy4 = fy * X * Y * d2,
              Jy5 = fy * X * d;
  const float Jzx = fx * (t[0] * d - t[2] * X * d2);
  const float Jzy = fy * (t[1] * d - t[2] * Y * d2);
  const float Jx[6] = {Jx0, 0.f, Jx2, Jx3, Jx4, Jx5};
  const float Jy[6] = {0.f, Jy1, Jy2, Jy3, Jy4, Jy5};
  const float wrx = wx * rx, wry = wy * ry;
  const float wzx = wx * Jzx, wzy = wy * Jzy;
#pragma unroll
  for (int a = 0; a < 6; ++a) {
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
struct LinSmem {
  float* sRt;      // [SMAX][12]  relative pose per slot (R row-major, t)
  float* sH;       // [SMAX][28]  H (21) + g (6) per slot, reduced over warps
  float* sPatch;   // [PMAX][4]   px, py, pd, -
  float* sC;       // [PMAX]
  float* sU;       // [PMAX]
  float* sQ;       // [PMAX]
  float* sEi;      // [PMAX][6]   source-frame column accumulators
  float* sBii;     // [36 + 6]    B_ii and v_i of the chunk
  float* sE;       // [EBUDGET]   E tile [patch][col][6]
  int* sFrame;     // [SMAX]
};
constexpr int LIN_SMEM_FLOATS = SMAX * 12 + SMAX * 28 + PMAX * 4 + PMAX * 3 + PMAX * 6 + 48 + EBUDGET + SMAX;
constexpr size_t LIN_SMEM_BYTES = sizeof(float) * LIN_SMEM_FLOATS;

__device__ __forceinline__ int pow2_ceil(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

// grid = (gx, batch), block = 256, dynamic smem = LIN_SMEM_BYTES
__global__ void __launch_bounds__(256, 2) linearize_kernel(Problem pb) {
  extern __shared__ float smem[];
  LinSmem s;
  s.sRt = smem;
  s.sH = s.sRt + SMAX * 12;
  s.sPatch = s.sH + SMAX * 28;
  s.sC = s.sPatch + PMAX * 4;
  s.sU = s.sC + PMAX;
  s.sQ = s.sU + PMAX;
  s.sEi = s.sQ + PMAX;
  s.sBii = s.sEi + PMAX * 6;
  s.sE = s.sBii + 48;
  s.sFrame = (int*)(s.sE + EBUDGET);

  const int w = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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

  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const Chunk ch = wp.chunks[c];
    if (ch.n_patches == 0) continue;
    __syncthreads();
    const int ns = ch.n_slots, ncols = ch.ncols, fi = ch.frame;
    const bool i_free = ch.icol >= 0;
    const int* cells = wp.cells + ch.cell_base;
    const int* kx = wp.kx + ch.patch_base;
    // ---- per-slot relative poses, zero the H accumulators
    for (int sl = tid; sl < ns; sl += 256) {
      const int fj = wp.slots[ch.slot_base + sl];
      s.sFrame[sl] = fj;
      float R[9], t[3];
      rel_pose(poses + 7 * (int64_t)fi, poses + 7 * (int64_t)fj, R, t);
#pragma unroll
      for (int x = 0; x < 9; ++x) s.sRt[sl * 12 + x] = R[x];
#pragma unroll
      for (int x = 0; x < 3; ++x) s.sRt[sl * 12 + 9 + x] = t[x];
    }
    for (int x = tid; x < ns * 28; x += 256) s.sH[x] = 0.f;
    if (tid < 48) s.sBii[tid] = 0.f;

    const int estride = ncols * 6;
    const int PB = (estride > 0) ? min(ch.n_patches, max(EBUDGET / estride, 1)) : ch.n_patches;
    for (int b0 = 0; b0 < ch.n_patches; b0 += PB) {
      const int b1 = min(b0 + PB, ch.n_patches);
      __syncthreads();
      // ---- stage patch centres, zero per-patch accumulators and the E tile
      for (int p = b0 + tid; p < b1; p += 256) {
        const float* pr = patches + (int64_t)kx[p] * pstride;
        const int q = p - b0;
        s.sPatch[q * 4 + 0] = pr[cidx];
        s.sPatch[q * 4 + 1] = pr[PP + cidx];
        s.sPatch[q * 4 + 2] = pr[2 * PP + cidx];
        s.sC[q] = 0.f;
        s.sU[q] = 0.f;
#pragma unroll
        for (int a = 0; a < 6; ++a) s.sEi[q * 6 + a] = 0.f;
      }
      for (int x = tid; x < (b1 - b0) * estride; x += 256) s.sE[x] = 0.f;
      __syncthreads();

      // ---- tile loop: lanes <-> slots (DW wide), PW patches per warp step
      for (int sb = 0; sb < ns; sb += 32) {
        const int ns_here = min(32, ns - sb);
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

        for (int p0 = b0 + warp * PW; p0 < b1; p0 += 8 * PW) {
          const int p = p0 + pl;
          const bool ok = slot_ok && p < b1;
          const int n = ok ? cells[p * ns + sl] : -1;
          float e[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, ck = 0.f, uk = 0.f;
          float ei[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (n >= 0) {
            const int q = p - b0;
            const float2 tg = __ldg(target + n), wt = __ldg(weight + n);
            edge_terms(s.sPatch[q * 4], s.sPatch[q * 4 + 1], s.sPatch[q * 4 + 2], fx, fy, cx, cy, R, t, tg, wt, H, g,
                       e, ck, uk);
            if (col_ok && schur) {
#pragma unroll
              for (int a = 0; a < 6; ++a) s.sE[q * estride + col * 6 + a] = e[a];
            }
            if (i_free) adj_map(R, t, e, ei);
          }
          // reduce the per-patch quantities over the DW lanes of this patch
          for (int o = DW >> 1; o > 0; o >>= 1) {
            ck += __shfl_xor_sync(0xffffffffu, ck, o);
            uk += __shfl_xor_sync(0xffffffffu, uk, o);
            if (i_free) {
#pragma unroll
              for (int a = 0; a < 6; ++a) ei[a] += __shfl_xor_sync(0xffffffffu, ei[a], o);
            }
          }
          if ((lane & (DW - 1)) == 0 && p < b1) {
            const int q = p - b0;
            atomicAdd(&s.sC[q], ck);
            atomicAdd(&s.sU[q], uk);
            if (i_free) {
#pragma unroll
              for (int a = 0; a < 6; ++a) atomicAdd(&s.sEi[q * 6 + a], -ei[a]);     // E_i -= w Jz Ji
            }
          }
        }
        if (slot_ok) {
#pragma unroll
          for (int x = 0; x < 21; ++x) atomicAdd(&s.sH[sl * 28 + x], H[x]);
#pragma unroll
          for (int x = 0; x < 6; ++x) atomicAdd(&s.sH[sl * 28 + 21 + x], g[x]);
        }
      }
      __syncthreads();

      // ---- duplicated (patch, slot) edges: rare slow path, shared-memory atomics
      if (n_dups > 0) {
        for (int dix = tid; dix < n_dups; dix += 256) {
          const DupEdge de = wp.dups[dix];
          if (de.chunk != c || de.p < b0 || de.p >= b1) continue;
          const int q = de.p - b0, sl = de.s;
          float R[9], t[3], H[21], g[6], e[6], ei[6], ck, uk;
#pragma unroll
          for (int x = 0; x < 9; ++x) R[x] = s.sRt[sl * 12 + x];
#pragma unroll
          for (int x = 0; x < 3; ++x) t[x] = s.sRt[sl * 12 + 9 + x];
#pragma unroll
          for (int x = 0; x < 21; ++x) H[x] = 0.f;
#pragma unroll
          for (int x = 0; x < 6; ++x) g[x] = 0.f;
          edge_terms(s.sPatch[q * 4], s.sPatch[q * 4 + 1], s.sPatch[q * 4 + 2], fx, fy, cx, cy, R, t,
                     __ldg(target + de.n), __ldg(weight + de.n), H, g, e, ck, uk);
          const int col = sl - ch.first_free;
          if (col >= 0 && col < ch.n_free && schur)
            for (int a = 0; a < 6; ++a) atomicAdd(&s.sE[q * estride + col * 6 + a], e[a]);
          if (i_free) {
            adj_map(R, t, e, ei);
            for (int a = 0; a < 6; ++a) atomicAdd(&s.sEi[q * 6 + a], -ei[a]);
          }
          atomicAdd(&s.sC[q], ck);
          atomicAdd(&s.sU[q], uk);
          for (int x = 0; x < 21; ++x) atomicAdd(&s.sH[sl * 28 + x], H[x]);
          for (int x = 0; x < 6; ++x) atomicAdd(&s.sH[sl * 28 + 21 + x], g[x]);
        }
        __syncthreads();
      }

      // ---- per patch: fold the source-frame column, Q = 1/(C + lambda), export Q, u, E
      for (int p = b0 + tid; p < b1; p += 256) {
        const int q = p - b0;
        const float Q = 1.0f / (s.sC[q] + lmbda);
        s.sQ[q] = Q;
        wp.Q[ch.patch_base + p] = Q;
        wp.u[ch.patch_base + p] = s.sU[q];
        if (i_free && schur) {
#pragma unroll
          for (int a = 0; a < 6; ++a) s.sE[q * estride + ch.icol * 6 + a] += s.sEi[q * 6 + a];
        }
      }
      __syncthreads();
      if (N > 0 && schur && ncols > 0) {
        float* eg = wp.ecells + 6 * ((int64_t)ch.ecell_base + (int64_t)b0 * ncols);
        for (int x = tid; x < (b1 - b0) * estride; x += 256) eg[x] = s.sE[x];

        // ---- Schur update of this batch: S[ca, cb] -= sum_p Q_p E_p[ca] E_p[cb]^T  (lower block triangle + mirror)
        const int npairs = ncols * (ncols + 1) / 2;
        for (int pr = tid; pr < npairs; pr += 256) {
          int ca = (int)((sqrtf(8.f * pr + 1.f) - 1.f) * 0.5f);
          while (ca * (ca + 1) / 2 > pr) --ca;
          while ((ca + 1) * (ca + 2) / 2 <= pr) ++ca;
          const int cb = pr - ca * (ca + 1) / 2;
          float acc[36];
#pragma unroll
          for (int x = 0; x < 36; ++x) acc[x] = 0.f;
          for (int q = 0; q < b1 - b0; ++q) {
            const float Q = s.sQ[q];
            const float* ea = s.sE + q * estride + ca * 6;
            const float* eb = s.sE + q * estride + cb * 6;
            float va[6], vb[6];
#pragma unroll
            for (int a = 0; a < 6; ++a) { va[a] = Q * ea[a]; vb[a] = eb[a]; }
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
              for (int b = 0; b < 6; ++b) acc[a * 6 + b] += va[a] * vb[b];
          }
          const int fa = ((ca < ch.n_free) ? s.sFrame[ch.first_free + ca] : fi) - t0;
          const int fb = ((cb < ch.n_free) ? s.sFrame[ch.first_free + cb] : fi) - t0;
#pragma unroll
          for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) {
              atomicAdd(&wp.S[(size_t)(6 * fa + a) * n6 + 6 * fb + b], -acc[a * 6 + b]);
              if (ca != cb) atomicAdd(&wp.S[(size_t)(6 * fb + b) * n6 + 6 * fa + a], -acc[a * 6 + b]);
            }
        }
        // y[ca] -= sum_p Q_p u_p E_p[ca]
        for (int x = tid; x < ncols * 6; x += 256) {
          const int ca = x / 6, a = x - ca * 6;
          float acc = 0.f;
          for (int q = 0; q < b1 - b0; ++q) acc += s.sQ[q] * s.sU[q] * s.sE[q * estride + x];
          const int fa = ((ca < ch.n_free) ? s.sFrame[ch.first_free + ca] : fi) - t0;
          atomicAdd(&wp.y[6 * fa + a], -acc);
        }
      }
    }
    __syncthreads();

    // ---- pose blocks of this chunk: one thread per target-frame slot
    if (N > 0) {
      for (int sl = tid; sl < ns; sl += 256) {
        const int fj = s.sFrame[sl];
        const bool j_free = (fj >= t0 && fj < pb.t1);
        if (!j_free && !i_free) continue;
        float Hm[36], g[6], R[9], t[3];
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int b = a; b < 6; ++b) {
            const float v = s.sH[sl * 28 + sym6(a, b)];
            Hm[a * 6 + b] = v;
            Hm[b * 6 + a] = v;
          }
#pragma unroll
        for (int a = 0; a < 6; ++a) g[a] = s.sH[sl * 28 + 21 + a];
#pragma unroll
        for (int x = 0; x < 9; ++x) R[x] = s.sRt[sl * 12 + x];
#pragma unroll
        for (int x = 0; x < 3; ++x) t[x] = s.sRt[sl * 12 + 9 + x];
        const int jo = 6 * (fj - t0), io = 6 * (fi - t0);
        if (j_free) {
#pragma unroll
          for (int a = 0; a < 6; ++a) {
#pragma unroll
            for (int b = 0; b < 6; ++b) atomicAdd(&wp.S[(size_t)(jo + a) * n6 + jo + b], Hm[a * 6 + b]);
            atomicAdd(&wp.y[jo + a], g[a]);                               // v_j += w r Jj
          }
        }
        if (i_free) {
          // AH = A * H (A applied to every column of H); B_ij -= AH ; B_ii += AH A^T ; v_i -= A g
          float AH[36];
#pragma unroll
          for (int b = 0; b < 6; ++b) {
            float colv[6], outv[6];
#pragma unroll
            for (int a = 0; a < 6; ++a) colv[a] = Hm[a * 6 + b];
            adj_map(R, t, colv, outv);
#pragma unroll
            for (int a = 0; a < 6; ++a) AH[a * 6 + b] = outv[a];
          }
          if (j_free) {
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
              for (int b = 0; b < 6; ++b) {
                atomicAdd(&wp.S[(size_t)(io + a) * n6 + jo + b], -AH[a * 6 + b]);
                atomicAdd(&wp.S[(size_t)(jo + b) * n6 + io + a], -AH[a * 6 + b]);
              }
          }
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            float rowv[6], outv[6];
#pragma unroll
            for (int b = 0; b < 6; ++b) rowv[b] = AH[a * 6 + b];
            adj_map(R, t, rowv, outv);                                    // (AH A^T)[a][:] = A * (AH[a][:])^T
#pragma unroll
            for (int b = 0; b < 6; ++b) atomicAdd(&s.sBii[a * 6 + b], outv[b]);
          }
          float Ag[6];
          adj_map(R, t, g, Ag);
#pragma unroll
          for (int a = 0; a < 6; ++a) atomicAdd(&s.sBii[36 + a], -Ag[a]);
        }
      }
      __syncthreads();
      if (i_free && tid < 42) {
        const int io = 6 * (fi - t0);
        if (tid < 36) atomicAdd(&wp.S[(size_t)(io + tid / 6) * n6 + io + tid % 6], s.sBii[tid]);
        else atomicAdd(&wp.y[io + tid - 36], s.sBii[tid]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Small dense solve: one CTA per window, fp64 in shared memory.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SOLVE_NMAX = 156;   // 6N <= 156 (N <= 26): 156*157*8 B = 196 KB of shared memory

__global__ void __launch_bounds__(256, 1) solve_small_kernel(Problem pb) {
  extern __shared__ double sd[];
  const int w = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  const int n = 6 * (pb.t1 - pb.t0), ld = n + 1;
  double* A = sd;            // [n][ld]
  double* b = sd + n * ld;   // [n]
  for (int x = tid; x < n * n; x += 256) {
    const int r = x / n, c = x - r * n;
    double v = (double)wp.S[x];
    if (r == c) v += 1e-4 * v + 1.0;                 // S += I * (1e-4 * S + 1)   (ba_cuda.cu:575/589)
    A[r * ld + c] = v;
  }
  for (int x = tid; x < n; x += 256) b[x] = (double)wp.y[x];
  __syncthreads();
  // right-looking Cholesky, lower triangle
  for (int k = 0; k < n; ++k) {
    const double dkk = sqrt(A[k * ld + k]);          // NaN for an indefinite matrix, like the reference (info ignored)
    __syncthreads();
    if (tid == 0) A[k * ld + k] = dkk;
    const double inv = 1.0 / dkk;
    for (int i = k + 1 + tid; i < n; i += 256) A[i * ld + k] *= inv;
    __syncthreads();
    const int m = n - k - 1;
    for (int x = tid; x < m * m; x += 256) {
      const int i = k + 1 + x / m, j = k + 1 + x % m;
      if (j <= i) A[i * ld + j] -= A[i * ld + k] * A[j * ld + k];
    }
    __syncthreads();
  }
  // forward / backward substitution by warp 0
  if (tid < 32) {
    for (int i = 0; i < n; ++i) {
      double sacc = 0.0;
      for (int j = lane; j < i; j += 32) sacc += A[i * ld + j] * b[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
      if (lane == 0) b[i] = (b[i] - sacc) / A[i * ld + i];
      __syncwarp();
    }
    for (int i = n - 1; i >= 0; --i) {
      double sacc = 0.0;
      for (int j = i + 1 + lane; j < n; j += 32) sacc += A[j * ld + i] * b[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
      if (lane == 0) b[i] = (b[i] - sacc) / A[i * ld + i];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int x = tid; x < n; x += 256) wp.dX[x] = (float)b[x];
}

// ---------------------------------------------------------------------------------------------------------------
// SE3 retraction of the free poses: poses[t] <- Exp(dX[t - t0]) * poses[t]   (ba_cuda.cu:88-206)
// ---------------------------------------------------------------------------------------------------------------
__global__ void pose_retr_kernel(Problem pb) {
  const int w = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = pb.t1 - pb.t0;
  if (i >= N) return;
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  float* P = pb.poses + (int64_t)w * pb.st.poses + 7 * (int64_t)(pb.t0 + i);
  const float* xi = wp.dX + 6 * i;
  const float tau[3] = {xi[0], xi[1], xi[2]}, phi[3] = {xi[3], xi[4], xi[5]};
  // expSO3 (ba_cuda.cu:88-110)
  const float theta_sq = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  const float theta_p4 = theta_sq * theta_sq;
  const float theta = sqrtf(theta_sq);
  float imag, real;
  if (theta_sq < 1e-8f) {
    imag = 0.5f - (1.0f / 48.0f) * theta_sq + (1.0f / 3840.0f) * theta_p4;
    real = 1.0f - (1.0f / 8.0f) * theta_sq + (1.0f / 384.0f) * theta_p4;
  } else {
    imag = sinf(0.5f * theta) / theta;
    real = cosf(0.5f * theta);
  }
  const float dq[4] = {imag * phi[0], imag * phi[1], imag * phi[2], real};
  // expSE3 translation part (ba_cuda.cu:125-153)
  float dt[3] = {tau[0], tau[1], tau[2]};
  if (theta > 1e-4f) {
    const float a = (1.0f - cosf(theta)) / theta_sq;
    const float c1[3] = {phi[1] * tau[2] - phi[2] * tau[1], phi[2] * tau[0] - phi[0] * tau[2],
                         phi[0] * tau[1] - phi[1] * tau[0]};
    const float b = (theta - sinf(theta)) / (theta * theta_sq);
    const float c2[3] = {phi[1] * c1[2] - phi[2] * c1[1], phi[2] * c1[0] - phi[0] * c1[2],
                         phi[0] * c1[1] - phi[1] * c1[0]};
#pragma unroll
    for (int x = 0; x < 3; ++x) dt[x] += a * c1[x] + b * c2[x];
  }
  // retrSE3 (ba_cuda.cu:156-174): no re-normalisation of the quaternion
  const float t[3] = {P[0], P[1], P[2]}, q[4] = {P[3], P[4], P[5], P[6]};
  float q1[4], t1[3];
  q1[0] = dq[3] * q[0] + dq[0] * q[3] + dq[1] * q[2] - dq[2] * q[1];
  q1[1] = dq[3] * q[1] + dq[1] * q[3] + dq[2] * q[0] - dq[0] * q[2];
  q1[2] = dq[3] * q[2] + dq[2] * q[3] + dq[0] * q[1] - dq[1] * q[0];
  q1[3] = dq[3] * q[3] - dq[0] * q[0] - dq[1] * q[1] - dq[2] * q[2];
  rot_q(dq, t, t1);
  P[0] = t1[0] + dt[0]; P[1] = t1[1] + dt[1]; P[2] = t1[2] + dt[2];
  P[3] = q1[0]; P[4] = q1[1]; P[5] = q1[2]; P[6] = q1[3];
}

// ---------------------------------------------------------------------------------------------------------------
// Back-substitution dZ = Q (u - E^T dX) and inverse-depth retraction (ba_cuda.cu:592, 209-229; block_e.cu:253-283)
// grid = (gx, batch), block = 128.  apply = 0 only computes dZ (debug export).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) update_kernel(Problem pb, int apply) {
  __shared__ float sdx[(SMAX + 1) * 6];
  const int w = blockIdx.y, tid = threadIdx.x;
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  float* patches = pb.patches + (int64_t)w * pb.st.patches;
  const int N = pb.t1 - pb.t0, t0 = pb.t0;
  const int PP = pb.P * pb.P, pstride = 3 * PP;
  const int n_chunks = wp.hdr->n_chunks;
  const bool schur = pb.with_schur != 0;
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const Chunk ch = wp.chunks[c];
    if (ch.n_patches == 0) continue;
    __syncthreads();
    const int ncols = (N > 0 && schur) ? ch.ncols : 0;
    for (int x = tid; x < ncols * 6; x += 128) {
      const int col = x / 6, a = x - col * 6;
      const int f = (col < ch.n_free) ? wp.slots[ch.slot_base + ch.first_free + col] : ch.frame;
      sdx[x] = wp.dX[6 * (f - t0) + a];
    }
    __syncthreads();
    for (int p = tid; p < ch.n_patches; p += 128) {
      const float Q = wp.Q[ch.patch_base + p];
      float acc = wp.u[ch.patch_base + p];
      const float* eg = wp.ecells + 6 * ((int64_t)ch.ecell_base + (int64_t)p * ch.ncols);
      for (int x = 0; x < ncols * 6; ++x) acc -= eg[x] * sdx[x];
      const float dz = Q * acc;
      wp.dZ[ch.patch_base + p] = dz;
      if (apply) {
        float* pr = patches + (int64_t)wp.kx[ch.patch_base + p] * pstride + 2 * PP;
        float d = pr[0] + dz;                   // reads [2][0][0] (ba_cuda.cu:218)
        d = (d > 20.f) ? 1.0f : d;
        d = fmaxf(d, 1e-4f);
        for (int x = 0; x < PP; ++x) pr[x] = d;
      }
    }
  }
}

__global__ void zero_kernel(Problem pb) {
  const int w = blockIdx.y;
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  const size_t n6 = (size_t)6 * (pb.t1 - pb.t0);
  const size_t total = n6 * n6;
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x)
    wp.S[x] = 0.f;
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < n6; x += (size_t)gridDim.x * blockDim.x)
    wp.y[x] = 0.f;
}

// ---------------------------------------------------------------------------------------------------------------
// Host-side launch sequence of one Gauss-Newton iteration
// ---------------------------------------------------------------------------------------------------------------
static int chunk_grid(const Problem& pb, int64_t batch) {
  int64_t g = pb.L.ch_max;
  const int64_t cap = batch > 1 ? 48 : 148 * 2;
  return (int)(g < cap ? g : cap);
}

// ev (optional): 6 events recorded before zero / linearize / solve / pose_retr / update and after update
cudaError_t launch_iteration(const Problem& pb, int64_t batch, bool apply, cudaStream_t stream, cudaEvent_t* ev) {
  const int N = pb.t1 - pb.t0;
  const int gx = chunk_grid(pb, batch);
  if (ev) cudaEventRecord(ev[0], stream);
  if (N > 0) {
    const size_t n6 = (size_t)6 * N;
    int zb = (int)((n6 * n6 + 255) / 256);
    if (zb > 148 * 8) zb = 148 * 8;
    if (zb < 1) zb = 1;
    zero_kernel<<<dim3((unsigned)zb, (unsigned)batch), 256, 0, stream>>>(pb);
    count_launch();
  }
  if (ev) cudaEventRecord(ev[1], stream);
  cudaFuncSetAttribute(linearize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIN_SMEM_BYTES);
  linearize_kernel<<<dim3((unsigned)gx, (unsigned)batch), 256, LIN_SMEM_BYTES, stream>>>(pb);
  count_launch();
  if (ev) cudaEventRecord(ev[2], stream);
  if (N > 0) {
    const int n = 6 * N;
    const size_t smem = sizeof(double) * ((size_t)n * (n + 1) + n);
    cudaFuncSetAttribute(solve_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    solve_small_kernel<<<(unsigned)batch, 256, smem, stream>>>(pb);
    count_launch();
    if (ev) cudaEventRecord(ev[3], stream);
    if (apply) {
      pose_retr_kernel<<<dim3((unsigned)((N + 63) / 64), (unsigned)batch), 64, 0, stream>>>(pb);
      count_launch();
    }
  } else if (ev) {
    cudaEventRecord(ev[3], stream);
  }
  if (ev) cudaEventRecord(ev[4], stream);
  update_kernel<<<dim3((unsigned)gx, (unsigned)batch), 128, 0, stream>>>(pb, apply ? 1 : 0);
  count_launch();
  if (ev) cudaEventRecord(ev[5], stream);
  return cudaGetLastError();
}

bool solve_small_supported(int N) { return 6 * N <= SOLVE_NMAX; }

}  // namespace pgba
