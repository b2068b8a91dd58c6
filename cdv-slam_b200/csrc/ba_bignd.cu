// Large solve of the global (loop-closure) bundle adjustment with a nested-dissection frame ordering.
// Replaces at::linalg_cholesky_ex + torch::cholesky_solve on the dense S of the reference
// (cdvslam/fastba/ba_cuda.cu:575-578 for eff_impl, :589-591 dense) for every system beyond the single-CTA solve (>= ND_MIN_N = 27 free poses).
//
// The natural-order blocked Cholesky of big_chol.cuh walks 6N / 48 dependent panel steps (125 for the 1000-frame global
// BA), three launches each, and nearly every step is latency: the pose graph is a chain (patches are seen by a few
// neighbouring frames) plus loop closures, so a panel has a handful of non-zero tiles.  Here the free frames are reordered
//
//     [ chain segment 0 | chain segment 1 | ... | chain segment P-1 | border ]
//
// where the border holds (a) every frame that is the target of a FAR edge (|j - i| > R: the loop-closure targets) and
// (b) for each of the P - 1 cut positions, the frames at or after the cut that share a patch with a frame before it (the
// separator; its width follows from the data).  No patch couples frames of two different segments, so the segments are
// eliminated IN PARALLEL, level by level (tile l of every segment in the same two launches), then the dense border:
// max_p T_p + Bt dependent steps instead of sum_p T_p + Bt (36 instead of 125 on the 1000-frame problem).  The ordering
// is computed on the device from the edge list (no host round trip); whatever the graph looks like the result is the
// same Cholesky solve of the same matrix under a symmetric permutation -- a graph that is not chain-like only makes the
// border large.  Segments and border are padded to whole tiles with identity rows.
//
//   nd_stats_kernel    per source frame: extent of its near edges; far edge targets -> border
//   nd_order_kernel    separators, segment / border positions of every free frame (one CTA per window)
//   nd_gather_kernel   S, y (natural order, written by linearize_kernel) -> permuted Sp, yp with the damping of
//                      ba_cuda.cu:575/589 applied; S, y are re-zeroed on the way
//   a panel step       nd_trsm_kernel: the row tiles below the diagonal tile, X L^T = A by blocked substitution with the
//                      inverses of the 6 x 6 diagonal blocks (all-zero tiles are skipped and stay inactive for the panel), the
//                      right-hand side as one more row, and W = L^-T of the diagonal tile for the backward substitution;
//                      nd_syrk_kernel: trailing update over pairs of active tiles (during the segment levels the border x
//                      border updates of different segments meet in the same tiles -> atomic adds there); its CTA 0 is the
//                      look-ahead: it updates the NEXT diagonal tile in shared memory and factors it (fp64), so the serial
//                      piece of a step never waits for a launch of its own.  nd_potf2_kernel only factors the first tile
//                      of a phase.
//   nd_border_kernel   the border's step count is only known on the device; as separate launches the worst case would have to
//                      be enqueued.  The first nd_nt / 4 border panels are launches (cheapest per panel), the rest ONE
//                      cooperative launch: the same step bodies with the panel loop on the device and grid-wide barriers;
//                      it returns at once when the launches covered the border.  All launches for batches of more than 4
//                      windows / PGBA_ND_COOP=0.
//   nd_backsolve_kernel  L^T x = z: border (one CTA), then the segments in parallel
//   nd_finish_kernel   x -> dX in frame order, pose retraction (ba_cuda.cu:88-206)
//   nd_cleanup_kernel  zeroes the tiles the factor touched (before the next iteration's gather)
// -DPGBA_ND_TIMING adds globaltimer stamps per launch (profiles/nd_timeline.py).
#include "big_chol.cuh"
#include <cooperative_groups.h>

#include "ba_cells.cuh"

namespace pgba {

#ifdef PGBA_ND_TIMING      // per-launch timeline of the factorisation (profiles/nd_timeline.py): entry, after pdl_wait, end
__device__ unsigned long long g_nd_ts[8 * 1024][3];
__device__ __forceinline__ unsigned long long nd_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
struct NdTs {
  int id;
  __device__ NdTs(int kind, int mode, int idx) : id(kind * 1024 + mode * 512 + idx) { if (threadIdx.x == 0) atomicMin(&g_nd_ts[id][0], nd_gtime()); }
  __device__ void waited() { if (threadIdx.x == 0) atomicMin(&g_nd_ts[id][1], nd_gtime()); }
  __device__ ~NdTs() { if (threadIdx.x == 0) atomicMax(&g_nd_ts[id][2], nd_gtime()); }
};
#define ND_TS(kind, mode, idx) NdTs nd_ts_(kind, mode, idx)
#define ND_TS_WAITED() nd_ts_.waited()
#else
#define ND_TS(kind, mode, idx) do { } while (0)
#define ND_TS_WAITED() do { } while (0)
#endif

struct NdSys {
  float* S; float* y; float* Sp; float* yp;
  int ld;                 // row stride of Sp (= capacity in unknowns)
  NdHeader* h;
  int* lminv; int* lmax1; int* border; int* pos; int* pfb; int* pfn; int* frame_at;
  float* winv; float* dinv; int* active; int* nact; int act_stride; int* chol_info;
};

__device__ __forceinline__ NdSys nd_sys(const Problem& pb, int w) {
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  char* base = (char*)pb.ws + pb.L.body0 + (size_t)w * pb.L.body_bytes;
  char* z = (char*)pb.ws + (size_t)w * pb.L.zero_bytes;
  NdSys s;
  s.S = wp.S; s.y = wp.y;
  s.Sp = (float*)(z + pb.L.z_Sp); s.yp = (float*)(z + pb.L.z_yp);
  s.ld = pb.L.nd_nt * NB;
  s.h = (NdHeader*)(z + pb.L.z_nd);
  int* f = (int*)(z + pb.L.z_ndf);
  s.lminv = f; s.lmax1 = f + pb.F; s.border = f + 2 * pb.F; s.pos = f + 3 * pb.F; s.pfb = f + 4 * pb.F; s.pfn = f + 5 * pb.F;
  s.frame_at = (int*)(base + pb.L.o_frame_at);
  s.winv = (float*)(base + pb.L.o_winv);
  s.dinv = (float*)(base + pb.L.o_dinv);
  s.active = (int*)(base + pb.L.o_active);
  s.nact = (int*)(z + pb.L.z_nact);
  s.act_stride = pb.L.big_tiles;
  s.chol_info = &wp.hdr->chol_info;
  return s;
}

// ---------------------------------------------------------------------------------------------------------------
// ordering
// ---------------------------------------------------------------------------------------------------------------
// grid = (gx, batch), block = 256
__global__ void __launch_bounds__(256) nd_stats_kernel(Problem pb) {
  pdl_wait();
  pdl_trigger();
  const int w = blockIdx.y + pb.w0, tid = threadIdx.x, lane = tid & 31;
  const NdSys sys = nd_sys(pb, w);
  const int64_t* ii = pb.ii + (int64_t)w * pb.st.ii;
  const int64_t* jj = pb.jj + (int64_t)w * pb.st.jj;
  const int64_t* kk = pb.kk + (int64_t)w * pb.st.kk;
  const int E = pb.n_edges_dev ? min((int)pb.E, max(pb.n_edges_dev[w], 0)) : (int)pb.E;
  const int R = pb.L.nd_R;
  const int stride = gridDim.x * blockDim.x;
  for (int e0 = blockIdx.x * blockDim.x; e0 < E; e0 += stride) {            // warp-uniform trip count
    const int e = e0 + tid;
    int i = -1, j = 0;
    if (e < E) {
      int64_t a, b, c;
      if (pb.idx32) {
        a = reinterpret_cast<const int32_t*>(ii)[e]; b = reinterpret_cast<const int32_t*>(jj)[e]; c = reinterpret_cast<const int32_t*>(kk)[e];
      } else {
        a = ii[e]; b = jj[e]; c = kk[e];
      }
      if (!(a < 0 || a >= pb.F || b < 0 || b >= pb.F || c < 0 || c >= pb.K)) { i = (int)a; j = (int)b; }   // as the plan
    }
    const bool ok = i >= 0;
    const bool far = ok && (j - i > R || i - j > R);
    if (far) sys.border[j] = 1;
    // near edges: extent per source frame, one atomic pair per run of equal source frames in the warp
    const int key = (ok && !far) ? i : -1;
    const unsigned grp = __match_any_sync(0xffffffffu, key);
    const int jmn = __reduce_min_sync(grp, j), jmx = __reduce_max_sync(grp, j);
    if (key >= 0 && lane == __ffs(grp) - 1) {
      atomicMax(&sys.lminv[i], 0x7fffffff - jmn);          // zero-initialised "min": stores max of (INT_MAX - j)
      atomicMax(&sys.lmax1[i], jmx + 1);
    }
  }
}

// grid = (1, batch), block = 1024
__global__ void __launch_bounds__(1024) nd_order_kernel(Problem pb) {
  pdl_wait();
  pdl_trigger();
  __shared__ int s_cut[ND_MAXP + 1], s_r[ND_MAXP + 1], s_T[ND_MAXP], s_base[ND_MAXP + 1], scratch[40];
  const int w = blockIdx.y + pb.w0, tid = threadIdx.x;
  const NdSys sys = nd_sys(pb, w);
  const int P = pb.L.nd_P, t0 = pb.t0, t1 = pb.t1, N = t1 - t0, F = pb.F;
  if (tid <= P) {
    s_cut[tid] = t0 + (int)(((int64_t)tid * N) / P);       // segment p = frames [cut[p], cut[p + 1]) minus the border
    s_r[tid] = s_cut[tid] - 1;                             // separator of cut p = [cut[p], r[p]] (empty so far)
  }
  __syncthreads();
  // a source frame whose near edges reach from before a cut to (or beyond) it: everything up to its far end is coupled
  // with frames before the cut
  for (int f = tid; f < F; f += 1024) {
    int lo = f, hi = f;
    const int a = __ldcg(&sys.lminv[f]), b = __ldcg(&sys.lmax1[f]);
    if (a) lo = min(lo, 0x7fffffff - a);
    if (b) hi = max(hi, b - 1);
    if (hi > lo)
      for (int p = 1; p < P; ++p) {
        const int c = s_cut[p];
        if (lo < c && c <= hi) atomicMax(&s_r[p], hi);
      }
  }
  __syncthreads();
  for (int x = tid; x < N; x += 1024) {
    const int f = t0 + x;
    int bd = __ldcg(&sys.border[f]) != 0;
    for (int p = 1; p < P && !bd; ++p) bd = (f >= s_cut[p] && f <= s_r[p]);
    sys.border[f] = bd;
    sys.pfb[x] = bd;
    sys.pfn[x] = !bd;
  }
  __syncthreads();
  const int nb = block_exclusive_scan(sys.pfb, N, scratch);
  __syncthreads();
  const int nn = block_exclusive_scan(sys.pfn, N, scratch);
  __syncthreads();
  if (tid < P) {
    const int hi = s_cut[tid + 1] - t0, lo = s_cut[tid] - t0;
    const int cnt = (hi >= N ? nn : sys.pfn[hi]) - sys.pfn[lo];
    s_T[tid] = (cnt + 7) / 8;
  }
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int p = 0; p < P; ++p) { s_base[p] = run; sys.h->T[p] = s_T[p]; sys.h->segbase[p] = run; run += s_T[p]; }
    s_base[P] = run;
    const int Bt = (nb + 7) / 8;
    sys.h->bbase = run; sys.h->Bt = Bt; sys.h->nt = run + Bt; sys.h->n_border = nb;
  }
  for (int x = tid; x < pb.L.nd_nt * 8; x += 1024) sys.frame_at[x] = -1;
  __syncthreads();
  for (int x = tid; x < N; x += 1024) {
    const int f = t0 + x;
    int ps;
    if (sys.border[f]) {
      ps = s_base[P] * 8 + sys.pfb[x];
    } else {
      int seg = 0;
      for (int p = 1; p < P; ++p) if (f >= s_cut[p]) seg = p;
      ps = s_base[seg] * 8 + (sys.pfn[x] - sys.pfn[s_cut[seg] - t0]);
    }
    sys.pos[f] = ps;
    sys.frame_at[ps] = f;
  }
}

// grid = (gx, batch), block = 256.  One row of the lower block triangle of S per CTA and trip.
__global__ void __launch_bounds__(256) nd_gather_kernel(Problem pb) {
  ND_TS(3, 0, 0);
  pdl_wait();
  pdl_trigger();
  ND_TS_WAITED();
  const int w = blockIdx.y + pb.w0, tid = threadIdx.x;
  const NdSys sys = nd_sys(pb, w);
  const int t0 = pb.t0, N = pb.t1 - t0, n6 = 6 * N;
  const size_t ld = (size_t)sys.ld;
  const int* __restrict__ pos = sys.pos + t0;
  for (int r = blockIdx.x; r < n6; r += gridDim.x) {
    const int fa = r / 6, ra = r - 6 * fa;
    const int pa = pos[fa];
    float2* row = reinterpret_cast<float2*>(sys.S + (size_t)r * n6);
    const int ncol2 = 3 * (fa + 1);
    for (int c2 = tid; c2 < ncol2; c2 += 256) {
      float2 v = row[c2];
      const int c = 2 * c2, fb = c / 6, cb = c - 6 * fb;
      const bool diag = fb == fa;
      if (!diag && v.x == 0.f && v.y == 0.f) continue;
      row[c2] = make_float2(0.f, 0.f);
      if (diag) {
        if (cb == ra) v.x = v.x + (1e-4f * v.x + 1.0f);           // S += I * (1e-4 * S + 1)   (ba_cuda.cu:575/589)
        else if (cb + 1 == ra) v.y = v.y + (1e-4f * v.y + 1.0f);
        *reinterpret_cast<float2*>(sys.Sp + (size_t)(6 * pa + ra) * ld + 6 * pa + cb) = v;
      } else {
        const int pbp = pos[fb];
        if (pa > pbp) {
          *reinterpret_cast<float2*>(sys.Sp + (size_t)(6 * pa + ra) * ld + 6 * pbp + cb) = v;
        } else {                                                  // the block lands above the diagonal: store its transpose
          sys.Sp[(size_t)(6 * pbp + cb) * ld + 6 * pa + ra] = v.x;
          sys.Sp[(size_t)(6 * pbp + cb + 1) * ld + 6 * pa + ra] = v.y;
        }
      }
    }
    if (tid == 0) { sys.yp[6 * pa + ra] = sys.y[r]; sys.y[r] = 0.f; }
  }
  // identity on the padding rows
  const int npos = sys.h->nt * 8;
  for (int x = blockIdx.x * 256 + tid; x < npos; x += gridDim.x * 256)
    if (sys.frame_at[x] < 0) {
#pragma unroll
      for (int a = 0; a < 6; ++a) sys.Sp[(size_t)(6 * x + a) * ld + 6 * x + a] = 1.0f;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// factorisation.  A panel step is (mode, idx): mode 0 = tile idx of segment blockIdx.x, mode 1 = tile idx of the border.
// Rows below panel tile k that can be non-zero: the later tiles of the same segment / of the border, then (segments
// only) all border tiles.
// ---------------------------------------------------------------------------------------------------------------
struct NdPanel { int k, nsb, nbelow, bbase; };

__device__ __forceinline__ bool nd_panel(const NdHeader* h, int mode, int idx, int p, NdPanel& pn) {
  pn.bbase = h->bbase;
  if (mode == 0) {
    const int T = h->T[p];
    if (idx >= T) return false;
    pn.k = h->segbase[p] + idx; pn.nsb = T - 1 - idx; pn.nbelow = pn.nsb + h->Bt;
  } else {
    const int Bt = h->Bt;
    if (idx >= Bt) return false;
    pn.k = pn.bbase + idx; pn.nsb = Bt - 1 - idx; pn.nbelow = pn.nsb;
  }
  return true;
}
__device__ __forceinline__ int nd_tile_of(const NdPanel& pn, int c) { return c < pn.nsb ? pn.k + 1 + c : pn.bbase + (c - pn.nsb); }

constexpr int ND_DINV = (NB / 6) * 36;    // floats of the 6 x 6 diagonal-block inverses of one tile

// Diagonal tile k: fp64 shared-memory Cholesky (ba_chol.cuh) and the inverses of the 6 x 6 diagonal blocks of L (what the
// row-tile solves substitute with); resets the panel's active list.  This is the serial piece of every panel step, so it
// carries nothing else: the explicit inverse W = L^-T of the whole tile, which only the backward substitution needs, is
// formed by one extra CTA of the row-tile solve (nd_trsm_kernel) -- with the inverse riding along as 48 extra rows the
// factorisation took 15 us instead of ~9.
// All 256 threads; sd: NB x (NB | 1) doubles + NB
__device__ __forceinline__ void nd_potf2_core(const NdSys& sys, int k, double* sd);

__device__ __forceinline__ void nd_potf2_dev(const NdSys& sys, int k, double* sd) {
  const int tid = threadIdx.x, kb = k * NB, ld = NB | 1;
  double* A = sd;
  const float* Sd = sys.Sp + (size_t)kb * sys.ld + kb;
  __syncthreads();
  {
    float v[NB * NB / 256];                     // all loads of the tile in flight before the first shared-memory store
#pragma unroll
    for (int i = 0; i < NB * NB / 256; ++i) {
      const int x = tid + 256 * i, r = x / NB, c = x - r * NB;
      v[i] = (c <= r) ? __ldcg(&Sd[(size_t)r * sys.ld + c]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NB * NB / 256; ++i) {
      const int x = tid + 256 * i, r = x / NB, c = x - r * NB;
      A[r * ld + c] = (double)v[i];
    }
  }
  nd_potf2_core(sys, k, sd);
}

// sd holds the lower triangle of diagonal tile k (zeros above) as doubles, row stride NB | 1
__device__ __forceinline__ void nd_potf2_core(const NdSys& sys, int k, double* sd) {
  const int tid = threadIdx.x, kb = k * NB, ld = NB | 1;
  double* A = sd;
  double* rd = sd + NB * ld;
  float* Sd = sys.Sp + (size_t)kb * sys.ld + kb;
  chol6_smem(A, rd, NB, NB - 1, ld);
  for (int x = tid; x < NB * NB; x += 256) {
    const int r = x / NB, c = x - r * NB;
    if (c <= r) Sd[(size_t)r * sys.ld + c] = (float)A[r * ld + c];
  }
  if (tid < NB) {                         // column c of the inverse of diagonal block b: forward substitution L_bb x = e_c
    const int b = tid / 6, c = tid - 6 * b, k6 = 6 * b;
    double x[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double s = (r == c) ? 1.0 : 0.0;
#pragma unroll
      for (int e = 0; e < r; ++e) s -= A[(k6 + r) * ld + k6 + e] * x[e];
      x[r] = s * rd[k6 + r];
    }
    float* D = sys.dinv + (size_t)k * ND_DINV + b * 36;
#pragma unroll
    for (int r = 0; r < 6; ++r) D[r * 6 + c] = (float)x[r];
    if (!(rd[tid] > 0.0) || !isfinite(rd[tid])) atomicCAS(sys.chol_info, 0, kb + tid + 1);
  }
  if (tid == 0) sys.nact[k] = 0;
}

// The first panel of a phase (the later ones are factored by the look-ahead CTA of the previous panel's trailing update).
// grid = (P | 1, batch), block = 256, dynamic smem: NB x (NB | 1) doubles + NB
__global__ void __launch_bounds__(256, 1) nd_potf2_kernel(Problem pb, int mode, int idx) {
  ND_TS(0, mode, idx);
  extern __shared__ double sd[];
  const NdSys sys = nd_sys(pb, blockIdx.y + pb.w0);
  NdPanel pn;
  const bool live = nd_panel(sys.h, mode, idx, blockIdx.x, pn);      // the ordering is final: read ahead of the wait
  pdl_wait();
  pdl_trigger();
  ND_TS_WAITED();
  if (!live) return;
  nd_potf2_dev(sys, pn.k, sd);
}

// In place sA <- sA L^-T (X L^T = A) for `rows` rows, block column by block column: X_C = (A_C - sum_{E < C} X_E L_CE^T) D_C^T
// with D_C the inverse of the diagonal block.  A row only ever reads its own earlier columns, so four adjacent lanes share a
// row (the inner sums split four ways, two shuffles to combine) and nothing but a __syncwarp separates the block steps.
// Threads 0 .. 4 * rows - 1 work (rows <= 48: warps 0..5); sL: the factor tile (lower triangle), sD: ND_DINV floats.
__device__ __forceinline__ void nd_solve_rows(float (*sA)[NB + 1], const float (*sL)[NB + 1], const float* sD, int rows) {
  const int tid = threadIdx.x, r = tid >> 2, part = tid & 3;
  if (r >= rows) return;
  const unsigned mask = rows >= 8 ? 0xffffffffu : ((1u << (4 * rows)) - 1u);      // rows is NB or 1
  for (int C = 0; C < NB / 6; ++C) {
    float t[6], t2[6], a6[6];
#pragma unroll
    for (int cp = 0; cp < 6; ++cp) { t[cp] = 0.f; t2[cp] = 0.f; a6[cp] = sA[r][6 * C + cp]; }
    // two columns per trip (e, e + 4), all 14 shared-memory loads ahead of the FMAs; a second column beyond 6 C reads
    // entries of the diagonal block and is multiplied by zero
    for (int e = part; e < 6 * C; e += 8) {
      const float x0 = sA[r][e], x1 = (e + 4 < 6 * C) ? sA[r][e + 4] : 0.f;
      float l0[6], l1[6];
#pragma unroll
      for (int cp = 0; cp < 6; ++cp) { l0[cp] = sL[6 * C + cp][e]; l1[cp] = sL[6 * C + cp][e + 4]; }
#pragma unroll
      for (int cp = 0; cp < 6; ++cp) { t[cp] = fmaf(-x0, l0[cp], t[cp]); t2[cp] = fmaf(-x1, l1[cp], t2[cp]); }
    }
#pragma unroll
    for (int cp = 0; cp < 6; ++cp) {
      t[cp] += t2[cp];
      t[cp] += __shfl_xor_sync(mask, t[cp], 1);
      t[cp] += __shfl_xor_sync(mask, t[cp], 2);
      t[cp] += a6[cp];
    }
    const float* D = sD + C * 36;
    // X[c] = sum_{cp <= c} t[cp] D[c][cp]: lane `part` takes c = part and c = part + 4
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = part + 4 * h;
      if (c < 6) {
        float x = 0.f;
#pragma unroll
        for (int cp = 0; cp < 6; ++cp) x = fmaf(t[cp], D[c * 6 + cp], x);      // D is lower triangular (zeros above)
        sA[r][6 * C + c] = x;
      }
    }
    __syncwarp(mask);
  }
}

// Row-tile solves of a panel.  grid = (P | 1, batch, nd_nt + 2) -- segment fastest, so the candidates beyond a segment's
// count (the grid is sized for the worst case) are dispatched last --, block = 256: blockIdx.z < nbelow: candidate row tile;
// == nbelow: the right-hand side; == nbelow + 1: the explicit inverse W = L^-T of the diagonal tile (the same solve applied
// to the identity), which the backward substitution uses later -- off the critical path of the factorisation.
// One item of the row-tile launch: c0 < nbelow: candidate row tile; == nbelow: the right-hand side; == nbelow + 1: the
// explicit inverse W = L^-T of the diagonal tile.  All 256 threads; starts with a barrier (the shared tiles may be in use).
__device__ __forceinline__ void nd_trsm_item(const NdSys& sys, const NdPanel& pn, int c0, float (*sA)[NB + 1],
                                             float (*sL)[NB + 1], float* sD) {
  __syncthreads();
  const int tid = threadIdx.x, kb = pn.k * NB;
  const bool rhs = (c0 == pn.nbelow), inv = (c0 == pn.nbelow + 1);
  const int t = (rhs || inv) ? -1 : nd_tile_of(pn, c0);
  const int rows = rhs ? 1 : NB;
  float* src = inv ? (sys.winv + (size_t)pn.k * NB * NB) : rhs ? (sys.yp + kb) : (sys.Sp + (size_t)t * NB * sys.ld + kb);
  const size_t rstride = inv ? (size_t)NB : rhs ? 0 : (size_t)sys.ld;
  // the row tile, the factor tile and the block inverses are requested together (one round trip instead of two; the
  // all-zero test of the row tile only decides whether the rest runs)
  const float* Ld = sys.Sp + (size_t)kb * sys.ld + kb;
  int nz = 0;
  {
    constexpr int NI = NB * NB / 256;
    float va[NI], vl[NI], vd[2];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int x = tid + 256 * i, r = x / NB, c = x - r * NB;
      va[i] = (r >= rows) ? 0.f : inv ? (r == c ? 1.f : 0.f) : src[r * rstride + c];
      vl[i] = (c <= r) ? Ld[(size_t)r * sys.ld + c] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) vd[i] = (tid + 256 * i < ND_DINV) ? sys.dinv[(size_t)pn.k * ND_DINV + tid + 256 * i] : 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int x = tid + 256 * i, r = x / NB, c = x - r * NB;
      sA[r][c] = va[i];
      sL[r][c] = vl[i];
      nz |= (va[i] != 0.f);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) if (tid + 256 * i < ND_DINV) sD[tid + 256 * i] = vd[i];
  }
  nz = __syncthreads_or(nz);
  if (!nz && !rhs) return;                       // an all-zero tile stays zero: inactive for this panel
  nd_solve_rows(sA, sL, sD, rows);
  __syncthreads();
  for (int x = tid; x < rows * NB; x += 256) {
    const int r = x / NB, c = x - r * NB;
    src[r * rstride + c] = sA[r][c];
  }
  if (tid == 0 && !inv) {
    const int slot = atomicAdd(&sys.nact[pn.k], 1);
    sys.active[(size_t)pn.k * sys.act_stride + slot] = t;          // -1 marks the rhs row
  }
}

__global__ void __launch_bounds__(256) nd_trsm_kernel(Problem pb, int mode, int idx) {
  ND_TS(1, mode, idx);
  __shared__ float sA[NB][NB + 1];
  __shared__ float sL[NB][NB + 1];
  __shared__ float sD[ND_DINV];
  const NdSys sys = nd_sys(pb, blockIdx.y + pb.w0);
  NdPanel pn;
  const bool live = nd_panel(sys.h, mode, idx, blockIdx.x, pn);      // the ordering is final: read ahead of the wait
  pdl_wait();
  pdl_trigger();
  ND_TS_WAITED();
  if (!live) return;
  const int c0 = blockIdx.z;
  if (c0 > pn.nbelow + 1) return;
  nd_trsm_item(sys, pn, c0, sA, sL, sD);
}

// Trailing update over pairs of active row tiles of the panel: Sp[tile a][tile b] -= X_a X_b^T (a below b).
// Look-ahead: the next panel's diagonal tile k + 1 (same segment / border) receives exactly one update from this panel --
// the pair (k + 1, k + 1), if tile k + 1 is active at all.  CTA 0 does that pair first and then factors the tile (the
// longest serial piece of a panel step) while the other CTAs work through the remaining pairs, so the next step starts
// with its row-tile solve instead of a separate factorisation launch.
// grid = (P | 1, batch, gx >= 2), block = 256 (16 x 16 threads, 3 x 3 outputs each), dynamic smem as nd_potf2_kernel
// The trailing update of one panel as seen by CTA `item` of `nitem` (>= 2): item 0 is the look-ahead CTA (when the phase
// has a next panel), the others share the pairs.  All 256 threads; starts with a barrier.
__device__ __forceinline__ void nd_syrk_items(const NdSys& sys, const NdPanel& pn, int mode, int item, int nitem,
                                              float (*sXa)[NB + 1], float (*sXb)[NB + 1], int& s_la, double* sd) {
  __syncthreads();
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int kb = pn.k * NB;
  const size_t ld = (size_t)sys.ld;
  const bool has_next = pn.nsb > 0;
  if (has_next && item == 0) {
    // ---- look-ahead CTA: D <- D - X X^T for the next diagonal tile D = (k + 1, k + 1), X = row tile (k + 1, k), applied in
    // shared memory on the way into the factorisation (the updated tile never goes back to global memory: the factor
    // overwrites it).  An inactive (all-zero) X changes nothing.  No look at the active list: this is the serial chain.
    constexpr int NI = NB * NB / 256;
    const int ld2 = NB | 1;
    double* A = sd;
    const float* X = sys.Sp + (size_t)(kb + NB) * ld + kb;
    const float* D = sys.Sp + (size_t)(kb + NB) * ld + kb + NB;
    float vx[NI], vd[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int x = tid + 256 * i, r = x / NB, c = x - r * NB;
      vx[i] = X[(size_t)r * ld + c];
      vd[i] = (c <= r) ? D[(size_t)r * ld + c] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int x = tid + 256 * i, r = x / NB, c = x - r * NB;
      sXa[c][r] = vx[i];
      A[r * ld2 + c] = (double)vd[i];
    }
    __syncthreads();
    float acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll 4
    for (int k = 0; k < NB; ++k) {
      float a[3], b[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) { a[i] = sXa[k][ty + 16 * i]; b[i] = sXa[k][tx + 16 * i]; }
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
        if (c <= r) A[r * ld2 + c] -= (double)acc[i][j];
      }
    nd_potf2_core(sys, pn.k + 1, sd);
    return;
  }
  const int na = sys.nact[pn.k];
  const int* act = sys.active + (size_t)pn.k * sys.act_stride;
  const int npairs = na * (na + 1) / 2;
  // the look-ahead pair (index of tile k + 1 in the active list) belongs to CTA 0
  if (tid == 0) s_la = -1;
  __syncthreads();
  if (has_next)
    for (int x = tid; x < na; x += 256) if (act[x] == pn.k + 1) s_la = x;
  __syncthreads();
  const int la = s_la;
  const int la_pair = la >= 0 ? la * (la + 1) / 2 + la : -1;
  const int pr0 = item - (has_next ? 1 : 0);
  const int prs = nitem - (has_next ? 1 : 0);
  for (int pr = pr0; pr < npairs; pr += prs) {
    if (pr == la_pair) continue;
    int ia, ib;
    pair_of(pr, ia, ib);
    int ta = act[ia], tb = act[ib];
    if (ta == -1 && tb == -1) continue;            // rhs x rhs: nothing to update
    // "a" is the lower tile (larger row index); the rhs row is below everything
    if (tb == -1 || (ta != -1 && tb > ta)) { const int s = ta; ta = tb; tb = s; }
    const bool rhs = (ta == -1);
    // segments are eliminated concurrently: their updates of border x border tiles (and of the border part of the rhs)
    // land in the same memory
    const bool shared_dst = (mode == 0) && tb >= pn.bbase;
    const int ra = rhs ? 0 : ta * NB, rb = tb * NB;
    const float* xa = rhs ? (sys.yp + kb) : (sys.Sp + (size_t)ra * ld + kb);
    const size_t sa = rhs ? 0 : ld;
    const int rows_a = rhs ? 1 : NB;
    const float* xb = sys.Sp + (size_t)rb * ld + kb;
    __syncthreads();
    {
      constexpr int NI = NB * NB / 256;
      float va[NI], vb[NI];                       // both tiles in flight before the first shared-memory store
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int x = tid + 256 * i, r = x / NB, k = x - r * NB;
        va[i] = (r < rows_a) ? xa[r * sa + k] : 0.f;
        vb[i] = xb[(size_t)r * ld + k];
      }
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int x = tid + 256 * i, r = x / NB, k = x - r * NB;
        sXa[k][r] = va[i];
        sXb[k][r] = vb[i];
      }
    }
    __syncthreads();
    float acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll 4
    for (int k = 0; k < NB; ++k) {
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
        if (r >= rows_a) continue;
        float* dst;
        if (rhs) dst = sys.yp + rb + c;
        else if (rb + c <= ra + r) dst = sys.Sp + (size_t)(ra + r) * ld + rb + c;     // lower triangle only
        else continue;
        if (shared_dst) atomicAdd(dst, -acc[i][j]); else *dst -= acc[i][j];
      }
  }
}

__global__ void __launch_bounds__(256) nd_syrk_kernel(Problem pb, int mode, int idx) {
  ND_TS(2, mode, idx);
  extern __shared__ double sd[];
  __shared__ float sXa[NB][NB + 1];     // [k][row]
  __shared__ float sXb[NB][NB + 1];
  __shared__ int s_la;
  const NdSys sys = nd_sys(pb, blockIdx.y + pb.w0);
  NdPanel pn;
  const bool live = nd_panel(sys.h, mode, idx, blockIdx.x, pn);      // the ordering is final: read ahead of the wait
  pdl_wait();
  pdl_trigger();
  ND_TS_WAITED();
  if (!live) return;
  nd_syrk_items(sys, pn, mode, (int)blockIdx.z, (int)gridDim.z, sXa, sXb, s_la, sd);
}

// The whole border phase in ONE cooperative launch: the number of border panels is only known on the device, and as
// separate launches the worst case (every frame in the border: (6N / 48) steps x 2 launches) has to be enqueued although the
// 1000-frame global BA needs 28 of 125 -- ~0.55 us per empty launch, 0.2 ms per call.  Here the panel loop runs on the
// device with grid-wide barriers (cooperative launch: co-residency guaranteed by the driver) between the row-tile solves and
// the trailing update.  grid = (G, batch), block = 256, dynamic smem as nd_potf2_kernel.
__global__ void __launch_bounds__(256, 2) nd_border_kernel(Problem pb, int first) {
  ND_TS(5, 0, 0);
  ND_TS_WAITED();
  extern __shared__ double sd[];
  __shared__ float sA[NB][NB + 1];
  __shared__ float sL[NB][NB + 1];
  __shared__ float sD[ND_DINV];
  __shared__ int s_la;
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const NdSys sys = nd_sys(pb, blockIdx.y + pb.w0);
  const int Bt = sys.h->Bt;
  int Bmax = 0;                                            // every window walks the same number of barriers
  for (int w = 0; w < (int)gridDim.y; ++w) Bmax = max(Bmax, nd_sys(pb, w + pb.w0).h->Bt);
  const int item = blockIdx.x, nitem = gridDim.x;
  if (first >= Bmax) return;                               // (every CTA of the grid sees the same counts)
  if (first == 0) {                                        // otherwise panel `first` was factored by the look-ahead before it
    if (item == 0 && Bt > 0) nd_potf2_dev(sys, sys.h->bbase, sd);
    grid.sync();
  }
  for (int b = first; b < Bmax; ++b) {
    NdPanel pn;
    const bool live = nd_panel(sys.h, 1, b, 0, pn);
    if (live)
      for (int c0 = item; c0 <= pn.nbelow + 1; c0 += nitem) nd_trsm_item(sys, pn, c0, sA, sL, sD);
    grid.sync();
    if (live) nd_syrk_items(sys, pn, 1, item, nitem, sA, sL, s_la, sd);
    grid.sync();
  }
}

// Backward substitution L^T x = z, one CTA of 1024 threads per (window, segment): the running solution lives in shared
// memory, indexed like yp; per panel the 32 warps split its active row tiles (see big_backsolve_kernel).
// mode 1: the border panels, last to first (grid = (1, batch)); mode 0: segment blockIdx.x, after the border
// (grid = (P, batch)).  dynamic smem: (nd_nt * NB + 33 * NB + NB (NB + 1)) floats
__global__ void __launch_bounds__(1024, 1) nd_backsolve_kernel(Problem pb, int mode) {
  ND_TS(4, mode, 0);
  pdl_wait();
  pdl_trigger();
  ND_TS_WAITED();
  extern __shared__ __align__(16) unsigned char bsm[];
  const NdSys sys = nd_sys(pb, blockIdx.y + pb.w0);
  const NdHeader* h = sys.h;
  const int bbase = h->bbase, nt = h->nt;
  const int p = blockIdx.x;
  const int first = mode ? bbase : h->segbase[p], count = mode ? h->Bt : h->T[p];
  if (count == 0) return;
  float* sx = reinterpret_cast<float*>(bsm);           // [nd_nt * NB]
  float* spart = sx + (size_t)pb.L.nd_nt * NB;         // [32][NB] per-warp partial sums
  float* sz = spart + 32 * NB;                         // [NB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t ld = (size_t)sys.ld;
  for (int x = bbase * NB + tid; x < nt * NB; x += 1024) sx[x] = sys.yp[x];
  if (!mode) for (int x = first * NB + tid; x < (first + count) * NB; x += 1024) sx[x] = sys.yp[x];
  __syncthreads();
  float* sW = sz + NB;                                 // [NB][NB + 1]: W = L^-T of the current diagonal tile
  for (int s = count - 1; s >= 0; --s) {
    const int k = first + s, kb = k * NB;
    const int na = sys.nact[k];
    const int* act = sys.active + (size_t)k * sys.act_stride;
    // W of this panel: requested together with the tile rows below (nothing here depends on x), used after the reduction
    float wv[3];
    const float* Wg = sys.winv + (size_t)k * NB * NB;
#pragma unroll
    for (int i = 0; i < 3; ++i) wv[i] = (tid + 1024 * i < NB * NB) ? Wg[tid + 1024 * i] : 0.f;
    float acc0 = 0, acc1 = 0;                          // columns lane and lane + 32 of the panel
    // work item = (active row tile, quarter of its rows): the 32 warps share the tiles of a panel evenly (a late panel has a
    // handful of active tiles), and an item is ONE batch of loads
    constexpr int RB = 12;                             // 24 loads in flight per lane
    for (int item = warp; item < 4 * na; item += 32) {
      const int t = act[item >> 2];
      if (t == -1) continue;                           // the right-hand-side row of the factorisation
      const int ra = t * NB, rb = (item & 3) * RB;
      const float* Lp = sys.Sp + (size_t)(ra + rb) * ld + kb;
      float v0[RB], v1[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const float* row = Lp + (size_t)r * ld;
        v0[r] = row[lane];
        v1[r] = (lane + 32 < NB) ? row[lane + 32] : 0.f;
      }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const float xr = sx[ra + rb + r];
        acc0 += v0[r] * xr;
        acc1 += v1[r] * xr;
      }
    }
    spart[warp * NB + lane] = acc0;
    if (lane + 32 < NB) spart[warp * NB + lane + 32] = acc1;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int x = tid + 1024 * i;
      if (x < NB * NB) sW[(x / NB) * (NB + 1) + (x % NB)] = wv[i];
    }
    __syncthreads();
    if (tid < NB) {
      float sum = 0;
#pragma unroll
      for (int wi = 0; wi < 32; ++wi) sum += spart[wi * NB + tid];
      sz[tid] = sx[kb + tid] - sum;
    }
    __syncthreads();
    if (tid < NB) {                                    // x[c] = sum_{e >= c} W[c][e] z[e]
      const float* W = sW + tid * (NB + 1);
      float a0 = 0, a1 = 0;
      int e = tid;
      for (; e + 1 < NB; e += 2) { a0 = fmaf(W[e], sz[e], a0); a1 = fmaf(W[e + 1], sz[e + 1], a1); }
      if (e < NB) a0 = fmaf(W[e], sz[e], a0);
      const float a = a0 + a1;
      sx[kb + tid] = a;
      sys.yp[kb + tid] = a;
    }
    __syncthreads();
  }
}

// grid = (ceil(N / 128), batch), block = 128
__global__ void nd_finish_kernel(Problem pb) {
  ND_TS(6, 0, 0);
  pdl_wait();
  pdl_trigger();
  ND_TS_WAITED();
  const int w = blockIdx.y + pb.w0;
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  const NdSys sys = nd_sys(pb, w);
  const int N = pb.t1 - pb.t0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int ps = sys.pos[pb.t0 + i];
  float xi[6];
#pragma unroll
  for (int a = 0; a < 6; ++a) { xi[a] = sys.yp[6 * ps + a]; wp.dX[6 * i + a] = xi[a]; }
  if (pb.apply) retract_pose(pb.poses + (int64_t)w * pb.st.poses + 7 * (int64_t)(pb.t0 + i), xi);
}

// grid = (nd_nt, batch), block = 256: panel blockIdx.x -- its diagonal tile and its active row tiles are the only tiles of
// that column the factorisation has written
__global__ void __launch_bounds__(256) nd_cleanup_kernel(Problem pb) {
  pdl_wait();
  pdl_trigger();
  const NdSys sys = nd_sys(pb, blockIdx.y + pb.w0);
  const int k = blockIdx.x, tid = threadIdx.x;
  if (k >= sys.h->nt) return;
  const size_t ld = (size_t)sys.ld;
  const int na = sys.nact[k];
  const int* act = sys.active + (size_t)k * sys.act_stride;
  for (int ia = -1; ia < na; ++ia) {
    const int t = ia < 0 ? k : act[ia];
    if (t < 0) continue;
    float* tile = sys.Sp + (size_t)t * NB * ld + (size_t)k * NB;
    for (int x = tid; x < NB * NB; x += 256) tile[(size_t)(x / NB) * ld + (x % NB)] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------
bool nd_active(const Problem& pb) { return pb.L.nd_P > 0; }

// after the plan, once per call
void launch_nd_order(const Problem& pb, int64_t batch, cudaStream_t stream) {
  const unsigned B = (unsigned)batch;
  int gx = (int)((pb.E + 255) / 256);
  if (gx > 148 * 8) gx = 148 * 8;
  if (gx < 1) gx = 1;
  launch_k(nd_stats_kernel, dim3(gx, B), dim3(256), 0, stream, pb);
  count_launch();
  launch_k(nd_order_kernel, dim3(1, B), dim3(1024), 0, stream, pb);
  count_launch();
}

// S, y -> dX, poses of one Gauss-Newton iteration; `more`: another iteration follows
cudaError_t launch_nd_solve(const Problem& pb, int64_t batch, cudaStream_t stream, bool more) {
  const unsigned B = (unsigned)batch;
  const int N = pb.t1 - pb.t0, P = pb.L.nd_P, nt = pb.L.nd_nt;
  launch_k(nd_gather_kernel, dim3(148 * 8, B), dim3(256), 0, stream, pb);
  count_launch();
  const size_t psm = sizeof(double) * ((size_t)NB * (NB | 1) + NB);
  cudaFuncSetAttribute(nd_potf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm);
  cudaFuncSetAttribute(nd_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm);
  const int border_steps = (N + 7) / 8;             // worst case: every frame in the border
  // border phase: one cooperative launch with the panel loop on the device when the grid fits (PGBA_ND_COOP=0: launches)
  int coop_g = 0;
  {
    const char* e = getenv("PGBA_ND_COOP");
    if (!(e && e[0] == '0') && batch <= 4) {
      static int slots[64];                          // co-resident CTAs of the cooperative kernel per device (0: not asked yet)
      int dev = 0;
      cudaGetDevice(&dev);
      if (dev >= 0 && dev < 64) {
        if (slots[dev] == 0) {
          int sms = 0, per_sm = 0;
          cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
          cudaFuncSetAttribute(nd_border_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm);
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nd_border_kernel, 256, psm);
          slots[dev] = (per_sm > 2 ? 2 : per_sm) * sms;
          if (slots[dev] <= 0) slots[dev] = -1;
        }
        const int g = slots[dev] / (int)batch;
        if (g >= 16) coop_g = g;
      }
    }
  }
  // Border panels: the first `lead` as launches (a kernel boundary with programmatic dependent launch costs ~0.7 us, a
  // grid-wide barrier ~2-3 us: 17.6 against 21.4 us per panel), the rest -- their number is only known on the device -- in the
  // cooperative kernel, which returns at once when the launches covered them.  lead = a quarter of the tile capacity: a
  // chain + loop-closure graph keeps 1 / 4 .. 1 / 5 of its frames in the border (c4: 28 of 142 tiles).
  const int lead = coop_g > 0 ? (nt / 4 > 4 ? nt / 4 : 4) : border_steps;
  for (int mode = 0; mode < 2; ++mode) {
    const int steps = mode == 0 ? pb.L.nd_tmax : (lead < border_steps ? lead : border_steps);
    const unsigned Z = mode == 0 ? (unsigned)P : 1u;
    // CTAs per segment / for the border in the trailing update: a segment panel has ~10-15 active tiles (50-120 pairs)
    const int gx = mode == 0 ? 64 : 296;
    for (int s = 0; s < steps; ++s) {
      if (s == 0) { launch_k(nd_potf2_kernel, dim3(Z, B), dim3(256), psm, stream, pb, mode, s); count_launch(); }
      launch_k(nd_trsm_kernel, dim3(Z, B, nt + 2), dim3(256), 0, stream, pb, mode, s);
      launch_k(nd_syrk_kernel, dim3(Z, B, gx), dim3(256), psm, stream, pb, mode, s);
      count_launch(); count_launch();
    }
    if (mode == 1 && coop_g > 0 && steps < border_steps) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)coop_g, B); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = psm; cfg.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeCooperative;
      at[0].val.cooperative = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      if (cudaLaunchKernelEx(&cfg, nd_border_kernel, pb, steps) == cudaSuccess) {
        count_launch();
      } else {                                         // e.g. the grid is not co-resident on this device: the rest as launches
        (void)cudaGetLastError();
        for (int s = steps; s < border_steps; ++s) {
          launch_k(nd_trsm_kernel, dim3(Z, B, nt + 2), dim3(256), 0, stream, pb, mode, s);
          launch_k(nd_syrk_kernel, dim3(Z, B, gx), dim3(256), psm, stream, pb, mode, s);
          count_launch(); count_launch();
        }
      }
    }
  }
  const size_t bsmem = sizeof(float) * ((size_t)nt * NB + 33 * NB + NB * (NB + 1));
  cudaFuncSetAttribute(nd_backsolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem);
  launch_k(nd_backsolve_kernel, dim3(1, B), dim3(1024), bsmem, stream, pb, 1);
  launch_k(nd_backsolve_kernel, dim3(P, B), dim3(1024), bsmem, stream, pb, 0);
  launch_k(nd_finish_kernel, dim3((N + 127) / 128, B), dim3(128), 0, stream, pb);
  count_launch(); count_launch(); count_launch();
  if (more) {
    launch_k(nd_cleanup_kernel, dim3(nt, B), dim3(256), 0, stream, pb);
    count_launch();
  }
  return cudaGetLastError();
}

#ifdef PGBA_ND_TIMING
void nd_timestamps(unsigned long long* out, int reset) {
  if (reset) {
    static unsigned long long init[8 * 1024][3];
    for (auto& r : init) { r[0] = ~0ull; r[1] = ~0ull; r[2] = 0ull; }
    cudaMemcpyToSymbol(g_nd_ts, init, sizeof(init));
  } else {
    cudaMemcpyFromSymbol(out, g_nd_ts, sizeof(unsigned long long) * 8 * 1024 * 3);
  }
}
#endif

}  // namespace pgba

#ifdef PGBA_ND_TIMING
extern "C" void pgba_nd_timestamps(unsigned long long* out, int reset) { pgba::nd_timestamps(out, reset); }
#endif
