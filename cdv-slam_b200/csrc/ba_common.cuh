// Shared device helpers and workspace layout of the patch-graph BA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/pgba.h"

namespace pgba {

constexpr int PMAX = 128;                 // upper bound of patches per chunk (one source frame, consecutive patch ids)
constexpr int SMAX = PGBA_MAX_SLOTS;      // distinct target frames per chunk
constexpr int EBUDGET = 12288;            // floats of shared memory for the per-batch E tile (48 KB)
constexpr int SOLVE_NMAX = 156;           // 6N handled by the single-CTA shared-memory solve (N <= 26)
constexpr int BIG_NB = 48;                // panel width of the blocked global-memory Cholesky (6N > SOLVE_NMAX)
constexpr int ND_MIN_N = 27;              // free poses from which the large solve reorders the frames (ba_bignd.cu): all of them

// ---------------------------------------------------------------------------------------------------------------
// Workspace.  [ zero region of window 0 | ... | zero region of window B-1 | body of window 0 | ... ]
// The zero regions are cleared by one memset at the start of every call; all offsets are 256-byte aligned.
// ---------------------------------------------------------------------------------------------------------------
struct WinHeader {      // first thing in the zero region
  int n_chunks;
  int n_patches;        // bump cursor: unique (source frame, patch) rows allocated so far
  int n_slots;          // bump cursor into slot_frames
  int n_cells;          // bump cursor into cells
  int n_ecells;         // bump cursor into ecells (units of 6 floats)
  int n_dups;           // duplicate (patch, target frame) edges
  int status;           // PGBA_ST_* bits
  int n_valid_edges;
  int ticket[4];        // "last block done" counters of the plan kernels
  int chol_info;        // 0, or 1 + index of the first non-positive pivot (big solve)
  int plan_hit;         // 1: this call found the window's tables valid (fingerprint match) and skipped the graph analysis
  int pad[2];
  // plan cache (single-launch cluster plans only; the memset of the grid-wide path clears it): 128-bit fingerprint of the
  // window's (ii, jj, kk, edge count) the tables below were built from (per-CTA partial sums, combined over the window's
  // cluster through distributed shared memory -- no global accumulator, so stale workspace contents cannot leak into it), and -- window 0 only, i.e. at byte offset 64 + 16 of
  // EVERY layout -- the descriptor of the call that last ran a plan on this workspace.  A window's tables are reused only if
  // both match, so a call with another layout in between (which rewrites the descriptor) invalidates every window.
  unsigned long long fp[2];
  int desc[8];          // magic, E, F, K, t0, t1, pc, batch
};
static_assert(sizeof(WinHeader) <= 256, "the header must fit the first 256-byte slot of the zero region");
constexpr int PLAN_DESC_MAGIC = 0x50474241;

struct Chunk {          // 64 bytes
  int frame;            // source frame i
  int kbase;            // patch ids covered: [kbase, kbase + pc)
  int edge_begin, edge_end;
  int n_patches, n_slots;
  int patch_base, slot_base, cell_base, ecell_base;
  int first_free;       // slots [first_free, first_free + n_free) are free poses (slots are sorted by frame)
  int n_free;
  int icol;             // E column of the source frame (-1 when the source pose is fixed)
  int ncols;            // n_free (+1 when the source frame is free and not itself a target)
  int pad0, pad1;
};

static_assert(sizeof(Chunk) == 64, "Chunk is written as four 16-byte words");

struct DupEdge { int chunk, p, s, n; };

// Nested-dissection ordering of the large (global BA) solve, see ba_bignd.cu.  Lives in the zero region.
constexpr int ND_MAXP = 32;
struct NdHeader {
  int nt;               // tiles (48 unknowns = 8 frames each) of the permuted system actually in use
  int bbase, Bt;        // first tile / number of tiles of the border (loop-closure targets + separators)
  int n_border;         // border frames
  int T[ND_MAXP];       // tiles of chain segment p
  int segbase[ND_MAXP]; // first tile of chain segment p
};
static_assert(sizeof(NdHeader) <= 512, "NdHeader slot");

inline bool nd_enabled() {                          // PGBA_BIG_ND=0: the natural-order blocked solver (A/B runs, tests)
  const char* e = getenv("PGBA_BIG_ND");
  return !(e && e[0] == '0');
}

struct Layout {         // host-computed
  int64_t E, F, K;      // edges (max per window), pose rows, patch rows
  int N;                // free poses
  int pc;               // patches per chunk (power of two, 8..128)
  int big;              // 1: dense S too large for the shared-memory solve -> blocked global-memory Cholesky
  int64_t ch_max, patch_max, slot_max, cell_cap, ecell_cap;
  // zero region (relative to the window's zero base)
  size_t z_hdr, z_fmaxinv, z_fkmax1, z_ccnt, z_nact, z_bs, z_y, z_S, zero_bytes;
  size_t o_rdiag, o_winv, o_active;   // big solve only
  int big_steps, big_tiles;
  // nested-dissection ordering (big solve with >= ND_MIN_N free poses): P chain segments + border, permuted system of
  // nd_nt tiles (capacity), nd_tmax = tiles of the longest possible segment, nd_R = "far edge" frame distance
  int nd_P, nd_nt, nd_tmax, nd_R;
  size_t z_nd, z_ndf, z_yp, z_Sp, o_frame_at, o_dinv;
  // body (relative to the window's body base)
  size_t o_fbase, o_ccur, o_chunks, o_perm, o_kx, o_slots, o_cells, o_dups, o_ecells, o_Q, o_u, o_dZ, o_dX, body_bytes;
  size_t body0;         // offset of the first body = batch * zero_bytes (aligned)
};

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

inline int choose_pc(int64_t E, int64_t batch, int64_t F, int64_t K) {
  if (batch >= 2 && F > 0 && K > 0) {
    // batches: the chunk-parallel kernels hold 2 CTAs per SM (296 slots).  Within one wave more, smaller chunks finish sooner;
    // beyond it the largest chunks win (the large-chunk code paths do less work per edge).  Measured on c2-shaped windows
    // (L2 flushed, us per call; chunks = windows x 22 frames x ceil(96 / pc)):
    //    4 windows: pc 32 (264 chunks) 94.6, pc 64 (176) 100.6      8 windows: pc 32 (528) 125.0, 64 (352) 116.8, 128 (176) 114.5
    //   12 windows: pc 64 (528) 127.0, pc 128 (264) 116.8          16 windows: pc 64 (704) 145.4, pc 128 (352) 129.0
    // -> the largest pc that still yields >= 200 chunks, estimating the active frames as E / 24 / (patches per frame)
    const int64_t M = K / F > 0 ? K / F : 1;
    int64_t frames = (E / 24 + M - 1) / M;
    if (frames < 1) frames = 1;
    for (int pc = PMAX; pc > 8; pc >>= 1)
      if (batch * frames * ((M + pc - 1) / pc) >= 200) return pc;
    return 8;
  }
  // single window, throughput regime (global BA): aim at >= ~1.5 CTAs per SM for the chunk-parallel kernels, assuming
  // ~24 edges per patch
  int64_t want = (E * batch) / (148 * 3 / 2 * 24);
  int pc = 8;
  while (pc * 2 <= want && pc < PMAX) pc *= 2;
  if (pc >= 32) return pc;
  // latency regime (single window): about one chunk per SM -- the power of two nearest to E / (148 * 16), at most 32.
  // Measured on c2 (37 824 edges, same box): pc = 8 (264 chunks) 86.0 us, 16 (132 chunks) 84.0 us, 32 (66 chunks) 92.1 us.
  want = (E * batch) / (148 * 16);
  pc = 8;
  while (pc * 2 * 1000 <= want * 1414 && pc < 32) pc *= 2;
  return pc;
}

inline Layout make_layout(int64_t E, int64_t F, int64_t K, int N, int64_t batch, int pc) {
  Layout L{};
  L.E = E; L.F = F; L.K = K; L.N = N; L.pc = pc;
  L.big = (6 * N > SOLVE_NMAX) ? 1 : 0;
  int64_t e1 = E > 0 ? E : 1;
  L.ch_max = F + K / pc + 1;
  if (L.ch_max > e1) L.ch_max = e1;
  L.patch_max = K < e1 ? K : e1;
  if (L.patch_max < 1) L.patch_max = 1;
  L.slot_max = e1;
  L.cell_cap = 8 * e1 + 1024;
  L.ecell_cap = N > 0 ? (L.cell_cap + L.patch_max) : 1;
  const size_t n6 = (size_t)6 * (size_t)(N > 0 ? N : 1);
  size_t o = 0;
  L.z_hdr = o;     o = align256(o + sizeof(WinHeader));
  L.z_fmaxinv = o; o = align256(o + 4 * (size_t)F);
  L.z_fkmax1 = o;  o = align256(o + 4 * (size_t)F);
  L.z_ccnt = o;    o = align256(o + 4 * (size_t)L.ch_max);
  L.big_steps = L.big ? (int)((n6 + BIG_NB - 1) / BIG_NB) : 0;
  L.big_tiles = L.big_steps + 1;
  L.nd_P = 0; L.nd_nt = 0; L.nd_tmax = 0; L.nd_R = 0;
  if (L.big && N >= ND_MIN_N && nd_enabled()) {
    int P = N / 60;
    if (P < 1) P = 1;
    if (P > ND_MAXP) P = ND_MAXP;
    const int nt = (N + 7) / 8 + P + 1;               // every segment and the border round up to whole tiles
    // the backward substitution keeps the whole solution in shared memory
    if (sizeof(float) * ((size_t)nt * BIG_NB + 33 * BIG_NB + BIG_NB * (BIG_NB + 1)) <= 200 * 1024) {
      L.nd_P = P; L.nd_nt = nt;
      const int lseg = (N + P - 1) / P;
      L.nd_tmax = (lseg + 7) / 8;
      L.nd_R = lseg / 4 > 4 ? lseg / 4 : 4;
      L.big_steps = nt; L.big_tiles = nt + 1;         // capacity of winv / active / nact
    }
  }
  const size_t np = (size_t)L.nd_nt * BIG_NB;
  L.z_nact = o;    o = align256(o + 4 * (size_t)(L.big_steps + 1));
  L.z_bs = o;      o = align256(o + 512);
  L.z_nd = o;      o = align256(o + (L.nd_P ? 512 : 0));
  L.z_ndf = o;     o = align256(o + (L.nd_P ? 6 * 4 * (size_t)F : 0));   // lminv, lmax1, border, pos, pfb, pfn
  L.z_yp = o;      o = align256(o + 4 * np);
  L.z_Sp = o;      o = align256(o + 4 * np * np);
  L.z_y = o;       o = align256(o + 4 * n6);
  L.z_S = o;       o = align256(o + 4 * n6 * n6);
  L.zero_bytes = o;
  o = 0;
  L.o_fbase = o;  o = align256(o + 4 * (size_t)F);
  L.o_ccur = o;   o = align256(o + 4 * (size_t)L.ch_max);
  L.o_chunks = o; o = align256(o + sizeof(Chunk) * (size_t)L.ch_max);
  L.o_perm = o;   o = align256(o + 16 * (size_t)e1);       // int4 records: edge, target frame, patch id, -
  L.o_kx = o;     o = align256(o + 4 * (size_t)L.patch_max);
  L.o_slots = o;  o = align256(o + 4 * (size_t)L.slot_max);
  L.o_cells = o;  o = align256(o + 4 * (size_t)L.cell_cap);
  L.o_dups = o;   o = align256(o + sizeof(DupEdge) * (size_t)e1);
  L.o_ecells = o; o = align256(o + 24 * (size_t)L.ecell_cap);
  L.o_Q = o;      o = align256(o + 4 * (size_t)L.patch_max);
  L.o_u = o;      o = align256(o + 4 * (size_t)L.patch_max);
  L.o_dZ = o;     o = align256(o + 4 * (size_t)L.patch_max);
  L.o_dX = o;     o = align256(o + 4 * n6);
  L.o_rdiag = o;  o = align256(o + (L.big ? 4 * n6 : 0));
  L.o_winv = o;   o = align256(o + (L.big ? 4 * (size_t)L.big_steps * BIG_NB * BIG_NB : 0));
  L.o_active = o; o = align256(o + (L.big ? 4 * (size_t)L.big_steps * L.big_tiles : 0));
  L.o_frame_at = o; o = align256(o + 4 * (size_t)L.nd_nt * 8);
  L.o_dinv = o;   o = align256(o + 4 * (size_t)L.nd_nt * (BIG_NB / 6) * 36);   // inverses of the 6 x 6 diagonal blocks of L
  L.body_bytes = o;
  L.body0 = L.zero_bytes * (size_t)batch;
  return L;
}

inline size_t total_bytes(const Layout& L, int64_t batch) { return (L.zero_bytes + L.body_bytes) * (size_t)batch; }

// Pointers of one window, resolved on the device from (workspace base, window index, layout).
struct WinPtrs {
  WinHeader* hdr; int* fmaxinv; int* fkmax1; int* ccnt; float* y; float* S;
  int* fbase; int* ccur; Chunk* chunks; int4* perm; int* kx; int* slots; int* cells; DupEdge* dups; float* ecells;
  float* Q; float* u; float* dZ; float* dX;
};

__host__ __device__ inline WinPtrs win_ptrs(void* ws, const Layout& L, int64_t b) {
  char* z = (char*)ws + (size_t)b * L.zero_bytes;
  char* base = (char*)ws + L.body0 + (size_t)b * L.body_bytes;
  WinPtrs p;
  p.hdr = (WinHeader*)(z + L.z_hdr);
  p.fmaxinv = (int*)(z + L.z_fmaxinv);
  p.fkmax1 = (int*)(z + L.z_fkmax1);
  p.ccnt = (int*)(z + L.z_ccnt);
  p.y = (float*)(z + L.z_y);
  p.S = (float*)(z + L.z_S);
  p.fbase = (int*)(base + L.o_fbase);
  p.ccur = (int*)(base + L.o_ccur);
  p.chunks = (Chunk*)(base + L.o_chunks);
  p.perm = (int4*)(base + L.o_perm);
  p.kx = (int*)(base + L.o_kx);
  p.slots = (int*)(base + L.o_slots);
  p.cells = (int*)(base + L.o_cells);
  p.dups = (DupEdge*)(base + L.o_dups);
  p.ecells = (float*)(base + L.o_ecells);
  p.Q = (float*)(base + L.o_Q);
  p.u = (float*)(base + L.o_u);
  p.dZ = (float*)(base + L.o_dZ);
  p.dX = (float*)(base + L.o_dX);
  return p;
}

// Problem description shared by all kernels (passed by value).
struct Problem {
  float* poses; float* patches; const float* intrinsics; const float* target; const float* weight; const float* lmbda;
  const int64_t* ii; const int64_t* jj; const int64_t* kk; const int32_t* n_edges_dev;
  pgba_strides st;
  int64_t E; int F; int K; int P; int t0; int t1; int with_schur; int apply;
  int plan_cache;       // 1: reuse a window's plan tables when its edge list is unchanged since the last call (see WinHeader)
  int batch;
  int w0;               // first window handled by this launch (window groups on separate streams; 0 otherwise)
  int idx32;            // 1: ii / jj / kk point to int32 arrays (host-buffer entry with 32-bit index uploads)
  void* ws; Layout L;
};

// ---------------------------------------------------------------------------------------------------------------
// SE3 helpers.  Same arithmetic as the reference's device helpers (cdvslam/fastba/ba_cuda.cu:36-174) but organised
// around the 3x3 matrix of the (possibly non-unit) relative quaternion so that it is computed once per frame pair.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rot_q(const float q[4], const float X[3], float Y[3]) {   // ba_cuda.cu:36-46
  float uv0 = 2.0f * (q[1] * X[2] - q[2] * X[1]);
  float uv1 = 2.0f * (q[2] * X[0] - q[0] * X[2]);
  float uv2 = 2.0f * (q[0] * X[1] - q[1] * X[0]);
  Y[0] = X[0] + q[3] * uv0 + (q[1] * uv2 - q[2] * uv1);
  Y[1] = X[1] + q[3] * uv1 + (q[2] * uv0 - q[0] * uv2);
  Y[2] = X[2] + q[3] * uv2 + (q[0] * uv1 - q[1] * uv0);
}

// Relative transform Gij = Gj * Gi^-1 as (R row-major [9], t [3])   (ba_cuda.cu:74-85)
__device__ __forceinline__ void rel_pose(const float* __restrict__ Pi, const float* __restrict__ Pj, float R[9],
                                         float t[3]) {
  const float qi[4] = {Pi[3], Pi[4], Pi[5], Pi[6]}, qj[4] = {Pj[3], Pj[4], Pj[5], Pj[6]};
  float q[4];
  q[0] = -qj[3] * qi[0] + qj[0] * qi[3] - qj[1] * qi[2] + qj[2] * qi[1];
  q[1] = -qj[3] * qi[1] + qj[1] * qi[3] - qj[2] * qi[0] + qj[0] * qi[2];
  q[2] = -qj[3] * qi[2] + qj[2] * qi[3] - qj[0] * qi[1] + qj[1] * qi[0];
  q[3] = qj[3] * qi[3] + qj[0] * qi[0] + qj[1] * qi[1] + qj[2] * qi[2];
  const float ti[3] = {Pi[0], Pi[1], Pi[2]};
  float r[3];
  rot_q(q, ti, r);
  t[0] = Pj[0] - r[0]; t[1] = Pj[1] - r[1]; t[2] = Pj[2] - r[2];
  const float ex[3] = {1.f, 0.f, 0.f}, ey[3] = {0.f, 1.f, 0.f}, ez[3] = {0.f, 0.f, 1.f};
  float c0[3], c1[3], c2[3];
  rot_q(q, ex, c0); rot_q(q, ey, c1); rot_q(q, ez, c2);
  R[0] = c0[0]; R[1] = c1[0]; R[2] = c2[0];
  R[3] = c0[1]; R[4] = c1[1]; R[5] = c2[1];
  R[6] = c0[2]; R[7] = c1[2]; R[8] = c2[2];
}

// The map the reference calls adjSE3 (ba_cuda.cu:57-72): Y[:3] = R^T a, Y[3:] = R^T (b + a x t), X = (a, b).
__device__ __forceinline__ void adj_map(const float R[9], const float t[3], const float X[6], float Y[6]) {
  const float a0 = X[0], a1 = X[1], a2 = X[2];
  const float b0 = X[3] + (a1 * t[2] - a2 * t[1]);
  const float b1 = X[4] + (a2 * t[0] - a0 * t[2]);
  const float b2 = X[5] + (a0 * t[1] - a1 * t[0]);
  Y[0] = R[0] * a0 + R[3] * a1 + R[6] * a2;
  Y[1] = R[1] * a0 + R[4] * a1 + R[7] * a2;
  Y[2] = R[2] * a0 + R[5] * a1 + R[8] * a2;
  Y[3] = R[0] * b0 + R[3] * b1 + R[6] * b2;
  Y[4] = R[1] * b0 + R[4] * b1 + R[7] * b2;
  Y[5] = R[2] * b0 + R[5] * b1 + R[8] * b2;
}

// SE3 retraction of one pose row: P <- Exp(xi) * P, xi = (tau, phi)   (ba_cuda.cu:88-174, used by :178-206)
__device__ __forceinline__ void retract_pose(float* P, const float* xi) {
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
// Programmatic dependent launch (PDL).  Every BA kernel starts with pdl_wait() -- it blocks until the previous
// kernel of the stream has completed and its writes are visible -- and is launched with the programmatic-stream-
// serialisation attribute, so its CTAs are scheduled (and run their prologue up to pdl_wait) while the previous
// kernel drains.  pdl_trigger() lets the next kernel start being scheduled.  PGBA_PDL=0 disables the attribute.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// Host-side count of kernel launches issued by this library (reported by bench.py as `gpu_launches`).
void count_launch();
long long launch_count();

// Upper-triangular packing of a symmetric 6x6: index of (r, c) with r <= c.
__host__ __device__ constexpr int sym6(int r, int c) { return r * 6 - (r * (r - 1)) / 2 + (c - r); }

}  // namespace pgba
