// extern "C" entry points of the BA part of libpgba.so (see include/pgba.h) + the reproject kernel.
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "ba_common.cuh"

namespace pgba {
void launch_plan(const Problem& pb, int64_t batch, cudaStream_t stream);
cudaError_t launch_iteration(const Problem& pb, int64_t batch, cudaStream_t stream, cudaEvent_t* ev, bool first, bool more);
void launch_linearize(const Problem& pb, int64_t batch, cudaStream_t stream, bool fuse_update, bool first);
void launch_solve(const Problem& pb, int64_t batch, cudaStream_t stream, bool more = false);
void launch_update(const Problem& pb, int64_t batch, cudaStream_t stream);
bool solve_supported(int N);
bool plan_clears_workspace(const Problem& pb, int64_t batch);
#ifdef PGBA_PLAN_TIMING
void plan_timestamps(unsigned long long* out);
#endif
#ifdef PGBA_LIN_TIMING
void lin_timestamps(long long* out);
void cta_timestamps(unsigned long long* out);
void plan_cta_timestamps(unsigned long long* out);
#endif

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PGBA_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// all PxP pixels of every edge's patch, frame ii -> jj; one thread per (edge, pixel); coords [E, 2, P, P]
//   clamp_depth == 0: the reference's reproject kernel (cdvslam/fastba/ba_cuda.cu:408-458): intrinsics row 0 for every
//                     edge, unguarded X / Z
//   clamp_depth != 0: pops.transform as slam.py:328 calls it (cdvslam/projective_ops.py:19-68): back-projection with the
//                     intrinsics of the SOURCE frame (intrinsics[:, ii], :57), projection with those of the TARGET frame
//                     (intrinsics[:, jj], :68) and d = 1 / Z.clamp(min=0.1) (:43)
__global__ void reproject_kernel(const float* __restrict__ poses, const float* __restrict__ patches,
                                 const float* __restrict__ intr, const int64_t* __restrict__ ii,
                                 const int64_t* __restrict__ jj, const int64_t* __restrict__ kk, int64_t E, int P,
                                 int clamp_depth, float* __restrict__ coords) {
  pdl_wait();
  pdl_trigger();
  const int PP = P * P;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= E * PP) return;
  const int64_t n = idx / PP;
  const int px = (int)(idx - n * PP);
  const int64_t fi = ii[n], fj = jj[n];
  const float* ki = clamp_depth ? intr + 4 * fi : intr;
  const float* kj = clamp_depth ? intr + 4 * fj : intr;
  float R[9], t[3];
  rel_pose(poses + 7 * fi, poses + 7 * fj, R, t);
  const float* pr = patches + kk[n] * 3 * PP;
  const float xi0 = (pr[px] - ki[2]) / ki[0], xi1 = (pr[PP + px] - ki[3]) / ki[1], pd = pr[2 * PP + px];
  const float X = R[0] * xi0 + R[1] * xi1 + R[2] + pd * t[0];
  const float Y = R[3] * xi0 + R[4] * xi1 + R[5] + pd * t[1];
  const float Z = R[6] * xi0 + R[7] * xi1 + R[8] + pd * t[2];
  const float fx = kj[0], fy = kj[1], cx = kj[2], cy = kj[3];
  float u, v;
  if (clamp_depth) {
    const float d = 1.0f / fmaxf(Z, 0.1f);
    u = fx * (d * X) + cx;
    v = fy * (d * Y) + cy;
  } else {                                 // ba_cuda.cu:451-452
    u = fx * (X / Z) + cx;
    v = fy * (Y / Z) + cy;
  }
  coords[n * 2 * PP + px] = u;
  coords[n * 2 * PP + PP + px] = v;
}

__global__ void export_debug_kernel(Problem pb, float* S, float* y, float* dX, int64_t* patch_ids, float* C, float* u,
                                    float* Q, float* dZ, int32_t* n_unique, int32_t* status) {
  pdl_wait();
  pdl_trigger();
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, 0);
  const size_t n6 = (size_t)6 * (pb.t1 - pb.t0);
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, T = (size_t)gridDim.x * blockDim.x;
  const float lm = pb.lmbda[0];
  if (S) for (size_t x = tid; x < n6 * n6; x += T) {       // only the lower block triangle is accumulated
    const size_t r = x / n6, c = x - r * n6;
    S[x] = (c / 6 <= r / 6) ? wp.S[x] : wp.S[c * n6 + r];
  }
  if (y) for (size_t x = tid; x < n6; x += T) y[x] = wp.y[x];
  if (dX) for (size_t x = tid; x < n6; x += T) dX[x] = wp.dX[x];
  const int M = wp.hdr->n_patches;
  for (size_t x = tid; x < (size_t)M; x += T) {
    if (patch_ids) patch_ids[x] = wp.kx[x];
    if (Q) Q[x] = wp.Q[x];
    if (C) C[x] = 1.0f / wp.Q[x] - lm;
    if (u) u[x] = wp.u[x];
    if (dZ) dZ[x] = wp.dZ[x];
  }
  if (tid == 0) {
    if (n_unique) *n_unique = M;
    if (status) *status = wp.hdr->status;
  }
}

static int check_common(const void* poses, const void* patches, const void* intrinsics, const void* target,
                        const void* weight, const void* lmbda, const void* ii, const void* jj, const void* kk,
                        int64_t E, int64_t F, int64_t K, int P, int t0, int t1) {
  if (!poses || !patches || !intrinsics || !lmbda) return PGBA_ERR_NULL;
  if (E > 0 && (!target || !weight || !ii || !jj || !kk)) return PGBA_ERR_NULL;
  if (E < 0 || F <= 0 || K <= 0 || P < 2 || t0 < 0 || t1 < t0 || t1 > F) return PGBA_ERR_SHAPE;
  if (E >= (int64_t)1 << 31 || K >= (int64_t)1 << 31) return PGBA_ERR_UNSUPPORTED;
  if (F > PGBA_MAX_POSE_ROWS) return PGBA_ERR_UNSUPPORTED;
  return PGBA_OK;
}

// PGBA_PLAN_CACHE=0: rebuild the plan tables on every call even when the edge list is unchanged (A/B runs, cold timings)
static bool plan_cache_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PGBA_PLAN_CACHE");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// patches per chunk: heuristic, or the PGBA_PC environment variable (8..128, power of two) for tuning
static int pick_pc(int64_t E, int64_t batch, int64_t F, int64_t K) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("PGBA_PC");
    int v = e ? atoi(e) : 0;
    forced = (v == 8 || v == 16 || v == 32 || v == 64 || v == 128) ? v : 0;
  }
  return forced ? forced : choose_pc(E, batch, F, K);
}

static Problem make_problem(float* poses, float* patches, const float* intrinsics, const float* target,
                            const float* weight, const float* lmbda, const int64_t* ii, const int64_t* jj,
                            const int64_t* kk, const int32_t* n_edges_dev, const pgba_strides* st, int64_t E,
                            int64_t F, int64_t K, int P, int t0, int t1, void* ws, int64_t batch) {
  Problem pb{};
  pb.poses = poses; pb.patches = patches; pb.intrinsics = intrinsics; pb.target = target; pb.weight = weight;
  pb.lmbda = lmbda; pb.ii = ii; pb.jj = jj; pb.kk = kk; pb.n_edges_dev = n_edges_dev;
  if (st) pb.st = *st;
  pb.E = E; pb.F = (int)F; pb.K = (int)K; pb.P = P; pb.t0 = t0; pb.t1 = t1; pb.with_schur = 1; pb.apply = 1;
  pb.plan_cache = plan_cache_enabled() ? 1 : 0;
  pb.batch = (int)batch;
  pb.ws = ws;
  pb.L = make_layout(E, F, K, t1 - t0, batch, pick_pc(E, batch, F, K));
  return pb;
}

// one memset for the zero regions of all windows (headers, frame statistics, chunk counters, S, y)
static cudaError_t clear_workspace(const Problem& pb, int64_t batch, cudaStream_t s) {
  if (plan_clears_workspace(pb, batch)) return cudaSuccess;          // done by the first phase of plan_cluster_kernel
  return cudaMemsetAsync(pb.ws, 0, pb.L.zero_bytes * (size_t)batch, s);
}

// ---- window groups.  A batched call runs plan -> k x (linearize -> solve -> update).  The linearisation is throughput
// bound (every SM busy); the small solve (one CTA per window) and the back-substitution are latency bound and leave most of
// the machine idle for ~35 us per iteration on the 64-window batch.  The windows are independent, so the call splits them
// into up to four groups after the plan: group 0 stays on the caller's stream, the others run their iterations on
// auxiliary streams (forked and joined with events, so the whole call is still one capturable unit of work on the
// caller's stream) and the solve / update of one group overlaps the linearisation of the next.
// PGBA_BATCH_GROUPS=1..4 overrides the group count (A/B runs).
constexpr int MAX_GROUPS = 4;
struct GroupCtx { int dev; cudaStream_t main; cudaStream_t aux[MAX_GROUPS - 1]; cudaEvent_t fork, join[MAX_GROUPS - 1]; };

static GroupCtx* group_ctx(cudaStream_t s) {
  static std::mutex mu;
  static std::vector<GroupCtx*> cache;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  for (GroupCtx* c : cache)
    if (c->dev == dev && c->main == s) return c;
  GroupCtx* c = new GroupCtx{};
  c->dev = dev; c->main = s;
  bool ok = cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming) == cudaSuccess;
  for (int g = 0; g < MAX_GROUPS - 1 && ok; ++g)
    ok = cudaStreamCreateWithFlags(&c->aux[g], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->join[g], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { (void)cudaGetLastError(); delete c; return nullptr; }      // (leaks the few handles created so far; never seen)
  cache.push_back(c);
  return c;
}

static int batch_groups(const Problem& pb, int64_t batch) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("PGBA_BATCH_GROUPS"); forced = e ? atoi(e) : 0; }
  int g = 1;
  // measured (us per call, 1 / 2 / 4 groups): 8 windows 112.7 / 108.5 / -, 12: 112.7 / 110.8 / -, 16: 129.0 / 129.1 / 135.2,
  // 24: - / 147.5 / 151.5, 64: 336 / 322 / 315
  if (!pb.L.big && pb.t1 > pb.t0) g = batch >= 32 ? 4 : (batch >= 4 ? 2 : 1);
  if (forced >= 1 && forced <= MAX_GROUPS) g = forced;
  if (g > batch) g = (int)batch;
  return g < 1 ? 1 : g;
}

// plan + iterations of a batch, on window groups when that pays.  The workspace's zero regions have been dealt with
// (clear_workspace) on `s`.  The plan of all windows runs first, on the caller's stream: its CTAs (1024 threads, the whole
// register file of an SM) cannot share an SM with the linearisation's, so a plan per group would not overlap anything
// (measured: 318 - 365 us against 316 us on the 64-window batch, profiles/README.md round 2).
static cudaError_t launch_planned(const Problem& pb, int64_t batch, int iterations, cudaStream_t s) {
  launch_plan(pb, batch, s);
  const int G = batch_groups(pb, batch);
  GroupCtx* gc = G > 1 ? group_ctx(s) : nullptr;
  cudaError_t e = cudaSuccess;
  if (!gc) {
    for (int it = 0; it < iterations && e == cudaSuccess; ++it) e = launch_iteration(pb, batch, s, nullptr, it == 0, it + 1 < iterations);
    return e;
  }
  e = cudaEventRecord(gc->fork, s);
  if (e != cudaSuccess) return e;
  cudaError_t first_err = cudaSuccess;
  for (int g = 0; g < G; ++g) {
    cudaStream_t sg = g == 0 ? s : gc->aux[g - 1];
    cudaError_t eg = g == 0 ? cudaSuccess : cudaStreamWaitEvent(sg, gc->fork, 0);
    Problem pg = pb;
    pg.w0 = (int)(batch * g / G);
    const int64_t nb = batch * (g + 1) / G - pg.w0;
    for (int it = 0; it < iterations && eg == cudaSuccess && nb > 0; ++it)
      eg = launch_iteration(pg, nb, sg, nullptr, it == 0, it + 1 < iterations);
    if (g > 0) {                                    // joined whatever happened: an unjoined fork would break a stream capture
      cudaError_t ej = cudaEventRecord(gc->join[g - 1], sg);
      if (ej == cudaSuccess) ej = cudaStreamWaitEvent(s, gc->join[g - 1], 0);
      if (eg == cudaSuccess) eg = ej;
    }
    if (first_err == cudaSuccess) first_err = eg;
  }
  return first_err;
}
}  // namespace pgba

using namespace pgba;

extern "C" {

const char* pgba_error_string(int code) {
  switch (code) {
    case PGBA_OK: return "ok";
    case PGBA_ERR_NULL: return "pgba: required pointer is NULL";
    case PGBA_ERR_SHAPE: return "pgba: invalid size / shape argument";
    case PGBA_ERR_WORKSPACE: return "pgba: workspace too small or misaligned";
    case PGBA_ERR_UNSUPPORTED: return "pgba: configuration outside the implemented range";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "pgba: unknown error";
  }
}

int pgba_version(void) { return 100; }

int pgba_ba_workspace_bytes(int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows, int t0, int t1,
                            int64_t batch, size_t* bytes) {
  if (!bytes) return PGBA_ERR_NULL;
  if (n_edges < 0 || n_pose_rows <= 0 || n_patch_rows <= 0 || t1 < t0 || batch <= 0) return PGBA_ERR_SHAPE;
  const Layout L = make_layout(n_edges, n_pose_rows, n_patch_rows, t1 - t0, batch, pick_pc(n_edges, batch, n_pose_rows, n_patch_rows));
  *bytes = total_bytes(L, batch);
  return PGBA_OK;
}

static int prepare(Problem& pb, float* poses, float* patches, const float* intrinsics, const float* target,
                   const float* weight, const float* lmbda, const int64_t* ii, const int64_t* jj, const int64_t* kk,
                   const int32_t* n_edges_dev, const pgba_strides* strides, int64_t batch, int64_t n_edges,
                   int64_t n_pose_rows, int64_t n_patch_rows, int P, int t0, int t1, void* workspace,
                   size_t workspace_bytes) {
  int rc = check_common(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, n_edges, n_pose_rows,
                        n_patch_rows, P, t0, t1);
  if (rc) return rc;
  if (batch <= 0) return PGBA_ERR_SHAPE;
  if (batch > 1 && !strides) return PGBA_ERR_NULL;
  if (batch > 65535) return PGBA_ERR_UNSUPPORTED;
  if (!workspace || ((uintptr_t)workspace & 255)) return PGBA_ERR_WORKSPACE;
  pb = make_problem(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, n_edges_dev, strides, n_edges,
                    n_pose_rows, n_patch_rows, P, t0, t1, workspace, batch);
  if (total_bytes(pb.L, batch) > workspace_bytes) return PGBA_ERR_WORKSPACE;
  if (t1 > t0 && !solve_supported(t1 - t0)) return PGBA_ERR_UNSUPPORTED;
  return PGBA_OK;
}

int pgba_ba_solve_batched(float* poses, float* patches, const float* intrinsics, const float* target,
                          const float* weight, const float* lmbda, const int64_t* ii, const int64_t* jj,
                          const int64_t* kk, const int32_t* n_edges_dev, const pgba_strides* strides, int64_t batch,
                          int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows, int P, int ppf, int t0, int t1,
                          int iterations, int eff_impl, void* workspace, size_t workspace_bytes,
                          pgba_stream_t stream) {
  (void)ppf; (void)eff_impl;
  if (iterations < 0) return PGBA_ERR_SHAPE;
  Problem pb;
  int rc = prepare(pb, poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, n_edges_dev, strides, batch,
                   n_edges, n_pose_rows, n_patch_rows, P, t0, t1, workspace, workspace_bytes);
  if (rc) return rc;
  if (iterations == 0 || n_edges == 0) return PGBA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = clear_workspace(pb, batch, s);
  if (e != cudaSuccess) return (int)e;
  e = launch_planned(pb, batch, iterations, s);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}

int pgba_ba_solve_profiled(float* poses, float* patches, const float* intrinsics, const float* target,
                           const float* weight, const float* lmbda, const int64_t* ii, const int64_t* jj,
                           const int64_t* kk, const pgba_strides* strides, int64_t batch, int64_t n_edges,
                           int64_t n_pose_rows, int64_t n_patch_rows, int P, int t0, int t1, int iterations,
                           void* workspace, size_t workspace_bytes, pgba_stream_t stream, float* stage_ms) {
  if (!stage_ms) return PGBA_ERR_NULL;
  if (iterations <= 0 || iterations > 16 || n_edges == 0) return PGBA_ERR_SHAPE;
  Problem pb;
  int rc = prepare(pb, poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, nullptr, strides, batch,
                   n_edges, n_pose_rows, n_patch_rows, P, t0, t1, workspace, workspace_bytes);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int nev = 2 + 4 * iterations;
  cudaEvent_t ev[2 + 4 * 16];
  for (int i = 0; i < nev; ++i) cudaEventCreate(&ev[i]);
  cudaEventRecord(ev[0], s);
  cudaError_t e = clear_workspace(pb, batch, s);
  launch_plan(pb, batch, s);
  cudaEventRecord(ev[1], s);
  for (int it = 0; it < iterations && e == cudaSuccess; ++it) e = launch_iteration(pb, batch, s, ev + 2 + 4 * it, it == 0, it + 1 < iterations);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) {
    cudaEventElapsedTime(&stage_ms[0], ev[0], ev[1]);
    for (int it = 0; it < iterations; ++it)
      for (int k = 0; k < 3; ++k) cudaEventElapsedTime(&stage_ms[1 + 3 * it + k], ev[2 + 4 * it + k], ev[2 + 4 * it + k + 1]);
  }
  for (int i = 0; i < nev; ++i) cudaEventDestroy(ev[i]);
  return (int)e;
}

long long pgba_launch_count(void) { return launch_count(); }

#ifdef PGBA_LIN_TIMING
void pgba_debug_lin_timestamps(long long* out32) { pgba::lin_timestamps(out32); }
void pgba_debug_cta_timestamps(unsigned long long* out6144) { pgba::cta_timestamps(out6144); }
void pgba_debug_plan_cta_timestamps(unsigned long long* out3072) { pgba::plan_cta_timestamps(out3072); }
#endif
#ifdef PGBA_PLAN_TIMING
void pgba_debug_plan_timestamps(unsigned long long* out16) { pgba::plan_timestamps(out16); }
#endif

int pgba_ba_solve(float* poses, float* patches, const float* intrinsics, const float* target, const float* weight,
                  const float* lmbda, const int64_t* ii, const int64_t* jj, const int64_t* kk, int64_t n_edges,
                  int64_t n_pose_rows, int64_t n_patch_rows, int P, int ppf, int t0, int t1, int iterations,
                  int eff_impl, void* workspace, size_t workspace_bytes, pgba_stream_t stream) {
  pgba_strides st{};
  return pgba_ba_solve_batched(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, nullptr, &st, 1,
                               n_edges, n_pose_rows, n_patch_rows, P, ppf, t0, t1, iterations, eff_impl, workspace,
                               workspace_bytes, stream);
}

// ---- host-buffer entry point -------------------------------------------------------------------------------------
namespace {
struct HostStage { size_t poses, patches, intr, target, weight, lmbda, ii, jj, kk, total; };
HostStage host_stage(int64_t E, int64_t F, int64_t K, int P, int index_bits = 64) {
  HostStage h{};
  size_t o = 0;
  h.poses = o;   o = align256(o + sizeof(float) * 7 * (size_t)F);
  h.patches = o; o = align256(o + sizeof(float) * 3 * P * P * (size_t)K);
  h.intr = o;    o = align256(o + sizeof(float) * 4 * (size_t)F);
  h.target = o;  o = align256(o + sizeof(float) * 2 * (size_t)E);
  h.weight = o;  o = align256(o + sizeof(float) * 2 * (size_t)E);
  h.lmbda = o;   o = align256(o + sizeof(float));
  const size_t ib = index_bits == 32 ? sizeof(int32_t) : sizeof(int64_t);
  h.ii = o;      o = align256(o + ib * (size_t)E);
  h.jj = o;      o = align256(o + ib * (size_t)E);
  h.kk = o;      o = align256(o + ib * (size_t)E);
  h.total = o;
  return h;
}
// fork / join events: one pair per (device, main stream), created on first use and never destroyed (they live as long as
// the library).  Callers on different streams therefore never share an event, so their record / wait pairs cannot
// interleave; two host threads driving the SAME stream concurrently are the caller's race, as with any stream.
struct StageEvents { int dev; cudaStream_t stream; cudaEvent_t ev[2]; };
cudaEvent_t* stage_events(cudaStream_t s) {
  static std::mutex mu;
  static std::vector<StageEvents*> cache;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  for (StageEvents* c : cache)
    if (c->dev == dev && c->stream == s) return c->ev;
  StageEvents* c = new StageEvents{dev, s, {nullptr, nullptr}};
  if (cudaEventCreateWithFlags(&c->ev[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev[1], cudaEventDisableTiming) != cudaSuccess) {
    delete c;
    return nullptr;
  }
  cache.push_back(c);
  return c->ev;
}
}  // namespace

int pgba_ba_host_staging_bytes(int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows, int P, size_t* bytes) {
  if (!bytes) return PGBA_ERR_NULL;
  if (n_edges < 0 || n_pose_rows <= 0 || n_patch_rows <= 0 || P < 2) return PGBA_ERR_SHAPE;
  *bytes = host_stage(n_edges, n_pose_rows, n_patch_rows, P).total;
  return PGBA_OK;
}

int pgba_ba_host_arena_offsets(int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows, int P, size_t* offsets9,
                               size_t* bytes) {
  if (!offsets9 || !bytes) return PGBA_ERR_NULL;
  if (n_edges < 0 || n_pose_rows <= 0 || n_patch_rows <= 0 || P < 2) return PGBA_ERR_SHAPE;
  const HostStage h = host_stage(n_edges, n_pose_rows, n_patch_rows, P);
  const size_t o[9] = {h.poses, h.patches, h.intr, h.target, h.weight, h.lmbda, h.ii, h.jj, h.kk};
  for (int i = 0; i < 9; ++i) offsets9[i] = o[i];
  *bytes = h.total;
  return PGBA_OK;
}

int pgba_ba_host_arena_offsets_i32(int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows, int P, size_t* offsets9,
                                   size_t* bytes) {
  if (!offsets9 || !bytes) return PGBA_ERR_NULL;
  if (n_edges < 0 || n_pose_rows <= 0 || n_patch_rows <= 0 || P < 2) return PGBA_ERR_SHAPE;
  const HostStage h = host_stage(n_edges, n_pose_rows, n_patch_rows, P, 32);
  const size_t o[9] = {h.poses, h.patches, h.intr, h.target, h.weight, h.lmbda, h.ii, h.jj, h.kk};
  for (int i = 0; i < 9; ++i) offsets9[i] = o[i];
  *bytes = h.total;
  return PGBA_OK;
}

static int solve_host_impl(float* poses_h, float* patches_h, const float* intrinsics_h, const float* target_h,
                           const float* weight_h, const float* lmbda_h, const void* ii_h, const void* jj_h,
                           const void* kk_h, int index_bits, int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows,
                           int P, int t0, int t1, int iterations, void* staging, size_t staging_bytes, void* workspace,
                           size_t workspace_bytes, pgba_stream_t stream, pgba_stream_t aux_stream) {
  if (iterations < 0) return PGBA_ERR_SHAPE;
  if (!staging || ((uintptr_t)staging & 255)) return PGBA_ERR_WORKSPACE;
  int rc = check_common(poses_h, patches_h, intrinsics_h, target_h, weight_h, lmbda_h, ii_h, jj_h, kk_h, n_edges,
                        n_pose_rows, n_patch_rows, P, t0, t1);
  if (rc) return rc;
  const HostStage h = host_stage(n_edges, n_pose_rows, n_patch_rows, P, index_bits);
  if (h.total > staging_bytes) return PGBA_ERR_WORKSPACE;
  const size_t ib = index_bits == 32 ? 4 : 8;
  char* sb = (char*)staging;
  float* d_poses = (float*)(sb + h.poses);
  float* d_patches = (float*)(sb + h.patches);
  Problem pb;
  pgba_strides st{};
  rc = prepare(pb, d_poses, d_patches, (const float*)(sb + h.intr), (const float*)(sb + h.target),
               (const float*)(sb + h.weight), (const float*)(sb + h.lmbda), (const int64_t*)(sb + h.ii),
               (const int64_t*)(sb + h.jj), (const int64_t*)(sb + h.kk), nullptr, &st, 1, n_edges, n_pose_rows,
               n_patch_rows, P, t0, t1, workspace, workspace_bytes);
  if (rc) return rc;
  pb.idx32 = index_bits == 32 ? 1 : 0;
  if (iterations == 0 || n_edges == 0) return PGBA_OK;
  cudaStream_t s = (cudaStream_t)stream, a = (cudaStream_t)aux_stream;
  cudaEvent_t* ev = stage_events(s);
  if (!ev) return PGBA_ERR_UNSUPPORTED;
  const size_t E = (size_t)n_edges;
  const bool fork = a != s;
  cudaError_t e = cudaSuccess;
#define PGBA_TRY(x) do { if (e == cudaSuccess) e = (x); } while (0)
  if (fork) {                                  // aux joins the capture / ordering of the main stream
    PGBA_TRY(cudaEventRecord(ev[0], s));
    PGBA_TRY(cudaStreamWaitEvent(a, ev[0], 0));
  }
  // arena mode: the caller's tensors are views of ONE host allocation laid out like the staging buffer
  // (pgba_ba_host_arena_offsets): two uploads (indices | everything else) and one download instead of nine + two
  const uintptr_t hb0 = (uintptr_t)poses_h - h.poses;              // integer arithmetic: the tensors may be unrelated
  const bool arena = (uintptr_t)patches_h == hb0 + h.patches && (uintptr_t)intrinsics_h == hb0 + h.intr &&
                     (uintptr_t)target_h == hb0 + h.target && (uintptr_t)weight_h == hb0 + h.weight &&
                     (uintptr_t)lmbda_h == hb0 + h.lmbda && (uintptr_t)ii_h == hb0 + h.ii &&
                     (uintptr_t)jj_h == hb0 + h.jj && (uintptr_t)kk_h == hb0 + h.kk;
  const char* hb = (const char*)hb0;
  if (arena) {
    PGBA_TRY(cudaMemcpyAsync(sb + h.ii, hb + h.ii, h.total - h.ii, cudaMemcpyHostToDevice, s));
    PGBA_TRY(cudaMemcpyAsync(sb, hb, h.ii, cudaMemcpyHostToDevice, a));
  } else {
    PGBA_TRY(cudaMemcpyAsync(sb + h.ii, ii_h, ib * E, cudaMemcpyHostToDevice, s));
    PGBA_TRY(cudaMemcpyAsync(sb + h.jj, jj_h, ib * E, cudaMemcpyHostToDevice, s));
    PGBA_TRY(cudaMemcpyAsync(sb + h.kk, kk_h, ib * E, cudaMemcpyHostToDevice, s));
    PGBA_TRY(cudaMemcpyAsync(sb + h.target, target_h, 8 * E, cudaMemcpyHostToDevice, a));
    PGBA_TRY(cudaMemcpyAsync(sb + h.weight, weight_h, 8 * E, cudaMemcpyHostToDevice, a));
    PGBA_TRY(cudaMemcpyAsync(d_patches, patches_h, sizeof(float) * 3 * P * P * (size_t)n_patch_rows, cudaMemcpyHostToDevice, a));
    PGBA_TRY(cudaMemcpyAsync(d_poses, poses_h, sizeof(float) * 7 * (size_t)n_pose_rows, cudaMemcpyHostToDevice, a));
    PGBA_TRY(cudaMemcpyAsync(sb + h.intr, intrinsics_h, sizeof(float) * 4, cudaMemcpyHostToDevice, a));
    PGBA_TRY(cudaMemcpyAsync(sb + h.lmbda, lmbda_h, sizeof(float), cudaMemcpyHostToDevice, a));
  }
  // from here on the auxiliary stream is forked off the main one: whatever fails, it is joined again before returning
  // (an unjoined fork would invalidate a stream capture in progress)
  PGBA_TRY(clear_workspace(pb, 1, s));
  if (e == cudaSuccess) launch_plan(pb, 1, s);
  if (fork) {
    cudaError_t ej = cudaEventRecord(ev[1], a);
    if (ej == cudaSuccess) ej = cudaStreamWaitEvent(s, ev[1], 0);
    if (e == cudaSuccess) e = ej;
  }
  if (e != cudaSuccess) return (int)e;
  for (int it = 0; it < iterations && e == cudaSuccess; ++it) e = launch_iteration(pb, 1, s, nullptr, it == 0, it + 1 < iterations);
  if (arena) {
    PGBA_TRY(cudaMemcpyAsync((char*)hb, sb, h.intr, cudaMemcpyDeviceToHost, s));          // poses | patches
  } else {
    PGBA_TRY(cudaMemcpyAsync(poses_h, d_poses, sizeof(float) * 7 * (size_t)n_pose_rows, cudaMemcpyDeviceToHost, s));
    PGBA_TRY(cudaMemcpyAsync(patches_h, d_patches, sizeof(float) * 3 * P * P * (size_t)n_patch_rows, cudaMemcpyDeviceToHost, s));
  }
#undef PGBA_TRY
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}

int pgba_ba_solve_host(float* poses_h, float* patches_h, const float* intrinsics_h, const float* target_h,
                       const float* weight_h, const float* lmbda_h, const int64_t* ii_h, const int64_t* jj_h,
                       const int64_t* kk_h, int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows, int P, int ppf,
                       int t0, int t1, int iterations, int eff_impl, void* staging, size_t staging_bytes,
                       void* workspace, size_t workspace_bytes, pgba_stream_t stream, pgba_stream_t aux_stream) {
  (void)ppf; (void)eff_impl;
  return solve_host_impl(poses_h, patches_h, intrinsics_h, target_h, weight_h, lmbda_h, ii_h, jj_h, kk_h, 64, n_edges,
                         n_pose_rows, n_patch_rows, P, t0, t1, iterations, staging, staging_bytes, workspace, workspace_bytes,
                         stream, aux_stream);
}

int pgba_ba_solve_host_i32(float* poses_h, float* patches_h, const float* intrinsics_h, const float* target_h,
                           const float* weight_h, const float* lmbda_h, const int32_t* ii_h, const int32_t* jj_h,
                           const int32_t* kk_h, int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows, int P, int ppf,
                           int t0, int t1, int iterations, int eff_impl, void* staging, size_t staging_bytes,
                           void* workspace, size_t workspace_bytes, pgba_stream_t stream, pgba_stream_t aux_stream) {
  (void)ppf; (void)eff_impl;
  return solve_host_impl(poses_h, patches_h, intrinsics_h, target_h, weight_h, lmbda_h, ii_h, jj_h, kk_h, 32, n_edges,
                         n_pose_rows, n_patch_rows, P, t0, t1, iterations, staging, staging_bytes, workspace, workspace_bytes,
                         stream, aux_stream);
}

int pgba_ba_linearize_debug(const float* poses, const float* patches, const float* intrinsics, const float* target,
                            const float* weight, const float* lmbda, const int64_t* ii, const int64_t* jj,
                            const int64_t* kk, int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows, int P,
                            int t0, int t1, int with_schur, float* S, float* y, float* dX, int64_t* patch_ids,
                            float* C, float* u, float* Q, float* dZ, int32_t* n_unique, int32_t* status,
                            void* workspace, size_t workspace_bytes, pgba_stream_t stream) {
  pgba_strides st{};
  Problem pb;
  int rc = prepare(pb, (float*)poses, (float*)patches, intrinsics, target, weight, lmbda, ii, jj, kk, nullptr, &st, 1,
                   n_edges, n_pose_rows, n_patch_rows, P, t0, t1, workspace, workspace_bytes);
  if (rc) return rc;
  pb.with_schur = with_schur ? 1 : 0;
  pb.plan_cache = 0;
  pb.apply = 0;                                  // nothing is modified; S, y stay in the workspace for the export
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = clear_workspace(pb, 1, s);
  if (e != cudaSuccess) return (int)e;
  launch_plan(pb, 1, s);
  launch_linearize(pb, 1, s, false, true);
  launch_k(export_debug_kernel, dim3(256), dim3(256), 0, s, pb, S, y, nullptr, patch_ids, C, u, Q, nullptr, n_unique, nullptr);   // before the solve
  count_launch();
  launch_solve(pb, 1, s);
  launch_update(pb, 1, s);
  launch_k(export_debug_kernel, dim3(64), dim3(256), 0, s, pb, nullptr, nullptr, dX, nullptr, nullptr, nullptr, nullptr, dZ, nullptr, status);
  count_launch();
  return (int)cudaGetLastError();
}

const int32_t* pgba_ba_status_ptr(const void* workspace, int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows,
                                  int t0, int t1, int64_t batch, int64_t b) {
  if (!workspace || b < 0 || b >= batch || n_pose_rows <= 0 || n_patch_rows <= 0 || t1 < t0) return nullptr;
  const Layout L = make_layout(n_edges, n_pose_rows, n_patch_rows, t1 - t0, batch, pick_pc(n_edges, batch, n_pose_rows, n_patch_rows));
  const WinHeader* h = (const WinHeader*)((const char*)workspace + (size_t)b * L.zero_bytes + L.z_hdr);
  return &h->status;
}

const int32_t* pgba_ba_plan_hit_ptr(const void* workspace, int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows,
                                    int t0, int t1, int64_t batch, int64_t b) {
  if (!workspace || b < 0 || b >= batch || n_pose_rows <= 0 || n_patch_rows <= 0 || t1 < t0) return nullptr;
  const Layout L = make_layout(n_edges, n_pose_rows, n_patch_rows, t1 - t0, batch, pick_pc(n_edges, batch, n_pose_rows, n_patch_rows));
  const WinHeader* h = (const WinHeader*)((const char*)workspace + (size_t)b * L.zero_bytes + L.z_hdr);
  return &h->plan_hit;
}

const int32_t* pgba_ba_order_ptr(const void* workspace, int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows,
                                 int t0, int t1, int64_t batch, int64_t b, int32_t* segments, int32_t* tile_capacity,
                                 const int32_t** header) {
  if (!workspace || b < 0 || b >= batch || n_pose_rows <= 0 || n_patch_rows <= 0 || t1 < t0) return nullptr;
  const Layout L = make_layout(n_edges, n_pose_rows, n_patch_rows, t1 - t0, batch, pick_pc(n_edges, batch, n_pose_rows, n_patch_rows));
  if (segments) *segments = L.nd_P;
  if (tile_capacity) *tile_capacity = L.nd_nt;
  if (L.nd_P == 0) return nullptr;
  const char* z = (const char*)workspace + (size_t)b * L.zero_bytes;
  if (header) *header = (const int32_t*)(z + L.z_nd);
  return (const int32_t*)(z + L.z_ndf) + 3 * n_pose_rows;            // lminv, lmax1, border, pos (ba_bignd.cu: nd_sys)
}

int pgba_reproject(const float* poses, const float* patches, const float* intrinsics, const int64_t* ii,
                   const int64_t* jj, const int64_t* kk, int64_t n_edges, int64_t n_pose_rows, int64_t n_patch_rows,
                   int P, int clamp_depth, float* coords, pgba_stream_t stream) {
  (void)n_pose_rows; (void)n_patch_rows;
  if (n_edges == 0) return PGBA_OK;
  if (!poses || !patches || !intrinsics || !ii || !jj || !kk || !coords) return PGBA_ERR_NULL;
  if (n_edges < 0 || P < 1) return PGBA_ERR_SHAPE;
  const int64_t total = n_edges * P * P;
  const int64_t blocks = (total + 255) / 256;
  if (blocks >= (int64_t)1 << 31) return PGBA_ERR_UNSUPPORTED;
  launch_k(reproject_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, poses, patches, intrinsics, ii, jj, kk,
                                                                      n_edges, P, clamp_depth, coords);
  count_launch();
  return (int)cudaGetLastError();
}

}  // extern "C"
