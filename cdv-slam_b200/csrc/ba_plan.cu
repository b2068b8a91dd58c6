// Graph analysis ("plan") of the patch-graph BA: turns the unordered (ii, jj, kk) edge list of the reference API
// into per-source-frame chunks with a dense [patch x target-frame] cell table.
//
// This replaces what the reference does with at::_unique(kk) (cdvslam/fastba/ba_cuda.cu:476-478) and, for
// eff_impl=True, with the EfficentE constructor on the CPU (cdvslam/fastba/block_e.cu:43-145: unique over frame
// pairs, patch_to_ku, index_tensor), without host synchronisation, sort or host tables.
//
//   plan_bucket_kernel   one CTA per window: counting sort of the edges by (source frame, 128-patch sub-range)
//   plan_cells_kernel    per chunk: compact patch list, sorted target-frame slots, cell table cells[p][s] = edge
#include "ba_common.cuh"

namespace pgba {

// In-place exclusive scan of a[0..n) by the whole block; returns the total.  scratch: >= 33 ints of shared memory.
__device__ int block_exclusive_scan(int* a, int n, int* scratch) {
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int per = (n + T - 1) / T;
  const int b = min(tid * per, n), e = min(b + per, n);
  int s = 0;
  for (int i = b; i < e; ++i) s += a[i];
  int x = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) scratch[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int w = (lane < (T >> 5)) ? scratch[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    scratch[lane] = w;
  }
  __syncthreads();
  int run = (wid > 0 ? scratch[wid - 1] : 0) + x - s;
  const int total = scratch[(T >> 5) - 1];
  for (int i = b; i < e; ++i) {
    int v = a[i];
    a[i] = run;
    run += v;
  }
  __syncthreads();
  return total;
}

// grid = (1, batch), block = 1024, dynamic smem = (3*F + 2*ch_max + 64) ints
__global__ void __launch_bounds__(1024, 1) plan_bucket_kernel(Problem pb) {
  extern __shared__ int sm[];
  const int w = blockIdx.y;
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  const int F = pb.F, K = pb.K;
  const int ch_max = (int)pb.L.ch_max;
  int* cnt = sm;                // [F]  count -> sub-chunk count -> chunk base
  int* kmin = cnt + F;          // [F]
  int* kmax = kmin + F;         // [F]
  int* ccnt = kmax + F;         // [ch_max] edges per chunk -> chunk edge offset
  int* ccur = ccnt + ch_max;    // [ch_max] fill cursor
  int* scratch = ccur + ch_max; // [64]
  const int tid = threadIdx.x, T = blockDim.x;
  const int64_t* ii = pb.ii + (int64_t)w * pb.st.ii;
  const int64_t* jj = pb.jj + (int64_t)w * pb.st.jj;
  const int64_t* kk = pb.kk + (int64_t)w * pb.st.kk;
  const int E = pb.n_edges_dev ? min((int)pb.E, max(pb.n_edges_dev[w], 0)) : (int)pb.E;

  if (tid == 0) {
    WinHeader h{};
    *wp.hdr = h;
  }
  for (int f = tid; f < F; f += T) { cnt[f] = 0; kmin[f] = 0x7fffffff; kmax[f] = -1; }
  __syncthreads();

  int bad = 0;
  for (int e = tid; e < E; e += T) {
    const int64_t i = ii[e], j = jj[e], k = kk[e];
    if (i < 0 || i >= F || j < 0 || j >= F || k < 0 || k >= K) { bad = 1; continue; }
    atomicAdd(&cnt[(int)i], 1);
    atomicMin(&kmin[(int)i], (int)k);
    atomicMax(&kmax[(int)i], (int)k);
  }
  if (bad) atomicOr(&wp.hdr->status, PGBA_ST_INDEX_RANGE);
  __syncthreads();
  for (int f = tid; f < F; f += T) cnt[f] = cnt[f] > 0 ? (kmax[f] - kmin[f]) / PMAX + 1 : 0;
  __syncthreads();
  int n_chunks = block_exclusive_scan(cnt, F, scratch);
  if (n_chunks > ch_max) {     // cannot happen for consistent sizes; keep the kernel memory-safe anyway
    if (tid == 0) atomicOr(&wp.hdr->status, PGBA_ST_CAPACITY);
    n_chunks = 0;
  }
  for (int c = tid; c < n_chunks; c += T) ccnt[c] = 0;
  __syncthreads();
  if (n_chunks > 0) {
    for (int e = tid; e < E; e += T) {
      const int64_t i = ii[e], j = jj[e], k = kk[e];
      if (i < 0 || i >= F || j < 0 || j >= F || k < 0 || k >= K) continue;
      atomicAdd(&ccnt[cnt[(int)i] + ((int)k - kmin[(int)i]) / PMAX], 1);
    }
  }
  __syncthreads();
  const int n_valid = block_exclusive_scan(ccnt, n_chunks, scratch);
  for (int c = tid; c < n_chunks; c += T) {
    ccur[c] = ccnt[c];
    Chunk ch{};
    ch.edge_begin = ccnt[c];
    ch.edge_end = (c + 1 < n_chunks) ? ccnt[c + 1] : n_valid;
    wp.chunks[c] = ch;
  }
  __syncthreads();
  for (int f = tid; f < F; f += T) {
    if (kmax[f] < 0) continue;
    const int nsub = (kmax[f] - kmin[f]) / PMAX + 1;
    for (int s = 0; s < nsub && cnt[f] + s < n_chunks; ++s) {
      wp.chunks[cnt[f] + s].frame = f;
      wp.chunks[cnt[f] + s].kbase = kmin[f] + s * PMAX;
    }
  }
  if (n_chunks > 0) {
    for (int e = tid; e < E; e += T) {
      const int64_t i = ii[e], j = jj[e], k = kk[e];
      if (i < 0 || i >= F || j < 0 || j >= F || k < 0 || k >= K) continue;
      const int pos = atomicAdd(&ccur[cnt[(int)i] + ((int)k - kmin[(int)i]) / PMAX], 1);
      wp.perm[pos] = e;
    }
  }
  if (tid == 0) {
    wp.hdr->n_chunks = n_chunks;
    wp.hdr->n_valid_edges = n_valid;
  }
}

// grid = (gx, batch), block = 256, static smem.  Chunks are taken round-robin by blockIdx.x.
__global__ void __launch_bounds__(256) plan_cells_kernel(Problem pb) {
  __shared__ unsigned pflag[PMAX / 32];
  __shared__ int ppref[PMAX / 32 + 1];
  __shared__ unsigned jflag[PGBA_MAX_POSE_ROWS / 32];
  __shared__ int jpref[PGBA_MAX_POSE_ROWS / 32 + 1];
  __shared__ int scratch[40];
  __shared__ Chunk sch;
  const int w = blockIdx.y, tid = threadIdx.x, T = blockDim.x;
  const WinPtrs wp = win_ptrs(pb.ws, pb.L, w);
  const int64_t* jj = pb.jj + (int64_t)w * pb.st.jj;
  const int64_t* kk = pb.kk + (int64_t)w * pb.st.kk;
  const int n_chunks = wp.hdr->n_chunks;
  const int nw = (pb.F + 31) / 32;
  const int t0 = pb.t0, t1 = pb.t1;

  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    __syncthreads();
    if (tid == 0) sch = wp.chunks[c];
    if (tid < PMAX / 32) pflag[tid] = 0;
    for (int x = tid; x < nw; x += T) jflag[x] = 0;
    __syncthreads();
    const int eb = sch.edge_begin, ee = sch.edge_end, kbase = sch.kbase, fi = sch.frame;
    for (int pos = eb + tid; pos < ee; pos += T) {
      const int n = wp.perm[pos];
      const int k = (int)kk[n] - kbase, j = (int)jj[n];
      atomicOr(&pflag[k >> 5], 1u << (k & 31));
      atomicOr(&jflag[j >> 5], 1u << (j & 31));
    }
    __syncthreads();
    for (int x = tid; x < nw; x += T) jpref[x] = __popc(jflag[x]);
    if (tid == 0) {
      int run = 0;
      for (int x = 0; x < PMAX / 32; ++x) { ppref[x] = run; run += __popc(pflag[x]); }
      ppref[PMAX / 32] = run;
    }
    __syncthreads();
    const int n_slots = block_exclusive_scan(jpref, nw, scratch);
    if (tid == 0) {
      const int n_patches = ppref[PMAX / 32];
      auto rank_j = [&](int f) {   // number of present frames < f
        if (f <= 0) return 0;
        if (f >= pb.F) return n_slots;
        return jpref[f >> 5] + __popc(jflag[f >> 5] & ((1u << (f & 31)) - 1u));
      };
      const int first_free = rank_j(t0);
      const int n_free = max(rank_j(t1) - first_free, 0);
      const bool i_free = (fi >= t0 && fi < t1);
      const bool i_is_slot = (jflag[fi >> 5] >> (fi & 31)) & 1u;
      int icol = -1, ncols = n_free;
      if (i_free) {
        if (i_is_slot) icol = rank_j(fi) - first_free;
        else { icol = n_free; ncols = n_free + 1; }
      }
      Chunk ch = sch;
      ch.n_patches = n_patches; ch.n_slots = n_slots; ch.first_free = first_free; ch.n_free = n_free;
      ch.icol = icol; ch.ncols = ncols;
      int st = 0;
      if (n_slots > SMAX) st |= PGBA_ST_TOO_MANY_SLOTS;
      if (!st) {
        ch.patch_base = atomicAdd(&wp.hdr->n_patches, n_patches);
        ch.slot_base = atomicAdd(&wp.hdr->n_slots, n_slots);
        ch.cell_base = atomicAdd(&wp.hdr->n_cells, n_patches * n_slots);
        ch.ecell_base = (pb.t1 > pb.t0) ? atomicAdd(&wp.hdr->n_ecells, n_patches * ncols) : 0;
        if ((int64_t)ch.patch_base + n_patches > pb.L.patch_max || (int64_t)ch.slot_base + n_slots > pb.L.slot_max ||
            (int64_t)ch.cell_base + (int64_t)n_patches * n_slots > pb.L.cell_cap ||
            (int64_t)ch.ecell_base + (int64_t)n_patches * ncols > pb.L.ecell_cap)
          st |= PGBA_ST_CAPACITY;
      }
      if (st) {
        atomicOr(&wp.hdr->status, st);
        ch.n_patches = 0; ch.n_slots = 0; ch.ncols = 0; ch.n_free = 0;
      }
      sch = ch;
      wp.chunks[c] = ch;
    }
    __syncthreads();
    const int n_patches = sch.n_patches, ns = sch.n_slots;
    if (n_patches == 0) continue;
    int* cells = wp.cells + sch.cell_base;
    for (int x = tid; x < n_patches * ns; x += T) cells[x] = -1;
    for (int b = tid; b < PMAX; b += T)
      if ((pflag[b >> 5] >> (b & 31)) & 1u)
        wp.kx[sch.patch_base + ppref[b >> 5] + __popc(pflag[b >> 5] & ((1u << (b & 31)) - 1u))] = kbase + b;
    for (int f = tid; f < pb.F; f += T)
      if ((jflag[f >> 5] >> (f & 31)) & 1u)
        wp.slots[sch.slot_base + jpref[f >> 5] + __popc(jflag[f >> 5] & ((1u << (f & 31)) - 1u))] = f;
    __syncthreads();
    for (int pos = eb + tid; pos < ee; pos += T) {
      const int n = wp.perm[pos];
      const int k = (int)kk[n] - kbase, j = (int)jj[n];
      const int p = ppref[k >> 5] + __popc(pflag[k >> 5] & ((1u << (k & 31)) - 1u));
      const int s = jpref[j >> 5] + __popc(jflag[j >> 5] & ((1u << (j & 31)) - 1u));
      const int old = atomicCAS(&cells[p * ns + s], -1, n);
      if (old != -1) {      // duplicated (patch, target frame) edge: handled by the slow path of the linearizer
        const int d = atomicAdd(&wp.hdr->n_dups, 1);
        DupEdge de; de.chunk = c; de.p = p; de.s = s; de.n = n;
        wp.dups[d] = de;
      }
    }
  }
}

void launch_plan(const Problem& pb, int64_t batch, cudaStream_t stream) {
  const size_t smem = sizeof(int) * (3 * (size_t)pb.F + 2 * (size_t)pb.L.ch_max + 64);
  cudaFuncSetAttribute(plan_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  plan_bucket_kernel<<<dim3(1, (unsigned)batch), 1024, smem, stream>>>(pb);
  count_launch();
  int gx = (int)(pb.L.ch_max < 148 * 4 ? pb.L.ch_max : 148 * 4);
  if (batch > 1) gx = (int)(pb.L.ch_max < 32 ? pb.L.ch_max : 32);
  plan_cells_kernel<<<dim3((unsigned)gx, (unsigned)batch), 256, 0, stream>>>(pb);
  count_launch();
}

size_t plan_bucket_smem_bytes(const Layout& L) {
  return sizeof(int) * (3 * (size_t)L.F + 2 * (size_t)L.ch_max + 64);
}

}  // namespace pgba
