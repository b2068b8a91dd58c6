// Per-chunk part of the BA plan: compact patch list, sorted target-frame slots, cell table cells[p][s] -> edge,
// duplicates list.  Shared by plan_cells_kernel (batched windows, global BA) and by the first linearisation of a call in
// the single-window regime, which builds the tables of its chunk itself (linearize_kernel<.., .., true>): the tables are
// then already on the SM when the linearisation needs them and one kernel with its dependent round trips disappears.
#pragma once
#include "ba_common.cuh"

namespace pgba {

// In-place exclusive scan of a[0..n) by the whole block; returns the total.  scratch: >= 33 ints of shared memory.
__device__ __forceinline__ int block_exclusive_scan(int* a, int n, int* scratch) {
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


struct CellScratch {           // shared memory of one CTA
  unsigned pflag[PMAX / 32];
  int ppref[PMAX / 32 + 1];
  unsigned jflag[PGBA_MAX_POSE_ROWS / 32];
  int jpref[PGBA_MAX_POSE_ROWS / 32 + 1];
  int scratch[40];
  int has_dup;
  Chunk sch;                   // the finished chunk descriptor
};

// Rank of patch bit k / target frame j among the present ones (valid after build_chunk_cells, until the next call).
__device__ __forceinline__ int cell_patch_rank(const CellScratch& sc, int k) {
  return sc.ppref[k >> 5] + __popc(sc.pflag[k >> 5] & ((1u << (k & 31)) - 1u));
}
__device__ __forceinline__ int cell_slot_rank(const CellScratch& sc, int j) {
  return sc.jpref[j >> 5] + __popc(sc.jflag[j >> 5] & ((1u << (j & 31)) - 1u));
}

// All 256 threads of the CTA.  Returns false when the chunk is empty / rejected (sc.sch.n_patches == 0).  Ends with the
// writes of the cell table in flight: __syncthreads() before reading them.
__device__ __forceinline__ bool build_chunk_cells(const Problem& pb, const WinPtrs& wp, int c, CellScratch& sc) {
  const int tid = threadIdx.x, T = blockDim.x;
  const int nw = (pb.F + 31) / 32;
  const int t0 = pb.t0, t1 = pb.t1;
  __syncthreads();
  if (tid == 0) sc.sch = wp.chunks[c];
  if (tid < PMAX / 32) sc.pflag[tid] = 0;
  if (tid == 0) sc.has_dup = 0;
  for (int x = tid; x < nw; x += T) sc.jflag[x] = 0;
  __syncthreads();
  const int eb = sc.sch.edge_begin, ee = sc.sch.edge_end, kbase = sc.sch.kbase, fi = sc.sch.frame;
  // presence bits of the chunk's patches / target frames.  All edges of a chunk hit the same few words, so the bits are
  // first OR-ed over the warp (one shared-memory atomic per distinct word and warp instead of one per edge: same-address
  // shared atomics serialise, 2 x 1700 of them per chunk was the bulk of this kernel on the batched shape)
  for (int pos0 = eb; pos0 < ee; pos0 += T) {                  // warp-uniform trip count
    const int pos = pos0 + tid;
    const bool in = pos < ee;
    int k = 0, j = 0;
    if (in) {
      const int4 rec = wp.perm[pos];
      k = rec.z - kbase; j = rec.y;
    }
    const int lane = tid & 31;
    const unsigned kbit = in ? 1u << (k & 31) : 0u, jbit = in ? 1u << (j & 31) : 0u;
#pragma unroll
    for (int wd = 0; wd < PMAX / 32; ++wd) {
      const unsigned m = __reduce_or_sync(0xffffffffu, (k >> 5) == wd ? kbit : 0u);
      if (m && lane == wd) atomicOr(&sc.pflag[wd], m);
    }
    const unsigned grp = __match_any_sync(0xffffffffu, in ? (j >> 5) : -1);
    const unsigned m = __reduce_or_sync(grp, jbit);
    if (in && lane == __ffs(grp) - 1) atomicOr(&sc.jflag[j >> 5], m);
  }
  __syncthreads();
  for (int x = tid; x < nw; x += T) sc.jpref[x] = __popc(sc.jflag[x]);
  if (tid == 0) {
    int run = 0;
    for (int x = 0; x < PMAX / 32; ++x) { sc.ppref[x] = run; run += __popc(sc.pflag[x]); }
    sc.ppref[PMAX / 32] = run;
  }
  __syncthreads();
  const int n_slots = block_exclusive_scan(sc.jpref, nw, sc.scratch);
  if (tid == 0) {
    const int n_patches = sc.ppref[PMAX / 32];
    auto rank_j = [&](int f) {   // number of present frames < f
      if (f <= 0) return 0;
      if (f >= pb.F) return n_slots;
      return sc.jpref[f >> 5] + __popc(sc.jflag[f >> 5] & ((1u << (f & 31)) - 1u));
    };
    const int first_free = rank_j(t0);
    const int n_free = max(rank_j(t1) - first_free, 0);
    const bool i_free = (fi >= t0 && fi < t1);
    const bool i_is_slot = (sc.jflag[fi >> 5] >> (fi & 31)) & 1u;
    int icol = -1, ncols = n_free;
    if (i_free) {
      if (i_is_slot) icol = rank_j(fi) - first_free;
      else { icol = n_free; ncols = n_free + 1; }
    }
    Chunk ch = sc.sch;
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
    sc.sch = ch;
    wp.chunks[c] = ch;
  }
  __syncthreads();
  const int n_patches = sc.sch.n_patches, ns = sc.sch.n_slots;
  if (n_patches == 0) return false;
  int* cells = wp.cells + sc.sch.cell_base;
  for (int x = tid; x < n_patches * ns; x += T) cells[x] = -1;
  for (int b = tid; b < PMAX; b += T)
    if ((sc.pflag[b >> 5] >> (b & 31)) & 1u)
      wp.kx[sc.sch.patch_base + sc.ppref[b >> 5] + __popc(sc.pflag[b >> 5] & ((1u << (b & 31)) - 1u))] = kbase + b;
  for (int f = tid; f < pb.F; f += T)
    if ((sc.jflag[f >> 5] >> (f & 31)) & 1u)
      wp.slots[sc.sch.slot_base + sc.jpref[f >> 5] + __popc(sc.jflag[f >> 5] & ((1u << (f & 31)) - 1u))] = f;
  __syncthreads();
  for (int pos = eb + tid; pos < ee; pos += T) {
    const int4 rec = wp.perm[pos];
    const int n = rec.x, k = rec.z - kbase, j = rec.y;
    const int p = sc.ppref[k >> 5] + __popc(sc.pflag[k >> 5] & ((1u << (k & 31)) - 1u));
    const int s = sc.jpref[j >> 5] + __popc(sc.jflag[j >> 5] & ((1u << (j & 31)) - 1u));
    const int old = atomicCAS(&cells[p * ns + s], -1, n);
    if (old != -1) {      // duplicated (patch, target frame) edge: handled by the slow path of the linearizer
      sc.has_dup = 1;
      const int d = atomicAdd(&wp.hdr->n_dups, 1);
      DupEdge de; de.chunk = c; de.p = p; de.s = s; de.n = n;
      wp.dups[d] = de;
    }
  }
  return true;
}

}  // namespace pgba
