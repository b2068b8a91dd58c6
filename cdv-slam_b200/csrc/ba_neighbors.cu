// cuda_ba.neighbors on the device (reference: cdvslam/fastba/ba.cpp:59-97; caller net_cdv.py:102-107, every update).
//
// The reference groups the edges by ii (at::_unique on the GPU), copies everything to the host, stable-sorts every group
// by jj and links each edge to the previous / next edge of its group (-1 at the ends), then copies ix, jx back: a device
// synchronisation and two transfers per network update.  Here: two launches, no host round trip, bit-identical links.
//
// neighbors_bin_kernel (ONE thread-block cluster of 8 CTAs x 1024 threads, hardware cluster barriers between the phases):
//   P0  zero the bin counters; every thread keeps its first edges in registers
//   P1  bin = ii mod NB (NB a power of two: keys inside a range <= NB never share a bin); rank[e] = arrival order in the bin
//   P2  exclusive scan of the bin counts (every CTA scans its own copy in shared memory with warp shuffles; CTA 0 also
//       writes the offsets to global memory for the second kernel)
//   P3  scatter the edge ids into the bins
// neighbors_link_kernel (all SMs): one warp per bin, bins dealt round-robin over the warps; per bin, order by (ii, jj, edge
//   id) -- what the reference's stable_sort of a group in input order produces -- by rank counting over the lanes, and link
//   consecutive edges with equal ii.  Bins of > 32 edges: the same in strips of 32.
#include <cooperative_groups.h>

#include "ba_common.cuh"

namespace cg = cooperative_groups;

namespace pgba {

constexpr int NBR_CL = 8, NBR_T = 1024, NBR_KEEP = 6, NBR_SMALL = 32, NBR_LT = 128;
constexpr int NBR_BINS_MAX = 32768;           // 128 KB of shared memory for the scanned offsets

static int nbr_bins(int64_t E) {
  int nb = 1024;
  while (nb < NBR_BINS_MAX && nb < E) nb <<= 1;
  return nb;
}

struct NbrKey { long long i, j; int e; };
__device__ __forceinline__ bool nbr_less(const NbrKey& a, const NbrKey& b) {
  if (a.i != b.i) return a.i < b.i;
  if (a.j != b.j) return a.j < b.j;
  return a.e < b.e;
}

// In-place exclusive scan of a[0..n), n a multiple of 1024, by a CTA of 1024 threads: warp w owns the contiguous segment
// [w n/32, (w+1) n/32) and walks it 32 elements at a time with shuffle scans (coalesced, bank-conflict free).  Returns the total.
__device__ __forceinline__ int scan_1024(int* a, int n, int* wsum /* [33] shared */) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seg = n >> 5;
  int carry = 0;
  for (int i = warp * seg + lane; i < (warp + 1) * seg; i += 32) {
    const int v = a[i];
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    a[i] = carry + x - v;
    carry += __shfl_sync(0xffffffffu, x, 31);
  }
  if (lane == 0) wsum[warp] = carry;
  __syncthreads();
  if (warp == 0) {
    const int v = wsum[lane];
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    wsum[lane] = x - v;
    if (lane == 31) wsum[32] = x;
  }
  __syncthreads();
  const int base = wsum[warp];
  for (int i = warp * seg + lane; i < (warp + 1) * seg; i += 32) a[i] += base;
  __syncthreads();
  return wsum[32];
}

__global__ void __launch_bounds__(NBR_T, 1) neighbors_bin_kernel(const int64_t* __restrict__ ii, int E, int nb,
                                                                int* __restrict__ cnt, int* __restrict__ rank,
                                                                int* __restrict__ slots, int* __restrict__ off) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ int s_off[];                    // [nb]
  __shared__ int wsum[33];
  cg::cluster_group cl = cg::this_cluster();
  const int tid = threadIdx.x, rk = blockIdx.x;
  const int gt = rk * NBR_T + tid, GT = NBR_CL * NBR_T;
  const unsigned mask = (unsigned)nb - 1u;

  for (int b = gt; b < nb; b += GT) cnt[b] = 0;
  long long ki[NBR_KEEP];
#pragma unroll
  for (int q = 0; q < NBR_KEEP; ++q) {
    const int e = gt + q * GT;
    ki[q] = e < E ? ii[e] : 0;
  }
  cl.sync();

  // ---- P1
  int rk_keep[NBR_KEEP];
#pragma unroll
  for (int q = 0; q < NBR_KEEP; ++q) {
    const int e = gt + q * GT;
    rk_keep[q] = 0;
    if (e < E) rk_keep[q] = atomicAdd(&cnt[(unsigned)ki[q] & mask], 1);
  }
  for (int e = gt + NBR_KEEP * GT; e < E; e += GT) rank[e] = atomicAdd(&cnt[(unsigned)ii[e] & mask], 1);
  cl.sync();

  // ---- P2
  for (int b = tid; b < nb; b += NBR_T) s_off[b] = __ldcg(&cnt[b]);
  __syncthreads();
  const int total = scan_1024(s_off, nb, wsum);
  if (rk == 0) {
    for (int b = tid; b < nb; b += NBR_T) off[b] = s_off[b];
    if (tid == 0) off[nb] = total;
  }

  // ---- P3
#pragma unroll
  for (int q = 0; q < NBR_KEEP; ++q) {
    const int e = gt + q * GT;
    if (e < E) slots[s_off[(unsigned)ki[q] & mask] + rk_keep[q]] = e;
  }
  for (int e = gt + NBR_KEEP * GT; e < E; e += GT) slots[s_off[(unsigned)ii[e] & mask] + rank[e]] = e;
}

// One WARP per bin, bins dealt round-robin over all warps of the grid (the non-empty bins are consecutive -- bin = ii mod nb
// and patch ids are dense -- so a contiguous assignment would leave the work on a dozen SMs).  A bin of n <= 32 edges (the
// normal case: a patch has ~20 edges) is ordered by rank counting over the lanes: lane x holds edge x of the bin, reads the
// other lanes' (ii, jj, edge id) by shuffles and counts the smaller keys -- (ii, jj, edge id) is what the reference's
// stable_sort of a group in input order produces; the ordered list goes through a per-warp shared-memory line and every
// lane links its edge to the entries before / after it when they belong to the same ii.  Larger bins (a group of more
// than 32 edges, or many keys folded into one bin): the same by the warp in strips of 32, positions by counting over the
// whole bin, ordered list in `tmp`.
__global__ void __launch_bounds__(NBR_LT) neighbors_link_kernel(const int64_t* __restrict__ ii, const int64_t* __restrict__ jj,
                                                               int nb, const int* __restrict__ off, const int* __restrict__ slots,
                                                               int* __restrict__ tmp, int64_t* __restrict__ ix,
                                                               int64_t* __restrict__ jx) {
  pdl_wait();
  pdl_trigger();
  __shared__ int s_e[NBR_LT / 32][32];
  __shared__ long long s_i[NBR_LT / 32][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = blockIdx.x * (NBR_LT / 32) + wib, nwarps = gridDim.x * (NBR_LT / 32);
  for (int b0 = warp; b0 < nb; b0 += 32 * nwarps) {
    // the warp's next 32 bins (stride nwarps): one lane looks at one bin, the non-empty ones are then processed in turn
    const int myb = b0 + lane * nwarps;
    int o_l = 0, n_l = 0;
    if (myb < nb) { o_l = off[myb]; n_l = off[myb + 1] - o_l; }
    unsigned todo = __ballot_sync(0xffffffffu, n_l > 0);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int o = __shfl_sync(0xffffffffu, o_l, src), n = __shfl_sync(0xffffffffu, n_l, src);
      if (n <= NBR_SMALL) {
        int e = 0;
        long long ki = 0, kj = 0;
        if (lane < n) { e = slots[o + lane]; ki = ii[e]; kj = jj[e]; }
        int pos = 0;
        for (int y = 0; y < n; ++y) {
          const long long yi = __shfl_sync(0xffffffffu, ki, y), yj = __shfl_sync(0xffffffffu, kj, y);
          const int ye = __shfl_sync(0xffffffffu, e, y);
          pos += (yi < ki || (yi == ki && (yj < kj || (yj == kj && ye < e)))) ? 1 : 0;
        }
        __syncwarp();
        if (lane < n) { s_e[wib][pos] = e; s_i[wib][pos] = ki; }
        __syncwarp();
        if (lane < n) {
          ix[e] = (pos > 0 && s_i[wib][pos - 1] == ki) ? (int64_t)s_e[wib][pos - 1] : -1;
          jx[e] = (pos + 1 < n && s_i[wib][pos + 1] == ki) ? (int64_t)s_e[wib][pos + 1] : -1;
        }
      } else {
        for (int a = lane; a < n; a += 32) {
          NbrKey ka; ka.e = slots[o + a]; ka.i = ii[ka.e]; ka.j = jj[ka.e];
          int pos = 0;
          for (int c = 0; c < n; ++c) {
            NbrKey kc; kc.e = slots[o + c]; kc.i = ii[kc.e]; kc.j = jj[kc.e];
            pos += nbr_less(kc, ka) ? 1 : 0;
          }
          tmp[o + pos] = ka.e;
        }
        __threadfence_block();
        __syncwarp();
        for (int a = lane; a < n; a += 32) {
          const int e = tmp[o + a];
          const long long gi = ii[e];
          int pe = -1, ne = -1;
          if (a > 0) { pe = tmp[o + a - 1]; if (ii[pe] != gi) pe = -1; }
          if (a + 1 < n) { ne = tmp[o + a + 1]; if (ii[ne] != gi) ne = -1; }
          ix[e] = pe; jx[e] = ne;
        }
        __syncwarp();
      }
    }
  }
}

}  // namespace pgba

using namespace pgba;

extern "C" {

int pgba_neighbors_workspace_bytes(int64_t n_edges, size_t* bytes) {
  if (!bytes) return PGBA_ERR_NULL;
  if (n_edges < 0) return PGBA_ERR_SHAPE;
  const size_t e1 = (size_t)(n_edges > 0 ? n_edges : 1);
  // bin counters | scanned offsets [nb + 1] | rank [E] (reused for the ordered lists of large bins) | slots [E]
  *bytes = align256(4 * (size_t)nbr_bins(n_edges)) + align256(4 * ((size_t)nbr_bins(n_edges) + 1)) + 2 * align256(4 * e1);
  return PGBA_OK;
}

int pgba_neighbors(const int64_t* ii, const int64_t* jj, int64_t n_edges, int64_t* ix, int64_t* jx, void* workspace,
                   size_t workspace_bytes, pgba_stream_t stream) {
  if (n_edges == 0) return PGBA_OK;
  if (!ii || !jj || !ix || !jx) return PGBA_ERR_NULL;
  if (n_edges < 0) return PGBA_ERR_SHAPE;
  if (n_edges >= (int64_t)1 << 31) return PGBA_ERR_UNSUPPORTED;
  size_t need = 0;
  pgba_neighbors_workspace_bytes(n_edges, &need);
  if (!workspace || ((uintptr_t)workspace & 255) || workspace_bytes < need) return PGBA_ERR_WORKSPACE;
  const int nb = nbr_bins(n_edges);
  char* w = (char*)workspace;
  int* cnt = (int*)w;
  int* off = (int*)(w + align256(4 * (size_t)nb));
  int* rank = (int*)((char*)off + align256(4 * ((size_t)nb + 1)));
  int* slots = (int*)((char*)rank + align256(4 * (size_t)n_edges));
  const size_t smem = sizeof(int) * (size_t)nb;
  cudaFuncSetAttribute(neighbors_bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(NBR_CL);
  cfg.blockDim = dim3(NBR_T);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = NBR_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, neighbors_bin_kernel, ii, (int)n_edges, nb, cnt, rank, slots, off);
  count_launch();
  if (e != cudaSuccess) return (int)e;
  const int link_grid = nb / 32 < 148 * 8 ? nb / 32 : 148 * 8;      // one bin per lane and trip; >= 1 (nb >= 1024)
  e = launch_k(neighbors_link_kernel, dim3((unsigned)link_grid), dim3(NBR_LT), 0, (cudaStream_t)stream, ii, jj, nb,
               (const int*)off, (const int*)slots, rank, ix, jx);
  count_launch();
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}

}  // extern "C"
