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
// neighbors_link_kernel (grid over the bins, all SMs): per bin, order by (ii, jj, edge id) -- what the reference's
//   stable_sort of a group in input order produces -- and link consecutive edges with equal ii.  Bins of <= 32 edges: one
//   thread, insertion sort in local memory (the normal case: a patch has ~20 edges); larger bins: the whole CTA, rank by
//   counting.
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

// grid = nb / NBR_LT CTAs of NBR_LT threads; `tmp` (E ints) receives the ordered list of the large bins
__global__ void __launch_bounds__(NBR_LT) neighbors_link_kernel(const int64_t* __restrict__ ii, const int64_t* __restrict__ jj,
                                                               int nb, const int* __restrict__ off, const int* __restrict__ slots,
                                                               int* __restrict__ tmp, int64_t* __restrict__ ix,
                                                               int64_t* __restrict__ jx) {
  pdl_wait();
  pdl_trigger();
  const int tid = threadIdx.x;
  const int b = blockIdx.x * NBR_LT + tid;
  const int o = off[b], n = off[b + 1] - o;
  // ---- small bins, one thread each: all edge ids first, then all keys (independent loads), then the sort
  if (n > 0 && n <= NBR_SMALL) {
    NbrKey k[NBR_SMALL];
    for (int x = 0; x < n; ++x) k[x].e = slots[o + x];
    for (int x = 0; x < n; ++x) { k[x].i = ii[k[x].e]; k[x].j = jj[k[x].e]; }
    for (int x = 1; x < n; ++x) {                     // insertion sort
      const NbrKey v = k[x];
      int y = x - 1;
      while (y >= 0 && nbr_less(v, k[y])) { k[y + 1] = k[y]; --y; }
      k[y + 1] = v;
    }
    for (int x = 0; x < n; ++x) {
      ix[k[x].e] = (x > 0 && k[x - 1].i == k[x].i) ? (int64_t)k[x - 1].e : -1;
      jx[k[x].e] = (x + 1 < n && k[x + 1].i == k[x].i) ? (int64_t)k[x + 1].e : -1;
    }
  }
  // ---- large bins (a group of more than 32 edges, or many keys folded into one bin): the whole CTA, one bin after the
  //      other: position = number of smaller keys, ordered list written to `tmp`, then linked
  if (!__syncthreads_or(n > NBR_SMALL)) return;
  for (int t = 0; t < NBR_LT; ++t) {
    const int bb = blockIdx.x * NBR_LT + t;
    const int ob = off[bb], nn = off[bb + 1] - ob;
    if (nn <= NBR_SMALL) continue;                    // uniform over the CTA
    for (int a = tid; a < nn; a += NBR_LT) {
      NbrKey ka; ka.e = slots[ob + a]; ka.i = ii[ka.e]; ka.j = jj[ka.e];
      int pos = 0;
      for (int c = 0; c < nn; ++c) {
        NbrKey kc; kc.e = slots[ob + c]; kc.i = ii[kc.e]; kc.j = jj[kc.e];
        pos += nbr_less(kc, ka) ? 1 : 0;
      }
      tmp[ob + pos] = ka.e;
    }
    __syncthreads();
    for (int a = tid; a < nn; a += NBR_LT) {
      const int e = tmp[ob + a];
      const long long gi = ii[e];
      int pe = -1, ne = -1;
      if (a > 0) { pe = tmp[ob + a - 1]; if (ii[pe] != gi) pe = -1; }
      if (a + 1 < nn) { ne = tmp[ob + a + 1]; if (ii[ne] != gi) ne = -1; }
      ix[e] = pe; jx[e] = ne;
    }
    __syncthreads();
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
  e = launch_k(neighbors_link_kernel, dim3((unsigned)(nb / NBR_LT)), dim3(NBR_LT), 0, (cudaStream_t)stream, ii, jj, nb,
               (const int*)off, (const int*)slots, rank, ix, jx);
  count_launch();
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}

}  // extern "C"
