// cuda_ba.neighbors on the device (reference: cdvslam/fastba/ba.cpp:59-97; caller net_cdv.py:102-107, every update).
//
// The reference groups the edges by ii (at::_unique on the GPU), copies everything to the host, stable-sorts every group
// by jj and links each edge to the previous / next edge of its group (-1 at the ends), then copies ix, jx back: a device
// synchronisation and two transfers per network update.  Here: ONE launch of a thread-block cluster (8 CTAs x 1024
// threads, hardware cluster barriers between the phases), no host round trip, bit-identical links.
//
//   P0  zero the bin counters; every thread keeps its first edges in registers
//   P1  bin = ii mod NB (NB a power of two: keys inside a range <= NB never share a bin); rank[e] = arrival order in the bin
//   P2  exclusive scan of the bin counts (every CTA scans its own copy in shared memory)
//   P3  scatter the edge ids into the bins
//   P4  per bin: order by (ii, jj, edge id) -- what the reference's stable_sort of a group in input order produces -- and
//       link consecutive edges with equal ii.  Bins of <= 32 edges: one thread, insertion sort in local memory (the normal
//       case: a patch has ~20 edges); larger bins: the whole CTA, rank by counting.
#include <cooperative_groups.h>

#include "ba_common.cuh"
#include "ba_cells.cuh"

namespace cg = cooperative_groups;

namespace pgba {

constexpr int NBR_CL = 8, NBR_T = 1024, NBR_KEEP = 6, NBR_SMALL = 32;
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

__global__ void __launch_bounds__(NBR_T, 1) neighbors_kernel(const int64_t* __restrict__ ii, const int64_t* __restrict__ jj,
                                                            int E, int nb, int* __restrict__ cnt, int* __restrict__ rank,
                                                            int* __restrict__ slots, int64_t* __restrict__ ix,
                                                            int64_t* __restrict__ jx) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ int s_off[];                    // [nb + 1]
  __shared__ int scratch[40];
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
  const int total = block_exclusive_scan(s_off, nb, scratch);
  if (tid == 0) s_off[nb] = total;
  __syncthreads();

  // ---- P3
#pragma unroll
  for (int q = 0; q < NBR_KEEP; ++q) {
    const int e = gt + q * GT;
    if (e < E) slots[s_off[(unsigned)ki[q] & mask] + rk_keep[q]] = e;
  }
  for (int e = gt + NBR_KEEP * GT; e < E; e += GT) slots[s_off[(unsigned)ii[e] & mask] + rank[e]] = e;
  cl.sync();

  // ---- P4a: small bins, one thread each
  for (int b = gt; b < nb; b += GT) {
    const int o = s_off[b], n = s_off[b + 1] - o;
    if (n == 0 || n > NBR_SMALL) continue;
    NbrKey k[NBR_SMALL];
    for (int x = 0; x < n; ++x) {
      const int e = __ldcg(&slots[o + x]);
      k[x].e = e; k[x].i = ii[e]; k[x].j = jj[e];
    }
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
  // ---- P4b: large bins (a group of more than 32 edges, or many keys folded into one bin), CTA rk takes every NBR_CL-th:
  //      position = number of smaller keys, ordered list written over `rank` (free since P3), then linked
  for (int b0 = 0; b0 < nb; b0 += NBR_T) {             // find them cooperatively, NBR_T bins at a time
    const int b = b0 + tid;
    const bool big = b < nb && (s_off[b + 1] - s_off[b]) > NBR_SMALL;
    if (!__syncthreads_or(big)) continue;
    for (int t = 0; t < NBR_T; ++t) {
      const int bb = b0 + t;
      if (bb >= nb) break;
      const int o = s_off[bb], n = s_off[bb + 1] - o;
      if (n <= NBR_SMALL || (bb % NBR_CL) != rk) continue;     // uniform over the CTA
      for (int a = tid; a < n; a += NBR_T) {
        NbrKey ka; ka.e = __ldcg(&slots[o + a]); ka.i = ii[ka.e]; ka.j = jj[ka.e];
        int pos = 0;
        for (int c = 0; c < n; ++c) {
          NbrKey kc; kc.e = __ldcg(&slots[o + c]); kc.i = ii[kc.e]; kc.j = jj[kc.e];
          pos += nbr_less(kc, ka) ? 1 : 0;
        }
        rank[o + pos] = ka.e;
      }
      __syncthreads();
      for (int a = tid; a < n; a += NBR_T) {
        const int e = rank[o + a];
        const long long gi = ii[e];
        int pe = -1, ne = -1;
        if (a > 0) { pe = rank[o + a - 1]; if (ii[pe] != gi) pe = -1; }
        if (a + 1 < n) { ne = rank[o + a + 1]; if (ii[ne] != gi) ne = -1; }
        ix[e] = pe; jx[e] = ne;
      }
      __syncthreads();
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
  *bytes = align256(4 * (size_t)nbr_bins(n_edges)) + 2 * align256(4 * e1);
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
  int* rank = (int*)(w + align256(4 * (size_t)nb));
  int* slots = (int*)(w + align256(4 * (size_t)nb) + align256(4 * (size_t)n_edges));
  const size_t smem = sizeof(int) * ((size_t)nb + 1);
  cudaFuncSetAttribute(neighbors_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, neighbors_kernel, ii, jj, (int)n_edges, nb, cnt, rank, slots, ix, jx);
  count_launch();
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}

}  // extern "C"
