// Tiled patch-correlation lookup for the production shape of CDV-SLAM (fp16 features, C <= 32 channels, P = 3,
// R = 3; reference: cdvslam/altcorr/correlation_kernel.cu:83-136, 193-233; call site cdvslam/slam.py:316-323).
//
// The per-edge kernels fetch the fmap2 neighbourhood of every edge from L2 (~10 KB per edge and level, ~0.7 GB per
// call at 38k edges).  Here the (edge, level) tasks are first binned by the fmap2 tile their windows fall into, then
// one CTA stages a tile (core 40 x 40 pixels + halo, all channels) in shared memory ONCE, transposed to
// [pixel][channel], and serves every task of the tile from it:
//   corr_classify_kernel  per task: 9 window origins -> 12 x 12 region -> tile id (or "slow" when the windows do not
//                         fit one region / tile), histogram of the tiles
//   corr_scan_kernel      offsets per tile + list of work items (tile, <= 64 tasks)
//   corr_scatter_kernel   task ids sorted by tile
//   corr_tile_kernel      per work item: tile -> shared memory; per task (one warp): the [9 patch pixels] x
//                         [12 x 16 region pixels] x [C] contraction on the tensor cores (mma.sync m16n8k16/k8, fp16
//                         inputs, fp32 accumulation), then window selection + bilinear blend + permute from the
//                         region volume in shared memory.  "Slow" tasks use the per-tap path (same arithmetic as
//                         corr_forward_kernel).
// Every task is computed independently of the binning order, so results are deterministic.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pcorr.h"

namespace pgba { void count_launch(); }

namespace pcorr_tiled {

constexpr int R = 3, D = 8, Do = 7, PP = 9;
constexpr int TC = 40;                 // tile core (pixels)
constexpr int HL = 6;                  // halo: region (12) centred in the core -> at most 6 pixels outside
constexpr int RG = 12;                 // region: RG x RG pixels hold the nine 8 x 8 windows
constexpr int TH = TC + 2 * HL;        // 52 tile rows staged in shared memory
constexpr int TW = 56;                 // tile columns: [tx0 - 2, tx0 + 54), tx0 = tile_x * TC - HL; multiple of 8 and
                                       // wide enough for the three 8-pixel n-tiles that cover any 12-wide region
constexpr int TXO = 2;                 // staged column 0 is map column tx0 - TXO (keeps 8-pixel groups 16 B aligned)
constexpr int PSH = (TH + 1) * TW;     // plane stride in halves (one spare row: conflict-free ldmatrix rows)
constexpr int TCH = 256;               // tasks per work item
constexpr int VS = RG * RG + 4;        // 148: row stride of the region volume
constexpr int NWARP = 8;

struct Level {
  const __half* fmap2;
  int H2, W2, ntx, nty, tile_base;
  float inv_scale;
};

struct Params {
  const __half* fmap1;
  Level lv[2];
  int nlev;
  const float* coords;
  const int64_t* us;
  const int64_t* vs;
  int B;
  int64_t E, K, F;
  int C;
  __half* out;
  int4* rec;          // per task: x0, y0 (region origin in the map), tile id, -
  int* tile_cnt;      // [ntiles + 1]   (last = slow tasks)   -- zeroed by the caller
  int* tile_off;      // [ntiles + 2]
  int* tile_cur;      // [ntiles + 1]
  int* sorted;        // [ntasks]
  int4* wi;           // work items: tile, begin, end, -
  int* n_wi;
  int ntiles;         // real tiles (the slow bin is index ntiles)
};

struct TaskGeom { int x0, y0; bool fits; int fx, fy; };

// lanes 0..8 hold the coordinates of the 9 patch pixels; returns the region origin and whether the 9 windows fit 12 x 12
__device__ __forceinline__ void task_coords(const Params& P, int64_t task, int& lev, int& b, int64_t& m) {
  lev = (int)(task % P.nlev);
  const int64_t be = task / P.nlev;
  b = (int)(be / P.E);
  m = be - (int64_t)b * P.E;
}

__device__ __forceinline__ int safe_floor(float v) {
  const float f = floorf(v);
  return (f > -1e6f && f < 1e6f) ? (int)f : -1000000;
}

// grid over tasks, one thread per task
__global__ void __launch_bounds__(256) corr_classify_kernel(Params P) {
  const int64_t ntask = (int64_t)P.B * P.E * P.nlev;
  const int64_t task = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (task >= ntask) return;
  int lev, b; int64_t m;
  task_coords(P, task, lev, b, m);
  const Level& lv = P.lv[lev];
  const float* cg = P.coords + ((int64_t)b * P.E + m) * 2 * PP;
  int xmin = 0x7fffffff, xmax = -0x7fffffff, ymin = 0x7fffffff, ymax = -0x7fffffff;
#pragma unroll
  for (int p = 0; p < PP; ++p) {
    const float x = (lev == 0) ? cg[p] : cg[p] * lv.inv_scale;
    const float y = (lev == 0) ? cg[PP + p] : cg[PP + p] * lv.inv_scale;
    const int fx = safe_floor(x), fy = safe_floor(y);
    xmin = min(xmin, fx); xmax = max(xmax, fx); ymin = min(ymin, fy); ymax = max(ymax, fy);
  }
  int x0 = xmin - R, y0 = ymin - R, flag = 0;
  int tile = P.ntiles;                                           // slow bin by default
  if ((xmax - xmin + D <= RG) && (ymax - ymin + D <= RG)) {
    // tile by the region centre, clamped to the map's tiles.  At the map border the region origin is clamped into the
    // tile's range: every pixel this drops is outside the map (zero), so the result is unchanged; the blend of such
    // tasks bounds-checks its taps against the region (flag).
    int tx = (x0 + HL) / TC, ty = (y0 + HL) / TC;
    if (x0 + HL < 0) tx = 0;
    if (y0 + HL < 0) ty = 0;
    tx = min(tx, lv.ntx - 1); ty = min(ty, lv.nty - 1);
    const int x0c = min(max(x0, tx * TC - HL), tx * TC - HL + TC);
    const int y0c = min(max(y0, ty * TC - HL), ty * TC - HL + TC);
    flag = (x0c != x0 || y0c != y0) ? 1 : 0;
    x0 = x0c; y0 = y0c;
    const int64_t jx = P.vs[m];
    tile = lv.tile_base + (int)((((int64_t)b * P.F + jx) * lv.nty + ty) * lv.ntx + tx);
  }
  P.rec[task] = make_int4(x0, y0, tile, flag);
  // warp-aggregated histogram: consecutive edges often share the target frame (hot level-1 tiles)
  const unsigned grp = __match_any_sync(__activemask(), tile);
  if ((int)(threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(&P.tile_cnt[tile], __popc(grp));
}

// one CTA: exclusive scan of the tile histogram, work-item list
__global__ void __launch_bounds__(1024) corr_scan_kernel(Params P) {
  __shared__ int s_part[1024];
  __shared__ int s_items[1024];
  const int tid = threadIdx.x, T = blockDim.x;
  const int n = P.ntiles + 1;
  const int per = (n + T - 1) / T;
  const int b0 = min(tid * per, n), b1 = min(b0 + per, n);
  int s = 0, it = 0;
  for (int i = b0; i < b1; ++i) {
    const int c = P.tile_cnt[i];
    s += c;
    it += (c + TCH - 1) / TCH;
  }
  s_part[tid] = s; s_items[tid] = it;
  __syncthreads();
  // Hillis-Steele inclusive scans over the per-thread partials
  for (int o = 1; o < T; o <<= 1) {
    const int a = (tid >= o) ? s_part[tid - o] : 0, c = (tid >= o) ? s_items[tid - o] : 0;
    __syncthreads();
    s_part[tid] += a; s_items[tid] += c;
    __syncthreads();
  }
  int run = s_part[tid] - s, irun = s_items[tid] - it;
  for (int i = b0; i < b1; ++i) {
    const int c = P.tile_cnt[i];
    P.tile_off[i] = run;
    P.tile_cur[i] = run;
    for (int k = 0; k < c; k += TCH) P.wi[irun++] = make_int4(i, run + k, run + min(k + TCH, c), 0);
    run += c;
  }
  if (tid == T - 1) { P.tile_off[n] = s_part[tid]; *P.n_wi = s_items[tid]; }
}

__global__ void __launch_bounds__(256) corr_scatter_kernel(Params P) {
  const int64_t ntask = (int64_t)P.B * P.E * P.nlev;
  const int64_t task = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (task >= ntask) return;
  const int tile = P.rec[task].z;
  const unsigned act = __activemask();
  const unsigned grp = __match_any_sync(act, tile);
  const int lane = threadIdx.x & 31, leader = __ffs(grp) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(&P.tile_cur[tile], __popc(grp));
  base = __shfl_sync(grp, base, leader);
  P.sorted[base + __popc(grp & ((1u << lane) - 1u))] = (int)task;
}

__device__ __forceinline__ void mma_k16(float d[4], const unsigned a[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_k8(float d[4], const unsigned a[2], unsigned b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(b0));
}

__device__ __forceinline__ unsigned pack2(__half lo, __half hi) {
  return (unsigned)__half_as_ushort(lo) | ((unsigned)__half_as_ushort(hi) << 16);
}

struct WarpScratch {
  float vol[PP][VS];
  float4 wgt[PP];                      // bilinear weights of pixel p: (1-dx)(1-dy), dx(1-dy), (1-dx)dy, dx dy
  float sx[PP], sy[PP];
  int wox[PP], woy[PP];
  int vbase[PP];                       // p * VS + woy * RG + wox
};

__device__ __forceinline__ void ldsm_x4_t(unsigned& r0, unsigned& r1, unsigned& r2, unsigned& r3, unsigned addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(unsigned& r0, unsigned& r1, unsigned addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

// dynamic smem: __half tile[C][TH + 1][TW] (natural layout, as in global memory) | WarpScratch[NWARP]
__global__ void __launch_bounds__(32 * NWARP, 1) corr_tile_kernel(Params P) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int C = P.C;
  __half* tile = reinterpret_cast<__half*>(smraw);
  WarpScratch* wsc = reinterpret_cast<WarpScratch*>(smraw + (((size_t)PSH * C * 2 + 127) & ~(size_t)127));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
  WarpScratch& S = wsc[warp];
  const unsigned tile_s = (unsigned)__cvta_generic_to_shared(tile);
  const int n_wi = *P.n_wi;
  // output o = lane + 32 t -> (p, yo, xo): packed p | (yo*RG + xo) << 8 | (yo*D + xo) << 16
  constexpr int NOUT = (Do * Do * PP + 31) / 32;
  int odec[NOUT];
#pragma unroll
  for (int t = 0; t < NOUT; ++t) {
    const int o = lane + 32 * t;
    const int p = o % PP, yo = (o / PP) % Do, xo = o / (PP * Do);
    odec[t] = p | ((yo * RG + xo) << 8) | ((yo * D + xo) << 16);
  }
  const bool has16 = C >= 16, has8 = (C & 8) != 0;             // C in {8, 16, 24}

  for (int w = blockIdx.x; w < n_wi; w += gridDim.x) {
    const int4 item = P.wi[w];
    const int tile_id = item.x;
    const bool slow = (tile_id == P.ntiles);
    int tx0 = 0, ty0 = 0;
    if (!slow) {
      const int lev = (P.nlev == 2 && tile_id >= P.lv[1].tile_base) ? 1 : 0;
      const Level& lv = P.lv[lev];
      int r = tile_id - lv.tile_base;
      const int tx = r % lv.ntx; r /= lv.ntx;
      const int ty = r % lv.nty; r /= lv.nty;                   // r = b * F + jx
      tx0 = tx * TC - HL; ty0 = ty * TC - HL;
      const int H2 = lv.H2, W2 = lv.W2;
      const int64_t plane = (int64_t)H2 * W2;
      const __half* f2g = lv.fmap2 + (int64_t)r * C * plane;
      __syncthreads();                                          // previous work item done with the tile
      // ---- stage the tile in its natural [c][y][x] layout: a warp copies row segments (28 pixel pairs, coalesced),
      //      four rows in flight per lane; zero outside the map
      const int gx = tx0 - TXO + 2 * lane;
      const bool lane_on = lane < TW / 2;
      const int nrows = C * TH;
      for (int r0 = warp * 4; r0 < nrows; r0 += NWARP * 4) {
        unsigned v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int row = r0 + k;
          const int c = row / TH, py = row - c * TH;
          const int gy = ty0 + py;
          v[k] = 0u;
          if (lane_on && row < nrows && gy >= 0 && gy < H2) {
            const __half* src = f2g + (int64_t)c * plane + (int64_t)gy * W2 + gx;
            if (gx >= 0 && gx + 1 < W2 && ((reinterpret_cast<uintptr_t>(src) & 3) == 0)) {
              v[k] = *reinterpret_cast<const unsigned*>(src);
            } else {
              const __half z = __float2half(0.f);
              v[k] = pack2((gx >= 0 && gx < W2) ? src[0] : z, (gx + 1 >= 0 && gx + 1 < W2) ? src[1] : z);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int row = r0 + k;
          const int c = row / TH, py = row - c * TH;
          if (lane_on && row < nrows) reinterpret_cast<unsigned*>(tile + (size_t)c * PSH + py * TW)[lane] = v[k];
        }
      }
      __syncthreads();
    }

    for (int t = item.y + warp; t < item.z; t += NWARP) {
      const int task = P.sorted[t];
      const int4 rec = P.rec[task];
      int tlev, b; int64_t m;
      task_coords(P, task, tlev, b, m);
      const Level& lv = P.lv[tlev];
      const int H2 = lv.H2, W2 = lv.W2;
      const int64_t plane = (int64_t)H2 * W2;
      const int64_t ix = P.us[m], jx = P.vs[m];
      const __half* f1g = P.fmap1 + ((int64_t)b * P.K + ix) * C * PP;
      const float* cg = P.coords + ((int64_t)b * P.E + m) * 2 * PP;
      __half* og = P.out + ((int64_t)b * P.E + m) * (int64_t)(Do * Do * PP) * P.nlev;
      __syncwarp();
      if (lane < PP) {
        const float x = (tlev == 0) ? cg[lane] : cg[lane] * lv.inv_scale;
        const float y = (tlev == 0) ? cg[PP + lane] : cg[PP + lane] * lv.inv_scale;
        S.sx[lane] = x; S.sy[lane] = y;
        const int wx = safe_floor(x) - R - rec.x, wy = safe_floor(y) - R - rec.y;
        S.wox[lane] = wx; S.woy[lane] = wy;
        S.vbase[lane] = lane * VS + wy * RG + wx;
        const float dx = x - floorf(x), dy = y - floorf(y);
        S.wgt[lane] = make_float4((1.f - dx) * (1.f - dy), dx * (1.f - dy), (1.f - dx) * dy, dx * dy);
      }
      __syncwarp();
      if (!slow) {
        // ---- A fragments: rows = patch pixels (9 of 16), k = channels; f1g is [C][9]
        unsigned a16[4] = {0u, 0u, 0u, 0u}, a8[2] = {0u, 0u};
        if (has16) {
          const int c = 2 * tig;
          a16[0] = pack2(f1g[c * PP + gid], f1g[(c + 1) * PP + gid]);
          a16[2] = pack2(f1g[(c + 8) * PP + gid], f1g[(c + 9) * PP + gid]);
          if (gid == 0) {                                       // row 8
            a16[1] = pack2(f1g[c * PP + 8], f1g[(c + 1) * PP + 8]);
            a16[3] = pack2(f1g[(c + 8) * PP + 8], f1g[(c + 9) * PP + 8]);
          }
        }
        if (has8) {
          const int c = (has16 ? 16 : 0) + 2 * tig;
          a8[0] = pack2(f1g[c * PP + gid], f1g[(c + 1) * PP + gid]);
          if (gid == 0) a8[1] = pack2(f1g[c * PP + 8], f1g[(c + 1) * PP + 8]);
        }
        // region origin in staged-tile coordinates; its columns are covered by three 8-pixel n-tiles from xa
        const int rx0 = rec.x - tx0 + TXO, ry0 = rec.y - ty0;
        const int xa = rx0 & ~7, xoff = rx0 - xa;               // xoff in [0, 8)
        // ldmatrix row addresses: lane l supplies row (l & 7) of matrix (l >> 3)
        //   k16, x4: matrices (ch 0-7, nt), (ch 8-15, nt), (ch 0-7, nt+1), (ch 8-15, nt+1)
        //   k8,  x4: matrices (ch k8+0-7, nt 0), (.., nt 1), (.., nt 2), (dummy = nt 2)
        const int mrow = lane & 7, mid = lane >> 3;
        const unsigned a16_01 = tile_s + 2u * (unsigned)(((mid & 1) * 8 + mrow) * PSH + xa + (mid >> 1) * 8);
        const unsigned a16_2 = tile_s + 2u * (unsigned)(((mid & 1) * 8 + mrow) * PSH + xa + 16);
        const unsigned a8_012 = tile_s + 2u * (unsigned)(((has16 ? 16 : 0) + mrow) * PSH + xa + (mid < 3 ? mid : 2) * 8);
        for (int r = 0; r < RG; ++r) {
          const unsigned rowoff = 2u * (unsigned)((ry0 + r) * TW);
          float d[3][4];
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) { d[nt][0] = d[nt][1] = d[nt][2] = d[nt][3] = 0.f; }
          if (has16) {
            unsigned b0, b1, b2, b3, b4, b5;
            ldsm_x4_t(b0, b1, b2, b3, a16_01 + rowoff);
            ldsm_x2_t(b4, b5, a16_2 + rowoff);
            mma_k16(d[0], a16, b0, b1);
            mma_k16(d[1], a16, b2, b3);
            mma_k16(d[2], a16, b4, b5);
          }
          if (has8) {
            unsigned b0, b1, b2, b3;
            ldsm_x4_t(b0, b1, b2, b3, a8_012 + rowoff);
            mma_k8(d[0], a8, b0);
            mma_k8(d[1], a8, b1);
            mma_k8(d[2], a8, b2);
          }
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) {
            const int col = nt * 8 + 2 * tig - xoff;            // region column of d[nt][0]
            if (col >= 0 && col < RG) { S.vol[gid][r * RG + col] = d[nt][0]; if (gid == 0) S.vol[8][r * RG + col] = d[nt][2]; }
            if (col + 1 >= 0 && col + 1 < RG) { S.vol[gid][r * RG + col + 1] = d[nt][1]; if (gid == 0) S.vol[8][r * RG + col + 1] = d[nt][3]; }
          }
        }
        __syncwarp();
        if (rec.w == 0) {
          const float* vol0 = &S.vol[0][0];
#pragma unroll
          for (int q = 0; q < NOUT; ++q) {
            const int o = lane + 32 * q;
            if (o >= Do * Do * PP) break;
            const int p = odec[q] & 0xff;
            const float4 wg = S.wgt[p];
            const float* v = vol0 + S.vbase[p] + ((odec[q] >> 8) & 0xff);
            const float res = wg.x * v[0] + wg.y * v[1] + wg.z * v[RG] + wg.w * v[RG + 1];
            og[(int64_t)o * P.nlev + tlev] = __float2half_rn(res);
          }
        } else {                                                // border task: taps outside the region are zero
          for (int o = lane; o < Do * Do * PP; o += 32) {
            const int p = o % PP, yo = (o / PP) % Do, xo = o / (PP * Do);
            const float xs = S.sx[p], ys = S.sy[p];
            const float dx = xs - floorf(xs), dy = ys - floorf(ys);
            const int cy = S.woy[p] + yo, cx = S.wox[p] + xo;
            auto tap = [&](int yy, int xx) { return (yy >= 0 && yy < RG && xx >= 0 && xx < RG) ? S.vol[p][yy * RG + xx] : 0.f; };
            const float res = (1.f - dx) * (1.f - dy) * tap(cy, cx) + dx * (1.f - dy) * tap(cy, cx + 1) +
                              (1.f - dx) * dy * tap(cy + 1, cx) + dx * dy * tap(cy + 1, cx + 1);
            og[(int64_t)o * P.nlev + tlev] = __float2half_rn(res);
          }
        }
      } else {
        // ---- per-tap path: same arithmetic as corr_forward_kernel
        const __half* f2g = lv.fmap2 + ((int64_t)b * P.F + jx) * C * plane;
        for (int o = lane; o < PP * D * D; o += 32) {
          const int p = o / (D * D), pos = o - p * (D * D);
          const int io = pos / D, jo = pos - io * D;
          const int fy = safe_floor(S.sy[p]), fx = safe_floor(S.sx[p]);
          float acc = 0.f;
          const int i1 = fy + (io - R), j1 = fx + (jo - R);
          if (fy > -1000000 && fx > -1000000 && i1 >= 0 && i1 < H2 && j1 >= 0 && j1 < W2) {
            const __half* src = f2g + (int64_t)i1 * W2 + j1;
            for (int c = 0; c < C; ++c) acc += __half2float(f1g[c * PP + p]) * __half2float(src[c * plane]);
          }
          S.vol[p][pos] = acc;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < NOUT; ++q) {
          const int o = lane + 32 * q;
          if (o >= Do * Do * PP) break;
          const int p = odec[q] & 0xff;
          const float xs = S.sx[p], ys = S.sy[p];
          const float dx = xs - floorf(xs), dy = ys - floorf(ys);
          const float* v = &S.vol[p][(odec[q] >> 16) & 0xff];
          const float res = (1.f - dx) * (1.f - dy) * v[0] + dx * (1.f - dy) * v[1] + (1.f - dx) * dy * v[D] +
                            dx * dy * v[D + 1];
          og[(int64_t)o * P.nlev + tlev] = __float2half_rn(res);
        }
      }
    }
  }
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct WsLayout { size_t rec, cnt, off, cur, sorted, wi, nwi, total; int ntiles; int64_t ntask; int64_t wi_cap; };

static WsLayout ws_layout(int nlev, int B, int64_t E, int64_t F, const int* H, const int* W) {
  WsLayout L{};
  L.ntask = (int64_t)B * E * nlev;
  int nt = 0;
  for (int l = 0; l < nlev; ++l) nt += (int)((int64_t)B * F * ((H[l] + TC - 1) / TC) * ((W[l] + TC - 1) / TC));
  L.ntiles = nt;
  L.wi_cap = L.ntask / TCH + nt + 2;
  size_t o = 0;
  L.cnt = o;    o = align256(o + 4 * (size_t)(nt + 1));       // zeroed by the caller: keep first
  L.nwi = o;    o = align256(o + 4);
  L.off = o;    o = align256(o + 4 * (size_t)(nt + 2));
  L.cur = o;    o = align256(o + 4 * (size_t)(nt + 1));
  L.rec = o;    o = align256(o + 16 * (size_t)(L.ntask > 0 ? L.ntask : 1));
  L.sorted = o; o = align256(o + 4 * (size_t)(L.ntask > 0 ? L.ntask : 1));
  L.wi = o;     o = align256(o + 16 * (size_t)L.wi_cap);
  L.total = o;
  return L;
}

}  // namespace pcorr_tiled

using namespace pcorr_tiled;

extern "C" {

int pcorr_tiled_supported(int C, int P, int radius, int dtype) {
  return (dtype == PCORR_F16 && P == 3 && radius == 3 && C >= 8 && C <= 24 && C % 8 == 0) ? 1 : 0;
}

int pcorr_tiled_workspace_bytes(int nlev, int B, int64_t E, int64_t F, int H0, int W0, int H1, int W1, size_t* bytes) {
  if (!bytes) return PCORR_ERR_NULL;
  if (nlev < 1 || nlev > 2 || B < 0 || E < 0 || F <= 0 || H0 <= 0 || W0 <= 0) return PCORR_ERR_SHAPE;
  const int H[2] = {H0, H1}, W[2] = {W0, W1};
  *bytes = ws_layout(nlev, B, E, F, H, W).total;
  return PCORR_OK;
}

int pcorr_forward_tiled(const void* fmap1, const void* fmap2_l0, const void* fmap2_l1, const float* coords,
                        const int64_t* ii, const int64_t* jj, int nlev, int B, int64_t E, int64_t K, int64_t F, int C,
                        int H0, int W0, int H1, int W1, int P, int radius, int dtype, void* out, void* workspace,
                        size_t workspace_bytes, pcorr_stream_t stream) {
  if (E == 0 || B == 0) return PCORR_OK;
  if (!fmap1 || !fmap2_l0 || (nlev == 2 && !fmap2_l1) || !coords || !ii || !jj || !out || !workspace) return PCORR_ERR_NULL;
  if (nlev < 1 || nlev > 2 || B < 0 || E < 0 || K <= 0 || F <= 0 || H0 <= 0 || W0 <= 0 || (nlev == 2 && (H1 <= 0 || W1 <= 0)))
    return PCORR_ERR_SHAPE;
  if (!pcorr_tiled_supported(C, P, radius, dtype)) return PCORR_ERR_UNSUPPORTED;
  const int H[2] = {H0, H1}, W[2] = {W0, W1};
  const WsLayout L = ws_layout(nlev, B, E, F, H, W);
  if (L.ntask >= ((int64_t)1 << 31) || (reinterpret_cast<uintptr_t>(workspace) & 255)) return PCORR_ERR_UNSUPPORTED;
  if (workspace_bytes < L.total) return PCORR_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  Params prm{};
  prm.fmap1 = (const __half*)fmap1;
  int base = 0;
  for (int l = 0; l < nlev; ++l) {
    Level& lv = prm.lv[l];
    lv.fmap2 = (const __half*)(l == 0 ? fmap2_l0 : fmap2_l1);
    lv.H2 = H[l]; lv.W2 = W[l];
    lv.ntx = (W[l] + TC - 1) / TC; lv.nty = (H[l] + TC - 1) / TC;
    lv.tile_base = base;
    lv.inv_scale = (l == 0) ? 1.f : 0.25f;
    base += (int)((int64_t)B * F * lv.nty * lv.ntx);
  }
  prm.nlev = nlev; prm.coords = coords; prm.us = ii; prm.vs = jj; prm.B = B; prm.E = E; prm.K = K; prm.F = F; prm.C = C;
  prm.out = (__half*)out;
  prm.rec = (int4*)(ws + L.rec); prm.tile_cnt = (int*)(ws + L.cnt); prm.tile_off = (int*)(ws + L.off);
  prm.tile_cur = (int*)(ws + L.cur); prm.sorted = (int*)(ws + L.sorted); prm.wi = (int4*)(ws + L.wi);
  prm.n_wi = (int*)(ws + L.nwi); prm.ntiles = L.ntiles;
  cudaError_t e = cudaMemsetAsync(ws + L.cnt, 0, L.off - L.cnt, s);       // histogram + work-item counter
  if (e != cudaSuccess) return (int)e;
  const unsigned gt = (unsigned)((L.ntask + 255) / 256);
  corr_classify_kernel<<<gt, 256, 0, s>>>(prm);
  pgba::count_launch();
  corr_scan_kernel<<<1, 1024, 0, s>>>(prm);
  pgba::count_launch();
  corr_scatter_kernel<<<gt, 256, 0, s>>>(prm);
  pgba::count_launch();
  const size_t smem = (((size_t)PSH * C * 2 + 127) & ~(size_t)127) + sizeof(WarpScratch) * NWARP;
  cudaFuncSetAttribute(corr_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int64_t grid = L.wi_cap < 148 ? L.wi_cap : 148;
  corr_tile_kernel<<<(unsigned)grid, 32 * NWARP, smem, s>>>(prm);
  pgba::count_launch();
  return (int)cudaGetLastError();
}

}  // extern "C"
