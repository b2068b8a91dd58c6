// Patch correlation lookup, production shape (fp16 features, C in {24, 32}, P = 3, radius = 3), sm_100a:
// channel-last maps + TMA region tiles + tensor-core contraction + fused blend.
//
// Reference semantics: cdvslam/altcorr/correlation_kernel.cu:83-136 (window dot products, zero outside the map) and
// :193-233 (4-tap bilinear blend, permute); two-level form of cdvslam/slam.py:316-323.
//
//   to_nhwc_kernel     fmap2 [B*F, C, H, W] -> [B*F, H, W, C]: every map pixel becomes one contiguous 2C-byte record,
//                      so the (<= 12 x 12)-pixel region that contains the nine 8x8 windows of an edge is a dense box.
//   corr_tma_kernel    one WARP per edge (both pyramid levels), persistent CTAs.  Per (edge, level): the region box is
//                      fetched by ONE TMA tile load (cp.async.bulk.tensor.3d, out-of-map pixels zero-filled by the
//                      hardware, completion on a per-warp mbarrier), issued one half-task ahead.  The contraction
//                      D[region pixel, patch pixel] = sum_c region[px][c] * patch[p][c] runs on the tensor cores
//                      (mma.sync m16n8k16 / m16n8k8, fp16 in, fp32 accumulate; operands by ldmatrix).  Each 16-pixel
//                      tile of D is written back IN PLACE over the region records it was computed from (9 floats fit a
//                      2C-byte record), then the per-pixel window selection + 4-tap blend + permute are done from
//                      shared memory and the result is stored once, in the final layout (levels interleaved).
//                      Edges whose nine windows do not fit one region (strong zoom) take a per-tap path in the kernel.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pcorr.h"

namespace pgba { void count_launch(); }

namespace pcorr_tma {

constexpr int R = 3, D = 8, Do = 7, PP = 9;
constexpr int RG = 12, RPX = RG * RG;       // region: 12 x 12 pixels
constexpr int WARPS = 4;

template <int C> struct __align__(128) WarpSmem {
  unsigned char reg[2][RPX * C * 2];   // TMA destinations; the region volume overwrites the records in place
  __half a[16 * C];                    // patch-feature record of the unit, [c][p] as in fmap1 (9C halfs used)
  float4 wgt[2][12];                   // bilinear weights of pixel p: (1-dx)(1-dy), dx(1-dy), (1-dx)dy, dx dy
  int vbase[2][12];                    // float offset of pixel p's window origin inside the volume (+ p)
  __half stage[7 * 136];               // blended results of the unit, [xo][yo * 9 + p][level] with padded rows
  unsigned long long bar[2];
};

struct Params {
  const __half* fmap1;                 // [B, K, C, 3, 3]
  const __half* nhwc[2];               // [B*F, H, W, C] per level
  int H[2], W[2];
  const float* coords;                 // [B, E, 2, 3, 3]
  const int64_t* us; const int64_t* vs;
  int B; int64_t E, K, F;
  __half* out;                         // [B, E, 7, 7, 3, 3, NLEV]
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// box [RG * C / 2, RG, 1] at (xw, y, frame) of the 3-D map [W * C / 2 (32-bit words), H, B*F]: the x and channel axes
// are merged so that one region row is ONE contiguous 24C-byte chunk (a [C, W, H, F] box would be fetched as 144 chunks of
// 2C bytes and is bound by the TMA unit's chunk rate).  Words outside the map arrive as zeros; pixel borders are word
// borders, so this is exactly "taps outside the map contribute 0".
__device__ __forceinline__ void tma_load_region(uint32_t dst, uint64_t tm, uint32_t bar, int xw, int y, int f) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(xw), "r"(y), "r"(f) : "memory");
}

__device__ __forceinline__ void ldsm_x4(unsigned& r0, unsigned& r1, unsigned& r2, unsigned& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x2(unsigned& r0, unsigned& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_k16(float d[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0,
                                        unsigned b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_k8(float d[4], unsigned a0, unsigned a1, unsigned b0) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0));
}

__device__ __forceinline__ int safe_floor(float v) {     // far-out / non-finite -> "entirely outside the map"
  const float f = floorf(v);
  return (f > -1e6f && f < 1e6f) ? (int)f : -1000000;
}

// Channel-last copy of the frame maps of up to two pyramid levels in ONE launch: [B*F, C, HW] -> [B*F, HW, C].
// A block handles PXT consecutive pixels of one frame: the threads read pairs of adjacent pixels of a channel (4-byte
// loads, coalesced), the [512][C] tile is transposed through shared memory and written with contiguous 16-byte stores.
// blockIdx.x enumerates (level, frame, pixel block).  HW must be even (the host falls back to scalar loads otherwise).
struct NhwcJob { const __half* src; __half* dst; int HW; int blocks_per_frame; int first_block; };

template <int C, bool PAIR, int PXT>
__global__ void __launch_bounds__(256) to_nhwc_kernel(NhwcJob j0, NhwcJob j1) {
  __shared__ __align__(16) __half tile[PXT * C];
  const bool second = (int)blockIdx.x >= j1.first_block;
  const NhwcJob J = second ? j1 : j0;
  const int blk = (int)blockIdx.x - J.first_block;
  const int frame = blk / J.blocks_per_frame;
  const int px0 = (blk - frame * J.blocks_per_frame) * PXT;
  const int npx = min(PXT, J.HW - px0);
  const __half* s = J.src + (int64_t)frame * C * J.HW + px0;
  const int t = threadIdx.x;
  if (PAIR) {
    // item = (pixel pair, channel): consecutive threads read consecutive pixel pairs of one channel; the trip count is
    // a compile-time constant, so all loads of a thread are in flight together
    constexpr int ITEMS = (PXT / 2) * C / 256;
    static_assert((PXT / 2) * C % 256 == 0, "tile items must divide evenly over the 256 threads");
    __half2 v[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int x = t + 256 * i;
      const int c = x / (PXT / 2), pp = x - c * (PXT / 2);
      v[i] = (2 * pp < npx) ? *reinterpret_cast<const __half2*>(s + (int64_t)c * J.HW + 2 * pp) : __half2();
    }
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int x = t + 256 * i;
      const int c = x / (PXT / 2), pp = x - c * (PXT / 2);
      tile[(2 * pp) * C + c] = __low2half(v[i]);
      tile[(2 * pp + 1) * C + c] = __high2half(v[i]);
    }
  } else {
    for (int px = t; px < npx; px += 256)
#pragma unroll
      for (int c = 0; c < C; ++c) tile[px * C + c] = s[(int64_t)c * J.HW + px];
  }
  __syncthreads();
  uint4* d = reinterpret_cast<uint4*>(J.dst + ((int64_t)frame * J.HW + px0) * C);
  const uint4* t4 = reinterpret_cast<const uint4*>(tile);
  for (int x = t; x < npx * C / 8; x += 256) d[x] = t4[x];
}

// Wide maps (C = 128): 128-pixel tiles through shared memory with rows padded by one word (conflict-free 4-byte
// copy-out, 2-way conflicts on the 2-byte transposing stores); reads are 4-byte (pixel pairs) and coalesced per channel
// plane, writes are 4-byte words, 128 contiguous bytes per warp.  H*W must be even.
template <int C>
__global__ void __launch_bounds__(256) to_nhwc_wide_kernel(NhwcJob j0, NhwcJob j1) {
  constexpr int PXT = 128, RS = C + 2;                               // tile pixels, row stride in halfs
  __shared__ __align__(16) __half tile[PXT * RS];
  const bool second = (int)blockIdx.x >= j1.first_block;
  const NhwcJob J = second ? j1 : j0;
  const int blk = (int)blockIdx.x - J.first_block;
  const int frame = blk / J.blocks_per_frame;
  const int px0 = (blk - frame * J.blocks_per_frame) * PXT;
  const int npx = min(PXT, J.HW - px0);
  const __half* s = J.src + (int64_t)frame * C * J.HW + px0;
  const int t = threadIdx.x;
  constexpr int ITEMS = (PXT / 2) * C / 256;                         // 32 for C = 128
#pragma unroll 8
  for (int i = 0; i < ITEMS; ++i) {
    const int x = t + 256 * i;
    const int c = x / (PXT / 2), pp = x - c * (PXT / 2);
    if (2 * pp < npx) {
      const __half2 v = *reinterpret_cast<const __half2*>(s + (int64_t)c * J.HW + 2 * pp);
      tile[(2 * pp) * RS + c] = __low2half(v);
      tile[(2 * pp + 1) * RS + c] = __high2half(v);
    }
  }
  __syncthreads();
  unsigned* d = reinterpret_cast<unsigned*>(J.dst + ((int64_t)frame * J.HW + px0) * C);
  const unsigned* tw = reinterpret_cast<const unsigned*>(tile);
  for (int x = t; x < npx * (C / 2); x += 256) {
    const int px = x / (C / 2), w = x - px * (C / 2);
    d[x] = tw[px * (RS / 2) + w];
  }
}

// Window selection + 4-tap bilinear blend of one (edge, level) from its volume in shared memory
// (correlation_kernel.cu:221-230).  Lane = (p & 3, x tap j): for a group of four patch pixels the lane reads the taps
// (iy, j) of all 8 window rows (tap (iy, j + 1) comes from the neighbouring lane) -- record addresses are 12 X + 16 Y + p (mod 32 banks), so the 4 x 8
// lanes of one load always hit 32 different banks whatever the per-pixel window origins are -- and produces the outputs
// (xo = j, yo = 0..6).  Results go to a small staging buffer [xo][yo * 9 + p][level] (row stride chosen so that these
// stores are conflict-free as well) from which the unit's output is written with fully coalesced stores.
// RS: records per volume row, SLOT: floats per record, SROW: halfs per staging row.
__device__ __forceinline__ void stage_store(__half* p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ void stage_store(float* p, float v) { *p = v; }

template <int RS, int SLOT, int NLEV, int SROW, typename OT>
__device__ __forceinline__ void blend_stage(const float* vol, const float4* wgt, const int* vbase, int lane, int lev,
                                            OT* stage) {
  const int p_lo = lane & 3, j = lane >> 2;
#pragma unroll
  for (int pg = 0; pg < 3; ++pg) {
    const int p = 4 * pg + p_lo;
    const int pc = min(p, PP - 1);                 // lanes beyond the last pixel repeat it (same address: broadcast)
    const float4 wg = wgt[pc];
    const float* v = vol + vbase[pc] + j * SLOT;
    float t0[8], t1[8];
#pragma unroll
    for (int iy = 0; iy < 8; ++iy) t0[iy] = v[iy * RS * SLOT];
#pragma unroll
    for (int iy = 0; iy < 8; ++iy) t1[iy] = __shfl_down_sync(0xffffffffu, t0[iy], 4);      // tap (iy, j + 1)
    if (p < PP && j < 7) {
      OT* st = stage + j * SROW + p * NLEV + lev;
#pragma unroll
      for (int yo = 0; yo < 7; ++yo) {
        const float r = wg.x * t0[yo] + wg.y * t1[yo] + wg.z * t0[yo + 1] + wg.w * t1[yo + 1];
        stage_store(st + yo * 9 * NLEV, r);
      }
    }
  }
}

template <int C, int NLEV>
__global__ void __launch_bounds__(32 * WARPS) corr_tma_kernel(const __grid_constant__ CUtensorMap tm0,
                                                              const __grid_constant__ CUtensorMap tm1, Params P) {
  extern __shared__ __align__(128) unsigned char smraw[];
  static_assert(C % 8 == 0 && 2 * C >= 4 * PP, "a 2C-byte pixel record must hold the 9 volume values");
  constexpr int SLOT = C / 2;                           // floats per pixel record
  constexpr int K16 = C / 16, K8 = (C % 16) / 8;
  constexpr uint32_t TX_BYTES = RPX * C * 2;
  constexpr int NA4 = C * PP / 8;                        // uint4 per patch-feature record
  constexpr int NAV = (NA4 + 31) / 32;
  using WS = WarpSmem<C>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WS& S = reinterpret_cast<WS*>(smraw)[warp];
  const uint32_t bar0 = smem_u32(&S.bar[0]);
  const uint32_t reg0 = smem_u32(&S.reg[0][0]);

  if (lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  __syncwarp();

  // ldmatrix lane offsets (bytes).  A operand = region records: matrices (rows 0-7 | 8-15) x (ch 0-7 | 8-15)
  // Row r of an 8-row fragment block is region record perm(r) = 0,2,4,6,1,3,5,7 of the block: with 12-float records
  // the D stores of lanes g = 0..3 (and 4..7) then fall into disjoint banks.
  const int rperm = (((lane & 7) & 3) << 1) | ((lane & 7) >> 2);
  const uint32_t a_off4 = (uint32_t)((((lane >> 3) & 1) * 8 + rperm) * (2 * C) + (lane >> 4) * 16);
  const uint32_t a_off2 = (uint32_t)((((lane >> 3) & 1) * 8 + rperm) * (2 * C) + K16 * 32);
  const int g = lane >> 2, tq = lane & 3;
  const int gperm = ((g & 3) << 1) | (g >> 2);
  const uint64_t tma0 = (uint64_t)&tm0, tma1 = (uint64_t)&tm1;

  const int64_t total = (int64_t)P.B * P.E;
  const int64_t ustride = (int64_t)gridDim.x * WARPS;
  const int64_t u_first = (int64_t)blockIdx.x * WARPS + warp;
  if (u_first >= total) return;
  const int64_t n_units = (total - u_first + ustride - 1) / ustride;
  const int64_t n_half = n_units * NLEV;

  // ---- pipeline state.  Raw per-unit values are loaded two units ahead and only used (arithmetic) one unit later, so
  //      no global-load latency sits on the warp's in-order instruction stream.
  struct Raw { float x, y; int jx, ix; };   // lanes 0..8: level-0 coordinates; all lanes: vs[m], us[m]
  Raw ra{0.f, 0.f, 0, 0}, rb{0.f, 0.f, 0, 0};
  int64_t ka = 0;                    // ordinal of the unit held by ra (unit = u_first + ka * ustride)
  uint4 apre[NAV];                   // prefetched patch features of the next unit
  bool fits_cur = false, fits_nxt = false;
  int frame_cur = 0, frame_nxt = 0;  // b * F + jj of the current / next half-task
  uint32_t phase = 0;                // bit b: parity to wait for on buffer b
  unsigned bf0[K16 * 2 + K8], bf1[K16 * 2 + K8];    // B operand (patch features), n-tiles p 0-7 / 8-15

  auto load_raw = [&](Raw& r, int64_t k) {
    const int64_t unit = u_first + k * ustride;
    const int64_t m = (P.B == 1) ? unit : unit % P.E;
    const float* cg = P.coords + unit * (2 * PP);
    if (lane < PP) { r.x = __ldg(cg + lane); r.y = __ldg(cg + PP + lane); }
    r.jx = (int)__ldg(P.vs + m);
    r.ix = (int)__ldg(P.us + m);
  };
  auto prefetch_a = [&]() {          // patch features of the unit held by ra
    const int64_t b = (P.B == 1) ? 0 : (u_first + ka * ustride) / P.E;
    const uint4* f1 = reinterpret_cast<const uint4*>(P.fmap1 + (b * P.K + ra.ix) * (int64_t)(C * PP));
#pragma unroll
    for (int k = 0; k < NAV; ++k)
      if (lane + 32 * k < NA4) apre[k] = __ldg(f1 + lane + 32 * k);
  };
  // geometry of the half-task (lev, buffer slot) of the unit held by ra; issues the region load
  auto prepare = [&](int lev, int slot, int& frame) -> bool {
    const float x = (lev == 0) ? ra.x : ra.x * 0.25f;     // level 1: coords / 4 in fp32 (slam.py:322)
    const float y = (lev == 0) ? ra.y : ra.y * 0.25f;
    const int fxp = safe_floor(x), fyp = safe_floor(y);
    const int xmin = __reduce_min_sync(0xffffffffu, lane < PP ? fxp : 0x7fffffff);
    const int xmax = __reduce_max_sync(0xffffffffu, lane < PP ? fxp : -0x7fffffff);
    const int ymin = __reduce_min_sync(0xffffffffu, lane < PP ? fyp : 0x7fffffff);
    const int ymax = __reduce_max_sync(0xffffffffu, lane < PP ? fyp : -0x7fffffff);
    const bool fits = (xmax - xmin + D <= RG) && (ymax - ymin + D <= RG);
    const int b = (P.B == 1) ? 0 : (int)((u_first + ka * ustride) / P.E);
    frame = b * (int)P.F + ra.jx;
    if (lane < PP) {
      const float dx = x - floorf(x), dy = y - floorf(y);
      S.wgt[slot][lane] = make_float4((1.f - dx) * (1.f - dy), dx * (1.f - dy), (1.f - dx) * dy, dx * dy);
      S.vbase[slot][lane] = fits ? ((fyp - ymin) * RG + (fxp - xmin)) * SLOT + lane : lane;
    }
    __syncwarp();
    if (fits && lane == 0) {
      const int H = lev == 0 ? P.H[0] : P.H[1], W = lev == 0 ? P.W[0] : P.W[1];
      // a region entirely outside the map may be fetched from any outside position (all zeros either way)
      const int x0 = min(max(xmin - R, -RG), W), y0 = min(max(ymin - R, -RG), H);
      fence_proxy_async();                    // generic-proxy accesses of this buffer precede the async-proxy write
      mbar_expect_tx(bar0 + 8 * slot, TX_BYTES);
      tma_load_region(reg0 + slot * TX_BYTES, lev == 0 ? tma0 : tma1, bar0 + 8 * slot, x0 * (C / 2), y0, frame);
    }
    return fits;
  };
  // after the last level of ra's unit has been prepared: ra <- rb, start loading the unit after that
  auto advance = [&]() {
    ra = rb;
    ++ka;
    if (ka + 1 < n_units) load_raw(rb, ka + 1);
  };

  // ---- prologue: half-task 0
  load_raw(ra, 0);
  if (n_units > 1) load_raw(rb, 1);
  prefetch_a();
  fits_cur = prepare(0, 0, frame_cur);
  if (NLEV == 1) advance();

  for (int64_t s = 0; s < n_half; ++s) {
    const int lev = (NLEV == 1) ? 0 : (int)(s & 1);
    const int slot = (int)(s & 1);
    const int64_t unit = u_first + (s / NLEV) * ustride;
    const bool have_next = s + 1 < n_half;
    const int lev_n = (NLEV == 1) ? 0 : (int)((s + 1) & 1);

    // ---- (0) new unit: its patch-feature record [c][p] to shared memory as it is (one 16-byte store per lane), then
    //      the B fragments (8 patch pixels x 8 channels, k contiguous) by 16-bit loads: b(k = c, n = p) = a[c * 9 + p]
    if (lev == 0) {
#pragma unroll
      for (int k = 0; k < NAV; ++k)
        if (lane + 32 * k < NA4) reinterpret_cast<uint4*>(&S.a[0])[lane + 32 * k] = apre[k];
      __syncwarp();
      const unsigned short* ar = reinterpret_cast<const unsigned short*>(&S.a[0]);
#pragma unroll
      for (int x = 0; x < K16 * 2 + K8; ++x) {
        const int c = 8 * x + 2 * tq;
        bf0[x] = (unsigned)ar[c * PP + g] | ((unsigned)ar[(c + 1) * PP + g] << 16);
        bf1[x] = (g == 0) ? ((unsigned)ar[c * PP + 8] | ((unsigned)ar[(c + 1) * PP + 8] << 16)) : 0u;
      }
    }

    // ---- (1) geometry + region load of the next half-task (its buffer was released by the previous epilogue), patch
    //      features of the next unit, raw values of the unit after it
    if (have_next) {
      fits_nxt = prepare(lev_n, slot ^ 1, frame_nxt);
      if (lev_n == 0) prefetch_a();
      if (lev_n == NLEV - 1) advance();
    }

    // ---- (2) contraction of the current half-task, volume written in place
    float* vol = reinterpret_cast<float*>(&S.reg[slot][0]);
    if (fits_cur) {
      mbar_wait(bar0 + 8 * slot, (phase >> slot) & 1u);
      phase ^= 1u << slot;
      const uint32_t rb_ = reg0 + slot * TX_BYTES;
      // three 16-pixel tiles at a time: their fragment loads, MMAs and in-place stores form independent chains
#pragma unroll
      for (int mg = 0; mg < RPX / 48; ++mg) {
        unsigned af[3][K16 * 4 + K8 * 2];
        float d0[3][4], d1[3][4];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const uint32_t tb = rb_ + (mg * 3 + q) * 16 * (2 * C);
#pragma unroll
          for (int ks = 0; ks < K16; ++ks)
            ldsm_x4(af[q][4 * ks], af[q][4 * ks + 1], af[q][4 * ks + 2], af[q][4 * ks + 3], tb + a_off4 + ks * 32);
          if (K8) ldsm_x2(af[q][4 * K16], af[q][4 * K16 + 1], tb + a_off2);
#pragma unroll
          for (int x = 0; x < 4; ++x) { d0[q][x] = 0.f; d1[q][x] = 0.f; }
        }
#pragma unroll
        for (int ks = 0; ks < K16; ++ks)
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            mma_k16(d0[q], af[q][4 * ks], af[q][4 * ks + 1], af[q][4 * ks + 2], af[q][4 * ks + 3], bf0[2 * ks], bf0[2 * ks + 1]);
            mma_k16(d1[q], af[q][4 * ks], af[q][4 * ks + 1], af[q][4 * ks + 2], af[q][4 * ks + 3], bf1[2 * ks], bf1[2 * ks + 1]);
          }
        if (K8) {
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            mma_k8(d0[q], af[q][4 * K16], af[q][4 * K16 + 1], bf0[2 * K16]);
            mma_k8(d1[q], af[q][4 * K16], af[q][4 * K16 + 1], bf1[2 * K16]);
          }
        }
        // D[px = 16 mt + perm(g) (+8)][p = 2 tq, 2 tq + 1] and p = 8 from the second n-tile (tq == 0)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          float* v0 = vol + ((mg * 3 + q) * 16 + gperm) * SLOT;
          *reinterpret_cast<float2*>(v0 + 2 * tq) = make_float2(d0[q][0], d0[q][1]);
          *reinterpret_cast<float2*>(v0 + 8 * SLOT + 2 * tq) = make_float2(d0[q][2], d0[q][3]);
          if (tq == 0) { v0[8] = d1[q][0]; v0[8 * SLOT + 8] = d1[q][2]; }
        }
      }
    } else {
      // per-tap path (windows too far apart for one region): vol[(io * 8 + jo)][p]
      const int H = lev == 0 ? P.H[0] : P.H[1], W = lev == 0 ? P.W[0] : P.W[1];
      const __half* f2 = (lev == 0 ? P.nhwc[0] : P.nhwc[1]) + (int64_t)frame_cur * H * W * C;
      const float* cg = P.coords + unit * (2 * PP);
      for (int o = lane; o < PP * D * D; o += 32) {
        const int p = o / (D * D), pos = o - p * (D * D);
        const int io = pos / D, jo = pos - io * D;
        const float xs = (lev == 0) ? cg[p] : cg[p] * 0.25f, ys = (lev == 0) ? cg[PP + p] : cg[PP + p] * 0.25f;
        const int i1 = safe_floor(ys) + (io - R), j1 = safe_floor(xs) + (jo - R);
        float acc = 0.f;
        if (i1 >= 0 && i1 < H && j1 >= 0 && j1 < W) {
          const __half* src = f2 + ((int64_t)i1 * W + j1) * C;
          for (int c = 0; c < C; ++c) acc += __half2float(S.a[c * PP + p]) * __half2float(src[c]);
        }
        vol[pos * SLOT + p] = acc;
      }
    }
    __syncwarp();

    // ---- (3) window selection + bilinear blend into the staging buffer; after the unit's last level: permuted,
    //      level-interleaved output out[xo][yo][p][lev] (correlation_kernel.cu:232, slam.py:323), coalesced
    {
      constexpr int SROW = (NLEV == 2) ? 136 : 72;          // halfs per staging row: 68 words (2 levels) / 36 words
      if (fits_cur) blend_stage<RG, SLOT, NLEV, SROW>(vol, S.wgt[slot], S.vbase[slot], lane, lev, S.stage);
      else blend_stage<D, SLOT, NLEV, SROW>(vol, S.wgt[slot], S.vbase[slot], lane, lev, S.stage);
      __syncwarp();
      if (lev == NLEV - 1) {
        __half* og = P.out + unit * (int64_t)(Do * Do * PP) * NLEV;
#pragma unroll
        for (int t = 0; t < (Do * Do * PP + 31) / 32; ++t) {
          const int o = lane + 32 * t;
          if (o < Do * Do * PP) {
            const int xo = (o * 1041) >> 16;                // o / 63 for o < 441
            const int r = o - xo * 63;
            if (NLEV == 2) reinterpret_cast<unsigned*>(og)[o] = reinterpret_cast<const unsigned*>(S.stage)[xo * (SROW / 2) + r];
            else og[o] = S.stage[xo * SROW + r];
          }
        }
        __syncwarp();
      }
    }
    fits_cur = fits_nxt;
    frame_cur = frame_nxt;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Wide feature maps (C a multiple of 32, e.g. the 128-channel maps of the DPVO-style network, net_dpv.py:99): same
// scheme, but the channel axis is split into chunks of 32.  Per (edge, level) the 12 x 12 region arrives as C / 32
// TMA boxes [32 ch, 12, 12] (64-byte pixel records, SWIZZLE_64B so that ldmatrix rows do not collide in the banks) in
// a ring of WNBUF buffers that is refilled chunk by chunk (two chunks ahead, across half-tasks); the accumulators of the
// nine region tiles stay in registers across the chunks and are then written to a separate volume buffer with the
// same 12-float records as the narrow kernel, so the blend / staging / output code is shared.
// ---------------------------------------------------------------------------------------------------------------
constexpr int WCK = 32;                      // channels per chunk
constexpr int WNBUF = 2;                     // chunk buffers per warp (ring)
constexpr int WWARPS = 7;                    // warps per CTA of the wide kernel (one CTA per SM)
constexpr int WCH_BYTES = RPX * WCK * 2;     // 9216

template <int C> struct __align__(512) WideSmem {
  unsigned char buf[WNBUF][WCH_BYTES];     // TMA destinations (64B-swizzled pixel records)
  float vol[RPX * 12];                     // region volume, 12-float records: vol[px * 12 + p]
  __half a[PP * C];                        // patch-feature record of the unit, [c][p] as in fmap1
  float4 wgt[2][12];
  int vbase[2][12];
  __half stage[7 * 136];
  unsigned long long bar[WNBUF];
};

template <int C, int NLEV>
__global__ void __launch_bounds__(32 * WWARPS, 1) corr_tma_wide_kernel(const __grid_constant__ CUtensorMap tm0,
                                                                      const __grid_constant__ CUtensorMap tm1, Params P) {
  extern __shared__ __align__(128) unsigned char smraw[];
  static_assert(C % WCK == 0, "channel count must be a multiple of the chunk size");
  static_assert((C / WCK) % WNBUF == 0, "the chunk ring must divide the chunks of a half-task");
  constexpr int NCH = C / WCK;                           // chunks per half-task
  constexpr int SLOT = 12;
  constexpr int NA4 = C * PP / 8;
  constexpr int NAV = (NA4 + 31) / 32;
  using WS = WideSmem<C>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // the swizzled TMA destinations need 512-byte alignment: align by hand (the launch adds 1 KB of slack)
  WS& S = reinterpret_cast<WS*>(smraw + ((512u - (smem_u32(smraw) & 511u)) & 511u))[warp];
  const uint32_t bar0 = smem_u32(&S.bar[0]);
  const uint32_t buf0 = smem_u32(&S.buf[0][0]);
  if (lane == 0) {
    for (int b = 0; b < WNBUF; ++b) mbar_init(bar0 + 8 * b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  __syncwarp();

  // ldmatrix lane offsets inside a chunk buffer: row = permuted region record of the 8-row block, 16-byte channel
  // group XOR-swizzled with bits 1..2 of the record index (SWIZZLE_64B of 64-byte records)
  const int rperm = (((lane & 7) & 3) << 1) | ((lane & 7) >> 2);
  const int rrow = ((lane >> 3) & 1) * 8 + rperm;                 // record inside the 16-record tile
  const int sw = (rrow >> 1) & 3;
  const uint32_t a_off_k0 = (uint32_t)(rrow * 64 + (((0 + (lane >> 4)) ^ sw) << 4));   // channels 0..15 of the chunk
  const uint32_t a_off_k1 = (uint32_t)(rrow * 64 + (((2 + (lane >> 4)) ^ sw) << 4));   // channels 16..31
  const int g = lane >> 2, tq = lane & 3;
  const int gperm = ((g & 3) << 1) | (g >> 2);
  const uint64_t tma0 = (uint64_t)&tm0, tma1 = (uint64_t)&tm1;

  const int64_t total = (int64_t)P.B * P.E;
  const int64_t ustride = (int64_t)gridDim.x * WWARPS;
  const int64_t u_first = (int64_t)blockIdx.x * WWARPS + warp;
  if (u_first >= total) return;
  const int64_t n_units = (total - u_first + ustride - 1) / ustride;
  const int64_t n_half = n_units * NLEV;

  struct Raw { float x, y; int jx, ix; };
  Raw ra{0.f, 0.f, 0, 0}, rb{0.f, 0.f, 0, 0};
  int64_t ka = 0;
  uint4 apre[NAV];
  struct Geo { int x0, y0, frame, lev; bool fits; };
  Geo gc{0, 0, 0, 0, false}, gn{0, 0, 0, 0, false};
  uint32_t phase = 0;
  unsigned bf0[C / 8], bf1[C / 8];

  auto load_raw = [&](Raw& r, int64_t k) {
    const int64_t unit = u_first + k * ustride;
    const int64_t m = (P.B == 1) ? unit : unit % P.E;
    const float* cg = P.coords + unit * (2 * PP);
    if (lane < PP) { r.x = __ldg(cg + lane); r.y = __ldg(cg + PP + lane); }
    r.jx = (int)__ldg(P.vs + m);
    r.ix = (int)__ldg(P.us + m);
  };
  auto prefetch_a = [&]() {
    const int64_t b = (P.B == 1) ? 0 : (u_first + ka * ustride) / P.E;
    const uint4* f1 = reinterpret_cast<const uint4*>(P.fmap1 + (b * P.K + ra.ix) * (int64_t)(C * PP));
#pragma unroll
    for (int k = 0; k < NAV; ++k)
      if (lane + 32 * k < NA4) apre[k] = __ldg(f1 + lane + 32 * k);
  };
  auto prepare = [&](int lev, int slot) -> Geo {
    const float x = (lev == 0) ? ra.x : ra.x * 0.25f;
    const float y = (lev == 0) ? ra.y : ra.y * 0.25f;
    const int fxp = safe_floor(x), fyp = safe_floor(y);
    const int xmin = __reduce_min_sync(0xffffffffu, lane < PP ? fxp : 0x7fffffff);
    const int xmax = __reduce_max_sync(0xffffffffu, lane < PP ? fxp : -0x7fffffff);
    const int ymin = __reduce_min_sync(0xffffffffu, lane < PP ? fyp : 0x7fffffff);
    const int ymax = __reduce_max_sync(0xffffffffu, lane < PP ? fyp : -0x7fffffff);
    Geo gq;
    gq.fits = (xmax - xmin + D <= RG) && (ymax - ymin + D <= RG);
    gq.lev = lev;
    const int b = (P.B == 1) ? 0 : (int)((u_first + ka * ustride) / P.E);
    gq.frame = b * (int)P.F + ra.jx;
    const int H = lev == 0 ? P.H[0] : P.H[1], W = lev == 0 ? P.W[0] : P.W[1];
    gq.x0 = min(max(xmin - R, -RG), W);
    gq.y0 = min(max(ymin - R, -RG), H);
    if (lane < PP) {
      const float dx = x - floorf(x), dy = y - floorf(y);
      S.wgt[slot][lane] = make_float4((1.f - dx) * (1.f - dy), dx * (1.f - dy), (1.f - dx) * dy, dx * dy);
      S.vbase[slot][lane] = gq.fits ? ((fyp - ymin) * RG + (fxp - xmin)) * SLOT + lane : lane;
    }
    __syncwarp();
    return gq;
  };
  // chunk q of half-task geometry gq into buffer q % WNBUF (the buffer must have been released)
  auto issue = [&](const Geo& gq, int q) {
    if (gq.fits && lane == 0) {
      const int b = q % WNBUF;
      fence_proxy_async();
      mbar_expect_tx(bar0 + 8 * b, WCH_BYTES);
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(buf0 + b * WCH_BYTES), "l"(gq.lev == 0 ? tma0 : tma1), "r"(bar0 + 8 * b), "r"(q * WCK), "r"(gq.x0),
                     "r"(gq.y0), "r"(gq.frame) : "memory");
    }
  };
  auto advance = [&]() {
    ra = rb;
    ++ka;
    if (ka + 1 < n_units) load_raw(rb, ka + 1);
  };

  // ---- prologue: half-task 0, all its chunks
  load_raw(ra, 0);
  if (n_units > 1) load_raw(rb, 1);
  prefetch_a();
  gc = prepare(0, 0);
#pragma unroll
  for (int q = 0; q < NCH && q < WNBUF; ++q) issue(gc, q);
  if (NLEV == 1) advance();

  for (int64_t s = 0; s < n_half; ++s) {
    const int lev = (NLEV == 1) ? 0 : (int)(s & 1);
    const int slot = (int)(s & 1);
    const int64_t unit = u_first + (s / NLEV) * ustride;
    const bool have_next = s + 1 < n_half;
    const int lev_n = (NLEV == 1) ? 0 : (int)((s + 1) & 1);

    // ---- (0) new unit: patch-feature record to shared memory, B fragments for all channels
    if (lev == 0) {
#pragma unroll
      for (int k = 0; k < NAV; ++k)
        if (lane + 32 * k < NA4) reinterpret_cast<uint4*>(&S.a[0])[lane + 32 * k] = apre[k];
      __syncwarp();
      const unsigned short* ar = reinterpret_cast<const unsigned short*>(&S.a[0]);
#pragma unroll
      for (int x = 0; x < C / 8; ++x) {
        const int c = 8 * x + 2 * tq;
        bf0[x] = (unsigned)ar[c * PP + g] | ((unsigned)ar[(c + 1) * PP + g] << 16);
        bf1[x] = (g == 0) ? ((unsigned)ar[c * PP + 8] | ((unsigned)ar[(c + 1) * PP + 8] << 16)) : 0u;
      }
    }

    // ---- (1) geometry of the next half-task, patch features of the next unit, raw values of the unit after it
    if (have_next) {
      gn = prepare(lev_n, slot ^ 1);
      if (lev_n == 0) prefetch_a();
      if (lev_n == NLEV - 1) advance();
      if (!gc.fits) {                     // the current half-task holds no buffers: fill them for the next one now
#pragma unroll
        for (int q = 0; q < NCH && q < WNBUF; ++q) issue(gn, q);
      }
    }

    // ---- (2) contraction, chunk by chunk; every consumed buffer is refilled for the next half-task
    float* vol = S.vol;
    if (gc.fits) {
      float d0[RPX / 16][4], d1[RPX / 16][4];
#pragma unroll
      for (int mt = 0; mt < RPX / 16; ++mt)
#pragma unroll
        for (int x = 0; x < 4; ++x) { d0[mt][x] = 0.f; d1[mt][x] = 0.f; }
#pragma unroll
      for (int q = 0; q < NCH; ++q) {
        const int b = q % WNBUF;
        mbar_wait(bar0 + 8 * b, (phase >> b) & 1u);
        phase ^= 1u << b;
        const uint32_t cb = buf0 + b * WCH_BYTES;
#pragma unroll
        for (int mg = 0; mg < RPX / 48; ++mg) {
          unsigned af[3][8];
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            const uint32_t tb = cb + (mg * 3 + t) * 16 * 64;
            ldsm_x4(af[t][0], af[t][1], af[t][2], af[t][3], tb + a_off_k0);
            ldsm_x4(af[t][4], af[t][5], af[t][6], af[t][7], tb + a_off_k1);
          }
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              const int x = 4 * q + 2 * ks;
              mma_k16(d0[mg * 3 + t], af[t][4 * ks], af[t][4 * ks + 1], af[t][4 * ks + 2], af[t][4 * ks + 3], bf0[x], bf0[x + 1]);
              mma_k16(d1[mg * 3 + t], af[t][4 * ks], af[t][4 * ks + 1], af[t][4 * ks + 2], af[t][4 * ks + 3], bf1[x], bf1[x + 1]);
            }
        }
        __syncwarp();
        // buffer b is free: it takes the chunk WNBUF positions further down the stream (this or the next half-task)
        if (q + WNBUF < NCH) issue(gc, q + WNBUF);
        else if (have_next) issue(gn, q + WNBUF - NCH);
      }
#pragma unroll
      for (int mt = 0; mt < RPX / 16; ++mt) {
        float* v0 = vol + (mt * 16 + gperm) * SLOT;
        *reinterpret_cast<float2*>(v0 + 2 * tq) = make_float2(d0[mt][0], d0[mt][1]);
        *reinterpret_cast<float2*>(v0 + 8 * SLOT + 2 * tq) = make_float2(d0[mt][2], d0[mt][3]);
        if (tq == 0) { v0[8] = d1[mt][0]; v0[8 * SLOT + 8] = d1[mt][2]; }
      }
    } else {
      const int H = lev == 0 ? P.H[0] : P.H[1], W = lev == 0 ? P.W[0] : P.W[1];
      const __half* f2 = (lev == 0 ? P.nhwc[0] : P.nhwc[1]) + (int64_t)gc.frame * H * W * C;
      const float* cg = P.coords + unit * (2 * PP);
      for (int o = lane; o < PP * D * D; o += 32) {
        const int p = o / (D * D), pos = o - p * (D * D);
        const int io = pos / D, jo = pos - io * D;
        const float xs = (lev == 0) ? cg[p] : cg[p] * 0.25f, ys = (lev == 0) ? cg[PP + p] : cg[PP + p] * 0.25f;
        const int i1 = safe_floor(ys) + (io - R), j1 = safe_floor(xs) + (jo - R);
        float acc = 0.f;
        if (i1 >= 0 && i1 < H && j1 >= 0 && j1 < W) {
          const __half* src = f2 + ((int64_t)i1 * W + j1) * C;
          for (int c = 0; c < C; ++c) acc += __half2float(S.a[c * PP + p]) * __half2float(src[c]);
        }
        vol[pos * SLOT + p] = acc;
      }
    }
    __syncwarp();

    // ---- (3) blend into the staging buffer; after the unit's last level: coalesced output
    {
      constexpr int SROW = (NLEV == 2) ? 136 : 72;
      if (gc.fits) blend_stage<RG, SLOT, NLEV, SROW>(vol, S.wgt[slot], S.vbase[slot], lane, lev, S.stage);
      else blend_stage<D, SLOT, NLEV, SROW>(vol, S.wgt[slot], S.vbase[slot], lane, lev, S.stage);
      __syncwarp();
      if (lev == NLEV - 1) {
        __half* og = P.out + unit * (int64_t)(Do * Do * PP) * NLEV;
#pragma unroll
        for (int t = 0; t < (Do * Do * PP + 31) / 32; ++t) {
          const int o = lane + 32 * t;
          if (o < Do * Do * PP) {
            const int xo = (o * 1041) >> 16;
            const int r = o - xo * 63;
            if (NLEV == 2) reinterpret_cast<unsigned*>(og)[o] = reinterpret_cast<const unsigned*>(S.stage)[xo * (SROW / 2) + r];
            else og[o] = S.stage[xo * SROW + r];
          }
        }
        __syncwarp();
      }
    }
    gc = gn;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 feature maps (training / parity path; C a multiple of 16, built for the 128-channel maps): the wide kernel's scheme
// with 16-channel chunks (64-byte pixel records again, SWIZZLE_64B, the same ldmatrix addresses: an 8x8 b16 matrix is
// 8 records x 4 floats, which is exactly the tf32 m16n8k8 A fragment) and the contraction as 3xTF32 on the tensor cores:
// a = hi + lo with hi = a rounded to tf32, lo = a - hi (see split_tf32); D += A_lo B_hi + A_hi B_lo + A_hi B_hi in fp32
// accumulators, i.e. the dropped terms are ~2^-21 relative -- fp32-class results (parity bar 1e-5 absolute, tests/test_parity_r2_gpu.py).  The patch
// fragments are re-read from shared memory per chunk (holding 128 channels x hi / lo in registers is not possible).
// ---------------------------------------------------------------------------------------------------------------
constexpr int FCK = 16;                      // channels per chunk (64-byte records)
constexpr int FWARPS = 6;                    // warps per CTA (one CTA per SM)

template <int C> struct __align__(512) Wide32Smem {
  unsigned char buf[WNBUF][WCH_BYTES];     // TMA destinations (64B-swizzled pixel records of 16 floats)
  float vol[RPX * 12];                     // region volume, 12-float records
  float a[PP * C];                         // patch-feature record of the unit, [c][p] as in fmap1
  float4 wgt[2][12];
  int vbase[2][12];
  float stage[7 * 136];
  unsigned long long bar[WNBUF];
};

struct Params32 {
  const float* fmap1;                  // [B, K, C, 3, 3]
  const float* nhwc[2];                // [B*F, H, W, C] per level
  int H[2], W[2];
  const float* coords;
  const int64_t* us; const int64_t* vs;
  int B; int64_t E, K, F;
  float* out;                          // [B, E, 7, 7, 3, 3, NLEV]
};

// a = hi + lo for the 3xTF32 contraction.  hi: a rounded to 10 mantissa bits by integer arithmetic (round half away, two
// instructions; `cvt.rna.tf32.f32` expands to a much longer sequence and was 65 % of this kernel's instructions, ncu source
// view profiles/r02/ncu_source_corr32_v15.txt); lo = a - hi, exact in fp32, handed to the tensor core as it is: the
// hardware ignores the low 13 mantissa bits of a tf32 operand, i.e. truncates lo to 10 bits -- the total representation
// error is <= 2^-11 * 2^-10 |a| = 2^-21 |a|.  (Values within 2^-11 of overflow would round up to inf: feature maps are O(1).)
__device__ __forceinline__ void split_tf32(float x, unsigned& hi, unsigned& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float d[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0,
                                         unsigned b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// fp32 maps [B*F, C, HW] -> [B*F, HW, C]: 64-pixel tiles through shared memory (rows padded by one float: both the
// transposing stores and the copy-out are conflict-free), coalesced 128-byte reads per channel plane, contiguous writes.
struct Nhwc32Job { const float* src; float* dst; int HW; int blocks_per_frame; int first_block; };

template <int C>
__global__ void __launch_bounds__(256) to_nhwc_f32_kernel(Nhwc32Job j0, Nhwc32Job j1) {
  constexpr int PXT = 64, RS = C + 1;
  extern __shared__ float tile32[];                     // [PXT][RS]
  const bool second = (int)blockIdx.x >= j1.first_block;
  const Nhwc32Job J = second ? j1 : j0;
  const int blk = (int)blockIdx.x - J.first_block;
  const int frame = blk / J.blocks_per_frame;
  const int px0 = (blk - frame * J.blocks_per_frame) * PXT;
  const int npx = min(PXT, J.HW - px0);
  const float* sp = J.src + (int64_t)frame * C * J.HW + px0;
  const int t = threadIdx.x;
  constexpr int ITEMS = PXT * C / 256;
#pragma unroll 8
  for (int i = 0; i < ITEMS; ++i) {
    const int x = t + 256 * i;
    const int c = x / PXT, px = x - c * PXT;
    if (px < npx) tile32[px * RS + c] = sp[(int64_t)c * J.HW + px];
  }
  __syncthreads();
  float* d = J.dst + ((int64_t)frame * J.HW + px0) * C;
  for (int x = t; x < npx * C; x += 256) {
    const int px = x / C, c = x - px * C;
    d[x] = tile32[px * RS + c];
  }
}

template <int C, int NLEV>
__global__ void __launch_bounds__(32 * FWARPS, 1) corr_tma_wide32_kernel(const __grid_constant__ CUtensorMap tm0,
                                                                        const __grid_constant__ CUtensorMap tm1, Params32 P) {
  extern __shared__ __align__(128) unsigned char smraw[];
  static_assert(C % FCK == 0 && (C / FCK) % WNBUF == 0, "channel count must be a multiple of 32");
  constexpr int NCH = C / FCK;                           // chunks per half-task
  constexpr int SLOT = 12;
  constexpr int NA4 = C * PP / 4;                        // uint4 per patch-feature record
  constexpr int NAV = (NA4 + 31) / 32;
  using WS = Wide32Smem<C>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WS& S = reinterpret_cast<WS*>(smraw + ((512u - (smem_u32(smraw) & 511u)) & 511u))[warp];
  const uint32_t bar0 = smem_u32(&S.bar[0]);
  const uint32_t buf0 = smem_u32(&S.buf[0][0]);
  if (lane == 0) {
    for (int b = 0; b < WNBUF; ++b) mbar_init(bar0 + 8 * b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  __syncwarp();

  // ldmatrix lane offsets inside a chunk buffer (identical to the fp16 wide kernel: 64-byte records, SWIZZLE_64B): the
  // x4 load of k-step s returns a0..a3 of the tf32 m16n8k8 fragment (rows g / g + 8, floats t / t + 4 of the step)
  const int rperm = (((lane & 7) & 3) << 1) | ((lane & 7) >> 2);
  const int rrow = ((lane >> 3) & 1) * 8 + rperm;
  const int sw = (rrow >> 1) & 3;
  const uint32_t a_off_k0 = (uint32_t)(rrow * 64 + (((0 + (lane >> 4)) ^ sw) << 4));   // floats 0..7 of the chunk
  const uint32_t a_off_k1 = (uint32_t)(rrow * 64 + (((2 + (lane >> 4)) ^ sw) << 4));   // floats 8..15
  const int g = lane >> 2, tq = lane & 3;
  const int gperm = ((g & 3) << 1) | (g >> 2);
  const uint64_t tma0 = (uint64_t)&tm0, tma1 = (uint64_t)&tm1;

  const int64_t total = (int64_t)P.B * P.E;
  const int64_t ustride = (int64_t)gridDim.x * FWARPS;
  const int64_t u_first = (int64_t)blockIdx.x * FWARPS + warp;
  if (u_first >= total) return;
  const int64_t n_units = (total - u_first + ustride - 1) / ustride;
  const int64_t n_half = n_units * NLEV;

  struct Raw { float x, y; int jx, ix; };
  Raw ra{0.f, 0.f, 0, 0}, rb{0.f, 0.f, 0, 0};
  int64_t ka = 0;
  uint4 apre[NAV];
  struct Geo { int x0, y0, frame, lev; bool fits; };
  Geo gc{0, 0, 0, 0, false}, gn{0, 0, 0, 0, false};
  uint32_t phase = 0;

  auto load_raw = [&](Raw& r, int64_t k) {
    const int64_t unit = u_first + k * ustride;
    const int64_t m = (P.B == 1) ? unit : unit % P.E;
    const float* cg = P.coords + unit * (2 * PP);
    if (lane < PP) { r.x = __ldg(cg + lane); r.y = __ldg(cg + PP + lane); }
    r.jx = (int)__ldg(P.vs + m);
    r.ix = (int)__ldg(P.us + m);
  };
  auto prefetch_a = [&]() {
    const int64_t b = (P.B == 1) ? 0 : (u_first + ka * ustride) / P.E;
    const uint4* f1 = reinterpret_cast<const uint4*>(P.fmap1 + (b * P.K + ra.ix) * (int64_t)(C * PP));
#pragma unroll
    for (int k = 0; k < NAV; ++k)
      if (lane + 32 * k < NA4) apre[k] = __ldg(f1 + lane + 32 * k);
  };
  auto prepare = [&](int lev, int slot) -> Geo {
    const float x = (lev == 0) ? ra.x : ra.x * 0.25f;
    const float y = (lev == 0) ? ra.y : ra.y * 0.25f;
    const int fxp = safe_floor(x), fyp = safe_floor(y);
    const int xmin = __reduce_min_sync(0xffffffffu, lane < PP ? fxp : 0x7fffffff);
    const int xmax = __reduce_max_sync(0xffffffffu, lane < PP ? fxp : -0x7fffffff);
    const int ymin = __reduce_min_sync(0xffffffffu, lane < PP ? fyp : 0x7fffffff);
    const int ymax = __reduce_max_sync(0xffffffffu, lane < PP ? fyp : -0x7fffffff);
    Geo gq;
    gq.fits = (xmax - xmin + D <= RG) && (ymax - ymin + D <= RG);
    gq.lev = lev;
    const int b = (P.B == 1) ? 0 : (int)((u_first + ka * ustride) / P.E);
    gq.frame = b * (int)P.F + ra.jx;
    const int H = lev == 0 ? P.H[0] : P.H[1], W = lev == 0 ? P.W[0] : P.W[1];
    gq.x0 = min(max(xmin - R, -RG), W);
    gq.y0 = min(max(ymin - R, -RG), H);
    if (lane < PP) {
      const float dx = x - floorf(x), dy = y - floorf(y);
      S.wgt[slot][lane] = make_float4((1.f - dx) * (1.f - dy), dx * (1.f - dy), (1.f - dx) * dy, dx * dy);
      S.vbase[slot][lane] = gq.fits ? ((fyp - ymin) * RG + (fxp - xmin)) * SLOT + lane : lane;
    }
    __syncwarp();
    return gq;
  };
  auto issue = [&](const Geo& gq, int q) {
    if (gq.fits && lane == 0) {
      const int b = q % WNBUF;
      fence_proxy_async();
      mbar_expect_tx(bar0 + 8 * b, WCH_BYTES);
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(buf0 + b * WCH_BYTES), "l"(gq.lev == 0 ? tma0 : tma1), "r"(bar0 + 8 * b), "r"(q * FCK), "r"(gq.x0),
                     "r"(gq.y0), "r"(gq.frame) : "memory");
    }
  };
  auto advance = [&]() {
    ra = rb;
    ++ka;
    if (ka + 1 < n_units) load_raw(rb, ka + 1);
  };

  load_raw(ra, 0);
  if (n_units > 1) load_raw(rb, 1);
  prefetch_a();
  gc = prepare(0, 0);
#pragma unroll
  for (int q = 0; q < NCH && q < WNBUF; ++q) issue(gc, q);
  if (NLEV == 1) advance();

  for (int64_t s = 0; s < n_half; ++s) {
    const int lev = (NLEV == 1) ? 0 : (int)(s & 1);
    const int slot = (int)(s & 1);
    const int64_t unit = u_first + (s / NLEV) * ustride;
    const bool have_next = s + 1 < n_half;
    const int lev_n = (NLEV == 1) ? 0 : (int)((s + 1) & 1);

    // ---- (0) new unit: patch-feature record to shared memory
    if (lev == 0) {
#pragma unroll
      for (int k = 0; k < NAV; ++k)
        if (lane + 32 * k < NA4) reinterpret_cast<uint4*>(&S.a[0])[lane + 32 * k] = apre[k];
      __syncwarp();
    }

    // ---- (1) geometry of the next half-task, patch features of the next unit, raw values of the unit after it
    if (have_next) {
      gn = prepare(lev_n, slot ^ 1);
      if (lev_n == 0) prefetch_a();
      if (lev_n == NLEV - 1) advance();
      if (!gc.fits) {
#pragma unroll
        for (int q = 0; q < NCH && q < WNBUF; ++q) issue(gn, q);
      }
    }

    // ---- (2) contraction, chunk by chunk (3xTF32); every consumed buffer is refilled for the next half-task
    float* vol = S.vol;
    if (gc.fits) {
      float d0[RPX / 16][4], d1[RPX / 16][4];
#pragma unroll
      for (int mt = 0; mt < RPX / 16; ++mt)
#pragma unroll
        for (int x = 0; x < 4; ++x) { d0[mt][x] = 0.f; d1[mt][x] = 0.f; }
#pragma unroll 1
      for (int q = 0; q < NCH; ++q) {
        // B fragments of this chunk: b(k = c, n = p) = a[c * 9 + p]; n-tile 0: p = g, n-tile 1: p = 8 (g == 0 only)
        unsigned bh0[4], bl0[4], bh1[4], bl1[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {                    // x = 2 * step + (0: k = t, 1: k = t + 4)
          const int c = q * FCK + 4 * x + tq;
          split_tf32(S.a[c * PP + g], bh0[x], bl0[x]);
          split_tf32(g == 0 ? S.a[c * PP + 8] : 0.f, bh1[x], bl1[x]);
        }
        const int b = q % WNBUF;
        mbar_wait(bar0 + 8 * b, (phase >> b) & 1u);
        phase ^= 1u << b;
        const uint32_t cb = buf0 + b * WCH_BYTES;
#pragma unroll
        for (int mt = 0; mt < RPX / 16; ++mt) {
          unsigned ar[8], ah[8], al[8];
          const uint32_t tb = cb + mt * 16 * 64;
          ldsm_x4(ar[0], ar[1], ar[2], ar[3], tb + a_off_k0);
          ldsm_x4(ar[4], ar[5], ar[6], ar[7], tb + a_off_k1);
#pragma unroll
          for (int x = 0; x < 8; ++x) split_tf32(__uint_as_float(ar[x]), ah[x], al[x]);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const int o = 4 * ks;
            mma_tf32(d0[mt], al[o], al[o + 1], al[o + 2], al[o + 3], bh0[2 * ks], bh0[2 * ks + 1]);
            mma_tf32(d0[mt], ah[o], ah[o + 1], ah[o + 2], ah[o + 3], bl0[2 * ks], bl0[2 * ks + 1]);
            mma_tf32(d0[mt], ah[o], ah[o + 1], ah[o + 2], ah[o + 3], bh0[2 * ks], bh0[2 * ks + 1]);
            mma_tf32(d1[mt], al[o], al[o + 1], al[o + 2], al[o + 3], bh1[2 * ks], bh1[2 * ks + 1]);
            mma_tf32(d1[mt], ah[o], ah[o + 1], ah[o + 2], ah[o + 3], bl1[2 * ks], bl1[2 * ks + 1]);
            mma_tf32(d1[mt], ah[o], ah[o + 1], ah[o + 2], ah[o + 3], bh1[2 * ks], bh1[2 * ks + 1]);
          }
        }
        __syncwarp();
        if (q + WNBUF < NCH) issue(gc, q + WNBUF);
        else if (have_next) issue(gn, q + WNBUF - NCH);
      }
#pragma unroll
      for (int mt = 0; mt < RPX / 16; ++mt) {
        float* v0 = vol + (mt * 16 + gperm) * SLOT;
        *reinterpret_cast<float2*>(v0 + 2 * tq) = make_float2(d0[mt][0], d0[mt][1]);
        *reinterpret_cast<float2*>(v0 + 8 * SLOT + 2 * tq) = make_float2(d0[mt][2], d0[mt][3]);
        if (tq == 0) { v0[8] = d1[mt][0]; v0[8 * SLOT + 8] = d1[mt][2]; }
      }
    } else {
      const int H = lev == 0 ? P.H[0] : P.H[1], W = lev == 0 ? P.W[0] : P.W[1];
      const float* f2 = (lev == 0 ? P.nhwc[0] : P.nhwc[1]) + (int64_t)gc.frame * H * W * C;
      const float* cg = P.coords + unit * (2 * PP);
      for (int o = lane; o < PP * D * D; o += 32) {
        const int p = o / (D * D), pos = o - p * (D * D);
        const int io = pos / D, jo = pos - io * D;
        const float xs = (lev == 0) ? cg[p] : cg[p] * 0.25f, ys = (lev == 0) ? cg[PP + p] : cg[PP + p] * 0.25f;
        const int i1 = safe_floor(ys) + (io - R), j1 = safe_floor(xs) + (jo - R);
        float acc = 0.f;
        if (i1 >= 0 && i1 < H && j1 >= 0 && j1 < W) {
          const float* src = f2 + ((int64_t)i1 * W + j1) * C;
          for (int c = 0; c < C; ++c) acc = fmaf(S.a[c * PP + p], src[c], acc);
        }
        vol[pos * SLOT + p] = acc;
      }
    }
    __syncwarp();

    // ---- (3) blend into the staging buffer; after the unit's last level: coalesced output
    {
      constexpr int SROW = (NLEV == 2) ? 136 : 72;
      if (gc.fits) blend_stage<RG, SLOT, NLEV, SROW>(vol, S.wgt[slot], S.vbase[slot], lane, lev, S.stage);
      else blend_stage<D, SLOT, NLEV, SROW>(vol, S.wgt[slot], S.vbase[slot], lane, lev, S.stage);
      __syncwarp();
      if (lev == NLEV - 1) {
        constexpr int ROWN = 63 * NLEV;                  // floats per staging row actually used
        float* og = P.out + unit * (int64_t)(Do * Do * PP) * NLEV;
#pragma unroll 4
        for (int o = lane; o < Do * Do * PP * NLEV; o += 32) {
          const int xo = o / ROWN, r = o - xo * ROWN;
          og[o] = S.stage[xo * SROW + r];
        }
        __syncwarp();
      }
    }
    gc = gn;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn resolve_encode_fn() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    return (EncodeTiledFn)p;
  return nullptr;
}

static EncodeTiledFn encode_fn() {
  static const EncodeTiledFn fn = resolve_encode_fn();      // thread-safe one-time initialisation
  return fn;
}

static int make_map(CUtensorMap* tm, const void* base, int C, int W, int H, int64_t frames) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return PCORR_ERR_UNSUPPORTED;
  const cuuint64_t dims[3] = {(cuuint64_t)W * C / 2, (cuuint64_t)H, (cuuint64_t)frames};
  const cuuint64_t strides[2] = {(cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[3] = {(cuuint32_t)(RG * C / 2), RG, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PCORR_OK : PCORR_ERR_UNSUPPORTED;
}

// 4-D map [C, W, H, frames] with box [32, 12, 12, 1] and SWIZZLE_64B (wide maps: one box per channel chunk)
static int make_map_wide(CUtensorMap* tm, const void* base, int C, int W, int H, int64_t frames) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return PCORR_ERR_UNSUPPORTED;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)frames};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {WCK, RG, RG, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PCORR_OK : PCORR_ERR_UNSUPPORTED;
}

// fp32 maps: 4-D map [C, W, H, frames] with box [16, 12, 12, 1] floats (64-byte records) and SWIZZLE_64B
static int make_map_wide32(CUtensorMap* tm, const void* base, int C, int W, int H, int64_t frames) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return PCORR_ERR_UNSUPPORTED;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)frames};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  const cuuint32_t box[4] = {FCK, RG, RG, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PCORR_OK : PCORR_ERR_UNSUPPORTED;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

template <int C, int NLEV>
static int launch_wide32(const CUtensorMap& tm0, const CUtensorMap& tm1, const Params32& P, cudaStream_t s) {
  auto kern = corr_tma_wide32_kernel<C, NLEV>;
  const size_t smem = sizeof(Wide32Smem<C>) * FWARPS + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t units = (int64_t)P.B * P.E;
  int64_t grid = (units + FWARPS - 1) / FWARPS;
  if (grid > 148) grid = 148;
  kern<<<(unsigned)grid, 32 * FWARPS, smem, s>>>(tm0, tm1, P);
  pgba::count_launch();
  return (int)cudaGetLastError();
}

static void transpose_maps32(int nlev, const float* src0, float* dst0, int HW0, const float* src1, float* dst1, int HW1,
                             int frames, cudaStream_t s) {
  constexpr int C = 128, PXT = 64;
  Nhwc32Job j0{src0, dst0, HW0, (HW0 + PXT - 1) / PXT, 0};
  Nhwc32Job j1{src1, dst1, HW1, (HW1 + PXT - 1) / PXT, j0.blocks_per_frame * frames};
  const int blocks = j1.first_block + (nlev == 2 ? j1.blocks_per_frame * frames : 0);
  if (nlev == 1) j1.first_block = 0x7fffffff;
  const size_t smem = sizeof(float) * PXT * (C + 1);
  cudaFuncSetAttribute(to_nhwc_f32_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  to_nhwc_f32_kernel<C><<<blocks, 256, smem, s>>>(j0, j1);
  pgba::count_launch();
}

template <int C, int NLEV>
static int launch_wide(const CUtensorMap& tm0, const CUtensorMap& tm1, const Params& P, cudaStream_t s) {
  auto kern = corr_tma_wide_kernel<C, NLEV>;
  const size_t smem = sizeof(WideSmem<C>) * WWARPS + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t units = (int64_t)P.B * P.E;
  int64_t grid = (units + WWARPS - 1) / WWARPS;
  if (grid > 148) grid = 148;
  kern<<<(unsigned)grid, 32 * WWARPS, smem, s>>>(tm0, tm1, P);
  pgba::count_launch();
  return (int)cudaGetLastError();
}

template <int C, int NLEV>
static int launch(const CUtensorMap& tm0, const CUtensorMap& tm1, const Params& P, cudaStream_t s) {
  auto kern = corr_tma_kernel<C, NLEV>;
  const size_t smem = sizeof(WarpSmem<C>) * WARPS;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t units = (int64_t)P.B * P.E;
  int64_t grid = (units + WARPS - 1) / WARPS;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  if (grid > 148 * per_sm) grid = 148 * per_sm;
  kern<<<(unsigned)grid, 32 * WARPS, smem, s>>>(tm0, tm1, P);
  pgba::count_launch();
  return (int)cudaGetLastError();
}

static void transpose_maps(int C, int nlev, const __half* src0, __half* dst0, int HW0, const __half* src1, __half* dst1,
                           int HW1, int frames, cudaStream_t s) {
  const bool pair = (HW0 % 2 == 0) && (nlev == 1 || HW1 % 2 == 0);
  if (C == 128 && pair) {                       // wide maps
    NhwcJob j0{src0, dst0, HW0, (HW0 + 127) / 128, 0};
    NhwcJob j1{src1, dst1, HW1, (HW1 + 127) / 128, j0.blocks_per_frame * frames};
    const int blocks = j1.first_block + (nlev == 2 ? j1.blocks_per_frame * frames : 0);
    if (nlev == 1) j1.first_block = 0x7fffffff;
    to_nhwc_wide_kernel<128><<<blocks, 256, 0, s>>>(j0, j1);
    pgba::count_launch();
    return;
  }
  const int pxt = C > 32 ? 128 : 512;
  NhwcJob j0{src0, dst0, HW0, (HW0 + pxt - 1) / pxt, 0};
  NhwcJob j1{src1, dst1, HW1, (HW1 + pxt - 1) / pxt, j0.blocks_per_frame * frames};
  int blocks = j1.first_block + (nlev == 2 ? j1.blocks_per_frame * frames : 0);
  if (nlev == 1) j1.first_block = 0x7fffffff;
  if (C == 24) {
    if (pair) to_nhwc_kernel<24, true, 512><<<blocks, 256, 0, s>>>(j0, j1);
    else to_nhwc_kernel<24, false, 512><<<blocks, 256, 0, s>>>(j0, j1);
  } else if (C == 32) {
    if (pair) to_nhwc_kernel<32, true, 512><<<blocks, 256, 0, s>>>(j0, j1);
    else to_nhwc_kernel<32, false, 512><<<blocks, 256, 0, s>>>(j0, j1);
  } else {
    to_nhwc_kernel<128, false, 128><<<blocks, 256, 0, s>>>(j0, j1);       // odd H*W only
  }
  pgba::count_launch();
}

}  // namespace pcorr_tma

using namespace pcorr_tma;

extern "C" {

int pcorr_tma_supported(int C, int P, int radius, int dtype) {
  if (P != 3 || radius != 3) return 0;
  if (dtype == PCORR_F16) return (C == 24 || C == 32 || C == 128) ? 1 : 0;
  if (dtype == PCORR_F32) return C == 128 ? 1 : 0;        // 3xTF32 tile kernel (corr_tma_wide32_kernel)
  return 0;
}

int pcorr_tma_workspace_bytes_dt(int nlev, int B, int64_t F, int C, int dtype, int H0, int W0, int H1, int W1, size_t* bytes) {
  if (!bytes) return PCORR_ERR_NULL;
  if (nlev < 1 || nlev > 2 || B <= 0 || F <= 0 || C <= 0 || H0 <= 0 || W0 <= 0 || (nlev == 2 && (H1 <= 0 || W1 <= 0)))
    return PCORR_ERR_SHAPE;
  const size_t es = dtype == PCORR_F32 ? 4 : 2;
  size_t n = align256((size_t)B * F * H0 * W0 * C * es);
  if (nlev == 2) n += align256((size_t)B * F * H1 * W1 * C * es);
  *bytes = n;
  return PCORR_OK;
}

int pcorr_tma_workspace_bytes(int nlev, int B, int64_t F, int C, int H0, int W0, int H1, int W1, size_t* bytes) {
  return pcorr_tma_workspace_bytes_dt(nlev, B, F, C, PCORR_F16, H0, W0, H1, W1, bytes);
}

// lookup on channel-last maps that are already in `workspace` (transpose == false) or are rebuilt from the NCHW maps first
static int forward_impl(const void* fmap1, const void* fmap2_l0, const void* fmap2_l1, const float* coords,
                        const int64_t* ii, const int64_t* jj, int nlev, int B, int64_t E, int64_t K, int64_t F, int C,
                        int H0, int W0, int H1, int W1, int P, int radius, int dtype, void* out, void* workspace,
                        size_t workspace_bytes, cudaStream_t s, bool transpose) {
  if (E == 0 || B == 0) return PCORR_OK;
  if (!fmap1 || (transpose && (!fmap2_l0 || (nlev == 2 && !fmap2_l1))) || !coords || !ii || !jj || !out || !workspace)
    return PCORR_ERR_NULL;
  if (nlev < 1 || nlev > 2 || B < 0 || E < 0 || K <= 0 || F <= 0 || H0 <= 0 || W0 <= 0) return PCORR_ERR_SHAPE;
  if (!pcorr_tma_supported(C, P, radius, dtype)) return PCORR_ERR_UNSUPPORTED;
  // the TMA box (12 x 12 pixels) must not exceed the map
  if (H0 < RG || W0 < RG || (nlev == 2 && (H1 < RG || W1 < RG))) return PCORR_ERR_UNSUPPORTED;
  if ((int64_t)B * F >= ((int64_t)1 << 31) || (int64_t)B * F > 65535 || ((uintptr_t)workspace & 255)) return PCORR_ERR_UNSUPPORTED;
  size_t need = 0;
  int rc = pcorr_tma_workspace_bytes_dt(nlev, B, F, C, dtype, H0, W0, H1, W1, &need);
  if (rc) return rc;
  if (need > workspace_bytes) return PCORR_ERR_SHAPE;
  if (dtype == PCORR_F32) {
    float* f0 = (float*)workspace;
    float* f1 = (float*)((char*)workspace + align256((size_t)B * F * H0 * W0 * C * 4));
    if (transpose)
      transpose_maps32(nlev, (const float*)fmap2_l0, f0, H0 * W0, (const float*)fmap2_l1, f1, nlev == 2 ? H1 * W1 : 0, B * (int)F, s);
    CUtensorMap t0m, t1m;
    rc = make_map_wide32(&t0m, f0, C, W0, H0, (int64_t)B * F);
    if (rc) return rc;
    if (nlev == 2) rc = make_map_wide32(&t1m, f1, C, W1, H1, (int64_t)B * F);
    else t1m = t0m;
    if (rc) return rc;
    Params32 Pf{};
    Pf.fmap1 = (const float*)fmap1;
    Pf.nhwc[0] = f0; Pf.nhwc[1] = f1;
    Pf.H[0] = H0; Pf.W[0] = W0; Pf.H[1] = H1; Pf.W[1] = W1;
    Pf.coords = coords; Pf.us = ii; Pf.vs = jj;
    Pf.B = B; Pf.E = E; Pf.K = K; Pf.F = F;
    Pf.out = (float*)out;
    return nlev == 2 ? launch_wide32<128, 2>(t0m, t1m, Pf, s) : launch_wide32<128, 1>(t0m, t1m, Pf, s);
  }
  __half* n0 = (__half*)workspace;
  __half* n1 = (__half*)((char*)workspace + align256((size_t)B * F * H0 * W0 * C * 2));
  if (transpose)
    transpose_maps(C, nlev, (const __half*)fmap2_l0, n0, H0 * W0, (const __half*)fmap2_l1, n1, nlev == 2 ? H1 * W1 : 0,
                   B * (int)F, s);
  CUtensorMap tm0, tm1;
  const bool wide = C > 32;
  rc = wide ? make_map_wide(&tm0, n0, C, W0, H0, (int64_t)B * F) : make_map(&tm0, n0, C, W0, H0, (int64_t)B * F);
  if (rc) return rc;
  if (nlev == 2) rc = wide ? make_map_wide(&tm1, n1, C, W1, H1, (int64_t)B * F) : make_map(&tm1, n1, C, W1, H1, (int64_t)B * F);
  else tm1 = tm0;
  if (rc) return rc;
  Params Pm{};
  Pm.fmap1 = (const __half*)fmap1;
  Pm.nhwc[0] = n0; Pm.nhwc[1] = n1;
  Pm.H[0] = H0; Pm.W[0] = W0; Pm.H[1] = H1; Pm.W[1] = W1;
  Pm.coords = coords; Pm.us = ii; Pm.vs = jj;
  Pm.B = B; Pm.E = E; Pm.K = K; Pm.F = F;
  Pm.out = (__half*)out;
  if (C == 128) return nlev == 2 ? launch_wide<128, 2>(tm0, tm1, Pm, s) : launch_wide<128, 1>(tm0, tm1, Pm, s);
  if (C == 24) return nlev == 2 ? launch<24, 2>(tm0, tm1, Pm, s) : launch<24, 1>(tm0, tm1, Pm, s);
  return nlev == 2 ? launch<32, 2>(tm0, tm1, Pm, s) : launch<32, 1>(tm0, tm1, Pm, s);
}

int pcorr_forward_tma(const void* fmap1, const void* fmap2_l0, const void* fmap2_l1, const float* coords,
                      const int64_t* ii, const int64_t* jj, int nlev, int B, int64_t E, int64_t K, int64_t F, int C,
                      int H0, int W0, int H1, int W1, int P, int radius, int dtype, void* out, void* workspace,
                      size_t workspace_bytes, pcorr_stream_t stream) {
  return forward_impl(fmap1, fmap2_l0, fmap2_l1, coords, ii, jj, nlev, B, E, K, F, C, H0, W0, H1, W1, P, radius, dtype, out,
                      workspace, workspace_bytes, (cudaStream_t)stream, true);
}

int pcorr_ring_update(const void* fmap2_l0, const void* fmap2_l1, int nlev, int B, int64_t F, int C, int H0, int W0, int H1,
                      int W1, int64_t first_frame, int64_t n_frames, void* ring, size_t ring_bytes, pcorr_stream_t stream) {
  if (n_frames == 0 || B == 0) return PCORR_OK;
  if (!fmap2_l0 || (nlev == 2 && !fmap2_l1) || !ring) return PCORR_ERR_NULL;
  if (nlev < 1 || nlev > 2 || B < 0 || F <= 0 || H0 <= 0 || W0 <= 0 || first_frame < 0 || n_frames < 0 ||
      first_frame + n_frames > F)
    return PCORR_ERR_SHAPE;
  if (!pcorr_tma_supported(C, 3, 3, PCORR_F16) || ((uintptr_t)ring & 255) || (int64_t)B * F > 65535) return PCORR_ERR_UNSUPPORTED;
  size_t need = 0;
  int rc = pcorr_tma_workspace_bytes(nlev, B, F, C, H0, W0, H1, W1, &need);
  if (rc) return rc;
  if (need > ring_bytes) return PCORR_ERR_SHAPE;
  const size_t hw0 = (size_t)H0 * W0, hw1 = nlev == 2 ? (size_t)H1 * W1 : 0;
  __half* n0 = (__half*)ring;
  __half* n1 = (__half*)((char*)ring + align256((size_t)B * F * hw0 * C * 2));
  for (int b = 0; b < B; ++b) {                   // frames [first, first + n) of every batch entry are contiguous
    const size_t f = (size_t)b * F + first_frame;
    transpose_maps(C, nlev, (const __half*)fmap2_l0 + f * hw0 * C, n0 + f * hw0 * C, (int)hw0,
                   nlev == 2 ? (const __half*)fmap2_l1 + f * hw1 * C : nullptr, n1 + f * hw1 * C, (int)hw1, (int)n_frames,
                   (cudaStream_t)stream);
  }
  return (int)cudaGetLastError();
}

int pcorr_forward_ring(const void* fmap1, const float* coords, const int64_t* ii, const int64_t* jj, int nlev, int B,
                       int64_t E, int64_t K, int64_t F, int C, int H0, int W0, int H1, int W1, int P, int radius, int dtype,
                       void* out, const void* ring, size_t ring_bytes, pcorr_stream_t stream) {
  return forward_impl(fmap1, nullptr, nullptr, coords, ii, jj, nlev, B, E, K, F, C, H0, W0, H1, W1, P, radius, dtype, out,
                      const_cast<void*>(ring), ring_bytes, (cudaStream_t)stream, false);
}

}  // extern "C"
