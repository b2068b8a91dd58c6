// tcgen05 (5th-generation tensor core) helpers for the chunk Schur product of linearize_kernel: raw PTX for sm_100a.
//
// The product  M = sum_q Q_q E_q E_q^T  (+ the gradient term sum_q Q_q u_q E_q) of a chunk of <= 96 patches and <= 11
// pose columns is one symmetric [67 x K] x [K x 67] contraction.  With X[n][q] = sqrt(Q_q) E_q[n] (row 6*ncols = sqrt(Q_q)
// u_q) it is D = X X^T, so ONE K-major operand array serves as both A and B of tcgen05.mma.  fp32 accuracy comes from the
// 3xTF32 split X = hi + lo (hi, lo representable in TF32): D = hi hi^T + lo hi^T + hi lo^T, fp32 accumulation in TMEM
// (the lo lo^T term is 2^-22 relative).  Unlike the legacy mma.sync path (round 1: slower than FFMA2 because of the
// fragment conversions and shared-memory wavefronts per fragment) the operands are read from shared memory by the tensor
// core itself, once, and the accumulator never touches the register file until the epilogue.
//
// Operand layout (UMMA canonical K-major, no swizzle; see cute/atom/mma_traits_sm100.hpp make_umma_desc<Major::K>): core
// matrices of 8 rows x 16 bytes (4 TF32 values) stored as 128 contiguous bytes; the core matrices of one 8-row group follow
// each other along K (LBO = 128 bytes), 8-row groups are UMMA_SBO bytes apart.
#pragma once
#include <stdint.h>

namespace pgba {
namespace umma {

constexpr int KMAX = 96;                       // patches per product (K extent of the operand array)
constexpr int NROWS = 72;                      // operand rows: 66 E columns + gradient row + padding (multiple of 8)
constexpr int LBO = 128;                       // bytes between the two K halves of one MMA (adjacent core matrices)
constexpr int SBO = (KMAX / 4) * 128;          // bytes between 8-row groups: 3072
constexpr int X_BYTES = (NROWS / 8) * SBO;     // 27 648 bytes per operand array (hi or lo)
constexpr int TMEM_COLS = 128;                 // allocation (power of two >= 72 accumulator columns)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, LBO, SBO in 16-byte units,
// version = 1 (sm_100), layout type 0 (no swizzle)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) | (1ull << 46);
}

// 32-bit instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32: D = F32, A = B = TF32, both K-major, M = 64
__device__ __forceinline__ uint32_t instr_desc_tf32_m64(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((64u >> 4) << 24);
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_in_smem) {          // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_in_smem)), "r"(TMEM_COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {               // whole warp (the one that allocated)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(TMEM_COLS) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (the tensor core's operand reads)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread -> one arrival on the mbarrier when they have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 / 8 consecutive columns (one row of D per thread)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// x = hi + lo: hi = x rounded to 10 mantissa bits by integer arithmetic (round half away; cvt.rna.tf32.f32 expands to a much
// longer instruction sequence), lo = x - hi exactly; the tensor core ignores the low 13 mantissa bits of a tf32 operand, so
// lo needs no rounding of its own (representation error <= 2^-21 |x|).
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = x - hi;
}

}  // namespace umma
}  // namespace pgba
