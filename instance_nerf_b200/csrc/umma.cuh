// Minimal tcgen05 / TMEM / mbarrier wrappers (inline PTX, sm_100a).
//
// Operand tiles live in shared memory in the canonical K-major, no-swizzle
// ("interleaved") UMMA layout: 8-row x 16-byte core matrices,
//     byte(r, k) = (r % 8) * 16 + (r / 8) * SBO + (k / 8) * LBO + (k % 8) * 2      (fp16)
// One tcgen05.mma (kind::f16) consumes K = 16, i.e. two core matrices along K.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=SWIZZLE_NONE [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// Instruction descriptor for kind::f16, fp16 A/B (K-major), fp32 accumulate (cute::UMMA::InstrDescriptor).
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4)            // c_format = F32
         | (0u << 7)            // a_format = F16
         | (0u << 10)           // b_format = F16
         | (0u << 15)           // a_major = K
         | (0u << 16)           // b_major = K
         | ((N >> 3) << 17)     // n_dim
         | ((M >> 4) << 24);    // m_dim
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

// Same, descriptors as (lo, hi) 32-bit halves: consecutive K chunks of one operand differ only by a constant added to the
// low word (start address >> 4; no carry out of its 14-bit field inside the 256 KB shared window), the high word (SBO, version)
// is a constant of the operand -- the issuing thread spends one add per operand per MMA instead of rebuilding both descriptors.
__device__ __forceinline__ void mma_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n"
        :: "r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// A operand in TENSOR MEMORY (lane = row, one 32-bit column = two consecutive fp16 K elements, 8 columns per K = 16 MMA;
// layout validated by scripts/probe/umma_probe_ts.cu), B through a shared-memory descriptor.
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) { return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }
// One lane of a converged warp (elect.sync): the MMA-issue branch is then uniform for the compiler -- a `tid == 0` branch makes
// it wrap every tcgen05 instruction in a per-active-thread loop (ELECT / BRA.U.ANY around each UTCHMMA).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(mbar)) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy smem writes visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(mbar)) : "memory");
}
// barrier over a subset of the CTA's warps (id 1..15; id 0 is __syncthreads)
template <uint32_t kId, uint32_t kThreads>
__device__ __forceinline__ void named_sync() { asm volatile("bar.sync %0, %1;" :: "n"(kId), "n"(kThreads) : "memory"); }
// Warpgroup-wide register re-budgeting (all 128 threads of an aligned warpgroup must execute it)
template <uint32_t kRegs>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(kRegs)); }
template <uint32_t kRegs>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(kRegs)); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* mbar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(ok) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
    return ok != 0;
}
// Blocking wait with a watchdog: a barrier that does not flip within 2^28 polls (tens of seconds: a failed try_wait suspends
// the thread for a hardware-defined interval before it returns) traps -- a mis-sized setmaxnreg budget or a missing arrive
// makes the launch fail with an error the C-ABI returns instead of hanging the GPU.  The counter is one register and one
// predicated add per poll (a clock64-based deadline cost 2.3 % of the render kernel on B200).
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(mbar, parity)) {
        if (++spins == (1u << 28)) __trap();
    }
}

// TMEM allocation: executed by ONE full warp; the base address lands in *slot (shared memory).
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "n"(kCols) : "memory");
}

// TMEM -> registers: this thread's lane (32 * (warp % 4) + laneid), 16 consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
// registers -> TMEM: this thread's lane, 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
           "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// byte offset of element (r, k) of an fp16 tile in the layout above
__device__ __host__ __forceinline__ uint32_t tile_off(uint32_t r, uint32_t k, uint32_t lbo, uint32_t sbo) {
    return (r & 7u) * 16u + (r >> 3) * sbo + (k >> 3) * lbo + (k & 7u) * 2u;
}

}  // namespace umma
