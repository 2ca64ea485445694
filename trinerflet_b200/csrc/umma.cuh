// sm_100a tensor-core plumbing used by the fused MLP kernels (mlp_tc.cu): tcgen05.mma with both operands in shared
// memory and the accumulator in tensor memory (TMEM), TMEM allocation, tcgen05.ld, mbarrier completion tracking.
//
// Operand tiles use the un-swizzled ("interleaved") canonical layout of the sm_100 shared-memory matrix descriptor:
// a tile of R rows x Ccols fp16 is stored as [Ccols/8][R][8] halves, i.e. element (r, c) lives at byte offset
//     ((c >> 3) * R + r) * 16 + (c & 7) * 2
// (8 rows x 16 bytes = one 128-byte core matrix; core matrices of consecutive 8-row groups are contiguous).
// The SAME bytes serve as
//   * a K-major  operand with MN = r, K = c:  SBO = 128        (8-row groups),  LBO = R * 16 (8-column groups);
//   * an MN-major operand with K  = r, MN = c: SBO = R * 16     (8-column groups), LBO = 128  (8-row groups);
// so an activation tile written once by its producer is the A operand of the next layer (K-major), and the A or B
// operand of a weight-gradient product that contracts over the points (MN-major); a weight tile is the B operand of
// the forward product (K-major) and of the input-gradient product (MN-major).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace tnl {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (r, c) in a tile of R rows
__host__ __device__ constexpr uint32_t tile_off(uint32_t R, uint32_t r, uint32_t c) { return ((c >> 3) * R + r) * 16u + (c & 7u) * 2u; }
__host__ __device__ constexpr uint32_t tile_bytes(uint32_t R, uint32_t Ccols) { return R * Ccols * 2u; }

// shared-memory matrix descriptor, no swizzle (layout_type 0), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// operand = rows [r0, ...) x columns [c0, ...) of a tile(R, *) at shared address `base`
//   K-major : MN = rows, K = columns; one MMA consumes 16 columns  -> advance c0 by 16 per K step
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t base, uint32_t R, uint32_t r0, uint32_t c0) {
    return make_desc(base + tile_off(R, r0, c0), R * 16u, 128u);
}
//   MN-major: K = rows, MN = columns; one MMA consumes 16 rows     -> advance r0 by 16 per K step
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t base, uint32_t R, uint32_t r0, uint32_t c0) {
    return make_desc(base + tile_off(R, r0, c0), 128u, R * 16u);
}

// instruction descriptor for kind::f16 : fp16 x fp16 -> fp32
__host__ __device__ constexpr uint32_t make_idesc(uint32_t M, uint32_t N, bool a_mn, bool b_mn) {
    return (1u << 4)                      // D format  = F32
           | (0u << 7) | (0u << 10)       // A, B format = F16
           | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they have completed
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes (st.shared) visible to the async proxy (the tensor core's operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Descriptor arithmetic in 16-byte units: the low word holds (address >> 4) in bits 0..13 and (LBO >> 4) in bits 16..29,
// the high word (SBO >> 4) and the version bit.  Stepping an operand along its contraction dimension only adds to the
// address field (shared memory is < 256 KB, so the field never carries).
struct Operand {
    uint32_t lo, hi, step;   // descriptor words for K step 0; address increment (16-byte units) per K step
};
// tile(R, *) at 16-byte-unit address a4 (+ element offset r0, c0), as a K-major operand (MN = rows, K = columns)
__device__ __forceinline__ Operand op_kmajor(uint32_t a4, uint32_t R, uint32_t r0, uint32_t c0) {
    Operand o;
    o.lo = ((a4 + (c0 >> 3) * R + r0) & 0x3FFFu) | (R << 16);   // LBO = R * 16 bytes
    o.hi = (128u >> 4) | (1u << 14);                             // SBO = 128 bytes, version 1
    o.step = 2 * R;                                              // 16 columns = two 8-column chunks
    return o;
}
// the same tile as an MN-major operand (K = rows, MN = columns)
__device__ __forceinline__ Operand op_mnmajor(uint32_t a4, uint32_t R, uint32_t r0, uint32_t c0) {
    Operand o;
    o.lo = ((a4 + (c0 >> 3) * R + r0) & 0x3FFFu) | ((128u >> 4) << 16);   // LBO = 128 bytes
    o.hi = R | (1u << 14);                                                 // SBO = R * 16 bytes
    o.step = 16;                                                           // 16 rows
    return o;
}
// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// KSTEPS products D (+)= A_k B_k^T and, optionally, the commit that tracks them.  Called by a CONVERGED warp: the
// descriptors are computed by all lanes (warp-uniform values the compiler can keep in uniform registers), one elected
// lane issues.  A single issuing thread pays ~85 cycles per tcgen05.mma when it also does the descriptor arithmetic in
// a divergent branch (measured, profiles/bench_umma.py), more than the 32-cycle tensor-pipe floor of a 128x64x16 product.
template <int KSTEPS>
__device__ __forceinline__ void mma_group(uint32_t d_tmem, Operand a, Operand b, uint32_t idesc, bool accumulate_first) {
    uint64_t ad[KSTEPS], bd[KSTEPS];
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
        ad[ks] = ((uint64_t)a.hi << 32) | (a.lo + ks * a.step);
        bd[ks] = ((uint64_t)b.hi << 32) | (b.lo + ks * b.step);
    }
    if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) mma_f16(d_tmem, ad[ks], bd[ks], idesc, (accumulate_first || ks > 0) ? 1u : 0u);
    }
    __syncwarp();
}
__device__ __forceinline__ void commit_elected(uint64_t* bar) {
    if (elect_one()) commit(bar);
    __syncwarp();
}
// runtime K-step count, issued by the calling thread alone (diagnostics)
__device__ __forceinline__ void mma_steps(uint32_t d_tmem, Operand a, Operand b, uint32_t idesc, int ksteps, bool accumulate_first) {
    for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t ad = ((uint64_t)a.hi << 32) | (a.lo + ks * a.step);
        const uint64_t bd = ((uint64_t)b.hi << 32) | (b.lo + ks * b.step);
        mma_f16(d_tmem, ad, bd, idesc, (accumulate_first || ks > 0) ? 1u : 0u);
    }
}

// TMEM allocation: executed by one full warp; the base address is written to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// TMEM -> registers: this warp's 32 lanes (lane field of taddr = 32 * (warp % 4)), 8 / 16 consecutive 32-bit columns.
// The destination registers are only defined after tmem_ld_wait(); pin() orders every later use behind it.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void pin(uint32_t (&r)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("" : "+r"(r[i]));
}
// load NCOLS (multiple of 8) accumulator columns of this thread's lane, as floats
template <int NCOLS>
__device__ __forceinline__ void tmem_load_row(uint32_t taddr, float (&v)[NCOLS]) {
    uint32_t r[NCOLS];
#pragma unroll
    for (int c = 0; c + 16 <= NCOLS; c += 16) tmem_ld16(taddr + c, r + c);
    if (NCOLS % 16) tmem_ld8(taddr + (NCOLS / 16) * 16, r + (NCOLS / 16) * 16);
    tmem_ld_wait();
    pin<NCOLS>(r);
#pragma unroll
    for (int i = 0; i < NCOLS; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace umma
}  // namespace tnl
