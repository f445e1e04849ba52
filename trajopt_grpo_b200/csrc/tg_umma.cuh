// tcgen05 (5th-gen tensor core) helpers for sm_100a: TMEM allocation, shared-memory
// matrix descriptors (K-major, SWIZZLE_128B), kind::tf32 MMA issue, commit, TMEM loads.
//
// Operand layout used everywhere in this engine (both A and B are K-major):
//   a matrix of R rows x K fp32/tf32 elements is stored as K/32 "k-blocks"; inside a
//   k-block every row is one 128-byte line (32 elements), rows are 128 B apart, and the
//   16-byte chunk index of a line is XOR-ed with (row % 8)  -- the SWIZZLE_128B canonical
//   layout (8 rows x 128 B = one 1024-byte swizzle atom; atoms of consecutive 8-row groups
//   are 1024 B apart = the descriptor's stride byte offset).  The buffer must be 1024-byte
//   aligned.  One tcgen05.mma of kind::tf32 consumes K = 8 elements = 32 bytes per row, so
//   stepping along K inside a k-block adds 32 B to the descriptor start address.
//
// 3xTF32: the tensor core reads only the top 19 bits of each 32-bit operand.  To keep
// fp32-level accuracy (the engine's 1e-5 parity tolerance) every operand x is split as
// hi = x & 0xffffe000 (exactly a tf32 number) and lo = x - hi (exact in fp32), and
// D = A_hi*B_hi + A_hi*B_lo + A_lo*B_hi is accumulated in the fp32 TMEM accumulator
// (the dropped lo*lo term is 2^-22 relative).
#pragma once
#include "tg_mlp.cuh"

// hi = x rounded to nearest tf32 (cvt.rna), so |lo| = |x - hi| <= 2^-11 |x| and the dropped
// lo*lo term is <= 2^-22 relative; the tensor core's own truncation of lo costs 2^-21.
TG_D float tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r & 0xffffe000u);
}

// byte offset of element (row, k) of a K-major SWIZZLE_128B operand with `rows` rows
TG_HD uint32_t sw128_offset(int rows, int row, int k) {
    const int kb = k >> 5, kk = k & 31;
    return (uint32_t)kb * (uint32_t)rows * 128u + (uint32_t)row * 128u + (uint32_t)(((kk >> 2) ^ (row & 7)) << 4) +
           (uint32_t)((kk & 3) << 2);
}

TG_D uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) /* LBO (unused for swizzled K-major) */ |
           (64ull << 32) /* SBO = 1024 B >> 4 */ | (1ull << 46) /* descriptor version (Blackwell) */ |
           (2ull << 61) /* SWIZZLE_128B */;
}

// instruction descriptor: D fp32, A/B tf32, both K-major, dense
TG_HD uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

TG_D void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {   // one full warp; ncols = power of two >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
TG_D void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
TG_D void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
TG_D void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (UMMA operand reads)
TG_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
TG_D void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
TG_D void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i)
TG_D void tmem_ld32(uint32_t taddr, float v[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Issue the MMAs of one GEMM  D[M x N] (+)= A[M x K] * B[N x K]^T  with the 3xTF32 split.
// a_hi/a_lo/b_hi/b_lo: shared-memory byte addresses (u32) of the four operand buffers in the
// layout above (A with a_rows rows, B with b_rows rows).  One thread.
TG_D void umma_gemm_3xtf32(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, int a_rows, uint32_t b_hi, uint32_t b_lo,
                           int b_rows, int K, uint32_t idesc, bool accumulate_first, int passes) {
    uint32_t acc = accumulate_first ? 1u : 0u;
    for (int pass = 0; pass < passes; ++pass) {
        const uint32_t a = (pass == 2) ? a_lo : a_hi;
        const uint32_t b = (pass == 1) ? b_lo : b_hi;
        for (int k = 0; k < K; k += 8) {
            const uint32_t ka = (uint32_t)(k >> 5) * (uint32_t)a_rows * 128u + (uint32_t)(k & 31) * 4u;
            const uint32_t kbo = (uint32_t)(k >> 5) * (uint32_t)b_rows * 128u + (uint32_t)(k & 31) * 4u;
            umma_tf32(tmem_d, umma_desc_sw128(a + ka), umma_desc_sw128(b + kbo), idesc, acc);
            acc = 1u;
        }
    }
}
