// tg_umma_selftest: D[128 x N] = A[128 x K] * B[N x K]^T on the tcgen05 tensor cores
// (kind::tf32, fp32 accumulation in TMEM) with 1 (plain TF32) or 3 (3xTF32) passes.
// Exercises exactly the descriptor / layout / TMEM helpers the fused kernels use, so the
// GPU test suite can pin them against a float64 matmul before they are trusted inside K1/K3.
#include "tg_umma.cuh"

__global__ void __launch_bounds__(128) umma_selftest_kernel(const float *__restrict__ A, const float *__restrict__ B,
                                                            float *__restrict__ D, int K, int N, int passes) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base;
    // 1024-byte aligned operand buffers
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int KB = K / 32;
    unsigned char *a_hi = base;
    unsigned char *a_lo = a_hi + (size_t)KB * 128 * 128;
    unsigned char *b_hi = a_lo + (size_t)KB * 128 * 128;
    unsigned char *b_lo = b_hi + (size_t)KB * N * 128;
    for (int idx = threadIdx.x; idx < 128 * K; idx += 128) {
        const int row = idx / K, k = idx % K;
        const float v = A[idx], hi = tf32_hi(v);
        const uint32_t off = sw128_offset(128, row, k);
        *reinterpret_cast<float *>(a_hi + off) = hi;
        *reinterpret_cast<float *>(a_lo + off) = v - hi;
    }
    for (int idx = threadIdx.x; idx < N * K; idx += 128) {
        const int row = idx / K, k = idx % K;
        const float v = B[idx], hi = tf32_hi(v);
        const uint32_t off = sw128_offset(N, row, k);
        *reinterpret_cast<float *>(b_hi + off) = hi;
        *reinterpret_cast<float *>(b_lo + off) = v - hi;
    }
    uint32_t ncols = 32;
    while ((int)ncols < N) ncols <<= 1;
    if (threadIdx.x == 0) {
        mbar_init(&mma_bar, 1);
        mbar_fence_init();
    }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base, ncols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (threadIdx.x == 0) {
        umma_gemm_3xtf32(tmem, smem_u32(a_hi), smem_u32(a_lo), 128, smem_u32(b_hi), smem_u32(b_lo), N, K,
                         umma_idesc_tf32(128, N), false, passes);
        umma_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = warp * 32 + lane;
    for (int c = 0; c < N; c += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (c + j < N) D[(size_t)row * N + c + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, ncols);
}

extern "C" int tg_umma_selftest(tg_ctx *ctx, const float *A, const float *B, float *D, int K, int N, int passes,
                                void *stream) {
    TG_REQUIRE(ctx && A && B && D, TG_ERR_ARG, "tg_umma_selftest: null argument");
    TG_REQUIRE(K >= 32 && K % 32 == 0 && K <= 256, TG_ERR_SHAPE, "K must be a multiple of 32 in [32,256]");
    TG_REQUIRE(N >= 16 && N % 16 == 0 && N <= 256, TG_ERR_SHAPE, "N must be a multiple of 16 in [16,256]");
    TG_REQUIRE(passes == 1 || passes == 3, TG_ERR_ARG, "passes must be 1 or 3");
    TG_CUDA(cudaSetDevice(ctx->device));
    const size_t smem = 1024 + 2 * (size_t)(K / 32) * 128 * 128 + 2 * (size_t)(K / 32) * N * 128;
    TG_REQUIRE(smem <= (size_t)ctx->smem_optin, TG_ERR_SHAPE, "operands need %zu B of shared memory", smem);
    TG_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, K, N, passes);
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}
