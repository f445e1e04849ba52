// tg_umma_selftest: small GEMMs on the tcgen05 tensor cores (kind::tf32, fp32 accumulation in
// TMEM) with 1 (plain TF32) or 3 (3xTF32) passes, in the three operand arrangements the fused
// kernels use.  It exercises exactly the descriptor / layout / TMEM helpers of tg_umma.cuh so
// the GPU test suite can pin them against a float64 matmul.
//   mode 0  D[128 x N] = A[128 x K] * B[N x K]^T      A K-major, B K-major   (forward layer)
//   mode 1  D[128 x N] = A[128 x K] * Bm[K x N]        A K-major, B MN-major  (backward-data)
//   mode 2  D[ 64 x N] = Am[K x 64]^T * Bm[K x N]      A MN-major, B MN-major (weight gradient,
//                                                       reduction over the K samples)
// All inputs/outputs are row-major fp32 in global memory.
#include "tg_umma.cuh"

__global__ void __launch_bounds__(128) umma_selftest_kernel(const float *__restrict__ A, const float *__restrict__ B,
                                                            float *__restrict__ D, int K, int N, int passes, int mode) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base;
    if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
    const int M = (mode == 2 || mode == 3) ? 64 : 128;
    // operand buffers: K-major in the core-matrix layout, MN-major in SW128_32B (tg_umma.cuh)
    const bool a_mn = (mode == 2 || mode == 3), b_mn = (mode >= 1 && mode <= 3);
    const int a_rows = a_mn ? K : M, a_cols = a_mn ? M : K;       // stored matrix [rows][cols]
    const int b_rows = b_mn ? K : N, b_cols = b_mn ? N : K;
    const size_t a_bytes = (size_t)a_rows * a_cols * 4;
    const size_t b_bytes = (size_t)b_rows * b_cols * 4;
    unsigned char *a_hi = smem_raw, *a_lo = a_hi + a_bytes, *b_hi = a_lo + a_bytes, *b_lo = b_hi + b_bytes;
    for (int idx = threadIdx.x; idx < a_rows * a_cols; idx += 128) {
        const int row = idx / a_cols, c = idx % a_cols;
        const float v = A[idx], hi = tf32_hi(v);
        const uint32_t off = a_mn ? mn32_offset(a_rows, row, c) : core_offset(a_cols, row, c);
        *reinterpret_cast<float *>(a_hi + off) = hi;
        *reinterpret_cast<float *>(a_lo + off) = v - hi;
    }
    for (int idx = threadIdx.x; idx < b_rows * b_cols; idx += 128) {
        const int row = idx / b_cols, c = idx % b_cols;
        const float v = B[idx], hi = tf32_hi(v);
        const uint32_t off = b_mn ? mn32_offset(b_rows, row, c) : core_offset(b_cols, row, c);
        *reinterpret_cast<float *>(b_hi + off) = hi;
        *reinterpret_cast<float *>(b_lo + off) = v - hi;
    }
    uint32_t ncols = 32;
    while ((int)ncols < N + (mode == 4 ? 2 * K : 0)) ncols <<= 1;
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
    if (mode == 4) {
        // A operand through tensor memory: thread i writes row i (hi at column N.., lo at N+K..)
        const uint32_t my = tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
        for (int c0 = 0; c0 < K; c0 += 32) {
            float hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float v = A[(size_t)threadIdx.x * K + c0 + j];
                hi[j] = tf32_hi(v);
                lo[j] = v - hi[j];
            }
            tmem_st32(my + (uint32_t)(N + c0), hi);
            tmem_st32(my + (uint32_t)(N + K + c0), lo);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (threadIdx.x == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, N, false, false);
            uint32_t acc = 0;
            for (int pass = 0; pass < passes; ++pass) {
                const uint32_t acol = tmem + (uint32_t)N + (pass == 2 ? (uint32_t)K : 0u);
                const uint32_t b = smem_u32(pass == 1 ? b_lo : b_hi);
                for (int k = 0; k < K; k += 8) {
                    umma_tf32_ts(tmem, acol + (uint32_t)k, umma_operand_desc(b, b_cols, false, k), idesc, acc);
                    acc = 1u;
                }
            }
            umma_commit(&mma_bar);
        }
    } else if (threadIdx.x == 0) {
        umma_gemm_3xtf32(tmem, smem_u32(a_hi), smem_u32(a_lo), a_mn ? a_rows : a_cols, a_mn, smem_u32(b_hi),
                         smem_u32(b_lo), b_mn ? b_rows : b_cols, b_mn, K, umma_idesc_tf32(M, N, a_mn, b_mn), false,
                         passes);
        umma_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = warp * 32 + lane;
    // M = 128: accumulator row r lives in TMEM lane r.  M = 64 (cta_group::1): row r lives in lane
    // 32*(r/16) + r%16, i.e. 16 rows in each warp's 32-lane quadrant, so all four warps take part.
    const int out_row = (M == 128) ? row : (lane < 16 ? warp * 16 + lane : -1);
    for (int c = 0; c < N; c += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        if (mode == 3) {          // raw dump of all 128 lanes (layout probing)
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c + j < N) D[(size_t)row * N + c + j] = v[j];
        } else if (out_row >= 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c + j < N) D[(size_t)out_row * N + c + j] = v[j];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, ncols);
}

extern "C" int tg_umma_selftest(tg_ctx *ctx, const float *A, const float *B, float *D, int K, int N, int passes,
                                int mode, void *stream) {
    TG_REQUIRE(ctx && A && B && D, TG_ERR_ARG, "tg_umma_selftest: null argument");
    TG_REQUIRE(mode >= 0 && mode <= 4, TG_ERR_ARG,
               "mode must be 0..4 (3 = mode 2 with a raw 128-lane dump, 4 = mode 0 with A in tensor memory)");
    TG_REQUIRE(K >= 32 && K % 32 == 0 && K <= 256, TG_ERR_SHAPE, "K must be a multiple of 32 in [32,256]");
    TG_REQUIRE(N >= 16 && N % 16 == 0 && N <= 256, TG_ERR_SHAPE, "N must be a multiple of 16 in [16,256]");
    TG_REQUIRE(passes == 1 || passes == 3, TG_ERR_ARG, "passes must be 1 or 3");
    TG_CUDA(cudaSetDevice(ctx->device));
    const int M = (mode == 2 || mode == 3) ? 64 : 128;
    const size_t a_bytes = (size_t)M * K * 4;
    const size_t b_bytes = (size_t)N * K * 4;
    const size_t smem = 2 * a_bytes + 2 * b_bytes;
    TG_REQUIRE(smem <= (size_t)ctx->smem_optin, TG_ERR_SHAPE, "operands need %zu B of shared memory", smem);
    TG_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, K, N, passes, mode);
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

// ---------------------------------------------------------------------------
// tg_tmem_probe: does a tensor-memory load / store issued while tcgen05.mma instructions are in flight wait for
// them?  One CTA: thread 0 issues n_mma MMAs (M = 128, N = 64, K = 8, zero operands) into columns 0..63 and
// commits; every warp then immediately does a tcgen05.ld (32 columns of an UNTOUCHED region) followed by a
// tcgen05.st, each timed with clock64; finally the commit barrier is awaited and timed as well.
//   out[0] = cycles of the ld, out[1] = cycles of the st (+wait::st), out[2] = cycles until the MMAs retired
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tmem_probe_kernel(int n_mma, int variant, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    float *A = reinterpret_cast<float *>(smem_raw);              // [128][64] K-major core-matrix layout, zeros
    float *B = A + 128 * 64;                                     // [64][64]
    for (int i = threadIdx.x; i < 128 * 64 + 64 * 64; i += 128) A[i] = 0.0f;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t my_tm = tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
    const uint32_t idesc = umma_idesc_tf32(128, 64);
    const uint32_t a_u = smem_u32(A), b_u = smem_u32(B);
    const long long t0 = clock64();
    if (variant == 0) {
        if (threadIdx.x == 0) {
            for (int i = 0; i < n_mma; ++i)
                umma_tf32(tmem, umma_operand_desc(a_u, 64, false, (i % 8) * 8), umma_operand_desc(b_u, 64, false, (i % 8) * 8),
                          idesc, i ? 1u : 0u);
            umma_commit(&bar);
        }
    } else if (threadIdx.x < 32) {
        // warp-uniform issue: descriptors advance in uniform registers, one elected lane issues.  A warp broadcast
        // (__shfl_sync from lane 0) is what tells ptxas that the base addresses are warp-uniform.
        const uint32_t au = variant == 2 ? __shfl_sync(0xffffffffu, a_u, 0) : a_u;
        const uint32_t bu = variant == 2 ? __shfl_sync(0xffffffffu, b_u, 0) : b_u;
        const uint32_t tmem_u = variant == 2 ? __shfl_sync(0xffffffffu, tmem, 0) : tmem;
        const uint64_t a0 = umma_operand_desc(au, 64, false, 0), b0 = umma_operand_desc(bu, 64, false, 0);
        for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                umma_tf32_w(tmem_u, a0 + (uint64_t)(k * 16), b0 + (uint64_t)(k * 16), idesc, (i | k) ? 1u : 0u);
        }
        umma_commit_w(&bar);
    }
    __syncthreads();
    const long long t1 = clock64();
    float v[32];
    tmem_ld32(my_tm + 256u, v);
    const long long t2 = clock64();
    tmem_st32(my_tm + 320u, v);
    tmem_st_wait();
    const long long t3 = clock64();
    mbar_wait(&bar, 0);
    const long long t4 = clock64();
    if (threadIdx.x == 0) {
        out[0] = t2 - t1;
        out[1] = t3 - t2;
        out[2] = t4 - t0;
        out[3] = t1 - t0;
    }
    if (v[0] == 123.0f) out[4] = 1;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

extern "C" int tg_tmem_probe(tg_ctx *ctx, int n_mma, long long *out_host4) {
    TG_REQUIRE(ctx && out_host4 && n_mma >= -20000 && n_mma <= 4096, TG_ERR_ARG, "tg_tmem_probe: bad argument");   // n_mma < 0: warp-uniform issue variant
    TG_CUDA(cudaSetDevice(ctx->device));
    long long *d = nullptr;
    TG_CUDA(cudaMalloc(&d, 64));
    TG_CUDA(cudaMemset(d, 0, 64));
    const size_t smem = (128 * 64 + 64 * 64) * 4;
    TG_CUDA(cudaFuncSetAttribute(tmem_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int rep = 0; rep < 2; ++rep) tmem_probe_kernel<<<1, 128, smem>>>(n_mma < 0 ? (-n_mma) % 10000 : n_mma, n_mma < -10000 ? 2 : (n_mma < 0 ? 1 : 0), d);
    TG_CUDA(cudaDeviceSynchronize());
    TG_CUDA(cudaMemcpy(out_host4, d, 32, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return TG_OK;
}
