// Device building blocks shared by the rollout (K1) and update (K3) kernels:
// Philox noise, TMA bulk staging of the policy weights into shared memory, and
// the register-tiled fp32 layer GEMMs over a tile of B environments/samples.
#pragma once
#include "tg_common.cuh"

// ---------------------------------------------------------------------------
// Philox4x32-10 keyed by the rollout seed, counter = (env lo, env hi, step, 0).
// One call yields the <=4 standard normals of one policy call (Box-Muller).
// oracle/philox.py restates this stream on the host.
// ---------------------------------------------------------------------------
TG_HD void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[0] = n0; c[1] = (uint32_t)p1; c[2] = n2; c[3] = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

TG_D void philox_normal4(uint64_t seed, uint64_t n, uint32_t t, float z[4]) {
    uint32_t c[4] = {(uint32_t)n, (uint32_t)(n >> 32), t, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float k24 = 5.9604644775390625e-08f;  // 2^-24
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float u1 = ((float)(c[2 * i] >> 8) + 0.5f) * k24;      // (0,1)
        const float u2 = ((float)(c[2 * i + 1] >> 8) + 0.5f) * k24;
        const float r = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        z[2 * i] = r * cs;
        z[2 * i + 1] = r * sn;
    }
}

// ---------------------------------------------------------------------------
// mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) helpers
// ---------------------------------------------------------------------------
TG_D uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

TG_D void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
TG_D void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
TG_D void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
TG_D void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy, bytes % 16 == 0, both pointers 16 B aligned
TG_D void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Stage `floats` fp32 values (multiple of 4) from global into shared memory with
// the TMA bulk engine; executed by the whole CTA, returns after the data landed.
TG_D void stage_weights_tma(float *dst, const float *src, int64_t floats, uint64_t *bar) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t total = (uint32_t)(floats * 4);
        mbar_expect_tx(bar, total);
        const uint32_t CH = 32768;
        for (uint32_t off = 0; off < total; off += CH) {
            const uint32_t n = (total - off) < CH ? (total - off) : CH;
            tma_bulk_g2s((char *)dst + off, (const char *)src + off, n, bar);
        }
    }
    mbar_wait(bar, 0);
}

// ---------------------------------------------------------------------------
// activations
// ---------------------------------------------------------------------------
TG_D float act_fwd(float z, int act) {
    if (act == TG_ACT_RELU) return fmaxf(z, 0.0f);
    if (act == TG_ACT_TANH) return tanhf(z);
    return 1.0f / (1.0f + expf(-z));
}
// derivative expressed through the activation OUTPUT h
TG_D float act_bwd_from_out(float h, int act) {
    if (act == TG_ACT_RELU) return h > 0.0f ? 1.0f : 0.0f;
    if (act == TG_ACT_TANH) return 1.0f - h * h;
    return h * (1.0f - h);
}

// ---------------------------------------------------------------------------
// Hidden layer over a tile:  Xout[n][b] = act( bias[n] + sum_k Xin[k][b] * Wt[k][n] )
// Xin/Xout: shared, rows of LDX = B+4 floats (env/sample index contiguous).
// Wt: shared, [K][NP] (neuron index contiguous, zero padded), bias [NP].
// Thread (tm, tn) owns envs {tm*4..+3, B/2+tm*4..+3} x neurons {tn*8..+7}.
// MODE 0: store act(z).  MODE 1 (backward-data): Xout = z * act'(Hout) where
// Hout currently holds the forward activation of that layer (read in place).
// ---------------------------------------------------------------------------
template <bool WG> TG_D float4 ldw4(const float *p) {
    if (WG) return __ldg(reinterpret_cast<const float4 *>(p));
    return *reinterpret_cast<const float4 *>(p);
}
template <bool WG> TG_D float ldw1(const float *p) {
    if (WG) return __ldg(p);
    return *p;
}

// WG: the weight operand lives in global memory (L1/L2-cached read-only path)
// instead of shared memory -- the fallback for policies whose staged weights do
// not fit next to the activation tiles.
template <int CFG, int MODE, bool WG>
TG_D void tile_layer(const float *__restrict__ Wt, const float *__restrict__ bias, const float *Xin, float *Xout,
                     int K, int act) {
    constexpr int B = TileCfg<CFG>::B, NP = TileCfg<CFG>::NP, LDX = B + 4, TMB = B / 8;
    const int tm = threadIdx.x % TMB, tn = threadIdx.x / TMB;
    float acc[8][8];
    {
        float bb[8];
        if (MODE == 0) {
            const float4 b0 = ldw4<WG>(bias + tn * 8);
            const float4 b1 = ldw4<WG>(bias + tn * 8 + 4);
            bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w;
            bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) bb[j] = 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = bb[j];
    }
    const float *xp = Xin + tm * 4;
    const float *wp = Wt + tn * 8;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 xa = *reinterpret_cast<const float4 *>(xp + k * LDX);
        const float4 xb = *reinterpret_cast<const float4 *>(xp + k * LDX + B / 2);
        const float4 wa = ldw4<WG>(wp + k * NP);
        const float4 wb = ldw4<WG>(wp + k * NP + 4);
        const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(x[i], w[j], acc[i][j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float *row = Xout + (tn * 8 + j) * LDX + tm * 4;
        float4 o0, o1;
        if (MODE == 0) {
            o0 = make_float4(act_fwd(acc[0][j], act), act_fwd(acc[1][j], act), act_fwd(acc[2][j], act),
                             act_fwd(acc[3][j], act));
            o1 = make_float4(act_fwd(acc[4][j], act), act_fwd(acc[5][j], act), act_fwd(acc[6][j], act),
                             act_fwd(acc[7][j], act));
        } else {
            const float4 h0 = *reinterpret_cast<const float4 *>(row);
            const float4 h1 = *reinterpret_cast<const float4 *>(row + B / 2);
            o0 = make_float4(acc[0][j] * act_bwd_from_out(h0.x, act), acc[1][j] * act_bwd_from_out(h0.y, act),
                             acc[2][j] * act_bwd_from_out(h0.z, act), acc[3][j] * act_bwd_from_out(h0.w, act));
            o1 = make_float4(acc[4][j] * act_bwd_from_out(h1.x, act), acc[5][j] * act_bwd_from_out(h1.y, act),
                             acc[6][j] * act_bwd_from_out(h1.z, act), acc[7][j] * act_bwd_from_out(h1.w, act));
        }
        *reinterpret_cast<float4 *>(row) = o0;
        *reinterpret_cast<float4 *>(row + B / 2) = o1;
    }
}

// ---------------------------------------------------------------------------
// Output layer (A <= 4 neurons): thread (b, part) sums its K-slice for env b.
// Wo [A][K] torch layout, bo [A].  Result for env b in mu[0..A) valid in the
// threads with threadIdx.x < B after the call (partials combined through P in a
// fixed order, so the sum is deterministic).  Contains one __syncthreads when
// NT > B.
// ---------------------------------------------------------------------------
template <int CFG, int A, bool WG>
TG_D void tile_output_layer(const float *__restrict__ Wo, const float *__restrict__ bo, const float *Xin, float *P,
                            int K, float mu[A]) {
    constexpr int B = TileCfg<CFG>::B, NT = TileCfg<CFG>::NT, LDX = B + 4, PARTS = NT / B;
    const int b = threadIdx.x % B, part = threadIdx.x / B;
    const int kc = (K + PARTS - 1) / PARTS;
    const int k0 = part * kc, k1 = min(K, k0 + kc);
    float s[A];
#pragma unroll
    for (int j = 0; j < A; ++j) s[j] = (part == 0) ? ldw1<WG>(bo + j) : 0.0f;
#pragma unroll 4
    for (int k = k0; k < k1; ++k) {
        const float x = Xin[k * LDX + b];
#pragma unroll
        for (int j = 0; j < A; ++j) s[j] = fmaf(x, ldw1<WG>(Wo + j * K + k), s[j]);
    }
    if (PARTS == 1) {
#pragma unroll
        for (int j = 0; j < A; ++j) mu[j] = s[j];
    } else {
        if (part > 0) {
#pragma unroll
            for (int j = 0; j < A; ++j) P[((part - 1) * A + j) * B + b] = s[j];
        }
        __syncthreads();
        if (part == 0) {
#pragma unroll
            for (int j = 0; j < A; ++j) {
                float v = s[j];
                for (int q = 1; q < PARTS; ++q) v += P[((q - 1) * A + j) * B + b];
                mu[j] = v;
            }
        }
    }
}

