// K3 on the tensor cores: fused clipped-surrogate objective + MLP backward for policies
// O -> 64 -> 64 -> A (two hidden layers of width TC_W; tg_policy_grad routes here under
// TG_MATH_AUTO / TG_MATH_3XTF32).  128 threads per CTA, thread i = sample i of the tile =
// TMEM lane i; persistent CTAs walk tiles of 128 samples (one step t x 128 consecutive envs).
//
// Per tile (all GEMMs are 3xTF32 tcgen05.mma with fp32 accumulation in TMEM):
//   P1  FP32 pipe: H1 = act(W0 x + b0) for the thread's own sample; the row is written twice,
//       hi/lo split: K-major (bufA, A operand of the forward GEMM) and MN-major (bufB, B operand
//       of the weight-gradient GEMM, reduction over samples)
//   MMA D_f[128x64] = H1 . W1^T
//   P3  tcgen05.ld -> H2 = act(D_f + b1); output Linear, log-prob, ratio, clipped surrogate,
//       d/dmu; dZ2 = (Wo^T dmu) * act'(H2) written K-major into bufA (forward GEMM is done)
//   MMA D_b[128x64] = dZ2 . W1            (B = W1 MN-major: reduction over its rows)
//       meanwhile: warp-shuffle column sums for dWo, db1
//   P4  dZ2 written MN-major into bufA (backward GEMM is done)
//   MMA D_w[64x64]  = dZ2^T . H1          (A, B MN-major, reduction over the 128 samples)
//       meanwhile: tcgen05.ld D_b -> dZ1 = D_b * act'(H1); column sums for db0, dW0
//   P5  tcgen05.ld D_w (M = 64: row r in lane 32*(r/16) + r%16) added to register accumulators
// Column sums over samples use a register butterfly (62 shuffles per 64-column matrix): after it
// lane l of a warp holds the warp's sums of columns 2l and 2l+1; per-warp partials live in
// registers for the whole kernel and are combined once at the end.
// Every CTA writes its partial gradient into its private copy (gpart); grad_reduce_kernel
// (tg_update.cu) sums the copies in a fixed order -> deterministic result in flat torch layout.
#include <math.h>

#include "tg_umma.cuh"

struct UpdTcArgs {
    tg_tc_layout lay;
    int64_t N;
    int T;
    const float *obs, *act, *adv, *oldlp;
    const int32_t *len;
    const float *packed;
    float inv_sd[TG_MAX_ACT], inv_var[TG_MAX_ACT], log_norm;
    float eps_clip, scale, kl_scale;
    float *gpart;   // [grid][n_params], zero-initialised
    double *spart;  // [grid][4]
};

// butterfly column sums: v[0..64) per lane -> (v[0], v[1]) = sums over the warp's 32 lanes of
// columns 2*lane and 2*lane+1.  Destroys v.
template <int HALF, int OFF> TG_D void colsum_step(float *v, int lane) {
    const bool up = (lane & OFF) != 0;
#pragma unroll
    for (int j = 0; j < HALF; ++j) {
        const float send = up ? v[j] : v[j + HALF];
        const float keep = up ? v[j + HALF] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
}
TG_D void colsum64(float *v, int lane) {
    colsum_step<32, 16>(v, lane);
    colsum_step<16, 8>(v, lane);
    colsum_step<8, 4>(v, lane);
    colsum_step<4, 2>(v, lane);
    colsum_step<2, 1>(v, lane);
}

template <int O, int A, bool RELU>
__global__ void __launch_bounds__(128) update_tc_kernel(const __grid_constant__ UpdTcArgs a) {
    constexpr int W = TC_W, O4 = (O + 1 + 3) / 4 * 4;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t wbar, mbar;
    __shared__ uint32_t tmem_slot;
    __shared__ double sred[4][4];
    if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
    float *Wsm = reinterpret_cast<float *>(smem_raw);
    unsigned char *bufA_hi = smem_raw + ((size_t)a.lay.total * 4 + 1023) / 1024 * 1024;
    unsigned char *bufA_lo = bufA_hi + 128 * W * 4;
    unsigned char *bufB_hi = bufA_lo + 128 * W * 4;
    unsigned char *bufB_lo = bufB_hi + 128 * W * 4;
    stage_weights_tma(Wsm, a.packed, a.lay.total, &wbar);
    if (threadIdx.x == 0) {
        mbar_init(&mbar, 1);
        mbar_fence_init();
    }
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t tm_f = tmem, tm_b = tmem + 64, tm_w = tmem + 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t idesc_f = umma_idesc_tf32(128, W, false, false);
    const uint32_t idesc_b = umma_idesc_tf32(128, W, false, true);
    const uint32_t idesc_w = umma_idesc_tf32(64, W, true, true);
    const uint32_t w_u = smem_u32(Wsm);
    const uint32_t wf_hi = w_u + (uint32_t)a.lay.whi[1] * 4u, wf_lo = w_u + (uint32_t)a.lay.wlo[1] * 4u;
    const uint32_t wb_hi = w_u + (uint32_t)a.lay.wbhi[1] * 4u, wb_lo = w_u + (uint32_t)a.lay.wblo[1] * 4u;
    const uint32_t A_hi = smem_u32(bufA_hi), A_lo = smem_u32(bufA_lo), B_hi = smem_u32(bufB_hi), B_lo = smem_u32(bufB_lo);
    // this thread's row in the K-major core-matrix layout and in the MN-major SW128_32B layout
    const uint32_t core_row = (uint32_t)(threadIdx.x >> 3) * (uint32_t)(W / 4) * 128u + (uint32_t)(threadIdx.x & 7) * 16u;
    const uint32_t mn_row = (uint32_t)threadIdx.x * 128u;
    const int rs = threadIdx.x & 3;
    const float *w1 = Wsm + a.lay.w1, *b1 = Wsm + a.lay.bias[1], *wo = Wsm + a.lay.wo, *bo = Wsm + a.lay.bo;
    const int act_kind = RELU ? TG_ACT_RELU : a.lay.act;

    // persistent per-thread gradient partials
    // dW1 accumulator [W][W] in shared memory, stored column-major (accS[k*W + r]) so that the 16
    // lanes that own consecutive rows r hit consecutive banks
    float *accS = reinterpret_cast<float *>(bufB_lo + 128 * W * 4);
    for (int i = threadIdx.x; i < W * W; i += 128) accS[i] = 0.0f;
    float c_wo[A][2], c_b1[2] = {0.f, 0.f}, c_b0[2] = {0.f, 0.f}, c_w0[O][2], c_bo[A];
#pragma unroll
    for (int j = 0; j < A; ++j) { c_wo[j][0] = c_wo[j][1] = 0.0f; c_bo[j] = 0.0f; }
#pragma unroll
    for (int o = 0; o < O; ++o) c_w0[o][0] = c_w0[o][1] = 0.0f;
    double s_obj = 0.0, s_cnt = 0.0, s_ratio = 0.0, s_clip = 0.0;

    const int64_t N = a.N;
    const int64_t NB = (N + 127) / 128;
    const int64_t ntiles = NB * a.T;
    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = (int)(tile / NB);
        const int64_t n = (tile % NB) * 128 + threadIdx.x;
        const bool valid = n < N && t < a.len[n];
        // also orders the previous tile's TMEM loads / smem reads before this tile's writes
        tc_fence_before();
        if (!__syncthreads_or(valid ? 1 : 0)) continue;
        tc_fence_after();
        // ---- P1: first Linear on the FP32 pipe
        float x[O];
#pragma unroll
        for (int o = 0; o < O; ++o) x[o] = valid ? a.obs[((int64_t)t * O + o) * N + n] : 0.0f;
        float h[W];
#pragma unroll
        for (int nn = 0; nn < W; ++nn) {
            float wrow[O4];
#pragma unroll
            for (int q = 0; q < O4; q += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(w1 + nn * O4 + q);
                wrow[q] = v.x; wrow[q + 1] = v.y; wrow[q + 2] = v.z; wrow[q + 3] = v.w;
            }
            float acc = wrow[O];
#pragma unroll
            for (int o = 0; o < O; ++o) acc = fmaf(wrow[o], x[o], acc);
            h[nn] = act_fwd(acc, act_kind);
        }
        // act'(H1) kept as a bit mask (ReLU) or re-read from bufB later (other activations)
        uint32_t m1lo = 0, m1hi = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            m1lo |= (h[j] > 0.0f ? 1u : 0u) << j;
            m1hi |= (h[32 + j] > 0.0f ? 1u : 0u) << j;
        }
#pragma unroll
        for (int c = 0; c < W / 4; ++c) {
            float4 hi, lo;
            hi.x = tf32_hi(h[4 * c]); hi.y = tf32_hi(h[4 * c + 1]); hi.z = tf32_hi(h[4 * c + 2]); hi.w = tf32_hi(h[4 * c + 3]);
            lo.x = h[4 * c] - hi.x; lo.y = h[4 * c + 1] - hi.y; lo.z = h[4 * c + 2] - hi.z; lo.w = h[4 * c + 3] - hi.w;
            const uint32_t oc = core_row + (uint32_t)c * 128u;
            *reinterpret_cast<float4 *>(bufA_hi + oc) = hi;
            *reinterpret_cast<float4 *>(bufA_lo + oc) = lo;
            // mn32_offset(128, row, 4c): 32-column block, 32-byte chunk XOR (row % 4), 16-byte half
            const uint32_t om = (uint32_t)(c >> 3) * (128u * 128u) + mn_row + (uint32_t)((((c & 7) >> 1) ^ rs) << 5) +
                                (uint32_t)(c & 1) * 16u;
            *reinterpret_cast<float4 *>(bufB_hi + om) = hi;
            *reinterpret_cast<float4 *>(bufB_lo + om) = lo;
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc_fence_after();
            umma_gemm_3xtf32(tm_f, A_hi, A_lo, W, false, wf_hi, wf_lo, W, false, W, idesc_f, false, 3);
            umma_commit(&mbar);
        }
        mbar_wait(&mbar, phase);
        phase ^= 1u;
        tc_fence_after();
        // ---- P3: H2, output Linear, objective, dZ2
#pragma unroll
        for (int c0 = 0; c0 < W; c0 += 32) {
            float z[32];
            tmem_ld32(tm_f + lane_base + (uint32_t)c0, z);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 b4 = *reinterpret_cast<const float4 *>(b1 + c0 + j);
                h[c0 + j] = act_fwd(z[j] + b4.x, act_kind);
                h[c0 + j + 1] = act_fwd(z[j + 1] + b4.y, act_kind);
                h[c0 + j + 2] = act_fwd(z[j + 2] + b4.z, act_kind);
                h[c0 + j + 3] = act_fwd(z[j + 3] + b4.w, act_kind);
            }
        }
        float mu[A], dmu[A];
#pragma unroll
        for (int j = 0; j < A; ++j) {
            float acc = bo[j];
#pragma unroll
            for (int q = 0; q < W; q += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(wo + j * W + q);
                acc = fmaf(h[q], v.x, acc); acc = fmaf(h[q + 1], v.y, acc);
                acc = fmaf(h[q + 2], v.z, acc); acc = fmaf(h[q + 3], v.w, acc);
            }
            mu[j] = acc;
            dmu[j] = 0.0f;
        }
        if (valid) {
            const int64_t row = (int64_t)t * N + n;
            float av[A], m2 = 0.0f;
#pragma unroll
            for (int j = 0; j < A; ++j) {
                av[j] = a.act[((int64_t)t * A + j) * N + n];
                const float z = (av[j] - mu[j]) * a.inv_sd[j];
                m2 += z * z;
            }
            const float lp = -0.5f * m2 - a.log_norm;
            const float adv = a.adv[row], olp = a.oldlp[row];
            const float ratio = expf(lp - olp);
            const float lo = 1.0f - a.eps_clip, hi = 1.0f + a.eps_clip;
            const float s1 = ratio * adv, s2 = fminf(fmaxf(ratio, lo), hi) * adv;
            const bool in_range = ratio >= lo && ratio <= hi;
            float g;
            if (s1 < s2) g = adv;
            else if (s1 > s2) g = in_range ? adv : 0.0f;
            else g = 0.5f * (adv + (in_range ? adv : 0.0f));
            s_obj += (double)fminf(s1, s2) * a.scale;
            float dlp = a.scale * g * ratio;
            if (a.kl_scale != 0.0f) {
                const float eo = expf(olp);
                s_obj += (double)a.kl_scale * eo * (olp - lp);
                dlp -= a.kl_scale * eo;
            }
            s_cnt += 1.0; s_ratio += ratio; s_clip += in_range ? 0.0 : 1.0;
#pragma unroll
            for (int j = 0; j < A; ++j) {
                dmu[j] = dlp * (av[j] - mu[j]) * a.inv_var[j];
                c_bo[j] += dmu[j];
            }
        }
        // dZ2 = (Wo^T dmu) * act'(H2), written K-major (hi/lo) into bufA; q_j = dmu_j * H2 kept for dWo
        float dz[W];
#pragma unroll
        for (int q = 0; q < W; q += 4) {
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < A; ++j) {
                const float4 v = *reinterpret_cast<const float4 *>(wo + j * W + q);
                s4[0] = fmaf(dmu[j], v.x, s4[0]); s4[1] = fmaf(dmu[j], v.y, s4[1]);
                s4[2] = fmaf(dmu[j], v.z, s4[2]); s4[3] = fmaf(dmu[j], v.w, s4[3]);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) dz[q + e] = s4[e] * act_bwd_from_out(h[q + e], act_kind);
        }
#pragma unroll
        for (int c = 0; c < W / 4; ++c) {
            float4 hi, lo;
            hi.x = tf32_hi(dz[4 * c]); hi.y = tf32_hi(dz[4 * c + 1]); hi.z = tf32_hi(dz[4 * c + 2]); hi.w = tf32_hi(dz[4 * c + 3]);
            lo.x = dz[4 * c] - hi.x; lo.y = dz[4 * c + 1] - hi.y; lo.z = dz[4 * c + 2] - hi.z; lo.w = dz[4 * c + 3] - hi.w;
            const uint32_t oc = core_row + (uint32_t)c * 128u;
            *reinterpret_cast<float4 *>(bufA_hi + oc) = hi;
            *reinterpret_cast<float4 *>(bufA_lo + oc) = lo;
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc_fence_after();
            umma_gemm_3xtf32(tm_b, A_hi, A_lo, W, false, wb_hi, wb_lo, W, true, W, idesc_b, false, 3);
            umma_commit(&mbar);
        }
        // while the backward-data GEMM runs: column sums for dWo (q_j = dmu_j * H2) and db1 (dZ2)
#pragma unroll
        for (int j = 0; j < A; ++j) {
            float q[W];
#pragma unroll
            for (int e = 0; e < W; ++e) q[e] = dmu[j] * h[e];
            colsum64(q, lane);
            c_wo[j][0] += q[0];
            c_wo[j][1] += q[1];
        }
        {
            float q[W];
#pragma unroll
            for (int e = 0; e < W; ++e) q[e] = dz[e];
            colsum64(q, lane);
            c_b1[0] += q[0];
            c_b1[1] += q[1];
        }
        mbar_wait(&mbar, phase);
        phase ^= 1u;
        tc_fence_after();
        // ---- P4: dZ2 again, MN-major (A operand of the weight-gradient GEMM), into bufA
#pragma unroll
        for (int c = 0; c < W / 4; ++c) {
            float4 hi, lo;
            hi.x = tf32_hi(dz[4 * c]); hi.y = tf32_hi(dz[4 * c + 1]); hi.z = tf32_hi(dz[4 * c + 2]); hi.w = tf32_hi(dz[4 * c + 3]);
            lo.x = dz[4 * c] - hi.x; lo.y = dz[4 * c + 1] - hi.y; lo.z = dz[4 * c + 2] - hi.z; lo.w = dz[4 * c + 3] - hi.w;
            const uint32_t om = (uint32_t)(c >> 3) * (128u * 128u) + mn_row + (uint32_t)((((c & 7) >> 1) ^ rs) << 5) +
                                (uint32_t)(c & 1) * 16u;
            *reinterpret_cast<float4 *>(bufA_hi + om) = hi;
            *reinterpret_cast<float4 *>(bufA_lo + om) = lo;
        }
        // dH1 = D_b ; dZ1 = dH1 * act'(H1)
        float d1[W];
#pragma unroll
        for (int c0 = 0; c0 < W; c0 += 32) tmem_ld32(tm_b + lane_base + (uint32_t)c0, d1 + c0);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc_fence_after();
            umma_gemm_3xtf32(tm_w, A_hi, A_lo, 128, true, B_hi, B_lo, 128, true, 128, idesc_w, false, 3);
            umma_commit(&mbar);
        }
        // while the weight-gradient GEMM runs: first-layer gradients
        if (RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                d1[j] = ((m1lo >> j) & 1u) ? d1[j] : 0.0f;
                d1[32 + j] = ((m1hi >> j) & 1u) ? d1[32 + j] : 0.0f;
            }
        } else {
            // H1 = hi + lo from this thread's own row of bufB (MN-major copy; the GEMM only reads it)
#pragma unroll
            for (int c = 0; c < W / 4; ++c) {
                const uint32_t om = (uint32_t)(c >> 3) * (128u * 128u) + mn_row +
                                    (uint32_t)((((c & 7) >> 1) ^ rs) << 5) + (uint32_t)(c & 1) * 16u;
                const float4 vh = *reinterpret_cast<const float4 *>(bufB_hi + om);
                const float4 vl = *reinterpret_cast<const float4 *>(bufB_lo + om);
                d1[4 * c] *= act_bwd_from_out(vh.x + vl.x, act_kind);
                d1[4 * c + 1] *= act_bwd_from_out(vh.y + vl.y, act_kind);
                d1[4 * c + 2] *= act_bwd_from_out(vh.z + vl.z, act_kind);
                d1[4 * c + 3] *= act_bwd_from_out(vh.w + vl.w, act_kind);
            }
        }
#pragma unroll
        for (int o = 0; o < O; ++o) {
            float q[W];
#pragma unroll
            for (int e = 0; e < W; ++e) q[e] = d1[e] * x[o];
            colsum64(q, lane);
            c_w0[o][0] += q[0];
            c_w0[o][1] += q[1];
        }
        colsum64(d1, lane);
        c_b0[0] += d1[0];
        c_b0[1] += d1[1];
        mbar_wait(&mbar, phase);
        phase ^= 1u;
        tc_fence_after();
        // ---- P5: dW1 tile partial out of TMEM (row r of the M=64 accumulator: lane 32*(r/16) + r%16)
        {
            float z[32];
#pragma unroll
            for (int c0 = 0; c0 < W; c0 += 32) {
                tmem_ld32(tm_w + lane_base + (uint32_t)c0, z);
                if (lane < 16) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) accS[(c0 + j) * W + warp * 16 + lane] += z[j];
                }
            }
        }
    }
    // ---- write this CTA's partial gradient (private copy, zero-initialised by the host)
    float *gp = a.gpart + (int64_t)blockIdx.x * a.lay.n_params;
    const int64_t f0 = a.lay.flat_w[0], f1 = a.lay.flat_w[1], f2 = a.lay.flat_w[2];
    if (lane < 16) {
        const int r = warp * 16 + lane;
#pragma unroll
        for (int k = 0; k < W; ++k) gp[f1 + (int64_t)r * W + k] = accS[k * W + r];
    }
    // column partials of the four warps: atomics on the CTA-private copy (sum of 4 floats starting
    // from 0: the result does not depend on the order only up to rounding, so fix the order instead)
    __syncthreads();
    for (int wsel = 0; wsel < 4; ++wsel) {
        if (warp == wsel) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = 2 * lane + e;
                gp[f1 + (int64_t)W * W + col] += c_b1[e];                       // b1
                gp[f0 + (int64_t)W * O + col] += c_b0[e];                       // b0
#pragma unroll
                for (int o = 0; o < O; ++o) gp[f0 + (int64_t)col * O + o] += c_w0[o][e];   // W0[col][o]
#pragma unroll
                for (int j = 0; j < A; ++j) gp[f2 + (int64_t)j * W + col] += c_wo[j][e];   // Wo[j][col]
            }
        }
        __syncthreads();
    }
    // dbo and the statistics: warp shuffle then a fixed-order sum over warps
    {
        double v[4] = {s_obj, s_cnt, s_ratio, s_clip};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            for (int off = 16; off > 0; off >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], off);
            if (lane == 0) sred[q][warp] = v[q];
        }
        __syncthreads();
        if (threadIdx.x < 4 && a.spart) {
            double s = 0.0;
            for (int w = 0; w < 4; ++w) s += sred[threadIdx.x][w];
            a.spart[(int64_t)blockIdx.x * 4 + threadIdx.x] = s;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < A; ++j) {
            float s = c_bo[j];
            for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
            if (lane == 0) sred[0][warp] = (double)s;
            __syncthreads();
            if (threadIdx.x == 0) {
                float tot = 0.0f;
                for (int w = 0; w < 4; ++w) tot += (float)sred[0][w];
                gp[f2 + (int64_t)A * W + j] = tot;
            }
            __syncthreads();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

template <int O, int A>
static int launch_update_tc(const UpdTcArgs &a, int grid, size_t smem, cudaStream_t st) {
    void (*kern)(const UpdTcArgs) =
        a.lay.act == TG_ACT_RELU ? update_tc_kernel<O, A, true> : update_tc_kernel<O, A, false>;
    TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 128, smem, st>>>(a);
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

bool tg_update_tc_eligible(const tg_mlp_cfg *mlp) {
    if (!tg_tc_eligible(mlp) || mlp->n_layers != 3) return false;
    const int O = mlp->dims[0], A = mlp->dims[3];
    return (O == 3 && A == 1) || (O == 5 && A == 1) || (O == 10 && A == 2) || (O == 20 && A == 4);
}

int tg_update_tc_grid(const tg_ctx *ctx) { return ctx->sm_count; }

// Launch the tensor-core update kernel.  gpart [grid][n_params] must be zeroed; spart [grid][4].
int tg_policy_grad_tc(tg_ctx *ctx, const tg_mlp_cfg *mlp, int64_t N, int T, const float *obs, const float *act,
                      const float *adv, const float *old_logp, const int32_t *len, const float *params,
                      const float *inv_sd, const float *inv_var, float log_norm, float eps_clip, float scale,
                      float kl_scale, float *gpart, double *spart, int grid, cudaStream_t st) {
    UpdTcArgs a;
    memset(&a, 0, sizeof(a));
    int rc = tg_build_tc_layout(mlp, &a.lay, true);
    if (rc) return rc;
    rc = tg_pack_weights_tc(ctx, a.lay, params, st);
    if (rc) return rc;
    a.N = N; a.T = T; a.obs = obs; a.act = act; a.adv = adv; a.oldlp = old_logp; a.len = len;
    a.packed = ctx->packed_tc;
    for (int j = 0; j < TG_MAX_ACT; ++j) { a.inv_sd[j] = inv_sd[j]; a.inv_var[j] = inv_var[j]; }
    a.log_norm = log_norm; a.eps_clip = eps_clip; a.scale = scale; a.kl_scale = kl_scale;
    a.gpart = gpart; a.spart = spart;
    const size_t smem = ((size_t)a.lay.total * 4 + 1023) / 1024 * 1024 + 4 * (size_t)128 * TC_W * 4 +
                        (size_t)TC_W * TC_W * 4;
    TG_REQUIRE(smem <= (size_t)ctx->smem_optin, TG_ERR_UNSUPPORTED, "tensor-core update needs %zu B of shared memory", smem);
    const int O = a.lay.O, A = a.lay.A;
    if (O == 3 && A == 1) return launch_update_tc<3, 1>(a, grid, smem, st);
    if (O == 5 && A == 1) return launch_update_tc<5, 1>(a, grid, smem, st);
    if (O == 10 && A == 2) return launch_update_tc<10, 2>(a, grid, smem, st);
    if (O == 20 && A == 4) return launch_update_tc<20, 4>(a, grid, smem, st);
    tg_set_error("no tensor-core update instance for obs %d / act %d", O, A);
    return TG_ERR_UNSUPPORTED;
}
