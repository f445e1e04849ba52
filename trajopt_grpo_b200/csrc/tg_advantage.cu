// K2: reward-to-go + advantage normalisation (include/trajopt_grpo.h: tg_advantage).
//
// Layout: rew/adv/rtg/values are [T][N] with the env index innermost, so a warp
// that owns 32 consecutive envs reads/writes one 128 B row segment per step.
// Each thread owns one env and runs the reverse discounted scan in a register
// (the time axis is inherently serial per env); group statistics are combined
// through shared memory in a fixed order (deterministic), in float64.
//
// GRPO (grpo.py:66-74,108-115): rtg[t] = r[t] + gamma*rtg[t+1] over the valid
// prefix; per group g the pool is every valid (episode, step) pair:
// A = (rtg - mean(rtg)) / std(rtg + 1e-8), unbiased std.  The scan is done twice
// (statistics pass, normalise pass) so that only rew is read and adv written:
// 8 B per slot-step when the second read hits L2, 12 B when it does not.
#include "tg_common.cuh"

#define ADV_THREADS 128
#define ADV_UNROLL 16

// rtg recurrence with torch's rounding (separate multiply and add, no FMA)
TG_D float rtg_step(float r, float gamma, float next) { return __fadd_rn(r, __fmul_rn(gamma, next)); }

// One env's reverse scan, statistics pass.  PRED = the warp's episode lengths differ: every lane walks the time axis
// from the warp's longest episode (Lw) so that each row access is one coalesced segment, and a lane whose episode has
// ended (t >= L) keeps rtg = 0 and contributes nothing.  PRED = false (all lanes have L == Lw): no predicates.
template <bool PRED>
TG_D void grpo_scan_stats(const float *__restrict__ rew, int64_t N, int64_t n, int L, int Lw, float gamma, double K,
                          double &ax, double &ay, double &ayy) {
    float rtg = 0.0f;
    // the recurrence is serial per env (bit-exact torch rounding forbids re-association), so the memory
    // parallelism comes from loading ADV_UNROLL reward rows ahead of the dependent chain
    int t = Lw - 1;
    for (; t >= ADV_UNROLL - 1; t -= ADV_UNROLL) {
        float r[ADV_UNROLL];
#pragma unroll
        for (int j = 0; j < ADV_UNROLL; ++j) r[j] = (!PRED || (t - j) < L) ? rew[(int64_t)(t - j) * N + n] : 0.0f;
#pragma unroll
        for (int j = 0; j < ADV_UNROLL; ++j) {
            if (!PRED || (t - j) < L) {
                rtg = rtg_step(r[j], gamma, rtg);
                const double y = (double)__fadd_rn(rtg, 1e-8f) - K;
                ax += (double)rtg;
                ay += y;
                ayy += y * y;
            }
        }
    }
    for (; t >= 0; --t) {
        if (!PRED || t < L) {
            rtg = rtg_step(rew[(int64_t)t * N + n], gamma, rtg);
            const double y = (double)__fadd_rn(rtg, 1e-8f) - K;
            ax += (double)rtg;
            ay += y;
            ayy += y * y;
        }
    }
}

// normalise pass: rescan, write the advantages (zeros past the end of the episode)
template <bool PRED>
TG_D void grpo_scan_norm(const float *__restrict__ rew, int64_t N, int64_t n, int L, int Lw, float gamma, float mean,
                         float sd, float *__restrict__ adv, float *__restrict__ rtg_out) {
    float rtg = 0.0f;
    int t = Lw - 1;
    for (; t >= ADV_UNROLL - 1; t -= ADV_UNROLL) {
        float r[ADV_UNROLL];
#pragma unroll
        for (int j = 0; j < ADV_UNROLL; ++j) r[j] = (!PRED || (t - j) < L) ? rew[(int64_t)(t - j) * N + n] : 0.0f;
#pragma unroll
        for (int j = 0; j < ADV_UNROLL; ++j) {
            const bool in = !PRED || (t - j) < L;
            if (in) rtg = rtg_step(r[j], gamma, rtg);
            adv[(int64_t)(t - j) * N + n] = in ? __fdiv_rn(__fsub_rn(rtg, mean), sd) : 0.0f;
            if (rtg_out) rtg_out[(int64_t)(t - j) * N + n] = in ? rtg : 0.0f;
        }
    }
    for (; t >= 0; --t) {
        const bool in = !PRED || t < L;
        if (in) rtg = rtg_step(rew[(int64_t)t * N + n], gamma, rtg);
        adv[(int64_t)t * N + n] = in ? __fdiv_rn(__fsub_rn(rtg, mean), sd) : 0.0f;
        if (rtg_out) rtg_out[(int64_t)t * N + n] = in ? rtg : 0.0f;
    }
}

__global__ void __launch_bounds__(ADV_THREADS)
adv_grpo_kernel(int64_t G, int E, int T, int GC, float gamma, const float *__restrict__ rew,
                const int32_t *__restrict__ len, float *__restrict__ adv, float *__restrict__ rtg_out) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int64_t N = G * (int64_t)E;
    const int64_t g0 = (int64_t)blockIdx.x * GC;
    const int ng = (int)min((int64_t)GC, G - g0);
    const int envs = ng * E;
    const int64_t n0 = g0 * E;
    double *sx = reinterpret_cast<double *>(sm_raw);  // [envs] sum rtg
    double *sy = sx + envs;                           // [envs] sum (rtg+1e-8)
    double *syy = sy + envs;                          // [envs] sum (rtg+1e-8)^2
    float *gmean = reinterpret_cast<float *>(syy + envs);  // [GC]
    float *gstd = gmean + GC;                              // [GC]

    // Both passes walk the time axis WARP-UNIFORMLY (see grpo_scan_stats): starting every lane at its own L-1 made
    // ragged warps touch 32 different rows per load (428 GB/s at the QuadPole2D shape, profiles/README_r2.md).
    // pass 1: per-env scan + statistics
    for (int i0 = (threadIdx.x & ~31); i0 < envs; i0 += ADV_THREADS) {
        const int i = i0 + (threadIdx.x & 31);
        const bool live = i < envs;
        const int64_t n = n0 + (live ? i : 0);
        const int L = live ? len[n] : 0;
        const int Lw = __reduce_max_sync(0xffffffffu, L);
        const bool ragged = __any_sync(0xffffffffu, live && L != Lw);
        if (!live) continue;
        // shift K = first scanned value of the group's first env: the one-pass variance of
        // (y - K) is exactly 0 for a constant group (std 0 -> NaN/inf like torch)
        // and well conditioned otherwise
        const int64_t e0 = n0 + (int64_t)(i / E) * E;
        const int L0 = len[e0];
        const double K = L0 > 0 ? (double)__fadd_rn(rew[(int64_t)(L0 - 1) * N + e0], 1e-8f) : 0.0;
        double ax = 0.0, ay = 0.0, ayy = 0.0;
        if (ragged) grpo_scan_stats<true>(rew, N, n, L, Lw, gamma, K, ax, ay, ayy);
        else grpo_scan_stats<false>(rew, N, n, L, Lw, gamma, K, ax, ay, ayy);
        sx[i] = ax; sy[i] = ay; syy[i] = ayy;
    }
    __syncthreads();
    if (threadIdx.x < ng) {
        const int g = threadIdx.x;
        double ax = 0.0, ay = 0.0, ayy = 0.0;
        int64_t cnt = 0;
        for (int e = 0; e < E; ++e) {
            ax += sx[g * E + e]; ay += sy[g * E + e]; ayy += syy[g * E + e];
            cnt += len[n0 + (int64_t)g * E + e];
        }
        const double mean = ax / (double)cnt;                  // cnt==0 -> NaN, as torch.mean of empty
        const double my = ay / (double)cnt;
        const double var = (ayy - ay * my) / (double)(cnt - 1);  // unbiased; cnt==1 -> 0/0 = NaN
        gmean[g] = (float)mean;
        gstd[g] = (float)sqrt(var > 0.0 || var != var ? var : 0.0);
    }
    __syncthreads();
    // pass 2: rescan and normalise; zero the padding
    for (int i0 = (threadIdx.x & ~31); i0 < envs; i0 += ADV_THREADS) {
        const int i = i0 + (threadIdx.x & 31);
        const bool live = i < envs;
        const int64_t n = n0 + (live ? i : 0);
        const int L = live ? len[n] : 0;
        const int Lw = __reduce_max_sync(0xffffffffu, L);
        const bool ragged = __any_sync(0xffffffffu, live && L != Lw);
        if (!live) continue;
        const float mean = gmean[i / E], sd = gstd[i / E];
        for (int t = T - 1; t >= Lw; --t) {
            adv[(int64_t)t * N + n] = 0.0f;
            if (rtg_out) rtg_out[(int64_t)t * N + n] = 0.0f;
        }
        if (ragged) grpo_scan_norm<true>(rew, N, n, L, Lw, gamma, mean, sd, adv, rtg_out);
        else grpo_scan_norm<false>(rew, N, n, L, Lw, gamma, mean, sd, adv, rtg_out);
    }
}

// ---------------------------------------------------------------------------
// PPO (ppo.py:93-139): advantages from MC returns or GAE with a critic baseline,
// then GLOBAL z-scores (unbiased std, +1e-8 on the denominator) of both the
// advantages and the returns over every valid step.
//   stage 1: per-env scan -> raw adv/rtg rows + per-block partial sums
//   stage 2: one block combines the partials in a fixed order -> 4 scalars
//   stage 3: normalise in place
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(ADV_THREADS)
adv_ppo_scan_kernel(int64_t N, int T, int gae, float gamma, float gamlam, const float *__restrict__ rew,
                    const int32_t *__restrict__ len, const float *__restrict__ val, float *__restrict__ adv,
                    float *__restrict__ rtg_out, double *__restrict__ partial) {
    __shared__ double red[5][ADV_THREADS];
    const int64_t n = (int64_t)blockIdx.x * ADV_THREADS + threadIdx.x;
    double sa = 0, saa = 0, sr = 0, srr = 0, cnt = 0;
    {
        // warp-uniform time axis (see adv_grpo_kernel): lanes past the end of their episode idle on the same row
        const bool live = n < N;
        const int64_t nn = live ? n : 0;
        const int L = live ? len[nn] : 0;
        const int Lw = __reduce_max_sync(0xffffffffu, L);
        if (live) {
            for (int t = T - 1; t >= Lw; --t) {
                adv[(int64_t)t * N + nn] = 0.0f;
                rtg_out[(int64_t)t * N + nn] = 0.0f;
            }
        }
        float rtg = 0.0f, a_next = 0.0f, v_next = 0.0f;
        constexpr int PU = 8;                 // rows loaded ahead of the serial recurrence
        for (int t0 = Lw - 1; t0 >= 0; t0 -= PU) {
            float rb[PU], vb[PU];
#pragma unroll
            for (int j = 0; j < PU; ++j) {
                const int tt = t0 - j;
                const bool in = tt >= 0 && tt < L;
                rb[j] = in ? rew[(int64_t)tt * N + nn] : 0.0f;
                vb[j] = in ? val[(int64_t)tt * N + nn] : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < PU; ++j) {
                const int t = t0 - j;
                if (t < 0) break;
                if (t >= L) {
                    if (live) { adv[(int64_t)t * N + nn] = 0.0f; rtg_out[(int64_t)t * N + nn] = 0.0f; }
                    continue;
                }
                const float r = rb[j], v = vb[j];
                float a, ret;
                if (!gae) {
                    rtg = rtg_step(r, gamma, rtg);            // ppo.py:103-108
                    ret = rtg;
                    a = __fsub_rn(rtg, v);                    // :111
                } else {
                    // :114-123 (masks are prefix masks: next value/advantage are 0 past the end)
                    const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(gamma, v_next)), v);
                    a = (t == T - 1) ? __fsub_rn(r, v) : __fadd_rn(delta, __fmul_rn(gamlam, a_next));
                    ret = __fadd_rn(v, a);                    // :124
                    a_next = a;
                    v_next = v;
                }
                adv[(int64_t)t * N + nn] = a;
                rtg_out[(int64_t)t * N + nn] = ret;
                sa += a; saa += (double)a * a; sr += ret; srr += (double)ret * ret;
            }
        }
        cnt = (double)L;
    }
    red[0][threadIdx.x] = sa; red[1][threadIdx.x] = saa; red[2][threadIdx.x] = sr; red[3][threadIdx.x] = srr;
    red[4][threadIdx.x] = cnt;
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0.0;
        for (int i = 0; i < ADV_THREADS; ++i) s += red[threadIdx.x][i];
        partial[(int64_t)blockIdx.x * 5 + threadIdx.x] = s;
    }
}

__global__ void adv_ppo_sums_kernel(int nblocks, const double *__restrict__ partial, double *__restrict__ sums) {
    // single thread, fixed order: nblocks is N/128, a few thousand adds.
    // sums = (sum adv, sum adv^2, sum ret, sum ret^2, n valid): additive over ranks, so a sharded run
    // allreduces these five doubles before normalising
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s[5] = {0, 0, 0, 0, 0};
        for (int b = 0; b < nblocks; ++b)
            for (int k = 0; k < 5; ++k) s[k] += partial[(int64_t)b * 5 + k];
        for (int k = 0; k < 5; ++k) sums[k] = s[k];
    }
}

__global__ void __launch_bounds__(ADV_THREADS)
adv_ppo_norm_kernel(int64_t N, int T, const int32_t *__restrict__ len, const double *__restrict__ sums,
                    float *__restrict__ adv, float *__restrict__ rtg) {
    const int64_t n = (int64_t)blockIdx.x * ADV_THREADS + threadIdx.x;
    if (n >= N) return;
    const int L = len[n];
    // mean / unbiased std in double from the five sums, then rounded to fp32 like torch's results
    const double cnt = sums[4];
    const double dma = sums[0] / cnt, dmr = sums[2] / cnt;
    const double va = (sums[1] - sums[0] * dma) / (cnt - 1.0), vr = (sums[3] - sums[2] * dmr) / (cnt - 1.0);
    const float ma = (float)dma, mr = (float)dmr;
    const float sa = __fadd_rn((float)sqrt(va > 0.0 || va != va ? va : 0.0), 1e-8f);
    const float sr = __fadd_rn((float)sqrt(vr > 0.0 || vr != vr ? vr : 0.0), 1e-8f);
    for (int t = 0; t < L; ++t) {
        const int64_t i = (int64_t)t * N + n;
        adv[i] = __fdiv_rn(__fsub_rn(adv[i], ma), sa);    // ppo.py:138
        rtg[i] = __fdiv_rn(__fsub_rn(rtg[i], mr), sr);    // ppo.py:139
    }
}

extern "C" int64_t tg_advantage_workspace_bytes(int64_t N, int G) {
    (void)G;
    const int64_t blocks = (N + ADV_THREADS - 1) / ADV_THREADS;
    return blocks * 5 * (int64_t)sizeof(double) + 5 * (int64_t)sizeof(double) + 64;
}

extern "C" int tg_advantage(tg_ctx *ctx, int mode, int64_t G, int E, int T, double gamma, double lam, const float *rew,
                            const int32_t *len, const float *values, float *out_adv, float *out_rtg, void *workspace,
                            void *stream) {
    TgRange nvtx_range("tg_advantage (K2: reward-to-go + advantages)");
    TG_REQUIRE(ctx && rew && len && out_adv, TG_ERR_ARG, "tg_advantage: null argument");
    TG_REQUIRE(G > 0 && E > 0 && T > 0, TG_ERR_SHAPE, "tg_advantage: G, E, T must be positive");
    TG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t N = G * (int64_t)E;
    if (mode == TG_ADV_GRPO) {
        const int GC = E >= ADV_THREADS ? 1 : ADV_THREADS / E;
        const size_t smem = (size_t)GC * E * 3 * sizeof(double) + (size_t)GC * 2 * sizeof(float);
        TG_REQUIRE(smem <= (size_t)ctx->smem_optin, TG_ERR_UNSUPPORTED, "group size %d too large", E);
        if (smem > 48 * 1024)
            TG_CUDA(cudaFuncSetAttribute(adv_grpo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)((G + GC - 1) / GC);
        adv_grpo_kernel<<<grid, ADV_THREADS, smem, st>>>(G, E, T, GC, (float)gamma, rew, len, out_adv, out_rtg);
        TG_CUDA(cudaGetLastError());
        return TG_OK;
    }
    TG_REQUIRE(mode == TG_ADV_PPO_MC || mode == TG_ADV_PPO_GAE, TG_ERR_ARG, "unknown advantage mode %d", mode);
    TG_REQUIRE(values && out_rtg && workspace, TG_ERR_ARG, "PPO advantages need values, out_rtg and workspace");
    const unsigned grid = (unsigned)((N + ADV_THREADS - 1) / ADV_THREADS);
    double *sums = reinterpret_cast<double *>(workspace) + (size_t)grid * 5;
    int rc = tg_advantage_ppo_raw(ctx, mode, G, E, T, gamma, lam, rew, len, values, out_adv, out_rtg, sums, workspace,
                                  stream);
    if (rc) return rc;
    return tg_advantage_ppo_normalize(ctx, N, T, len, sums, out_adv, out_rtg, stream);
}

extern "C" int tg_advantage_ppo_raw(tg_ctx *ctx, int mode, int64_t G, int E, int T, double gamma, double lam,
                                    const float *rew, const int32_t *len, const float *values, float *out_adv,
                                    float *out_rtg, double *out_sums, void *workspace, void *stream) {
    TgRange nvtx_range("tg_advantage_ppo_raw (K2)");
    TG_REQUIRE(ctx && rew && len && values && out_adv && out_rtg && out_sums && workspace, TG_ERR_ARG,
               "tg_advantage_ppo_raw: null argument");
    TG_REQUIRE(mode == TG_ADV_PPO_MC || mode == TG_ADV_PPO_GAE, TG_ERR_ARG, "unknown PPO advantage mode %d", mode);
    TG_REQUIRE(G > 0 && E > 0 && T > 0, TG_ERR_SHAPE, "tg_advantage_ppo_raw: G, E, T must be positive");
    TG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t N = G * (int64_t)E;
    const unsigned grid = (unsigned)((N + ADV_THREADS - 1) / ADV_THREADS);
    double *partial = reinterpret_cast<double *>(workspace);
    adv_ppo_scan_kernel<<<grid, ADV_THREADS, 0, st>>>(N, T, mode == TG_ADV_PPO_GAE, (float)gamma,
                                                       (float)(gamma * lam), rew, len, values,
                                                       out_adv, out_rtg, partial);
    adv_ppo_sums_kernel<<<1, 32, 0, st>>>((int)grid, partial, out_sums);
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

extern "C" int tg_advantage_ppo_normalize(tg_ctx *ctx, int64_t N, int T, const int32_t *len, const double *sums,
                                          float *adv, float *rtg, void *stream) {
    TgRange nvtx_range("tg_advantage_ppo_normalize (K2)");
    TG_REQUIRE(ctx && len && sums && adv && rtg, TG_ERR_ARG, "tg_advantage_ppo_normalize: null argument");
    TG_REQUIRE(N > 0 && T > 0, TG_ERR_SHAPE, "tg_advantage_ppo_normalize: N, T must be positive");
    TG_CUDA(cudaSetDevice(ctx->device));
    const unsigned grid = (unsigned)((N + ADV_THREADS - 1) / ADV_THREADS);
    adv_ppo_norm_kernel<<<grid, ADV_THREADS, 0, (cudaStream_t)stream>>>(N, T, len, sums, adv, rtg);
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}
