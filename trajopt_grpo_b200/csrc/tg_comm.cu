// Gradient allreduce fused with the Adam step over NVLink peer memory (include/trajopt_grpo.h: tg_comm_*,
// tg_allreduce_adam_step).
//
// The GRPO / PPO path has exactly one exchange per optimizer step: the flat policy gradient (<= 290 KB) summed
// over the ranks.  Issued through NCCL from Python that is a collective launch + a separate Adam
// launch per update; here every rank owns one cudaMalloc'ed window that its peers map through CUDA IPC
// (one process per GPU, NVSwitch gives every pair full bandwidth):
//
//     window = [ grad slot 0 | grad slot 1 | flags[world] ]        (slots alternate with the step parity)
//
//   1. K3's grad_reduce_kernel writes the rank's gradient straight into its own slot (the caller passes the slot as
//      tg_policy_grad's out_grad: no copy);
//   2. comm_post_kernel: one thread per peer publishes "my gradient of step e is complete" with a system-scope
//      release store into the PEER's flag array (a 4-byte NVLink write);
//   3. allreduce_adam_kernel: every CTA waits (system-scope acquire loads on its OWN flags) until all ranks have
//      published step e, then each thread sums its elements over the ranks' slots with peer loads IN RANK ORDER --
//      the same fp32 sum on every rank, so the weights stay bit-identical -- and applies Adam in the same pass.
//
// One-shot allreduce: every rank reads world x n floats (2.3 MB at world = 8), latency-bound like the NCCL call it
// replaces but without the extra launches and without a reduce-scatter/all-gather round trip.  Two slots make the
// barrier one-sided: a rank may run one step ahead of a slow peer (it then writes the OTHER slot); it cannot run
// two ahead because step e+1 cannot complete before the slow peer has published e+1, which it does only after it
// finished reading step e.
#include "tg_common.cuh"

struct tg_comm {
    tg_ctx *ctx;
    int rank, world;
    int64_t n;                       // floats per gradient slot
    size_t slot_bytes, window_bytes;
    unsigned char *local;            // this rank's window (cudaMalloc)
    unsigned char *peer[16];         // every rank's window as mapped here (peer[rank] == local)
    bool opened[16];
    uint32_t epoch;                  // steps completed
};

#define TG_COMM_MAX_WORLD 16
#define TG_COMM_FLAG_STRIDE 32       // uint32 per flag (128 B apart)

struct CommPtrs {
    const float *slot[TG_COMM_MAX_WORLD];
    uint32_t *flags[TG_COMM_MAX_WORLD];
};

TG_D void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
TG_D uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// thread q tells rank q that this rank's gradient of step `epoch` is in place
__global__ void comm_post_kernel(CommPtrs P, int rank, int world, uint32_t epoch) {
    const int q = threadIdx.x;
    if (q < world) {
        __threadfence_system();      // the gradient slot was written by the previous kernel on this stream
        st_release_sys(P.flags[q] + (size_t)rank * TG_COMM_FLAG_STRIDE, epoch);
    }
}

__global__ void __launch_bounds__(256)
allreduce_adam_kernel(CommPtrs P, int rank, int world, uint32_t epoch, int64_t n, float *__restrict__ p,
                      float *__restrict__ m, float *__restrict__ v, float *__restrict__ gsum, float step_size,
                      float sqrt_bc2, float w1, float b2, float w2, float eps) {
    if (threadIdx.x < world) {
        const uint32_t *f = P.flags[rank] + (size_t)threadIdx.x * TG_COMM_FLAG_STRIDE;
        while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) __nanosleep(64);
    }
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = 0.0f;
        for (int q = 0; q < world; ++q) gi = __fadd_rn(gi, __ldcv(P.slot[q] + i));     // rank order: identical everywhere
        if (gsum) gsum[i] = gi;
        // torch.optim.Adam, single-tensor formula (tg_adam_step)
        const float mi = __fadd_rn(m[i], __fmul_rn(w1, __fsub_rn(gi, m[i])));
        const float vi = __fadd_rn(__fmul_rn(v[i], b2), __fmul_rn(__fmul_rn(w2, gi), gi));
        m[i] = mi;
        v[i] = vi;
        const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), sqrt_bc2), eps);
        p[i] = __fsub_rn(p[i], __fmul_rn(step_size, __fdiv_rn(mi, denom)));
    }
}

extern "C" int tg_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int tg_comm_create(tg_ctx *ctx, int rank, int world, int64_t n_floats, tg_comm **out, void *handle_out) {
    TG_REQUIRE(ctx && out && handle_out, TG_ERR_ARG, "tg_comm_create: null argument");
    TG_REQUIRE(world >= 2 && world <= TG_COMM_MAX_WORLD && rank >= 0 && rank < world, TG_ERR_ARG,
               "tg_comm_create: rank %d / world %d out of range (2..%d ranks)", rank, world, TG_COMM_MAX_WORLD);
    TG_REQUIRE(n_floats > 0, TG_ERR_SHAPE, "tg_comm_create: n_floats must be positive");
    TG_CUDA(cudaSetDevice(ctx->device));
    tg_comm *c = new tg_comm();
    c->ctx = ctx; c->rank = rank; c->world = world; c->n = n_floats; c->epoch = 0;
    c->slot_bytes = ((size_t)n_floats * 4 + 255) / 256 * 256;
    c->window_bytes = 2 * c->slot_bytes + (size_t)TG_COMM_MAX_WORLD * TG_COMM_FLAG_STRIDE * 4;
    cudaError_t e = cudaMalloc((void **)&c->local, c->window_bytes);
    if (e != cudaSuccess) {
        delete c;
        tg_set_error("tg_comm_create: cudaMalloc(%zu) -> %s", c->window_bytes, cudaGetErrorString(e));
        return TG_ERR_CUDA;
    }
    TG_CUDA(cudaMemset(c->local, 0, c->window_bytes));
    TG_CUDA(cudaDeviceSynchronize());
    for (int q = 0; q < TG_COMM_MAX_WORLD; ++q) { c->peer[q] = nullptr; c->opened[q] = false; }
    c->peer[rank] = c->local;
    TG_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle_out, c->local));
    *out = c;
    return TG_OK;
}

extern "C" int tg_comm_connect(tg_comm *c, const void *all_handles) {
    TG_REQUIRE(c && all_handles, TG_ERR_ARG, "tg_comm_connect: null argument");
    TG_CUDA(cudaSetDevice(c->ctx->device));
    const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *)all_handles;
    for (int q = 0; q < c->world; ++q) {
        if (q == c->rank) continue;
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h[q], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            tg_set_error("tg_comm_connect: cannot map the window of rank %d (%s); peers need NVLink/PCIe P2P access", q,
                         cudaGetErrorString(e));
            return TG_ERR_CUDA;
        }
        c->peer[q] = (unsigned char *)p;
        c->opened[q] = true;
    }
    return TG_OK;
}

extern "C" int tg_comm_destroy(tg_comm *c) {
    if (!c) return TG_OK;
    cudaSetDevice(c->ctx->device);
    cudaDeviceSynchronize();
    for (int q = 0; q < c->world; ++q)
        if (c->opened[q]) cudaIpcCloseMemHandle(c->peer[q]);
    if (c->local) cudaFree(c->local);
    delete c;
    return TG_OK;
}

// device pointer of the gradient slot the NEXT tg_allreduce_adam_step will sum: pass it as tg_policy_grad's out_grad
extern "C" int tg_comm_grad_slot(tg_comm *c, float **out) {
    TG_REQUIRE(c && out, TG_ERR_ARG, "tg_comm_grad_slot: null argument");
    *out = (float *)(c->local + (size_t)((c->epoch + 1) & 1u) * c->slot_bytes);
    return TG_OK;
}

extern "C" int tg_allreduce_adam_step(tg_ctx *ctx, tg_comm *c, int64_t n, float *params, float *exp_avg, float *exp_avg_sq,
                                      int64_t step, double lr, double beta1, double beta2, double eps, float *out_grad_sum,
                                      void *stream) {
    TgRange nvtx_range("tg_allreduce_adam_step (peer-memory gradient allreduce + Adam)");
    TG_REQUIRE(ctx && c && params && exp_avg && exp_avg_sq, TG_ERR_ARG, "tg_allreduce_adam_step: null argument");
    TG_REQUIRE(n > 0 && n <= c->n && step >= 1, TG_ERR_SHAPE, "tg_allreduce_adam_step: n must be in (0, %lld], step >= 1",
               (long long)c->n);
    for (int q = 0; q < c->world; ++q)
        TG_REQUIRE(c->peer[q] != nullptr, TG_ERR_ARG, "tg_allreduce_adam_step: rank %d is not connected (tg_comm_connect)", q);
    TG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    c->epoch += 1;
    const uint32_t epoch = c->epoch;
    CommPtrs P;
    for (int q = 0; q < TG_COMM_MAX_WORLD; ++q) {
        const int qq = q < c->world ? q : c->rank;
        P.slot[q] = (const float *)(c->peer[qq] + (size_t)(epoch & 1u) * c->slot_bytes);
        P.flags[q] = (uint32_t *)(c->peer[qq] + 2 * c->slot_bytes);
    }
    comm_post_kernel<<<1, 32, 0, st>>>(P, c->rank, c->world, epoch);
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    int64_t blocks = (n + 255) / 256;
    if (blocks > 4 * ctx->sm_count) blocks = 4 * ctx->sm_count;
    allreduce_adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(P, c->rank, c->world, epoch, n, params, exp_avg, exp_avg_sq,
                                                            out_grad_sum, (float)(lr / bc1), (float)sqrt(bc2),
                                                            (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                                            (float)eps);
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}
