// Trajectory export (buffers/rollout_buffer.py:72-102, Rollout_Buffer.save_trajectory): the reference
// walks every episode on the host and stacks its valid steps.  Here the valid (episode, step) rows are
// compacted on the device into one dense table -- row off[n] + t holds step t of episode n -- so the
// host copies exactly the bytes the CSV needs (one D2H) and never touches the zero padding.
#include "tg_common.cuh"

// One thread per (t, n) slot.  obs [T][O][N], act [T][A][N] are read coalesced along n; a thread writes its
// row of O + A floats (contiguous per thread).
__global__ void __launch_bounds__(256) export_rows_kernel(int64_t N, int T, int O, int A, const float *__restrict__ obs,
                                                          const float *__restrict__ act, const int32_t *__restrict__ len,
                                                          const int64_t *__restrict__ row0, int32_t *__restrict__ ids,
                                                          float *__restrict__ rows) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (n >= N || t >= len[n]) return;
    const int64_t r = row0[n] + t;
    ids[r] = (int32_t)n;
    float *dst = rows + r * (O + A);
    for (int i = 0; i < O; ++i) dst[i] = obs[((int64_t)t * O + i) * N + n];
    for (int j = 0; j < A; ++j) dst[O + j] = act[((int64_t)t * A + j) * N + n];
}

extern "C" int tg_export_trajectory(tg_ctx *ctx, int64_t N, int T, int O, int A, const float *obs, const float *act,
                                    const int32_t *len, const int64_t *row0, int32_t *out_episode_id, float *out_rows,
                                    void *stream) {
    TG_REQUIRE(ctx && obs && act && len && row0 && out_episode_id && out_rows, TG_ERR_ARG,
               "tg_export_trajectory: null argument");
    TG_REQUIRE(N > 0 && T > 0 && T <= 65535 && O > 0 && A > 0, TG_ERR_SHAPE, "tg_export_trajectory: bad shape");
    TG_CUDA(cudaSetDevice(ctx->device));
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)T);
    export_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(N, T, O, A, obs, act, len, row0, out_episode_id, out_rows);
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}
