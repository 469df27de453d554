// HOST-buffer boundary: the end-to-end entry point a non-torch caller of the reference's training path would bind.
// One call = K TRAIN steps on K*B explicit triples that live in HOST memory: H2D copies, plan construction, K steps,
// D2H copy of the K losses -- what K iterations of `session.run(train_op)` fed by input_fn do in the reference
// [ref src/models/estimator.py:79-95, src/models/data_utils.py:4-26].
#include "glove_common.cuh"

namespace glove {
__global__ void iota64_kernel(int64_t *out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = i;
}
__global__ void read_step_kernel(const glove_scalars *sc, int32_t *out) { *out = sc->step; }
struct HostStaging {
    int32_t *row, *col;
    float *a, *b;
    int64_t *idx;
    float *losses;
    int32_t *step;
    size_t bytes;
};
static HostStaging staging_view(void *base, int32_t K, int32_t B) {
    HostStaging v;
    char *p = (char *)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char *r = p ? p + off : nullptr; off += align_up(bytes); return r; };
    const int64_t N = (int64_t)K * B;
    v.row = (int32_t *)take(4 * N); v.col = (int32_t *)take(4 * N);
    v.a = (float *)take(4 * N); v.b = (float *)take(4 * N);
    v.idx = (int64_t *)take(8 * N);
    v.losses = (float *)take(4 * (size_t)K);
    v.step = (int32_t *)take(4);
    v.bytes = off;
    return v;
}
}  // namespace glove
using namespace glove;

extern "C" {

size_t glove_host_staging_bytes(int32_t K, int32_t B) {
    if (K <= 0 || B <= 0) return 0;
    return staging_view(nullptr, K, B).bytes;
}

int glove_train_steps_host(const glove_step_args *args, void *plan, void *prepare_ws, size_t prepare_ws_bytes,
                           void *staging, size_t staging_bytes, const int32_t *host_row, const int32_t *host_col,
                           const float *host_colA, const float *host_colB, int32_t K, float *host_losses,
                           void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GLOVE_REQUIRE(args && plan && prepare_ws && staging && host_row && host_col && host_colA && host_colB && host_losses,
                  "glove_train_steps_host: null pointer");
    GLOVE_REQUIRE(K > 0 && K == args->plan_K, "glove_train_steps_host: K must equal args->plan_K");
    HostStaging st = staging_view(staging, K, args->B);
    if (staging_bytes < st.bytes)
        return set_error(GLOVE_EWORKSPACE, "glove_train_steps_host: staging %zu < required %zu", staging_bytes, st.bytes);
    const int64_t N = (int64_t)K * args->B;
    GLOVE_CHECK_CUDA(cudaMemcpyAsync(st.row, host_row, 4 * N, cudaMemcpyHostToDevice, stream));
    GLOVE_CHECK_CUDA(cudaMemcpyAsync(st.col, host_col, 4 * N, cudaMemcpyHostToDevice, stream));
    GLOVE_CHECK_CUDA(cudaMemcpyAsync(st.a, host_colA, 4 * N, cudaMemcpyHostToDevice, stream));
    GLOVE_CHECK_CUDA(cudaMemcpyAsync(st.b, host_colB, 4 * N, cudaMemcpyHostToDevice, stream));
    iota64_kernel<<<kNumSMs, 256, 0, stream>>>(st.idx, N);
    // the plan's first_step must equal the device step counter: read it back (tiny D2H, part of the e2e cost)
    int32_t first_step = 0;
    read_step_kernel<<<1, 1, 0, stream>>>(args->scalars, st.step);
    GLOVE_CHECK_CUDA(cudaMemcpyAsync(&first_step, st.step, 4, cudaMemcpyDeviceToHost, stream));
    GLOVE_CHECK_CUDA(cudaStreamSynchronize(stream));
    int rc = glove_prepare_batches(plan, prepare_ws, prepare_ws_bytes, st.row, st.col, st.a, st.b, N, st.idx, 0, 0,
                                   first_step, K, args->B, (int32_t)args->V, stream);
    if (rc != GLOVE_OK) return rc;
    glove_step_args a = *args;
    a.plan = plan;
    a.loss_out = st.losses;
    a.loss_cap = K;  // loss of step s lands in losses[s % K]
    for (int32_t k = 0; k < K; ++k) {
        rc = glove_train_step(&a, stream);
        if (rc != GLOVE_OK) return rc;
    }
    GLOVE_CHECK_CUDA(cudaMemcpyAsync(host_losses, st.losses, 4 * (size_t)K, cudaMemcpyDeviceToHost, stream));
    GLOVE_CHECK_CUDA(cudaStreamSynchronize(stream));
    if (first_step % K != 0) {  // rotate so that host_losses[k] is the loss of the k-th step of this call
        float tmp[1024];
        if (K <= 1024) {
            for (int k = 0; k < K; ++k) tmp[k] = host_losses[(first_step + k) % K];
            memcpy(host_losses, tmp, 4 * (size_t)K);
        }
    }
    return GLOVE_OK;
}

}  // extern "C"
