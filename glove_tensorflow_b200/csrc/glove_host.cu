// HOST-buffer boundary: the end-to-end entry point a non-torch caller of the reference's training path would bind.
// One call = K TRAIN steps on K*B explicit triples that live in HOST memory: H2D copies, plan construction, K steps,
// D2H copy of the K losses -- what K iterations of `session.run(train_op)` fed by input_fn do in the reference
// [ref src/models/estimator.py:79-95, src/models/data_utils.py:4-26].
#include <new>

#include "glove_common.cuh"

namespace glove {
__global__ void iota64_kernel(int64_t *out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = i;
}
__global__ void read_step_kernel(const glove_scalars *sc, int32_t *out) { *out = sc->step; }
constexpr int kHostLossCap = 4096;   // most steps one call may run
struct HostStaging {
    int32_t *row[2], *col[2];   // double-buffered COO of one plan chunk (plan_K * B triples)
    float *a[2], *b[2];
    int64_t *idx;
    float *losses;              // [kHostLossCap]
    int32_t *step;
    size_t bytes;
};
static HostStaging staging_view(void *base, int32_t K, int32_t B) {
    HostStaging v;
    char *p = (char *)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char *r = p ? p + off : nullptr; off += align_up(bytes); return r; };
    const int64_t N = (int64_t)K * B;
    for (int i = 0; i < 2; ++i) {
        v.row[i] = (int32_t *)take(4 * N); v.col[i] = (int32_t *)take(4 * N);
        v.a[i] = (float *)take(4 * N); v.b[i] = (float *)take(4 * N);
    }
    v.idx = (int64_t *)take(8 * N);
    v.losses = (float *)take(4 * (size_t)kHostLossCap);
    v.step = (int32_t *)take(4);
    v.bytes = off;
    return v;
}

}  // namespace glove
using namespace glove;

// helper streams / events of the host entry: CALLER-OWNED (glove_host_pipe_create / _destroy), one per concurrent caller
// and device -- the library keeps no process-global state, so two pipes can drive two streams or devices at once
struct glove_host_pipe {
    int device = -1;
    cudaStream_t copy = nullptr, side = nullptr;
    cudaEvent_t plan_ready[2] = {nullptr, nullptr}, chunk_done[2] = {nullptr, nullptr}, step_done[2] = {nullptr, nullptr};
    cudaEvent_t caught_up = nullptr;
};

extern "C" {

int glove_host_pipe_create(glove_host_pipe **out) {
    GLOVE_REQUIRE(out, "glove_host_pipe_create: null output");
    *out = nullptr;
    glove_host_pipe *hp = new (std::nothrow) glove_host_pipe();
    GLOVE_REQUIRE(hp, "glove_host_pipe_create: out of memory");
    cudaError_t e = cudaGetDevice(&hp->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hp->copy, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hp->side, cudaStreamNonBlocking);
    cudaEvent_t *evs[] = {&hp->plan_ready[0], &hp->plan_ready[1], &hp->chunk_done[0], &hp->chunk_done[1],
                          &hp->step_done[0], &hp->step_done[1], &hp->caught_up};
    for (cudaEvent_t *ev : evs)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        glove_host_pipe_destroy(hp);
        return set_error(GLOVE_ECUDA, "glove_host_pipe_create: %s", cudaGetErrorString(e));
    }
    *out = hp;
    return GLOVE_OK;
}

int glove_host_pipe_destroy(glove_host_pipe *hp) {
    if (!hp) return GLOVE_OK;
    cudaEvent_t evs[] = {hp->plan_ready[0], hp->plan_ready[1], hp->chunk_done[0], hp->chunk_done[1],
                         hp->step_done[0], hp->step_done[1], hp->caught_up};
    for (cudaEvent_t ev : evs)
        if (ev) cudaEventDestroy(ev);
    if (hp->copy) cudaStreamDestroy(hp->copy);
    if (hp->side) cudaStreamDestroy(hp->side);
    delete hp;
    return GLOVE_OK;
}

size_t glove_host_staging_bytes(int32_t K, int32_t B) {
    if (K <= 0 || B <= 0) return 0;
    return staging_view(nullptr, K, B).bytes;
}
size_t glove_host_plan_bytes(int32_t K, int32_t B) {
    if (K <= 0 || B <= 0) return 0;
    return 2 * align_up(glove_plan_bytes(K, B));
}

// Pipeline over the n = K / plan_K chunks of a call (chunk c uses plan / staging buffer c & 1):
//   copy stream : H2D(c) -> glove_prepare_batches(c)            (waits until the steps of chunk c-2 released the buffers)
//   main stream : plan_K x glove_train_step(c)                   (waits for plan_ready[c])
//   side stream : glove_catchup_step(s+1) while step s runs      (exact-replay Adam only; needs step s-1 finished)
int glove_train_steps_host(glove_host_pipe *hp, const glove_step_args *args, void *plan, void *prepare_ws, size_t prepare_ws_bytes,
                           void *staging, size_t staging_bytes, const int32_t *host_row, const int32_t *host_col,
                           const float *host_colA, const float *host_colB, int32_t K, float *host_losses,
                           void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GLOVE_REQUIRE(hp && args && plan && prepare_ws && staging && host_row && host_col && host_colA && host_colB && host_losses,
                  "glove_train_steps_host: null pointer");
    GLOVE_REQUIRE(args->struct_size == sizeof(glove_step_args), "glove_train_steps_host: glove_step_args.struct_size mismatch");
    {
        int dev = -1;
        GLOVE_CHECK_CUDA(cudaGetDevice(&dev));
        GLOVE_REQUIRE(dev == hp->device, "glove_train_steps_host: pipe was created on device %d, current device is %d", hp->device, dev);
    }
    const int32_t PK = args->plan_K;
    GLOVE_REQUIRE(PK > 0 && K > 0 && K % PK == 0 && K <= kHostLossCap,
                  "glove_train_steps_host: K must be a multiple of args->plan_K, at most %d", kHostLossCap);
    HostStaging st = staging_view(staging, PK, args->B);
    if (staging_bytes < st.bytes)
        return set_error(GLOVE_EWORKSPACE, "glove_train_steps_host: staging %zu < required %zu", staging_bytes, st.bytes);
    const int64_t N = (int64_t)PK * args->B;
    const int n_chunks = K / PK;
    void *plans[2] = {plan, (char *)plan + align_up(glove_plan_bytes(PK, args->B))};
    iota64_kernel<<<kNumSMs, 256, 0, stream>>>(st.idx, N);
    // the plans' first_step must equal the device step counter: read it back (tiny D2H, part of the e2e cost)
    int32_t first_step = 0;
    read_step_kernel<<<1, 1, 0, stream>>>(args->scalars, st.step);
    GLOVE_CHECK_CUDA(cudaMemcpyAsync(&first_step, st.step, 4, cudaMemcpyDeviceToHost, stream));
    GLOVE_CHECK_CUDA(cudaStreamSynchronize(stream));   // everything enqueued before this call has finished
    const bool catchup = args->optimizer == GLOVE_OPT_ADAM && args->adam_mode == GLOVE_ADAM_REPLAY_EXACT && args->n_shards <= 1 &&
                         args->dp_world <= 1;
    glove_step_args a = *args;
    a.loss_out = st.losses;
    a.loss_cap = K;  // loss of step s lands in losses[s % K]
    bool have_prev = false, pending_catchup = false;
    auto enqueue_plan = [&](int c) -> int {
        const int b = c & 1;
        if (c >= 2) GLOVE_CHECK_CUDA(cudaStreamWaitEvent(hp->copy, hp->chunk_done[b], 0));
        const int64_t o = (int64_t)c * N;
        GLOVE_CHECK_CUDA(cudaMemcpyAsync(st.row[b], host_row + o, 4 * N, cudaMemcpyHostToDevice, hp->copy));
        GLOVE_CHECK_CUDA(cudaMemcpyAsync(st.col[b], host_col + o, 4 * N, cudaMemcpyHostToDevice, hp->copy));
        GLOVE_CHECK_CUDA(cudaMemcpyAsync(st.a[b], host_colA + o, 4 * N, cudaMemcpyHostToDevice, hp->copy));
        GLOVE_CHECK_CUDA(cudaMemcpyAsync(st.b[b], host_colB + o, 4 * N, cudaMemcpyHostToDevice, hp->copy));
        int rc = glove_prepare_batches(plans[b], prepare_ws, prepare_ws_bytes, st.row[b], st.col[b], st.a[b], st.b[b], N,
                                       st.idx, 0, 0, first_step + c * PK, PK, args->B, (int32_t)args->V, hp->copy);
        if (rc != GLOVE_OK) return rc;
        GLOVE_CHECK_CUDA(cudaEventRecord(hp->plan_ready[b], hp->copy));
        return GLOVE_OK;
    };
    if (int rc = enqueue_plan(0)) return rc;
    for (int c = 0; c < n_chunks; ++c) {
        if (c + 1 < n_chunks)
            if (int rc = enqueue_plan(c + 1)) return rc;
        a.plan = plans[c & 1];
        GLOVE_CHECK_CUDA(cudaStreamWaitEvent(stream, hp->plan_ready[c & 1], 0));
        for (int32_t k = 0; k < PK; ++k) {
            const int32_t s = first_step + c * PK + k;
            if (pending_catchup) {
                GLOVE_CHECK_CUDA(cudaStreamWaitEvent(stream, hp->caught_up, 0));
                pending_catchup = false;
            }
            if (catchup && k + 1 < PK && s + 1 < args->alpha_len) {
                if (have_prev) GLOVE_CHECK_CUDA(cudaStreamWaitEvent(hp->side, hp->step_done[(s - 1) & 1], 0));
                else {   // first step of the call: the plan must be there; earlier steps were synchronised above
                    GLOVE_CHECK_CUDA(cudaStreamWaitEvent(hp->side, hp->plan_ready[c & 1], 0));
                }
                if (k == 0) GLOVE_CHECK_CUDA(cudaStreamWaitEvent(hp->side, hp->plan_ready[c & 1], 0));
                int rc = glove_catchup_step(&a, s + 1, hp->side);
                if (rc != GLOVE_OK) return rc;
                GLOVE_CHECK_CUDA(cudaEventRecord(hp->caught_up, hp->side));
                pending_catchup = true;
            }
            int rc = glove_train_step(&a, stream);
            if (rc != GLOVE_OK) return rc;
            GLOVE_CHECK_CUDA(cudaEventRecord(hp->step_done[s & 1], stream));
            have_prev = true;
        }
        GLOVE_CHECK_CUDA(cudaEventRecord(hp->chunk_done[c & 1], stream));
    }
    if (pending_catchup) GLOVE_CHECK_CUDA(cudaStreamWaitEvent(stream, hp->caught_up, 0));
    GLOVE_CHECK_CUDA(cudaMemcpyAsync(host_losses, st.losses, 4 * (size_t)K, cudaMemcpyDeviceToHost, stream));
    GLOVE_CHECK_CUDA(cudaStreamSynchronize(stream));
    if (first_step % K != 0) {  // rotate so that host_losses[k] is the loss of the k-th step of this call
        float tmp[kHostLossCap];
        for (int k = 0; k < K; ++k) tmp[k] = host_losses[(first_step + k) % K];
        memcpy(host_losses, tmp, 4 * (size_t)K);
    }
    return GLOVE_OK;
}

}  // extern "C"
