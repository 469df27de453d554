// Packed-table utilities: init / pack / unpack / last_step access / lazy-state flush.
#include <stdarg.h>

#include "glove_common.cuh"

namespace glove {

static thread_local char g_err[512] = "";
char *err_buf() { return g_err; }
int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

__global__ void table_init_kernel(float *table, int64_t V, int32_t d, int32_t S, int32_t P, int32_t opt, int32_t side) {
    const int64_t total = V * P * S;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int32_t c = (int32_t)(i % S);
        int32_t p = (int32_t)((i / S) % P);
        float val = 0.0f;
        if (opt == GLOVE_OPT_ADAGRAD && p == 1 && (c < d || c == bias_col(d, side))) val = kAdagradInit;
        table[i] = val;
    }
}

// one thread per (row, column<=d) element; coalesced on the packed side
__global__ void pack_plane_kernel(float *table, int64_t V, int32_t d, int32_t S, int32_t P, int32_t plane, int32_t bcol,
                                  const float *__restrict__ emb, const float *__restrict__ bias) {
    const int64_t total = V * (d + 1);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / (d + 1);
        int32_t c = (int32_t)(i % (d + 1));
        float *dst = table + (r * P + plane) * S;
        if (c < d) dst[c] = emb[r * d + c];
        else if (bias) dst[bcol] = bias[r];
    }
}
__global__ void unpack_plane_kernel(const float *__restrict__ table, int64_t V, int32_t d, int32_t S, int32_t P,
                                    int32_t plane, int32_t bcol, float *emb, float *bias) {
    const int64_t total = V * (d + 1);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / (d + 1);
        int32_t c = (int32_t)(i % (d + 1));
        const float *src = table + (r * P + plane) * S;
        if (c < d) { if (emb) emb[r * d + c] = src[c]; }
        else if (bias) bias[r] = src[bcol];
    }
}
__global__ void get_ls_kernel(const float *table, int64_t V, int32_t lcol, int32_t S, int32_t P, int32_t *out) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < V; r += (int64_t)gridDim.x * blockDim.x)
        out[r] = __float_as_int(table[r * P * S + lcol]);
}
__global__ void set_ls_kernel(float *table, int64_t V, int32_t lcol, int32_t S, int32_t P, const int32_t *in) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < V; r += (int64_t)gridDim.x * blockDim.x)
        table[r * P * S + lcol] = __int_as_float(in[r]);
}

// Flush: one warp per row, lanes stride the columns.  Rows with last_step == 0 were never updated (m = v = 0): the
// idle step is an exact no-op on them, so they are skipped and keep last_step == 0.
__global__ void __launch_bounds__(256) flush_kernel(float *table, int64_t V, int32_t d, int32_t S, int32_t side,
                                                    const float *__restrict__ alpha, int32_t to_step, float b1,
                                                    float b2, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < V; r += nwarps) {
        float *x = table + r * 3 * S, *m = x + S, *v = m + S;
        const int32_t lcol = ls_col(d, side);
        const int32_t ls = __float_as_int(x[lcol]);
        if (ls <= 0 || ls >= to_step) continue;
        for (int32_t cc = lane; cc <= d; cc += 32) {
            const int32_t c = cc < d ? cc : bias_col(d, side);
            float xv = x[c], mv = m[c], vv = v[c];
            for (int32_t s = ls; s < to_step; ++s) adam_idle_step(xv, mv, vv, alpha[s], b1, b2, eps);
            x[c] = xv; m[c] = mv; v[c] = vv;
        }
        __syncwarp();
        if (lane == 0) x[lcol] = __int_as_float(to_step);
    }
}

// Closed-form flush (GLOVE_ADAM_REPLAY): same per-element arithmetic as stage_closed_kernel (replay_x4), written back
// to the table together with the decayed moments.  One warp per row, lanes stride the float4 columns.
__global__ void __launch_bounds__(256) flush_closed_kernel(float *table, int64_t V, int32_t d, int32_t S, int32_t side,
                                                           const float *__restrict__ alpha, int32_t to_step, float b1,
                                                           float b2, float l2b1, float l2b2, float eps) {
    __shared__ ReplayTables tabs;
    replay_tables_init(tabs, b1, b2);
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int32_t lcol = ls_col(d, side), S4 = S >> 2;
    for (int64_t r = warp; r < V; r += nwarps) {
        float *x = table + r * 3 * S, *m = x + S, *v = m + S;
        const int32_t ls = __float_as_int(x[lcol]);
        if (ls <= 0 || ls >= to_step) continue;
        const int gap = to_step - ls;
        const ReplayCoef c = replay_coef(tabs, alpha, ls, gap, l2b1, l2b2, lane);
        for (int32_t f = lane; f < S4; f += 32) {
            float4 xv = ld4(x + 4 * f);
            const float4 mv = ld4(m + 4 * f), vv = ld4(v + 4 * f);
            xv = replay_x4(xv, mv, vv, c, eps);
            if ((lcol >> 2) == f) f4c(xv, lcol & 3) = __int_as_float(to_step);
            st4(x + 4 * f, xv);
            st4(m + 4 * f, scale4(mv, c.dm));
            st4(v + 4 * f, scale4(vv, c.dv));
        }
    }
}

}  // namespace glove

using namespace glove;

extern "C" {

const char *glove_last_error(void) { return err_buf(); }
int32_t glove_abi_version(void) { return GLOVE_B200_ABI_VERSION; }
size_t glove_step_args_size(void) { return sizeof(glove_step_args); }
int32_t glove_table_stride(int32_t d) { return table_stride(d); }
int32_t glove_table_planes(int32_t optimizer) { return table_planes(optimizer); }

static inline int grid_for(int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads;
    int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

int glove_table_init(float *table, int64_t V, int32_t d, int32_t optimizer, int32_t side, void *stream) {
    GLOVE_REQUIRE(table && V > 0 && d > 0 && (side == 0 || side == 1), "glove_table_init: bad arguments");
    GLOVE_REQUIRE(optimizer >= 0 && optimizer <= 2, "glove_table_init: unsupported optimizer %d", optimizer);
    const int32_t S = table_stride(d), P = table_planes(optimizer);
    table_init_kernel<<<grid_for(V * P * S, 256), 256, 0, (cudaStream_t)stream>>>(table, V, d, S, P, optimizer, side);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}

int glove_pack_plane(float *table, int64_t V, int32_t d, int32_t planes, int32_t plane, int32_t side, const float *emb,
                     const float *bias, void *stream) {
    GLOVE_REQUIRE(table && emb && V > 0 && d > 0 && plane >= 0 && plane < planes && (side == 0 || side == 1),
                  "glove_pack_plane: bad arguments");
    pack_plane_kernel<<<grid_for(V * (d + 1), 256), 256, 0, (cudaStream_t)stream>>>(
        table, V, d, table_stride(d), planes, plane, bias_col(d, side), emb, bias);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}

int glove_unpack_plane(const float *table, int64_t V, int32_t d, int32_t planes, int32_t plane, int32_t side,
                       float *emb, float *bias, void *stream) {
    GLOVE_REQUIRE(table && (emb || bias) && V > 0 && d > 0 && plane >= 0 && plane < planes && (side == 0 || side == 1),
                  "glove_unpack_plane: bad arguments");
    unpack_plane_kernel<<<grid_for(V * (d + 1), 256), 256, 0, (cudaStream_t)stream>>>(
        table, V, d, table_stride(d), planes, plane, bias_col(d, side), emb, bias);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}

int glove_get_last_step(const float *table, int64_t V, int32_t d, int32_t planes, int32_t side, int32_t *out,
                        void *stream) {
    GLOVE_REQUIRE(table && out && V > 0 && d > 0 && planes >= 1 && (side == 0 || side == 1), "glove_get_last_step: bad arguments");
    get_ls_kernel<<<grid_for(V, 256), 256, 0, (cudaStream_t)stream>>>(table, V, ls_col(d, side), table_stride(d), planes, out);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}
int glove_set_last_step(float *table, int64_t V, int32_t d, int32_t planes, int32_t side, const int32_t *in,
                        void *stream) {
    GLOVE_REQUIRE(table && in && V > 0 && d > 0 && planes >= 1 && (side == 0 || side == 1), "glove_set_last_step: bad arguments");
    set_ls_kernel<<<grid_for(V, 256), 256, 0, (cudaStream_t)stream>>>(table, V, ls_col(d, side), table_stride(d), planes, in);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}

int glove_flush_lazy_state(float *table, int64_t V, int32_t d, int32_t optimizer, int32_t side, const float *alpha,
                           int32_t alpha_len, int32_t to_step, float beta1, float beta2, float epsilon, void *stream) {
    GLOVE_REQUIRE(table && V > 0 && d > 0 && (side == 0 || side == 1), "glove_flush_lazy_state: bad arguments");
    if (optimizer != GLOVE_OPT_ADAM) return GLOVE_OK;  // Adagrad / SGD are truly sparse: nothing to replay
    GLOVE_REQUIRE(alpha && to_step <= alpha_len, "glove_flush_lazy_state: alpha table too short (%d < %d)", alpha_len,
                  to_step);
    if (to_step <= 0) return GLOVE_OK;
    flush_closed_kernel<<<grid_for(V * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        table, V, d, table_stride(d), side, alpha, to_step, beta1, beta2, replay_log2(beta1), replay_log2(beta2), epsilon);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}

int glove_flush_lazy_state_exact(float *table, int64_t V, int32_t d, int32_t optimizer, int32_t side, const float *alpha,
                                 int32_t alpha_len, int32_t to_step, float beta1, float beta2, float epsilon, void *stream) {
    GLOVE_REQUIRE(table && V > 0 && d > 0 && (side == 0 || side == 1), "glove_flush_lazy_state: bad arguments");
    if (optimizer != GLOVE_OPT_ADAM) return GLOVE_OK;  // Adagrad / SGD are truly sparse: nothing to replay
    GLOVE_REQUIRE(alpha && to_step <= alpha_len, "glove_flush_lazy_state: alpha table too short (%d < %d)", alpha_len,
                  to_step);
    if (to_step <= 0) return GLOVE_OK;
    flush_kernel<<<grid_for(V * 32, 256), 256, 0, (cudaStream_t)stream>>>(table, V, d, table_stride(d), side, alpha,
                                                                         to_step, beta1, beta2, epsilon);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}

}  // extern "C"
