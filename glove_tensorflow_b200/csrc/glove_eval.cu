// EVAL: forward-only pass, model_fn(mode=EVAL) + RegressionHead / BinaryClassHead metric sums
// [ref src/models/estimator.py:87-92, src/models/train_utils.py:47-54].  Deterministic two-level reduction:
// one warp per tile of <= 32 consecutive triples (tiles never straddle a batch), then one warp per batch.
#include "glove_common.cuh"

namespace glove {

constexpr int kEvalTile = 32;

__device__ __forceinline__ float softplus_e(float x) { return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x))); }

template <int NV>
__global__ void __launch_bounds__(256) eval_tiles_kernel(const float *__restrict__ rt, const float *__restrict__ ct,
                                                         const glove_scalars *sc, int32_t P, int32_t d, int32_t S,
                                                         const int32_t *__restrict__ row, const int32_t *__restrict__ col,
                                                         const float *__restrict__ colA, const float *__restrict__ colB,
                                                         int64_t first, int64_t count, int32_t B, int32_t head,
                                                         int64_t tiles_per_batch, int64_t n_tiles, double *tile_sums) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int S4 = S >> 2;
    const float g = sc->g;
    for (int64_t t = warp; t < n_tiles; t += nwarps) {
        const int64_t batch = t / tiles_per_batch, within = t % tiles_per_batch;
        const int64_t lo = batch * B + within * kEvalTile;
        int64_t hi = lo + kEvalTile;
        if (hi > (batch + 1) * (int64_t)B) hi = (batch + 1) * (int64_t)B;
        if (hi > count) hi = count;
        double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int64_t n = lo; n < hi; ++n) {
            const int32_t i = row[first + n], j = col[first + n];
            const float a = colA[first + n], b = colB[first + n];
            const float *R = rt + (int64_t)i * P * S, *C = ct + (int64_t)j * P * S;
            float dot = 0.f, nr = 0.f, nc = 0.f, rb = 0.f, cb = 0.f;
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                const int f = lane + 32 * r;
                if (f < S4) {
                    const float4 x = ld4(R + 4 * f), y = ld4(C + 4 * f);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int cc = 4 * f + c;
                        const float xv = f4v(x, c), yv = f4v(y, c);
                        if (cc < d) { dot += xv * yv; nr += xv * xv; nc += yv * yv; }
                        else if (cc == d) rb = xv;          // row table: bias in column d
                        else if (cc == d + 1) cb = yv;      // col table: bias in column d + 1
                    }
                }
            }
            dot = warp_sum(dot); nr = warp_sum(nr); nc = warp_sum(nc); rb = warp_sum(rb); cb = warp_sum(cb);
            const float z = ((dot + rb) + cb) + g;
            if (head == GLOVE_HEAD_GLOVE) {
                const float r_ = z - a;
                s[0] += (double)(b * r_ * r_); s[1] += b; s[2] += (double)b * a; s[3] += (double)b * z;
            } else {
                s[0] += (double)(a * softplus_e(-z)); s[1] += a; s[2] += (double)(b * softplus_e(z)); s[3] += b;
            }
            s[4] += nr; s[5] += nc; s[6] += (double)rb * rb; s[7] += (double)cb * cb;
        }
        if (lane < 8) {
            double v = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) if (lane == q) v = s[q];
            tile_sums[t * 8 + lane] = v;
        }
    }
}

__global__ void eval_reduce_kernel(const double *__restrict__ tile_sums, int64_t tiles_per_batch, int64_t n_tiles,
                                   int64_t n_batches, double *out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp; b < n_batches; b += nwarps) {
        const int64_t t0 = b * tiles_per_batch;
        int64_t t1 = t0 + tiles_per_batch;
        if (t1 > n_tiles) t1 = n_tiles;
        double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int64_t t = t0 + lane; t < t1; t += 32)
#pragma unroll
            for (int q = 0; q < 8; ++q) s[q] += tile_sums[t * 8 + q];
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
        if (lane == 0)
#pragma unroll
            for (int q = 0; q < 8; ++q) out[b * 8 + q] = s[q];
    }
}

}  // namespace glove

using namespace glove;

extern "C" {

size_t glove_eval_workspace_bytes(int64_t count, int32_t batch_size) {
    if (count <= 0 || batch_size <= 0) return 0;
    const int64_t tpb = (batch_size + kEvalTile - 1) / kEvalTile;
    const int64_t nb = (count + batch_size - 1) / batch_size;
    return align_up((size_t)(nb * tpb) * 8 * sizeof(double));
}

int glove_eval_loss(const float *row_table, const float *col_table, const glove_scalars *scalars, int32_t planes,
                    int32_t d, const int32_t *row, const int32_t *col, const float *colA, const float *colB,
                    int64_t first, int64_t count, int32_t batch_size, int32_t head, double *out, void *workspace,
                    size_t workspace_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GLOVE_REQUIRE(row_table && col_table && scalars && row && col && colA && colB && out && workspace,
                  "glove_eval_loss: null pointer");
    GLOVE_REQUIRE(planes >= 1 && planes <= 3 && d > 0 && first >= 0 && count > 0 && batch_size > 0,
                  "glove_eval_loss: bad sizes");
    GLOVE_REQUIRE(head == GLOVE_HEAD_GLOVE || head == GLOVE_HEAD_LOGISTIC, "glove_eval_loss: unsupported head %d", head);
    const size_t need = glove_eval_workspace_bytes(count, batch_size);
    if (workspace_bytes < need)
        return set_error(GLOVE_EWORKSPACE, "glove_eval_loss: workspace %zu < required %zu", workspace_bytes, need);
    const int32_t S = table_stride(d);
    const int nv = (S / 4 + 31) / 32;
    if (nv > 4) return set_error(GLOVE_EUNSUPPORTED, "glove_eval_loss: embedding size %d > 510 not supported", d);
    const int64_t tpb = (batch_size + kEvalTile - 1) / kEvalTile;
    const int64_t nb = (count + batch_size - 1) / batch_size;
    const int64_t n_tiles = nb * tpb;
    double *tiles = (double *)workspace;
    int64_t blocks = (n_tiles + 7) / 8;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
#define LAUNCH_EVAL(NV)                                                                                             \
    eval_tiles_kernel<NV><<<(int)blocks, 256, 0, stream>>>(row_table, col_table, scalars, planes, d, S, row, col, colA, \
                                                           colB, first, count, batch_size, head, tpb, n_tiles, tiles)
    switch (nv) {
        case 1: LAUNCH_EVAL(1); break;
        case 2: LAUNCH_EVAL(2); break;
        case 3: LAUNCH_EVAL(3); break;
        default: LAUNCH_EVAL(4); break;
    }
#undef LAUNCH_EVAL
    GLOVE_CHECK_LAUNCH();
    int64_t rblocks = (nb + 7) / 8;
    if (rblocks > kNumSMs * 4) rblocks = kNumSMs * 4;
    eval_reduce_kernel<<<(int)rblocks, 256, 0, stream>>>(tiles, tpb, n_tiles, nb, out);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}

}  // extern "C"
