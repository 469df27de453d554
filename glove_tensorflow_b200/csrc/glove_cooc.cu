// PREPROCESS: token-id stream -> symmetric co-occurrence table with the GloVe columns, i.e.
// create_interaction_dataframe + create_glove_dataframe of the reference preprocessor [ref src/data/text8.py:84-139]
// without the position cross-join (2-3 min and 21 GB of pandas at text8 scale):
//
//   cooc_chunk    emit, for every position p and distance k = 1..context, the pair (id[p], id[p+k]) (right context only,
//                 equal ids dropped) [ref text8.py:90-95]; radix-sort the packed (row, col) keys with k as payload;
//                 reduce-by-key to {count, numer} where numer = sum L/k, L = lcm(1..context): the sum of 1/k is carried
//                 as an exact integer [ref text8.py:98-102]
//   cooc_merge    union of two sorted partial tables (chunks of a corpus larger than one sort)
//   cooc_finish   union with the transposed table and sum [ref text8.py:105-110], count >= count_minimum
//                 [ref text8.py:129], value = numer / L (one rounding), neg_weight = count_row * proportion_col
//                 [ref text8.py:113-117], glove_weight = clip((count/100)^0.75, 0, 1), glove_value = log(value)
//                 [ref text8.py:130-139]; records come out in a keyed-hash order (the reference orders them by Python's
//                 per-process salted hash() of the token pair, i.e. randomly [ref text8.py:119-124])
//
// Integer / sort work, HBM-bound: CUB radix sort + reduce-by-key around small hand-written kernels; no tensor cores.
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>

#include "glove_common.cuh"

namespace glove {

struct Agg { long long count, numer; };
struct AggSum {
    __host__ __device__ __forceinline__ Agg operator()(const Agg &a, const Agg &b) const {
        return Agg{a.count + b.count, a.numer + b.numer};
    }
};
// payload k-1 of a sorted pair -> its contribution {1, L/k}
struct KToAgg {
    long long w[8];
    __host__ __device__ __forceinline__ Agg operator()(const uint8_t &k1) const { return Agg{1, w[k1 & 7]}; }
};

static int vbits_of(int64_t V) {
    int b = 1;
    while ((1ll << b) < V) ++b;
    return b;
}
static long long lcm_upto(int n) {
    long long l = 1;
    for (int k = 2; k <= n; ++k) {
        long long a = l, b = k;
        while (b) { const long long t = a % b; a = b; b = t; }
        l = l / a * k;
    }
    return l;
}

struct CoocWs {
    uint64_t *keys[4];
    uint32_t *idx[2];
    uint8_t *kk[2];
    Agg *agg, *agg2;
    uint8_t *flags;
    long long *count_out;   // device scalar written by CUB
    void *cub_temp;
    size_t cub_bytes;
    size_t bytes;
};
static CoocWs cooc_ws_view(void *base, int64_t n) {
    CoocWs w;
    char *p = (char *)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char *r = p ? p + off : nullptr; off += align_up(bytes); return r; };
    for (int i = 0; i < 4; ++i) w.keys[i] = (uint64_t *)take(sizeof(uint64_t) * n);
    for (int i = 0; i < 2; ++i) w.idx[i] = (uint32_t *)take(sizeof(uint32_t) * n);
    for (int i = 0; i < 2; ++i) w.kk[i] = (uint8_t *)take(n);
    w.agg = (Agg *)take(sizeof(Agg) * n);
    w.agg2 = (Agg *)take(sizeof(Agg) * n);
    w.flags = (uint8_t *)take(n);
    w.count_out = (long long *)take(sizeof(long long));
    size_t a = 0, b = 0, c = 0, d = 0, e = 0;
    const int ni = (int)n;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (uint64_t *)nullptr, (uint64_t *)nullptr, (uint8_t *)nullptr, (uint8_t *)nullptr, ni);
    cub::DeviceRadixSort::SortPairs(nullptr, b, (uint64_t *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, ni);
    thrust::transform_iterator<KToAgg, const uint8_t *> it((const uint8_t *)nullptr, KToAgg());
    cub::DeviceReduce::ReduceByKey(nullptr, c, (uint64_t *)nullptr, (uint64_t *)nullptr, it, (Agg *)nullptr, (long long *)nullptr, AggSum(), ni);
    cub::DeviceReduce::ReduceByKey(nullptr, d, (uint64_t *)nullptr, (uint64_t *)nullptr, (Agg *)nullptr, (Agg *)nullptr, (long long *)nullptr, AggSum(), ni);
    cub::DeviceSelect::Flagged(nullptr, e, (uint32_t *)nullptr, (uint8_t *)nullptr, (uint32_t *)nullptr, (long long *)nullptr, ni);
    w.cub_bytes = a;
    if (b > w.cub_bytes) w.cub_bytes = b;
    if (c > w.cub_bytes) w.cub_bytes = c;
    if (d > w.cub_bytes) w.cub_bytes = d;
    if (e > w.cub_bytes) w.cub_bytes = e;
    w.cub_temp = take(w.cub_bytes);
    w.bytes = off;
    return w;
}

// pair (p, k): row = id[p], col = id[p + k]; ids outside [0, V) are an error of the caller (flag)
__global__ void __launch_bounds__(256) emit_pairs_kernel(const int32_t *__restrict__ ids, int64_t n_positions, int64_t n_tokens,
                                                         int32_t context, int vbits, int32_t V, uint64_t *keys, uint8_t *kk,
                                                         int32_t *bad) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_positions * context) return;
    const int64_t p = t / context;
    const int k = (int)(t - p * context) + 1;
    const uint64_t none = (1ull << (2 * vbits)) - 1;   // row == col == 2^vbits - 1: never a valid pair, sorts last
    uint64_t key = none;
    if (p + k < n_tokens) {
        const int32_t a = ids[p], b = ids[p + k];
        if (a < 0 || a >= V || b < 0 || b >= V) *bad = 1;
        else if (a != b) key = ((uint64_t)(uint32_t)a << vbits) | (uint32_t)b;
    }
    keys[t] = key;
    kk[t] = (uint8_t)(k - 1);
}

__global__ void __launch_bounds__(256) iota_kernel(uint32_t *idx, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = (uint32_t)i;
}
// concatenated tables A | B addressed by one index
__global__ void __launch_bounds__(256) gather_agg_kernel(const uint32_t *__restrict__ idx, int64_t n, const Agg *__restrict__ a,
                                                         int64_t na, const Agg *__restrict__ b, Agg *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = idx[i];
    out[i] = j < na ? a[j] : b[j - na];
}
__global__ void __launch_bounds__(256) transpose_keys_kernel(const uint64_t *__restrict__ keys, int64_t n, int vbits,
                                                             uint64_t *out, uint32_t *idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = keys[i], mask = (1ull << vbits) - 1;
    out[i] = k;
    out[n + i] = ((k & mask) << vbits) | (k >> vbits);
    idx[i] = (uint32_t)i;
    idx[n + i] = (uint32_t)i;   // both orientations carry the same {count, numer}
}
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
__global__ void __launch_bounds__(256) flag_min_count_kernel(const Agg *__restrict__ agg, int64_t n, long long count_min,
                                                             uint8_t *flags, uint32_t *idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = agg[i].count >= count_min && agg[i].count != 0 && agg[i].numer != 0;
    idx[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(256) order_keys_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ sel,
                                                         int64_t n, uint64_t salt, uint64_t *hashed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) hashed[i] = mix64(keys[sel[i]] + salt * 0x9e3779b97f4a7c15ull);
}
__global__ void __launch_bounds__(256) columns_kernel(const uint64_t *__restrict__ keys, const Agg *__restrict__ agg,
                                                      const uint32_t *__restrict__ order, int64_t n, int vbits, double inv_unit,
                                                      long long unit, const long long *__restrict__ vocab_count,
                                                      double total, int32_t *row, int32_t *col, long long *count,
                                                      double *value, double *neg_weight, double *glove_weight,
                                                      double *glove_value) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = order[i];
    const uint64_t k = keys[j];
    const int32_t a = (int32_t)(k >> vbits), b = (int32_t)(k & ((1ull << vbits) - 1));
    const Agg g = agg[j];
    const double v = (double)g.numer / (double)unit;          // exact integer / small integer: one rounding
    row[i] = a;
    col[i] = b;
    count[i] = g.count;
    value[i] = v;
    neg_weight[i] = (double)vocab_count[a] * ((double)vocab_count[b] / total);
    glove_weight[i] = fmin(fmax(pow((double)g.count / 100.0, 0.75), 0.0), 1.0);
    glove_value[i] = log(v);
}

static int read_count(const long long *dev, long long *host, cudaStream_t stream) {
    GLOVE_CHECK_CUDA(cudaMemcpyAsync(host, dev, sizeof(long long), cudaMemcpyDeviceToHost, stream));
    GLOVE_CHECK_CUDA(cudaStreamSynchronize(stream));
    return GLOVE_OK;
}

}  // namespace glove

using namespace glove;

extern "C" {

size_t glove_cooc_workspace_bytes(int64_t n_items) { return cooc_ws_view(nullptr, n_items < 1 ? 1 : n_items).bytes; }

int glove_cooc_chunk(const int32_t *token_ids, int64_t n_positions, int64_t n_tokens, int32_t V, int32_t context,
                     void *workspace, size_t workspace_bytes, uint64_t *out_keys, int64_t *out_agg, int64_t capacity,
                     int64_t *n_unique_host, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GLOVE_REQUIRE(token_ids && workspace && out_keys && out_agg && n_unique_host, "glove_cooc_chunk: null pointer");
    GLOVE_REQUIRE(n_positions > 0 && n_tokens >= n_positions && V > 1 && context >= 1 && context <= 8,
                  "glove_cooc_chunk: bad sizes (context must be 1..8)");
    const int64_t n = n_positions * context;
    GLOVE_REQUIRE(n < (1ll << 31), "glove_cooc_chunk: more than 2^31 pairs in one chunk");
    const int vbits = vbits_of((int64_t)V + 1);
    GLOVE_REQUIRE(2 * vbits <= 62, "glove_cooc_chunk: vocabulary too large");
    const CoocWs w = cooc_ws_view(workspace, n);
    if (workspace_bytes < w.bytes)
        return set_error(GLOVE_EWORKSPACE, "glove_cooc_chunk: workspace %zu < required %zu", workspace_bytes, w.bytes);
    int32_t *bad = (int32_t *)w.flags;
    GLOVE_CHECK_CUDA(cudaMemsetAsync(bad, 0, sizeof(int32_t), stream));
    emit_pairs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(token_ids, n_positions, n_tokens, context, vbits, V,
                                                                        w.keys[0], w.kk[0], bad);
    GLOVE_CHECK_LAUNCH();
    size_t tb = w.cub_bytes;
    GLOVE_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, tb, w.keys[0], w.keys[1], w.kk[0], w.kk[1], (int)n, 0,
                                                     2 * vbits, stream));
    KToAgg conv;
    const long long unit = lcm_upto(context);
    for (int k = 0; k < 8; ++k) conv.w[k] = k < context ? unit / (k + 1) : 0;
    thrust::transform_iterator<KToAgg, const uint8_t *> vals(w.kk[1], conv);
    tb = w.cub_bytes;
    GLOVE_CHECK_CUDA(cub::DeviceReduce::ReduceByKey(w.cub_temp, tb, w.keys[1], w.keys[0], vals, w.agg, w.count_out, AggSum(),
                                                    (int)n, stream));
    long long runs = 0;
    int32_t bad_h = 0;
    GLOVE_CHECK_CUDA(cudaMemcpyAsync(&bad_h, bad, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    if (int rc = read_count(w.count_out, &runs, stream)) return rc;
    if (bad_h) return set_error(GLOVE_EINVAL, "glove_cooc_chunk: token id outside [0, %d)", V);
    // the sentinel run (dropped pairs), when present, is the last one
    uint64_t last_key = 0;
    if (runs > 0) {
        GLOVE_CHECK_CUDA(cudaMemcpyAsync(&last_key, w.keys[0] + runs - 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
        GLOVE_CHECK_CUDA(cudaStreamSynchronize(stream));
        if (last_key == (1ull << (2 * vbits)) - 1) --runs;
    }
    if (runs > capacity)
        return set_error(GLOVE_EINVAL, "glove_cooc_chunk: %lld distinct pairs > capacity %lld", runs, (long long)capacity);
    GLOVE_CHECK_CUDA(cudaMemcpyAsync(out_keys, w.keys[0], sizeof(uint64_t) * runs, cudaMemcpyDeviceToDevice, stream));
    GLOVE_CHECK_CUDA(cudaMemcpyAsync(out_agg, w.agg, sizeof(Agg) * runs, cudaMemcpyDeviceToDevice, stream));
    GLOVE_CHECK_CUDA(cudaStreamSynchronize(stream));
    *n_unique_host = runs;
    return GLOVE_OK;
}

int glove_cooc_merge(const uint64_t *keys_a, const int64_t *agg_a, int64_t n_a, const uint64_t *keys_b, const int64_t *agg_b,
                     int64_t n_b, int32_t V, void *workspace, size_t workspace_bytes, uint64_t *out_keys, int64_t *out_agg,
                     int64_t capacity, int64_t *n_unique_host, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GLOVE_REQUIRE(workspace && out_keys && out_agg && n_unique_host && n_a >= 0 && n_b >= 0 && n_a + n_b > 0,
                  "glove_cooc_merge: bad arguments");
    const int64_t n = n_a + n_b;
    GLOVE_REQUIRE(n < (1ll << 31), "glove_cooc_merge: more than 2^31 entries");
    const int vbits = vbits_of((int64_t)V + 1);
    const CoocWs w = cooc_ws_view(workspace, n);
    if (workspace_bytes < w.bytes)
        return set_error(GLOVE_EWORKSPACE, "glove_cooc_merge: workspace %zu < required %zu", workspace_bytes, w.bytes);
    if (n_a) GLOVE_CHECK_CUDA(cudaMemcpyAsync(w.keys[0], keys_a, sizeof(uint64_t) * n_a, cudaMemcpyDeviceToDevice, stream));
    if (n_b) GLOVE_CHECK_CUDA(cudaMemcpyAsync(w.keys[0] + n_a, keys_b, sizeof(uint64_t) * n_b, cudaMemcpyDeviceToDevice, stream));
    iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(w.idx[0], n);
    GLOVE_CHECK_LAUNCH();
    size_t tb = w.cub_bytes;
    GLOVE_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, tb, w.keys[0], w.keys[1], w.idx[0], w.idx[1], (int)n, 0,
                                                     2 * vbits, stream));
    gather_agg_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(w.idx[1], n, (const Agg *)agg_a, n_a, (const Agg *)agg_b,
                                                                        w.agg);
    GLOVE_CHECK_LAUNCH();
    // reduce straight into the caller's buffers when they can hold the worst case, else refuse
    GLOVE_REQUIRE(capacity >= n, "glove_cooc_merge: capacity %lld < n_a + n_b = %lld", (long long)capacity, (long long)n);
    tb = w.cub_bytes;
    GLOVE_CHECK_CUDA(cub::DeviceReduce::ReduceByKey(w.cub_temp, tb, w.keys[1], out_keys, w.agg, (Agg *)out_agg, w.count_out,
                                                    AggSum(), (int)n, stream));
    long long runs = 0;
    if (int rc = read_count(w.count_out, &runs, stream)) return rc;
    *n_unique_host = runs;
    return GLOVE_OK;
}

int glove_cooc_finish(const uint64_t *keys, const int64_t *agg, int64_t n, int32_t V, int32_t context, int64_t count_min,
                      const int64_t *vocab_count, int64_t total_tokens, uint64_t order_key, void *workspace,
                      size_t workspace_bytes, int32_t *row, int32_t *col, int64_t *count, double *value, double *neg_weight,
                      double *glove_weight, double *glove_value, int64_t capacity, int64_t *n_out_host, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GLOVE_REQUIRE(keys && agg && vocab_count && workspace && n_out_host && n > 0, "glove_cooc_finish: bad arguments");
    GLOVE_REQUIRE(context >= 1 && context <= 8 && total_tokens > 0 && V > 1, "glove_cooc_finish: bad sizes");
    const int64_t n2 = 2 * n;
    GLOVE_REQUIRE(n2 < (1ll << 31), "glove_cooc_finish: more than 2^30 distinct pairs");
    const int vbits = vbits_of((int64_t)V + 1);
    const CoocWs w = cooc_ws_view(workspace, n2);
    if (workspace_bytes < w.bytes)
        return set_error(GLOVE_EWORKSPACE, "glove_cooc_finish: workspace %zu < required %zu", workspace_bytes, w.bytes);
    const unsigned g2 = (unsigned)((n2 + 255) / 256);
    // union with the transposed table, summed per (row, col)
    transpose_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(keys, n, vbits, w.keys[0], w.idx[0]);
    GLOVE_CHECK_LAUNCH();
    size_t tb = w.cub_bytes;
    GLOVE_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, tb, w.keys[0], w.keys[1], w.idx[0], w.idx[1], (int)n2, 0,
                                                     2 * vbits, stream));
    gather_agg_kernel<<<g2, 256, 0, stream>>>(w.idx[1], n2, (const Agg *)agg, n2, (const Agg *)agg, w.agg);
    GLOVE_CHECK_LAUNCH();
    tb = w.cub_bytes;
    GLOVE_CHECK_CUDA(cub::DeviceReduce::ReduceByKey(w.cub_temp, tb, w.keys[1], w.keys[0], w.agg, w.agg2, w.count_out, AggSum(),
                                                    (int)n2, stream));
    long long m = 0;
    if (int rc = read_count(w.count_out, &m, stream)) return rc;
    // count >= count_minimum
    flag_min_count_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(w.agg2, m, count_min, w.flags, w.idx[0]);
    GLOVE_CHECK_LAUNCH();
    tb = w.cub_bytes;
    GLOVE_CHECK_CUDA(cub::DeviceSelect::Flagged(w.cub_temp, tb, w.idx[0], w.flags, w.idx[1], w.count_out, (int)m, stream));
    long long n_out = 0;
    if (int rc = read_count(w.count_out, &n_out, stream)) return rc;
    *n_out_host = n_out;
    if (n_out > capacity)
        return set_error(GLOVE_EINVAL, "glove_cooc_finish: %lld records > capacity %lld", n_out, (long long)capacity);
    if (n_out == 0) return GLOVE_OK;
    GLOVE_REQUIRE(row && col && count && value && neg_weight && glove_weight && glove_value, "glove_cooc_finish: null output");
    // keyed-hash record order
    const unsigned go = (unsigned)((n_out + 255) / 256);
    order_keys_kernel<<<go, 256, 0, stream>>>(w.keys[0], w.idx[1], n_out, order_key, w.keys[2]);
    GLOVE_CHECK_LAUNCH();
    tb = w.cub_bytes;
    GLOVE_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, tb, w.keys[2], w.keys[3], w.idx[1], w.idx[0], (int)n_out, 0, 64,
                                                     stream));
    const long long unit = lcm_upto(context);
    columns_kernel<<<go, 256, 0, stream>>>(w.keys[0], w.agg2, w.idx[0], n_out, vbits, 1.0 / (double)unit, unit,
                                           (const long long *)vocab_count, (double)total_tokens, row, col,
                                           (long long *)count, value, neg_weight, glove_weight, glove_value);
    GLOVE_CHECK_LAUNCH();
    GLOVE_CHECK_CUDA(cudaStreamSynchronize(stream));
    return GLOVE_OK;
}

}  // extern "C"
