// tensor-core candidate pass for the cosine top-k (placeholder: routes to the exact fp32 scan until the tcgen05 kernel
// below is enabled)
#include "glove_topk.cuh"
using namespace glove;
extern "C" {
size_t glove_topk_workspace_bytes(int64_t V, int32_t d, int32_t n_queries, int32_t k) {
    if (V <= 0 || d <= 0 || n_queries <= 0 || k <= 0) return 0;
    return scan_fp32_workspace(n_queries, k);
}
int glove_topk_cosine(const float *table, int64_t V, int32_t d, int32_t planes, const void *norm_bf16,
                      const float *inv_norm, const int32_t *query_ids, int32_t n_queries, int32_t k, float *out_sim,
                      int32_t *out_idx, void *workspace, size_t workspace_bytes, void *stream) {
    return glove_topk_cosine_fp32(table, V, d, planes, inv_norm, query_ids, n_queries, k, out_sim, out_idx, workspace,
                                  workspace_bytes, stream);
}
}
