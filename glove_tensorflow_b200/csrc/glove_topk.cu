// PREDICT: cosine top-k.  (placeholder until the train path is validated on the GPU; replaced below in this round)
#include "glove_common.cuh"
using namespace glove;
extern "C" {
int32_t glove_topk_kpad(int32_t d) { return (d + 63) / 64 * 64; }
int64_t glove_topk_vpad(int64_t V) { return (V + 255) / 256 * 256; }
int glove_normalize_rows(const float *, int64_t, int32_t, int32_t, void *, float *, void *) {
    return set_error(GLOVE_EUNSUPPORTED, "glove_normalize_rows: not built yet");
}
size_t glove_topk_workspace_bytes(int64_t, int32_t, int32_t, int32_t) { return 0; }
int glove_topk_cosine(const float *, int64_t, int32_t, int32_t, const void *, const float *, const int32_t *, int32_t,
                      int32_t, float *, int32_t *, void *, size_t, void *) {
    return set_error(GLOVE_EUNSUPPORTED, "glove_topk_cosine: not built yet");
}
int glove_topk_cosine_fp32(const float *, int64_t, int32_t, int32_t, const float *, const int32_t *, int32_t, int32_t,
                           float *, int32_t *, void *, size_t, void *) {
    return set_error(GLOVE_EUNSUPPORTED, "glove_topk_cosine_fp32: not built yet");
}
}
