// Input pipeline: keyed on-GPU shuffle + batch plan construction (sort-by-token, segment and work-item lists).
// Replaces tf.data make_csv_dataset shuffle/batch [ref src/models/data_utils.py:4-26] and the per-step
// Unique + UnsortedSegmentSum de-duplication that Keras OptimizerV2 does inside the step: here the
// de-duplication structure of K steps is built once, off the critical path, and reused by both sides of each step.
#include <cub/cub.cuh>

#include "glove_common.cuh"

namespace glove {

struct PrepSide {
    uint32_t *keys_in, *keys_out, *vals_out;
    int32_t *f_seg, *f_item, *f_long, *f_part;
    int32_t *e_seg, *e_item, *e_long, *e_part;
    int32_t *slot_of_p;
};
struct PrepWs {
    uint32_t *vals_in;
    float *A, *Bv;
    PrepSide side[2];
    void *cub_temp;
    size_t cub_bytes;
    size_t bytes;
};

static size_t cub_temp_bytes(int64_t N) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                    (uint32_t *)nullptr, (int)N, 0, 32);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (int32_t *)nullptr, (int32_t *)nullptr, (int)N);
    return align_up(a > b ? a : b);
}

static PrepWs prep_view(void *base, int32_t K, int32_t B) {
    PrepWs w;
    char *p = (char *)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char *r = p ? p + off : nullptr; off += align_up(bytes); return r; };
    const int64_t N = (int64_t)K * B;
    w.vals_in = (uint32_t *)take(4 * N);
    w.A = (float *)take(4 * N);
    w.Bv = (float *)take(4 * N);
    for (int s = 0; s < 2; ++s) {
        PrepSide &ps = w.side[s];
        ps.keys_in = (uint32_t *)take(4 * N);
        ps.keys_out = (uint32_t *)take(4 * N);
        ps.vals_out = (uint32_t *)take(4 * N);
        ps.f_seg = (int32_t *)take(4 * N);
        ps.f_item = (int32_t *)take(4 * N);
        ps.f_long = (int32_t *)take(4 * N);
        ps.f_part = (int32_t *)take(4 * N);
        ps.e_seg = (int32_t *)take(4 * N);
        ps.e_item = (int32_t *)take(4 * N);
        ps.e_long = (int32_t *)take(4 * N);
        ps.e_part = (int32_t *)take(4 * N);
        ps.slot_of_p = (int32_t *)take(4 * N);
    }
    w.cub_bytes = cub_temp_bytes(N);
    w.cub_temp = take(w.cub_bytes);
    w.bytes = off;
    return w;
}

__global__ void shuffle_indices_kernel(uint32_t key, int64_t nnz, int64_t first, int64_t count, int h, int64_t *out) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = first + k;
        const uint32_t epoch = (uint32_t)(n / nnz);
        out[k] = (int64_t)feistel_permute((uint64_t)(n % nnz), (uint64_t)nnz, key + epoch, h);
    }
}

__global__ void gather_kernel(const int32_t *__restrict__ row, const int32_t *__restrict__ col,
                              const float *__restrict__ colA, const float *__restrict__ colB, int64_t nnz,
                              const int64_t *__restrict__ sample_idx, int64_t first_sample, uint32_t key, int h,
                              int32_t N, int32_t B, int vbits, int32_t n_shards, int32_t v_loc, PrepWs w) {
    for (int32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
        int64_t src;
        if (sample_idx) {
            src = sample_idx[p];
        } else {
            const int64_t n = first_sample + p;
            src = (int64_t)feistel_permute((uint64_t)(n % nnz), (uint64_t)nnz, key + (uint32_t)(n / nnz), h);
        }
        const uint32_t kb = (uint32_t)(p / B) << vbits;
        uint32_t i = (uint32_t)row[src], j = (uint32_t)col[src];
        if (n_shards > 1) {  // owner-major id: rows of one owner are contiguous in every sorted list
            i = (i % n_shards) * v_loc + i / n_shards;
            j = (j % n_shards) * v_loc + j / n_shards;
        }
        w.side[0].keys_in[p] = kb | i;
        w.side[1].keys_in[p] = kb | j;
        w.vals_in[p] = (uint32_t)p;
        w.A[p] = colA[src];
        w.Bv[p] = colB[src];
    }
}

__global__ void heads_kernel(const uint32_t *__restrict__ keys, int32_t N, int32_t *f_seg) {
    for (int32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < N; q += gridDim.x * blockDim.x)
        f_seg[q] = (q == 0 || keys[q] != keys[q - 1]) ? 1 : 0;
}

__global__ void segs_kernel(const uint32_t *__restrict__ keys, int32_t N, uint32_t idmask, PrepSide ps, PlanSide out,
                            PlanHeader *hdr, int side) {
    for (int32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < N; q += gridDim.x * blockDim.x) {
        if (ps.f_seg[q]) {
            out.seg_start[ps.e_seg[q]] = q;
            out.seg_id[ps.e_seg[q]] = (int32_t)(keys[q] & idmask);
        }
        if (q == N - 1) {
            const int32_t nseg = ps.e_seg[q] + ps.f_seg[q];
            out.seg_start[nseg] = N;
            hdr->n_seg[side] = nseg;
        }
    }
}

__global__ void itemflags_kernel(int32_t N, PrepSide ps, PlanSide out) {
    for (int32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < N; q += gridDim.x * blockDim.x) {
        const int32_t g = ps.e_seg[q] + ps.f_seg[q] - 1;
        const int32_t s0 = out.seg_start[g], len = out.seg_start[g + 1] - s0;
        const int32_t fi = ((q - s0) % kItemMax == 0) ? 1 : 0;
        const int32_t lg = len > kItemMax ? 1 : 0;
        ps.f_item[q] = fi;
        ps.f_long[q] = ps.f_seg[q] & lg;
        ps.f_part[q] = fi & lg;
    }
}

__global__ void items_kernel(int32_t N, int32_t B, int32_t K, PrepSide ps, PlanSide out, PlanHeader *hdr, int side) {
    for (int32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < N; q += gridDim.x * blockDim.x) {
        const int32_t g = ps.e_seg[q] + ps.f_seg[q] - 1;
        if (ps.f_item[q]) {
            const int32_t it = ps.e_item[q];
            out.item_seg[it] = g;
            out.item_start[it] = q;
            out.item_part[it] = ps.f_part[q] ? ps.e_part[q] : -1;
        }
        if (ps.f_seg[q]) out.seg_long[g] = ps.f_long[q] ? ps.e_long[q] : -1;
        if (ps.f_long[q]) {
            const int32_t l = ps.e_long[q];
            out.long_seg[l] = g;
            out.long_item[l] = ps.e_item[q];
        }
        if (q % B == 0) {
            const int32_t k = q / B;
            out.b_seg[k] = ps.e_seg[q];
            out.b_item[k] = ps.e_item[q];
            out.b_long[k] = ps.e_long[q];
            out.b_part[k] = ps.e_part[q];
        }
        if (q == N - 1) {
            out.b_seg[K] = ps.e_seg[q] + ps.f_seg[q];
            out.b_item[K] = ps.e_item[q] + ps.f_item[q];
            out.b_long[K] = ps.e_long[q] + ps.f_long[q];
            out.b_part[K] = ps.e_part[q] + ps.f_part[q];
            hdr->n_item[side] = out.b_item[K];
            hdr->n_long[side] = out.b_long[K];
            hdr->n_part[side] = out.b_part[K];
        }
    }
}

// per batch: first slot owned by each shard (binary search of owner * v_loc in the sorted ids) and the padded block size
__global__ void owners_kernel(int32_t K, int32_t n_shards, int32_t v_loc, PlanSide out) {
    for (int32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < K * (kMaxShards + 1); t += gridDim.x * blockDim.x) {
        const int32_t k = t / (kMaxShards + 1), r = t % (kMaxShards + 1);
        int32_t lo = out.b_seg[k], hi = out.b_seg[k + 1];
        const int32_t end = hi;
        if (r >= n_shards) { out.b_own[t] = end - out.b_seg[k]; continue; }
        const int32_t key = r * v_loc;
        while (lo < hi) {
            const int32_t mid = (lo + hi) >> 1;
            if (out.seg_id[mid] < key) lo = mid + 1; else hi = mid;
        }
        out.b_own[t] = lo - out.b_seg[k];
    }
}
// first work item of each shard's block: items are sorted by segment, so it is a lower bound over item_seg
__global__ void owner_items_kernel(int32_t K, PlanSide out) {
    for (int32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < K * (kMaxShards + 1); t += gridDim.x * blockDim.x) {
        const int32_t k = t / (kMaxShards + 1);
        const int32_t g = out.b_seg[k] + out.b_own[t];     // first segment of the block (or one past the last)
        int32_t lo = out.b_item[k], hi = out.b_item[k + 1];
        while (lo < hi) {
            const int32_t mid = (lo + hi) >> 1;
            if (out.item_seg[mid] < g) lo = mid + 1; else hi = mid;
        }
        out.b_own_item[t] = lo - out.b_item[k];
    }
}
__global__ void upad_kernel(int32_t K, int32_t n_shards, PlanSide out) {
    for (int32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
        int32_t m = 0;
        for (int r = 0; r < n_shards; ++r)
            m = max(m, out.b_own[k * (kMaxShards + 1) + r + 1] - out.b_own[k * (kMaxShards + 1) + r]);
        out.b_upad[k] = m;
    }
}

__global__ void slots_kernel(int32_t N, int32_t B, int32_t n_shards, int32_t v_loc, PrepSide ps, PlanSide out) {
    for (int32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < N; q += gridDim.x * blockDim.x) {
        const int32_t g = ps.e_seg[q] + ps.f_seg[q] - 1;
        ps.slot_of_p[ps.vals_out[q]] = plan_pos(out, q / B, g, n_shards, v_loc);
    }
}

__global__ void fill_kernel(int32_t N, int32_t B, PrepSide ps, const int32_t *__restrict__ other_slot_of_p,
                            const float *__restrict__ A, const float *__restrict__ Bv, PlanSide out) {
    for (int32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < N; q += gridDim.x * blockDim.x) {
        const uint32_t p = ps.vals_out[q];
        out.rec[q] = make_int4(other_slot_of_p[p], __float_as_int(A[p]), __float_as_int(Bv[p]), (int32_t)(p % (uint32_t)B));
    }
}

// seg_prev[g] = 1 when the id of segment g (batch k) also occurs in batch k-1 of the same plan: binary search in the
// sorted id list of that batch.  The first batch of a plan gets 1 ("unknown": it is never pre-replayed).
__global__ void segprev_kernel(int32_t N, int32_t B, PrepSide ps, PlanSide out) {
    for (int32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < N; q += gridDim.x * blockDim.x) {
        if (!ps.f_seg[q]) continue;
        const int32_t g = ps.e_seg[q], k = q / B;
        int32_t flag = 1;
        if (k > 0) {
            const int32_t id = out.seg_id[g];
            int32_t lo = out.b_seg[k - 1], hi = out.b_seg[k];
            while (lo < hi) {
                const int32_t mid = (lo + hi) >> 1;
                if (out.seg_id[mid] < id) lo = mid + 1; else hi = mid;
            }
            flag = (lo < out.b_seg[k] && out.seg_id[lo] == id) ? 1 : 0;
        }
        out.seg_prev[g] = flag;
    }
}

__global__ void itemrec_kernel(int32_t B, int32_t n_shards, int32_t v_loc, const int32_t *__restrict__ n_items, PlanSide out) {
    const int32_t NI = *n_items;
    for (int32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < NI; it += gridDim.x * blockDim.x) {
        const int32_t g = out.item_seg[it], start = out.item_start[it];
        const int32_t k = start / B;
        const int32_t seg_end = out.seg_start[g + 1], seg_len = seg_end - out.seg_start[g];
        const int32_t n = min(start + kItemMax, seg_end) - start;
        const int32_t part = seg_len > kItemMax ? out.item_part[it] - out.b_part[k] + 1 : 0;
        out.item_rec[it] = make_int4(part ? out.seg_long[g] - out.b_long[k] : out.seg_id[g],
                                     plan_pos(out, k, g, n_shards, v_loc), start, n | (part << 8));
    }
}

__global__ void longrec_kernel(int32_t B, int32_t n_shards, int32_t v_loc, const int32_t *__restrict__ n_long, PlanSide out) {
    const int32_t NL = *n_long;
    for (int32_t l = blockIdx.x * blockDim.x + threadIdx.x; l < NL; l += gridDim.x * blockDim.x) {
        const int32_t g = out.long_seg[l], it0 = out.long_item[l];
        const int32_t k = out.seg_start[g] / B;
        const int32_t len = out.seg_start[g + 1] - out.seg_start[g];
        out.long_rec[l] = make_int4(out.seg_id[g], plan_pos(out, k, g, n_shards, v_loc), out.item_part[it0] - out.b_part[k],
                                    (len + kItemMax - 1) / kItemMax);
    }
}

// ---- request lists of the row-sharded exchange -------------------------------------------------------------------------
// key of sorted position q of side s: (batch, shard that owns q's segment, position of the opposite row)
__global__ void need_keys_kernel(int32_t N, int32_t B, int32_t v_loc, int pbits, PrepSide ps, PlanSide out) {
    for (int32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < N; q += gridDim.x * blockDim.x) {
        const int32_t g = ps.e_seg[q] + ps.f_seg[q] - 1;
        const uint32_t r = (uint32_t)(out.seg_id[g] / v_loc);
        ps.keys_in[q] = ((uint32_t)(q / B) << (pbits + 3)) | (r << pbits) | (uint32_t)out.rec[q].x;
    }
}
__global__ void need_compact_kernel(int32_t N, uint32_t posmask, const uint32_t *__restrict__ keys,
                                    const int32_t *__restrict__ flag, const int32_t *__restrict__ excl, uint32_t *uniq,
                                    int32_t *need_pos, int32_t *n_uniq) {
    for (int32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < N; q += gridDim.x * blockDim.x) {
        if (flag[q]) { uniq[excl[q]] = keys[q]; need_pos[excl[q]] = (int32_t)(keys[q] & posmask); }
        if (q == N - 1) *n_uniq = excl[q] + flag[q];
    }
}
// seg_push of the OPPOSITE side from the request lists of this side: block (k, r) walks the rows shard r needs in batch k
// and sets bit r on the segment that owns each of them (rows r owns itself are not requests)
__global__ void push_mask_kernel(int32_t n_shards, PlanSide need, PlanSide target) {
    const int32_t k = blockIdx.x / n_shards, r = blockIdx.x % n_shards;
    constexpr int W = kMaxShards + 1;
    const int32_t *off = need.need_off + ((int64_t)k * kMaxShards + r) * W;
    const int32_t upad = max(target.b_upad[k], 1);
    for (int32_t e = off[0] + threadIdx.x; e < off[n_shards]; e += blockDim.x) {
        const int32_t pos = need.need_pos[e], owner = pos / upad;
        if (owner == r || owner >= n_shards) continue;
        const int32_t g = target.b_seg[k] + target.b_own[k * W + owner] + (pos - owner * upad);
        atomicOr(target.seg_push + g, 1 << r);
    }
}
__global__ void need_offsets_kernel(int32_t K, int32_t n_shards, int pbits, const uint32_t *__restrict__ uniq,
                                    const int32_t *__restrict__ n_uniq, const int32_t *__restrict__ upad_other, int32_t *need_off) {
    const int32_t total = K * kMaxShards * (kMaxShards + 1);
    const int32_t nu = *n_uniq;
    for (int32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int32_t q = t % (kMaxShards + 1), r = (t / (kMaxShards + 1)) % kMaxShards, k = t / (kMaxShards * (kMaxShards + 1));
        // first unique key >= (k, r, q * upad): for r or q beyond n_shards this is the end of the previous block
        uint32_t key;
        if (r >= n_shards) key = (uint32_t)(k + 1) << (pbits + 3);
        else if (q >= n_shards) key = ((uint32_t)k << (pbits + 3)) + ((uint32_t)(r + 1) << pbits);   // + : r + 1 may carry into k
        else key = ((uint32_t)k << (pbits + 3)) | ((uint32_t)r << pbits) | (uint32_t)(q * upad_other[k]);
        int32_t lo = 0, hi = nu;
        while (lo < hi) {
            const int32_t mid = (lo + hi) >> 1;
            if (uniq[mid] < key) lo = mid + 1; else hi = mid;
        }
        need_off[t] = lo;
    }
}

// ---- shared plan construction (row-sharded tables): pull one shard's slice of a plan built by another rank -------------
// A plan of K steps of the GLOBAL batch costs O(K * B_global) to build and every rank needs only the ~1/n_shards of it that
// describes the segments it owns.  Ranks therefore take turns building whole plans (rank c % n_shards builds the plan of
// chunk c into peer-mapped memory) and everybody copies its own slice into a local plan buffer of the same layout, so all
// indices stay valid and the step kernels do not know the difference: per rank the construction cost no longer grows with
// the number of GPUs.  Slice of shard r = the per-batch offset tables and request offsets (whole: they are tiny), and of
// every batch the segment records, triple records, work-item records, long-segment records and request-list entries of
// r's block (what the stage / update / exchange kernels of shard r dereference).
struct PullArgs {
    const PlanHeader *src_hdr;
    PlanHeader *dst_hdr;
    PlanSide src[2], dst[2];
    int32_t K, n_shards, shard;
};
__device__ __forceinline__ void copy_i32(int32_t *dst, const int32_t *src, int64_t n) {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}
__global__ void __launch_bounds__(256) plan_pull_head_kernel(const PullArgs a) {
    constexpr int W = kMaxShards + 1;
    if (blockIdx.x == 0) copy_i32((int32_t *)a.dst_hdr, (const int32_t *)a.src_hdr, sizeof(PlanHeader) / 4);
    const int s = blockIdx.x & 1;
    const PlanSide &S = a.src[s], &D = a.dst[s];
    switch (blockIdx.x >> 1) {
        case 0: copy_i32(D.b_seg, S.b_seg, a.K + 1); copy_i32(D.b_item, S.b_item, a.K + 1); break;
        case 1: copy_i32(D.b_long, S.b_long, a.K + 1); copy_i32(D.b_part, S.b_part, a.K + 1); copy_i32(D.b_upad, S.b_upad, a.K); break;
        case 2: copy_i32(D.b_own, S.b_own, (int64_t)a.K * W); copy_i32(D.b_own_item, S.b_own_item, (int64_t)a.K * W); break;
        default: copy_i32(D.need_off, S.need_off, (int64_t)a.K * kMaxShards * W); break;
    }
}
// element range [lo, hi) of an array, this CTA's share: 4 independent loads per thread in flight (remote reads over
// NVLink cost ~2-3 us each; the copy is bound by how many of them are outstanding)
template <typename T>
__device__ __forceinline__ void copy_part(T *dst, const T *src, int lo, int hi, int part, int nparts) {
    const int stride = nparts * blockDim.x;
    int i = lo + part * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < hi; i += 4 * stride) {
        const T a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < hi; i += stride) dst[i] = src[i];
}
// int32 arrays: source and destination have the same layout (same offset from a 256-byte aligned base), so the 16-byte
// aligned middle of the range goes as int4
__device__ __forceinline__ void copy_part32(int32_t *dst, const int32_t *src, int lo, int hi, int part, int nparts) {
    const int lo4 = (lo + 3) & ~3, hi4 = hi & ~3;
    if (lo4 >= hi4) { copy_part(dst, src, lo, hi, part, nparts); return; }
    copy_part(reinterpret_cast<int4 *>(dst), reinterpret_cast<const int4 *>(src), lo4 >> 2, hi4 >> 2, part, nparts);
    if (part == 0) {
        for (int i = lo + threadIdx.x; i < lo4; i += blockDim.x) dst[i] = src[i];
        for (int i = hi4 + threadIdx.x; i < hi; i += blockDim.x) dst[i] = src[i];
    }
}
// grid = K * 2 * nparts CTAs: CTA (k, side, part) copies every nparts-th 256-element block of each range of batch k
__global__ void __launch_bounds__(256) plan_pull_body_kernel(const PullArgs a, int nparts) {
    constexpr int W = kMaxShards + 1;
    const int part = blockIdx.x % nparts, t = blockIdx.x / nparts, k = t >> 1, s = t & 1;
    const PlanSide &S = a.src[s], &D = a.dst[s];
    const int r = a.shard;
    // offsets come from the local copy (written by the head kernel, one launch earlier on the same stream)
    const int g0 = D.b_seg[k] + D.b_own[k * W + r], g1 = D.b_seg[k] + D.b_own[k * W + r + 1];
    if (g1 > g0) {
        copy_part32(D.seg_id, S.seg_id, g0, g1, part, nparts);
        copy_part32(D.seg_prev, S.seg_prev, g0, g1, part, nparts);
        copy_part32(D.seg_push, S.seg_push, g0, g1, part, nparts);
        copy_part32(D.seg_long, S.seg_long, g0, g1, part, nparts);
        copy_part32(D.seg_start, S.seg_start, g0, g1 + 1, part, nparts);
        const int q0 = S.seg_start[g0], q1 = S.seg_start[g1];
        copy_part(D.rec, S.rec, q0, q1, part, nparts);
    }
    copy_part(D.item_rec, S.item_rec, D.b_item[k] + D.b_own_item[k * W + r], D.b_item[k] + D.b_own_item[k * W + r + 1], part, nparts);
    copy_part(D.long_rec, S.long_rec, D.b_long[k], D.b_long[k + 1], part, nparts);
    // request lists: the rows shard r needs (from every owner), and the rows every other shard needs from owner r
    const int32_t *off = D.need_off + (int64_t)k * kMaxShards * W;
    for (int q = 0; q < a.n_shards; ++q) {
        if (q == r) copy_part32(D.need_pos, S.need_pos, off[r * W], off[r * W + a.n_shards], part, nparts);
        else copy_part32(D.need_pos, S.need_pos, off[q * W + r], off[q * W + r + 1], part, nparts);
    }
}

__global__ void header_kernel(PlanHeader *hdr, int32_t K, int32_t B, int32_t first_step, int32_t n_shards, int32_t v_loc) {
    hdr->magic = kPlanMagic;
    hdr->K = K;
    hdr->B = B;
    hdr->first_step = first_step;
    hdr->n_shards = n_shards;
    hdr->v_loc = v_loc;
}

}  // namespace glove

using namespace glove;

extern "C" {

int glove_shuffle_indices(uint32_t key, int64_t nnz, int64_t first, int64_t count, int64_t *out, void *stream) {
    GLOVE_REQUIRE(out && nnz > 0 && first >= 0 && count >= 0, "glove_shuffle_indices: bad arguments");
    if (count == 0) return GLOVE_OK;
    int64_t blocks = (count + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    shuffle_indices_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(key, nnz, first, count,
                                                                         feistel_half_bits((uint64_t)nnz), out);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}

size_t glove_plan_bytes(int32_t K, int32_t B) {
    if (K <= 0 || B <= 0) return 0;
    return plan_view(nullptr, K, B).bytes;
}
size_t glove_prepare_workspace_bytes(int32_t K, int32_t B) {
    if (K <= 0 || B <= 0) return 0;
    return prep_view(nullptr, K, B).bytes;
}

int glove_prepare_batches(void *plan, void *workspace, size_t workspace_bytes, const int32_t *row, const int32_t *col,
                          const float *colA, const float *colB, int64_t nnz, const int64_t *sample_idx,
                          int64_t first_sample, uint32_t shuffle_key, int32_t first_step, int32_t K, int32_t B,
                          int32_t V, void *stream_) {
    return glove_prepare_batches_sharded(plan, workspace, workspace_bytes, row, col, colA, colB, nnz, sample_idx,
                                         first_sample, shuffle_key, first_step, K, B, V, 1, stream_);
}

int glove_prepare_batches_sharded(void *plan, void *workspace, size_t workspace_bytes, const int32_t *row,
                                  const int32_t *col, const float *colA, const float *colB, int64_t nnz,
                                  const int64_t *sample_idx, int64_t first_sample, uint32_t shuffle_key,
                                  int32_t first_step, int32_t K, int32_t B, int32_t V, int32_t n_shards, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GLOVE_REQUIRE(n_shards >= 1 && n_shards <= kMaxShards, "glove_prepare_batches: n_shards %d not in [1, %d]", n_shards, kMaxShards);
    const int32_t v_loc = (V + n_shards - 1) / n_shards;
    const int64_t V_ids = (int64_t)v_loc * n_shards;  // range of the (remapped) ids
    GLOVE_REQUIRE(plan && workspace && row && col && colA && colB, "glove_prepare_batches: null pointer");
    GLOVE_REQUIRE(K > 0 && B > 0 && V > 0 && nnz > 0 && first_sample >= 0, "glove_prepare_batches: bad sizes");
    const int64_t N64 = (int64_t)K * B;
    GLOVE_REQUIRE(N64 < (1ll << 30), "glove_prepare_batches: K*B = %lld too large (max 2^30)", (long long)N64);
    int vbits = 1;
    while ((1ll << vbits) < V_ids) ++vbits;
    int kbits = 0;
    while ((1ll << kbits) < K) ++kbits;
    GLOVE_REQUIRE(vbits + kbits <= 32, "glove_prepare_batches: K=%d batches x V=%d ids do not fit 32-bit sort keys", K, V);
    PrepWs w = prep_view(workspace, K, B);
    if (workspace_bytes < w.bytes)
        return set_error(GLOVE_EWORKSPACE, "glove_prepare_batches: workspace %zu < required %zu", workspace_bytes, w.bytes);
    PlanView pv = plan_view(plan, K, B);
    const int32_t N = (int32_t)N64;
    const int threads = 256;
    int blocks = (N + threads - 1) / threads;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;

    header_kernel<<<1, 1, 0, stream>>>(pv.hdr, K, B, first_step, n_shards, v_loc);
    gather_kernel<<<blocks, threads, 0, stream>>>(row, col, colA, colB, nnz, sample_idx, first_sample, shuffle_key,
                                                  feistel_half_bits((uint64_t)nnz), N, B, vbits, n_shards, v_loc, w);
    GLOVE_CHECK_LAUNCH();
    const uint32_t idmask = (uint32_t)((1ull << vbits) - 1);
    for (int s = 0; s < 2; ++s) {
        PrepSide &ps = w.side[s];
        size_t tb = w.cub_bytes;
        GLOVE_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, tb, ps.keys_in, ps.keys_out, w.vals_in,
                                                         ps.vals_out, N, 0, vbits + kbits, stream));
        heads_kernel<<<blocks, threads, 0, stream>>>(ps.keys_out, N, ps.f_seg);
        tb = w.cub_bytes;
        GLOVE_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_temp, tb, ps.f_seg, ps.e_seg, N, stream));
        segs_kernel<<<blocks, threads, 0, stream>>>(ps.keys_out, N, idmask, ps, pv.side[s], pv.hdr, s);
        itemflags_kernel<<<blocks, threads, 0, stream>>>(N, ps, pv.side[s]);
        tb = w.cub_bytes;
        GLOVE_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_temp, tb, ps.f_item, ps.e_item, N, stream));
        tb = w.cub_bytes;
        GLOVE_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_temp, tb, ps.f_long, ps.e_long, N, stream));
        tb = w.cub_bytes;
        GLOVE_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_temp, tb, ps.f_part, ps.e_part, N, stream));
        items_kernel<<<blocks, threads, 0, stream>>>(N, B, K, ps, pv.side[s], pv.hdr, s);
        owners_kernel<<<(K * (kMaxShards + 1) + 255) / 256, 256, 0, stream>>>(K, n_shards, v_loc, pv.side[s]);
        upad_kernel<<<(K + 255) / 256, 256, 0, stream>>>(K, n_shards, pv.side[s]);
        owner_items_kernel<<<(K * (kMaxShards + 1) + 255) / 256, 256, 0, stream>>>(K, pv.side[s]);
        slots_kernel<<<blocks, threads, 0, stream>>>(N, B, n_shards, v_loc, ps, pv.side[s]);
        itemrec_kernel<<<blocks, threads, 0, stream>>>(B, n_shards, v_loc, &pv.hdr->n_item[s], pv.side[s]);
        longrec_kernel<<<blocks, threads, 0, stream>>>(B, n_shards, v_loc, &pv.hdr->n_long[s], pv.side[s]);
        segprev_kernel<<<blocks, threads, 0, stream>>>(N, B, ps, pv.side[s]);
        GLOVE_CHECK_LAUNCH();
    }
    for (int s = 0; s < 2; ++s)
        fill_kernel<<<blocks, threads, 0, stream>>>(N, B, w.side[s], w.side[1 - s].slot_of_p, w.A, w.Bv, pv.side[s]);
    GLOVE_CHECK_LAUNCH();
    if (n_shards > 1) {
        // request lists: which opposite rows does each shard need from each owner (sort + unique of (batch, shard, position))
        int pbits = 1;
        while ((1ll << pbits) < (int64_t)B + B / 4 + 64) ++pbits;   // positions < snapshot rows
        GLOVE_REQUIRE(pbits + 3 + kbits <= 32, "glove_prepare_batches: K=%d x B=%d does not fit the 32-bit request keys", K, B);
        for (int s = 0; s < 2; ++s) {
            PrepSide &ps = w.side[s];
            need_keys_kernel<<<blocks, threads, 0, stream>>>(N, B, v_loc, pbits, ps, pv.side[s]);
            size_t tb = w.cub_bytes;
            GLOVE_CHECK_CUDA(cub::DeviceRadixSort::SortKeys(w.cub_temp, tb, ps.keys_in, ps.keys_out, N, 0, pbits + 3 + kbits, stream));
            heads_kernel<<<blocks, threads, 0, stream>>>(ps.keys_out, N, ps.f_item);
            tb = w.cub_bytes;
            GLOVE_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_temp, tb, ps.f_item, ps.e_item, N, stream));
            need_compact_kernel<<<blocks, threads, 0, stream>>>(N, (1u << pbits) - 1, ps.keys_out, ps.f_item, ps.e_item,
                                                                ps.keys_in, pv.side[s].need_pos, ps.f_long);
            need_offsets_kernel<<<(K * kMaxShards * (kMaxShards + 1) + 255) / 256, 256, 0, stream>>>(
                K, n_shards, pbits, ps.keys_in, ps.f_long, pv.side[1 - s].b_upad, pv.side[s].need_off);
        }
        // the same lists seen from the owners: which shards does each staged row have to be pushed to
        for (int s = 0; s < 2; ++s) GLOVE_CHECK_CUDA(cudaMemsetAsync(pv.side[s].seg_push, 0, sizeof(int32_t) * (size_t)N, stream));
        for (int s = 0; s < 2; ++s) push_mask_kernel<<<K * n_shards, 256, 0, stream>>>(n_shards, pv.side[s], pv.side[1 - s]);
        GLOVE_CHECK_LAUNCH();
    }
    return GLOVE_OK;
}

int glove_plan_pull_slice(void *dst_plan, const void *src_plan, int32_t K, int32_t B, int32_t n_shards, int32_t shard,
                          void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GLOVE_REQUIRE(dst_plan && src_plan && K > 0 && B > 0, "glove_plan_pull_slice: bad arguments");
    GLOVE_REQUIRE(n_shards > 1 && n_shards <= kMaxShards && shard >= 0 && shard < n_shards,
                  "glove_plan_pull_slice: bad shard %d of %d", shard, n_shards);
    PlanView sv = plan_view(const_cast<void *>(src_plan), K, B), dv = plan_view(dst_plan, K, B);
    PullArgs a;
    a.src_hdr = sv.hdr; a.dst_hdr = dv.hdr;
    for (int s = 0; s < 2; ++s) { a.src[s] = sv.side[s]; a.dst[s] = dv.side[s]; }
    a.K = K; a.n_shards = n_shards; a.shard = shard;
    plan_pull_head_kernel<<<8, 256, 0, stream>>>(a);
    int nparts = (2 * kNumSMs + 2 * K - 1) / (2 * K);
    if (nparts < 1) nparts = 1;
    plan_pull_body_kernel<<<2 * K * nparts, 256, 0, stream>>>(a, nparts);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}

int glove_plan_need_info(const void *plan, int32_t K, int32_t B, int32_t *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GLOVE_REQUIRE(plan && out && K > 0, "glove_plan_need_info: bad arguments");
    PlanView pv = plan_view(const_cast<void *>(plan), K, B);
    const size_t n = (size_t)K * kMaxShards * (kMaxShards + 1);
    for (int s = 0; s < 2; ++s)
        GLOVE_CHECK_CUDA(cudaMemcpyAsync(out + s * n, pv.side[s].need_off, 4 * n, cudaMemcpyDeviceToHost, stream));
    GLOVE_CHECK_CUDA(cudaStreamSynchronize(stream));
    return GLOVE_OK;
}

int glove_plan_shard_info(const void *plan, int32_t K, int32_t B, int32_t k, int32_t *out20, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GLOVE_REQUIRE(plan && out20 && k >= 0 && k < K, "glove_plan_shard_info: bad arguments");
    PlanView pv = plan_view(const_cast<void *>(plan), K, B);
    for (int s = 0; s < 2; ++s) {
        GLOVE_CHECK_CUDA(cudaMemcpyAsync(out20 + 10 * s, pv.side[s].b_own + (size_t)k * (kMaxShards + 1),
                                         4 * (kMaxShards + 1), cudaMemcpyDeviceToHost, stream));
        GLOVE_CHECK_CUDA(cudaMemcpyAsync(out20 + 10 * s + 9, pv.side[s].b_upad + k, 4, cudaMemcpyDeviceToHost, stream));
    }
    GLOVE_CHECK_CUDA(cudaStreamSynchronize(stream));
    return GLOVE_OK;
}

int glove_plan_batch_counts(const void *plan, int32_t K, int32_t B, int32_t k, int32_t *out4, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GLOVE_REQUIRE(plan && out4 && k >= 0 && k < K, "glove_plan_batch_counts: bad arguments");
    PlanView pv = plan_view(const_cast<void *>(plan), K, B);
    int32_t tmp[2][2][2];
    for (int s = 0; s < 2; ++s) {
        GLOVE_CHECK_CUDA(cudaMemcpyAsync(tmp[s][0], pv.side[s].b_seg + k, 8, cudaMemcpyDeviceToHost, stream));
        GLOVE_CHECK_CUDA(cudaMemcpyAsync(tmp[s][1], pv.side[s].b_item + k, 8, cudaMemcpyDeviceToHost, stream));
    }
    GLOVE_CHECK_CUDA(cudaStreamSynchronize(stream));
    out4[0] = tmp[0][0][1] - tmp[0][0][0];
    out4[1] = tmp[1][0][1] - tmp[1][0][0];
    out4[2] = tmp[0][1][1] - tmp[0][1][0];
    out4[3] = tmp[1][1][1] - tmp[1][1][0];
    return GLOVE_OK;
}

}  // extern "C"
