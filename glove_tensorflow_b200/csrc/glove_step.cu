// TRAIN step: model_fn(mode=TRAIN) of the reference [ref src/models/estimator.py:13-56] as three sm_100a kernels.
//
//   stage_kernel   one warp per distinct row / col id of the batch: (replay missed idle Adam steps,) publish the
//                  pre-step row into a compact L2-resident snapshot cache[side][slot][S].
//   update_kernel  one warp per work item (<= kItemMax triples of one id, both sides in one launch): gather the
//                  opposite rows from the snapshot with 128-bit loads, warp-shuffle dot product, residual, loss,
//                  gradient accumulation in registers, and -- when the item is the whole segment -- the fused sparse
//                  optimizer update written in place to the packed table.  Because all forward reads come from the
//                  snapshot, the row side and the col side never race (SURVEY §7 hard part 2).
//   fix_kernel     segments longer than kItemMax: partial sums are combined in fixed order (one CTA per segment, one
//                  thread per column) and updated; the last CTA to finish reduces the per-item loss terms in fixed
//                  order, updates the scalar global bias, publishes the loss and increments the device step counter.
//
// Everything is deterministic: no floating-point atomics, fixed summation order given (B, kItemMax).
#include "glove_common.cuh"

namespace glove {

enum { MODE_TRAIN = 0, MODE_GRAD = 1, MODE_APPLY = 2 };

struct StepParams {
    float *table[2];
    glove_scalars *sc;
    const PlanHeader *hdr;
    PlanSide side[2];
    float *cache[2];
    float *partial[2];
    float4 *item_out[2];
    float *grad[2];       // MODE_GRAD / MODE_APPLY: dense per-slot gradient buffers
    float *grad_scalars;  // [4]
    const float *alpha;
    float *loss_out;
    int32_t alpha_len, loss_cap;
    int64_t V;
    int32_t d, S, P, B, K;
    int32_t head, opt, adam_mode, mode;
    float lr, l2, rs, nf, b1, b2, eps;
    int32_t dp_rank, dp_world, dp_block;
};

struct StepWs {
    float *cache[2];
    float *partial[2];
    float4 *item_out[2];
    size_t bytes;
};
static inline int64_t max_items_per_batch(int32_t B) { return (int64_t)B + B / kItemMax + 2; }
static inline int64_t max_parts_per_batch(int32_t B) { return 2 * (int64_t)B / kItemMax + 2; }
static StepWs step_ws_view(void *base, int32_t B, int32_t d) {
    StepWs w;
    char *p = (char *)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char *r = p ? p + off : nullptr; off += align_up(bytes); return r; };
    const int32_t S = table_stride(d);
    for (int s = 0; s < 2; ++s) w.cache[s] = (float *)take(sizeof(float) * (size_t)B * S);
    for (int s = 0; s < 2; ++s) w.partial[s] = (float *)take(sizeof(float) * (size_t)max_parts_per_batch(B) * S);
    for (int s = 0; s < 2; ++s) w.item_out[s] = (float4 *)take(sizeof(float4) * (size_t)max_items_per_batch(B));
    w.bytes = off;
    return w;
}

// ---- small device helpers ---------------------------------------------------------------------------------------
// value of column `col` of a row held as NV float4 per lane (float4 index f = lane + 32 r), broadcast to all lanes
template <int NV>
__device__ __forceinline__ float row_col(const float4 (&x)[NV], int col, int lane) {
    const int f = col >> 2, c = col & 3, r = f >> 5, src = f & 31;
    float v = 0.0f;
#pragma unroll
    for (int rr = 0; rr < NV; ++rr)
        if (rr == r) v = (c == 0 ? x[rr].x : c == 1 ? x[rr].y : c == 2 ? x[rr].z : x[rr].w);
    return __shfl_sync(0xffffffffu, v, src);
}

__device__ __forceinline__ bool batch_index(const StepParams &p, int &k, int &step) {
    step = p.sc->step;
    k = step - p.hdr->first_step;
    if (p.hdr->magic != kPlanMagic || k < 0 || k >= p.hdr->K || p.hdr->B != p.B ||
        (p.opt == GLOVE_OPT_ADAM && step >= p.alpha_len)) {
        if (threadIdx.x == 0 && blockIdx.x == 0) p.sc->error = 1;
        return false;
    }
    return true;
}

__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// residual e = dL_data/dz and the (un-normalised) data loss term of one triple
__device__ __forceinline__ void head_eval(int head, float z, float a, float b, float invB, float nf, float &e, float &l) {
    if (head == GLOVE_HEAD_GLOVE) {  // a = target, b = weight
        const float r = z - a;
        l = b * r * r;
        e = (2.0f * invB) * b * r;
    } else {  // a = pos weight (value), b = neg weight
        const float sg = sigmoid_f(z);
        l = a * softplus_f(-z) + nf * (b * softplus_f(z));
        e = (a * (sg - 1.0f) + nf * b * sg) * invB;
    }
}

// ---- K1: stage -----------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) stage_kernel(const StepParams p) {
    int k, step;
    if (!batch_index(p, k, step)) return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int seg0[2] = {p.side[0].b_seg[k], p.side[1].b_seg[k]};
    const int U0 = p.side[0].b_seg[k + 1] - seg0[0], U1 = p.side[1].b_seg[k + 1] - seg0[1];
    const int S4 = p.S >> 2;
    const bool replay = (p.opt == GLOVE_OPT_ADAM && p.adam_mode == GLOVE_ADAM_REPLAY);
    for (int w = warp; w < U0 + U1; w += nwarps) {
        const int s = w >= U0 ? 1 : 0;
        const int slot = s ? w - U0 : w;
        const int id = p.side[s].seg_id[seg0[s] + slot];
        const float *row = p.table[s] + (int64_t)id * p.P * p.S;
        float4 x[NV];
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int f = lane + 32 * r;
            x[r] = f < S4 ? ld4(row + 4 * f) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (replay) {
            const int ls = __float_as_int(row_col<NV>(x, p.d + 1, lane));
            if (ls > 0 && ls < step) {
                // Padding columns (col > d, and lanes beyond the row) hold m = v = 0: a 0 / eps division sends the whole
                // warp through the IEEE slow path on every step.  They get dummy operands (m = v = 1) in registers and
                // are rebuilt when the row is stored.
                float4 m[NV], v[NV];
#pragma unroll
                for (int r = 0; r < NV; ++r) {
                    const int f = lane + 32 * r;
                    m[r] = f < S4 ? ld4(row + p.S + 4 * f) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[r] = f < S4 ? ld4(row + 2 * p.S + 4 * f) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (4 * f + c > p.d) { f4c(m[r], c) = 1.0f; f4c(v[r], c) = 1.0f; }
                }
                for (int t = ls; t < step; ++t) {
                    const float a = __ldg(p.alpha + t);
                    bool changed = false;
#pragma unroll
                    for (int r = 0; r < NV; ++r) {
                        const int f = lane + 32 * r;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float old = f4c(x[r], c);
                            adam_idle_step(f4c(x[r], c), f4c(m[r], c), f4c(v[r], c), a, p.b1, p.b2, p.eps);
                            changed |= (4 * f + c <= p.d) && (f4c(x[r], c) != old);
                        }
                    }
                    // |increment| shrinks monotonically (x0.9 per step from m, at most x1.012 from alpha and sqrt(v)):
                    // once no element of the row moves, none ever will again -> the remaining steps are exact no-ops on x
                    if (!__any_sync(0xffffffffu, changed)) break;
                }
#pragma unroll
                for (int r = 0; r < NV; ++r) {
                    const int f = lane + 32 * r;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int col = 4 * f + c;
                        if (col == p.d + 1) f4c(x[r], c) = __int_as_float(ls);
                        else if (col > p.d + 1) f4c(x[r], c) = 0.0f;
                    }
                }
            }
        }
        float *dst = p.cache[s] + (int64_t)slot * p.S;
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int f = lane + 32 * r;
            if (f < S4) st4(dst + 4 * f, x[r]);
        }
    }
}

// ---- optimizer epilogue on a row held in registers ---------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void apply_row(const StepParams &p, float *row, float4 (&x)[NV], const float4 (&G)[NV],
                                          int ls, int step, int lane) {
    const int S4 = p.S >> 2;
    // Padding columns (col > d) carry G = m = v = 0; 0 / eps would drag the warp through the IEEE-division slow path,
    // so they run on dummy operands (1) and are written back as zeros.
    if (p.opt == GLOVE_OPT_ADAM) {
        float4 m[NV], v[NV];
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int f = lane + 32 * r;
            m[r] = f < S4 ? ld4(row + p.S + 4 * f) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[r] = f < S4 ? ld4(row + 2 * p.S + 4 * f) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (p.adam_mode == GLOVE_ADAM_REPLAY && ls > 0) {
            for (int t = ls; t < step; ++t) {
#pragma unroll
                for (int r = 0; r < NV; ++r) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        f4c(m[r], c) = __fmul_rn(f4c(m[r], c), p.b1);
                        f4c(v[r], c) = __fmul_rn(f4c(v[r], c), p.b2);
                    }
                }
            }
        }
        const float a = __ldg(p.alpha + step);
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int f = lane + 32 * r;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const bool pad = 4 * f + c > p.d;
                float mm = pad ? 1.0f : f4c(m[r], c), vv = pad ? 1.0f : f4c(v[r], c);
                adam_update(f4c(x[r], c), mm, vv, pad ? 1.0f : f4v(G[r], c), a, p.b1, p.b2, p.eps);
                f4c(m[r], c) = pad ? 0.0f : mm;
                f4c(v[r], c) = pad ? 0.0f : vv;
            }
            if (f < S4) { st4(row + p.S + 4 * f, m[r]); st4(row + 2 * p.S + 4 * f, v[r]); }
        }
    } else if (p.opt == GLOVE_OPT_ADAGRAD) {
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int f = lane + 32 * r;
            float4 acc = f < S4 ? ld4(row + p.S + 4 * f) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const bool pad = 4 * f + c > p.d;
                float aa = pad ? 1.0f : f4c(acc, c);
                adagrad_update(f4c(x[r], c), aa, pad ? 1.0f : f4v(G[r], c), p.lr, p.eps);
                f4c(acc, c) = pad ? 0.0f : aa;
            }
            if (f < S4) st4(row + p.S + 4 * f, acc);
        }
    } else {
#pragma unroll
        for (int r = 0; r < NV; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) sgd_update(f4c(x[r], c), f4v(G[r], c), p.lr);
    }
    // plane 0: new x, bias, last_step = step + 1, zero padding
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        const int f = lane + 32 * r;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int col = 4 * f + c;
            if (col == p.d + 1) f4c(x[r], c) = __int_as_float(step + 1);
            else if (col > p.d + 1) f4c(x[r], c) = 0.0f;
        }
        if (f < S4) st4(row + 4 * f, x[r]);
    }
}

// ---- K2: update ----------------------------------------------------------------------------------------------------
template <int NV, int G>
__global__ void __launch_bounds__(128) update_kernel(const StepParams p) {
    int k, step;
    if (!batch_index(p, k, step)) return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int it0[2] = {p.side[0].b_item[k], p.side[1].b_item[k]};
    const int nI0 = p.side[0].b_item[k + 1] - it0[0], nI1 = p.side[1].b_item[k + 1] - it0[1];
    const int S4 = p.S >> 2;
    const float gbias = p.sc->g;
    const float invB = 1.0f / (float)p.B;
    const float ce = (2.0f * p.rs * p.l2) / ((float)p.d * (float)p.B), cbias = (2.0f * p.rs * p.l2) / (float)p.B;

    for (int w = warp; w < nI0 + nI1; w += nwarps) {
        const int s = w >= nI0 ? 1 : 0;
        const int itl = s ? w - nI0 : w;
        const PlanSide &ps = p.side[s];
        const int it = it0[s] + itl;
        const int g = ps.item_seg[it];
        const int slot = g - ps.b_seg[k];
        const int start = ps.item_start[it];
        const int seg_begin = ps.seg_start[g], seg_end = ps.seg_start[g + 1];
        const int end = min(start + kItemMax, seg_end);
        const float *own = p.cache[s] + (int64_t)slot * p.S;
        const float *opp_base = p.cache[1 - s];

        float4 x[NV], dv[NV], acc[NV];
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int f = lane + 32 * r;
            x[r] = f < S4 ? ld4(own + 4 * f) : make_float4(0.f, 0.f, 0.f, 0.f);
            acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int col = 4 * f + c;
                f4c(dv[r], c) = col < p.d ? f4c(x[r], c) : (col == p.d ? 1.0f : 0.0f);
            }
        }
        const float bias_own = row_col<NV>(x, p.d, lane);
        const int ls = __float_as_int(row_col<NV>(x, p.d + 1, lane));
        float sum_e = 0.0f, loss_d = 0.0f;
        int n_eff = 0;

        for (int q = start; q < end; q += G) {
            float4 o[G][NV];
            float a[G], b[G];
            bool valid[G];
#pragma unroll
            for (int t = 0; t < G; ++t) {
                valid[t] = (q + t) < end;
                int os = 0;
                a[t] = 0.0f; b[t] = 0.0f;
                if (valid[t] && p.dp_world > 1) valid[t] = (ps.owner[q + t] / p.dp_block) == p.dp_rank;
                if (valid[t]) { os = ps.oslot[q + t]; a[t] = ps.a[q + t]; b[t] = ps.b[q + t]; }
                const float *orow = opp_base + (int64_t)os * p.S;
#pragma unroll
                for (int r = 0; r < NV; ++r) {
                    const int f = lane + 32 * r;
                    o[t][r] = (valid[t] && f < S4) ? ld4_nc(orow + 4 * f) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            float dot[G];
#pragma unroll
            for (int t = 0; t < G; ++t) {
                float acc_d = 0.0f;
#pragma unroll
                for (int r = 0; r < NV; ++r)
                    acc_d += dv[r].x * o[t][r].x + dv[r].y * o[t][r].y + dv[r].z * o[t][r].z + dv[r].w * o[t][r].w;
                dot[t] = acc_d;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                for (int t = 0; t < G; ++t) dot[t] += __shfl_xor_sync(0xffffffffu, dot[t], off);
            }
#pragma unroll
            for (int t = 0; t < G; ++t) {
                if (!valid[t]) continue;  // warp-uniform
                const float z = (dot[t] + bias_own) + gbias;  // dot already holds the opposite bias (dv[col d] = 1)
                float e, l;
                head_eval(p.head, z, a[t], b[t], invB, p.nf, e, l);
                sum_e += e;
                loss_d += l;
                ++n_eff;
#pragma unroll
                for (int r = 0; r < NV; ++r) {
                    acc[r].x += e * o[t][r].x; acc[r].y += e * o[t][r].y;
                    acc[r].z += e * o[t][r].z; acc[r].w += e * o[t][r].w;
                }
            }
        }

        // activity-L2 gradient (n occurrences of this id) and loss term; bias gradient goes to column d
        const float fn = (float)n_eff;
        float sq = 0.0f;
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int f = lane + 32 * r;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int col = 4 * f + c;
                const float xv = f4c(x[r], c);
                if (col < p.d) { f4c(acc[r], c) += fn * ce * xv; sq += xv * xv; }
                else if (col == p.d) f4c(acc[r], c) = sum_e + fn * cbias * xv;
                else f4c(acc[r], c) = 0.0f;
            }
        }
        sq = warp_sum(sq);
        if (lane == 0) {
            const float regp = fn * ((p.l2 / (float)p.d) * sq + p.l2 * bias_own * bias_own);
            p.item_out[s][itl] = make_float4(s == 0 ? loss_d : 0.0f, s == 0 ? sum_e : 0.0f, regp, 0.0f);
        }

        const bool whole = (seg_end - seg_begin) <= kItemMax;
        if (!whole) {
            float *dst = p.partial[s] + (int64_t)(ps.item_part[it] - ps.b_part[k]) * p.S;
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                const int f = lane + 32 * r;
                if (f < S4) st4(dst + 4 * f, acc[r]);
            }
        } else if (p.mode == MODE_GRAD) {
            float *dst = p.grad[s] + (int64_t)slot * p.S;
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                const int f = lane + 32 * r;
                if (f < S4) st4(dst + 4 * f, acc[r]);
            }
        } else {
            float *row = p.table[s] + (int64_t)ps.seg_id[g] * p.P * p.S;
            apply_row<NV>(p, row, x, acc, ls, step, lane);
        }
    }
}

// ---- DP apply: one warp per segment, gradient comes from the all-reduced dense buffer -------------------------------
template <int NV>
__global__ void __launch_bounds__(128) apply_kernel(const StepParams p) {
    int k, step;
    if (!batch_index(p, k, step)) return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int seg0[2] = {p.side[0].b_seg[k], p.side[1].b_seg[k]};
    const int U0 = p.side[0].b_seg[k + 1] - seg0[0], U1 = p.side[1].b_seg[k + 1] - seg0[1];
    const int S4 = p.S >> 2;
    for (int w = warp; w < U0 + U1; w += nwarps) {
        const int s = w >= U0 ? 1 : 0;
        const int slot = s ? w - U0 : w;
        const int id = p.side[s].seg_id[seg0[s] + slot];
        float *row = p.table[s] + (int64_t)id * p.P * p.S;
        float4 x[NV], G[NV];
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int f = lane + 32 * r;
            x[r] = f < S4 ? ld4(p.cache[s] + (int64_t)slot * p.S + 4 * f) : make_float4(0.f, 0.f, 0.f, 0.f);
            G[r] = f < S4 ? ld4(p.grad[s] + (int64_t)slot * p.S + 4 * f) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const int ls = __float_as_int(row_col<NV>(x, p.d + 1, lane));
        apply_row<NV>(p, row, x, G, ls, step, lane);
    }
}

// ---- K3: long segments + finish --------------------------------------------------------------------------------------
__device__ void finish_step(const StepParams &p, int k, int step, const float *reduced /* MODE_APPLY */) {
    __shared__ double sh[3][256];
    const int tid = threadIdx.x;
    double ld = 0.0, se = 0.0, rg = 0.0;
    if (reduced == nullptr) {
        const int nI0 = p.side[0].b_item[k + 1] - p.side[0].b_item[k];
        const int nI1 = p.side[1].b_item[k + 1] - p.side[1].b_item[k];
        int i = tid;
        for (; i + 7 * 256 < nI0; i += 8 * 256) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = p.item_out[0][i + u * 256];
#pragma unroll
            for (int u = 0; u < 8; ++u) { ld += v[u].x; se += v[u].y; rg += v[u].z; }
        }
        for (; i < nI0; i += 256) { const float4 v = p.item_out[0][i]; ld += v.x; se += v.y; rg += v.z; }
        i = tid;
        for (; i + 7 * 256 < nI1; i += 8 * 256) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = p.item_out[1][i + u * 256];
#pragma unroll
            for (int u = 0; u < 8; ++u) rg += v[u].z;
        }
        for (; i < nI1; i += 256) { const float4 v = p.item_out[1][i]; rg += v.z; }
    }
    sh[0][tid] = ld; sh[1][tid] = se; sh[2][tid] = rg;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) { sh[0][tid] += sh[0][tid + o]; sh[1][tid] += sh[1][tid + o]; sh[2][tid] += sh[2][tid + o]; }
        __syncthreads();
    }
    if (tid != 0) return;
    if (reduced) { ld = reduced[0]; se = reduced[1]; rg = reduced[2]; } else { ld = sh[0][0]; se = sh[1][0]; rg = sh[2][0]; }
    if (p.mode == MODE_GRAD) {
        p.grad_scalars[0] = (float)ld; p.grad_scalars[1] = (float)se; p.grad_scalars[2] = (float)rg; p.grad_scalars[3] = 0.0f;
        p.sc->ticket = 0;
        return;
    }
    const double B = (double)p.B;
    float g = p.sc->g;
    const double reg = (double)p.rs * (rg / B + (double)p.l2 * (double)g * (double)g);
    const float loss = (float)(ld / B + reg);
    const float dg = (float)se + (2.0f * p.rs * p.l2) * g;
    if (p.opt == GLOVE_OPT_ADAM) {  // dense ResourceApplyAdam form for the scalar variable
        const float a = p.alpha[step];
        float m = p.sc->g_s0, v = p.sc->g_s1;
        m = __fadd_rn(m, __fmul_rn(__fsub_rn(dg, m), __fsub_rn(1.0f, p.b1)));
        v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(dg, dg), v), __fsub_rn(1.0f, p.b2)));
        g = __fsub_rn(g, __fdiv_rn(__fmul_rn(a, m), __fadd_rn(__fsqrt_rn(v), p.eps)));
        p.sc->g_s0 = m; p.sc->g_s1 = v;
    } else if (p.opt == GLOVE_OPT_ADAGRAD) {
        float acc = p.sc->g_s0;
        adagrad_update(g, acc, dg, p.lr, p.eps);
        p.sc->g_s0 = acc;
    } else {
        sgd_update(g, dg, p.lr);
    }
    p.sc->g = g;
    p.sc->loss = loss;
    if (p.loss_out) p.loss_out[step % p.loss_cap] = loss;
    p.sc->ticket = 0;
    __threadfence();
    p.sc->step = step + 1;
}

__global__ void __launch_bounds__(256) fix_kernel(const StepParams p) {
    int k, step;
    if (!batch_index(p, k, step)) return;
    const int L0[2] = {p.side[0].b_long[k], p.side[1].b_long[k]};
    const int nL0 = p.side[0].b_long[k + 1] - L0[0], nL1 = p.side[1].b_long[k + 1] - L0[1];
    const float a = p.opt == GLOVE_OPT_ADAM ? p.alpha[step] : 0.0f;
    for (int l = blockIdx.x; l < nL0 + nL1; l += gridDim.x) {
        const int s = l >= nL0 ? 1 : 0;
        const PlanSide &ps = p.side[s];
        const int ll = L0[s] + (s ? l - nL0 : l);
        const int g = ps.long_seg[ll], it0 = ps.long_item[ll];
        const int slot = g - ps.b_seg[k];
        const int len = ps.seg_start[g + 1] - ps.seg_start[g];
        const int npieces = (len + kItemMax - 1) / kItemMax;
        const float *part = p.partial[s] + (int64_t)(ps.item_part[it0] - ps.b_part[k]) * p.S;
        float *row = p.table[s] + (int64_t)ps.seg_id[g] * p.P * p.S;
        const float *crow = p.cache[s] + (int64_t)slot * p.S;
        const int ls = __float_as_int(crow[p.d + 1]);
        for (int c = threadIdx.x; c < p.S; c += blockDim.x) {
            float G = 0.0f;
            int q = 0;
            for (; q + 8 <= npieces; q += 8) {  // 8 independent loads in flight, additions stay in piece order
                float t[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] = part[(int64_t)(q + u) * p.S + c];
#pragma unroll
                for (int u = 0; u < 8; ++u) G += t[u];
            }
            for (; q < npieces; ++q) G += part[(int64_t)q * p.S + c];
            if (p.mode == MODE_GRAD) { p.grad[s][(int64_t)slot * p.S + c] = G; continue; }
            if (c <= p.d) {
                float x = crow[c];
                if (p.opt == GLOVE_OPT_ADAM) {
                    float m = row[p.S + c], v = row[2 * p.S + c];
                    if (p.adam_mode == GLOVE_ADAM_REPLAY && ls > 0)
                        for (int t = ls; t < step; ++t) { m = __fmul_rn(m, p.b1); v = __fmul_rn(v, p.b2); }
                    adam_update(x, m, v, G, a, p.b1, p.b2, p.eps);
                    row[p.S + c] = m; row[2 * p.S + c] = v;
                } else if (p.opt == GLOVE_OPT_ADAGRAD) {
                    float acc = row[p.S + c];
                    adagrad_update(x, acc, G, p.lr, p.eps);
                    row[p.S + c] = acc;
                } else {
                    sgd_update(x, G, p.lr);
                }
                row[c] = x;
            } else if (c == p.d + 1) {
                row[c] = __int_as_float(step + 1);
            }
        }
    }
    // last CTA to arrive finishes the step (fixed-order reduction of the per-item loss terms)
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&p.sc->ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (is_last) {
        __threadfence();
        finish_step(p, k, step, nullptr);
    }
}

__global__ void __launch_bounds__(256) apply_finish_kernel(const StepParams p, const float *reduced) {
    int k, step;
    if (!batch_index(p, k, step)) return;
    finish_step(p, k, step, reduced);
}

// ---- host side -------------------------------------------------------------------------------------------------------
static int fill_params(const glove_step_args *a, StepParams &p, int mode) {
    GLOVE_REQUIRE(a, "step: null args");
    GLOVE_REQUIRE(a->row_table && a->col_table && a->scalars && a->plan && a->workspace, "step: null pointer in args");
    GLOVE_REQUIRE(a->V > 0 && a->d > 0 && a->B > 0 && a->plan_K > 0, "step: bad sizes");
    GLOVE_REQUIRE(a->head == GLOVE_HEAD_GLOVE || a->head == GLOVE_HEAD_LOGISTIC, "step: unsupported head %d", a->head);
    GLOVE_REQUIRE(a->optimizer >= 0 && a->optimizer <= 2, "step: unsupported optimizer %d", a->optimizer);
    if (a->optimizer == GLOVE_OPT_ADAM) GLOVE_REQUIRE(a->alpha && a->alpha_len > 0, "step: Adam needs the alpha table");
    const int32_t S = table_stride(a->d);
    if (S / 4 > 32 * 4) return set_error(GLOVE_EUNSUPPORTED, "step: embedding size %d > 510 not supported", a->d);
    StepWs w = step_ws_view(a->workspace, a->B, a->d);
    if (a->workspace_bytes < w.bytes)
        return set_error(GLOVE_EWORKSPACE, "step: workspace %zu < required %zu", a->workspace_bytes, w.bytes);
    PlanView pv = plan_view(const_cast<void *>(a->plan), a->plan_K, a->B);
    p.table[0] = a->row_table; p.table[1] = a->col_table;
    p.sc = a->scalars;
    p.hdr = pv.hdr;
    for (int s = 0; s < 2; ++s) {
        p.side[s] = pv.side[s];
        p.cache[s] = w.cache[s]; p.partial[s] = w.partial[s]; p.item_out[s] = w.item_out[s];
        p.grad[s] = nullptr;
    }
    p.grad_scalars = nullptr;
    p.alpha = a->alpha; p.alpha_len = a->alpha_len;
    p.loss_out = a->loss_cap > 0 ? a->loss_out : nullptr; p.loss_cap = a->loss_cap > 0 ? a->loss_cap : 1;
    p.V = a->V; p.d = a->d; p.S = S; p.P = table_planes(a->optimizer); p.B = a->B; p.K = a->plan_K;
    p.head = a->head; p.opt = a->optimizer; p.adam_mode = a->adam_mode; p.mode = mode;
    p.lr = a->learning_rate; p.l2 = a->l2_reg; p.rs = a->reg_scale; p.nf = a->neg_factor;
    p.b1 = a->beta1; p.b2 = a->beta2; p.eps = a->epsilon;
    p.dp_world = mode == MODE_TRAIN ? 1 : (a->dp_world > 1 ? a->dp_world : 1);
    p.dp_rank = a->dp_rank;
    if (p.dp_world > 1) GLOVE_REQUIRE(a->B % p.dp_world == 0 && a->dp_rank >= 0 && a->dp_rank < p.dp_world, "step: bad dp split");
    p.dp_block = a->B / p.dp_world;
    return GLOVE_OK;
}

template <typename Kern>
static int occupancy_grid(Kern kern, int threads) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    int dev = 0, sms = kNumSMs;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return per_sm * sms;
}

template <int NV>
static int launch_step(const StepParams &p, cudaStream_t stream, cudaEvent_t *ev = nullptr) {
    static int g_stage = 0, g_update = 0, g_apply = 0;
    if (!g_stage) g_stage = occupancy_grid(stage_kernel<NV>, 256);
    if (!g_update) g_update = occupancy_grid(update_kernel<NV, 4>, 128);
    if (!g_apply) g_apply = occupancy_grid(apply_kernel<NV>, 128);
    if (p.mode == MODE_TRAIN || p.mode == MODE_GRAD) {
        if (ev) cudaEventRecord(ev[0], stream);
        stage_kernel<NV><<<g_stage, 256, 0, stream>>>(p);
        if (ev) cudaEventRecord(ev[1], stream);
        update_kernel<NV, 4><<<g_update, 128, 0, stream>>>(p);
        if (ev) cudaEventRecord(ev[2], stream);
        fix_kernel<<<kNumSMs, 256, 0, stream>>>(p);
        if (ev) cudaEventRecord(ev[3], stream);
    } else {
        apply_kernel<NV><<<g_apply, 128, 0, stream>>>(p);
        apply_finish_kernel<<<1, 256, 0, stream>>>(p, p.grad_scalars);
    }
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}

static int dispatch(const StepParams &p, cudaStream_t stream, cudaEvent_t *ev = nullptr) {
    const int nv = (p.S / 4 + 31) / 32;
    switch (nv) {
        case 1: return launch_step<1>(p, stream, ev);
        case 2: return launch_step<2>(p, stream, ev);
        case 3: return launch_step<3>(p, stream, ev);
        case 4: return launch_step<4>(p, stream, ev);
    }
    return set_error(GLOVE_EUNSUPPORTED, "step: stride %d not supported", p.S);
}

}  // namespace glove

using namespace glove;

extern "C" {

size_t glove_step_workspace_bytes(int32_t B, int32_t d) {
    if (B <= 0 || d <= 0) return 0;
    return step_ws_view(nullptr, B, d).bytes;
}

int glove_train_step(const glove_step_args *args, void *stream) {
    StepParams p;
    int rc = fill_params(args, p, MODE_TRAIN);
    if (rc != GLOVE_OK) return rc;
    return dispatch(p, (cudaStream_t)stream);
}

int glove_train_step_profiled(const glove_step_args *args, void *stream_, float *ms3) {
    cudaStream_t stream = (cudaStream_t)stream_;
    StepParams p;
    int rc = fill_params(args, p, MODE_TRAIN);
    if (rc != GLOVE_OK) return rc;
    GLOVE_REQUIRE(ms3, "glove_train_step_profiled: null output");
    cudaEvent_t ev[4];
    for (int i = 0; i < 4; ++i) GLOVE_CHECK_CUDA(cudaEventCreate(&ev[i]));
    rc = dispatch(p, stream, ev);
    if (rc == GLOVE_OK) {
        GLOVE_CHECK_CUDA(cudaEventSynchronize(ev[3]));
        for (int i = 0; i < 3; ++i) cudaEventElapsedTime(&ms3[i], ev[i], ev[i + 1]);
    }
    for (int i = 0; i < 4; ++i) cudaEventDestroy(ev[i]);
    return rc;
}

int glove_grad_step(const glove_step_args *args, float *grad_rows, float *grad_cols, float *grad_scalars, void *stream) {
    StepParams p;
    int rc = fill_params(args, p, MODE_GRAD);
    if (rc != GLOVE_OK) return rc;
    GLOVE_REQUIRE(grad_rows && grad_cols && grad_scalars, "glove_grad_step: null gradient buffer");
    p.grad[0] = grad_rows; p.grad[1] = grad_cols; p.grad_scalars = grad_scalars;
    return dispatch(p, (cudaStream_t)stream);
}

int glove_apply_step(const glove_step_args *args, const float *grad_rows, const float *grad_cols,
                     const float *grad_scalars, void *stream) {
    StepParams p;
    int rc = fill_params(args, p, MODE_APPLY);
    if (rc != GLOVE_OK) return rc;
    GLOVE_REQUIRE(grad_rows && grad_cols && grad_scalars, "glove_apply_step: null gradient buffer");
    p.grad[0] = const_cast<float *>(grad_rows); p.grad[1] = const_cast<float *>(grad_cols);
    p.grad_scalars = const_cast<float *>(grad_scalars);
    return dispatch(p, (cudaStream_t)stream);
}

}  // extern "C"
