// K2, lane-group version (the default update kernel): G = 8 / 16 / 32 lanes per work item, 32 / G items per warp side by side.
//
// What ncu shows for the one-warp-per-item kernels (update_kernel / update_kernel3; profiles/r02_update_kernels.md): 127
// registers -> 16 warps per SM, each of them walking one item's chain of dependent instructions at ~1 instruction per 13
// cycles (5-step shuffle reductions, L2 round trips), ~840 warp-instructions per item of which ~680 are per-item overhead
// rather than per-triple work, and a quarter of every row-wide instruction wasted (a 304-float row fills 76 of the 96
// float4 slots of a warp).  The kernel is bound by instruction latency at low occupancy, not by HBM or L2 bandwidth.
// Here a d = 300 row is spread over 16 lanes x 5 float4 (76 of 80 slots), so one warp-instruction serves TWO items: the
// same 16 warps per SM keep twice as many items in flight, every per-item instruction is amortised over two items, and a
// dot product needs 4 shuffle steps instead of 5.  The optimizer slot planes are no longer parked in registers for the
// whole item: the NEXT item's [m | v] lines are prefetched into L2 at item start (prefetch.global.L2, no register, no
// shared memory) and read chunk by chunk in the epilogue.
//
// Work split and arithmetic per element are those of update_kernel; the summation order inside a dot product / a row
// reduction differs (16-lane tree), so results agree with update_kernel to rounding, not bit for bit.  Deterministic
// (fixed order given B, kItemMax and the grid).
#pragma once

namespace glove {

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int G, int NV>
__device__ __forceinline__ void g_load_row(float4 (&x)[NV], const float *row, int l, int S4, bool on) {
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        const int f = l + G * r;
        x[r] = (on && (r < NV - 1 || f < S4)) ? ld4(row + 4 * f) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
template <int G, int NV>
__device__ __forceinline__ void g_store_row(float *row, const float4 (&x)[NV], int l, int S4) {
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        const int f = l + G * r;
        if (r < NV - 1 || f < S4) st4(row + 4 * f, x[r]);
    }
}
// acc = sum_{i < n} rows[i * stride_rows] in index order (L2-coherent loads: the rows were written by other SMs)
template <int G, int NV>
__device__ __forceinline__ void g_sum_partials(float4 (&acc)[NV], const float *base, int n, int stride_rows, int S,
                                               float4 (&b0)[NV], int l, int S4) {
#pragma unroll
    for (int r = 0; r < NV; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int q = 0; q < n; ++q) {
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int f = l + G * r;
            const float *p0 = base + (int64_t)q * stride_rows * S + 4 * f;
            b0[r] = (r < NV - 1 || f < S4) ? __ldcg(reinterpret_cast<const float4 *>(p0)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int r = 0; r < NV; ++r) { acc[r].x += b0[r].x; acc[r].y += b0[r].y; acc[r].z += b0[r].z; acc[r].w += b0[r].w; }
    }
}

// store without the compiler-level memory barrier of st4_hint (a group never re-reads a table row it has written, so
// the next chunk's loads may be scheduled above it)
__device__ __forceinline__ void st4_stream(float *p, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol));
}
// optimizer epilogue of one row: x = pre-step snapshot row (registers), acc = de-duplicated gradient; the slot planes
// (prefetched into L2 while the previous item ran) are loaded into the two dead gather buffers, all chunks back to back,
// then x / m / v are updated and written once
template <int G, int NV>
__device__ __forceinline__ void g_apply_row(const StepParams &p, float *row, const float4 (&x)[NV], const float4 (&acc)[NV],
                                            float4 (&m)[NV], float4 (&v)[NV], int s, int step, int l, int gap, uint64_t pol) {
    const int S = p.S, S4 = S >> 2;
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        const int f = l + G * r;
        const bool in = r < NV - 1 || f < S4;
        m[r] = (in && p.P >= 2 && !(p.ablate & 4)) ? __ldcg(reinterpret_cast<const float4 *>(row + S + 4 * f)) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[r] = (in && p.P >= 3 && !(p.ablate & 4)) ? __ldcg(reinterpret_cast<const float4 *>(row + 2 * S + 4 * f)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float na = p.opt == GLOVE_OPT_ADAM ? -__ldg(p.alpha + step) : 0.0f;
    float dm = 1.0f, dv = 1.0f;
    if (gap > 0) replay_decay(gap, p.l2b1, p.l2b2, dm, dv);
    const int lcol = ls_col(p.d, s);
    const int f_ls = lcol >> 2, c_ls = lcol & 3;
    const float lsv = __int_as_float(step + 1);
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        const int f = l + G * r;
        if (r < NV - 1 || f < S4) {
            float4 x4 = x[r];
            optimizer_chunk(p, x4, m[r], v[r], acc[r], na, dm, dv, gap > 0);
            if (p.P >= 2) st4_stream(row + S + 4 * f, m[r], pol);
            if (p.P >= 3) st4_stream(row + 2 * S + 4 * f, v[r], pol);
            if (f == f_ls) {
                x4.x = c_ls == 0 ? lsv : x4.x; x4.y = c_ls == 1 ? lsv : x4.y;
                x4.z = c_ls == 2 ? lsv : x4.z; x4.w = c_ls == 3 ? lsv : x4.w;
            }
            st4_stream(row + 4 * f, x4, pol);
        }
    }
}

template <int G, int NV, int HEAD, bool DP, int SIDE>
__device__ __forceinline__ void k5_side(const StepParams &p, const int k, const int step, const int lane, const int warp,
                                        const int nwarps, double *const wa) {
    constexpr int s = SIDE;
    constexpr int NG = 32 / G;                         // items per warp, side by side
    const int grp = lane / G, l = lane % G;
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u) << (grp * G));
    const PlanSide &ps = p.side[s];
    const int S = p.S, S4 = S >> 2;
    const bool train = p.mode == MODE_TRAIN || p.mode == MODE_SHARD;
    const bool closed = p.opt == GLOVE_OPT_ADAM && p.adam_mode == GLOVE_ADAM_REPLAY;
    const int64_t id0 = (int64_t)p.shard * p.v_loc;
    const int oi0 = p.mode == MODE_SHARD ? ps.b_own_item[k * (kMaxShards + 1) + p.shard] : 0;
    const int4 *const irec = ps.item_rec + ps.b_item[k] + oi0;
    const int nI = (p.mode == MODE_SHARD ? ps.b_own_item[k * (kMaxShards + 1) + p.shard + 1] : ps.b_item[k + 1] - ps.b_item[k]) - oi0;
    const int stride = nwarps * NG;
    const int bcol = bias_col(p.d, s);
    const float gbias = p.sc->g;
    const int slot_bytes = (p.P - 1) * S * 4;
    const int4 zero4 = make_int4(0, 0, 0, 0);
    auto load_opp = [&](float4 (&buf)[NV], int pos, bool on) {
        g_load_row<G, NV>(buf, p.snap[1 - s] + (int64_t)((p.ablate & 2) ? (pos & 63) : pos) * S, l, S4, on);
    };
    // L2 prefetch of the slot planes an item will read in its epilogue (whole 128-byte lines)
    auto prefetch_slots = [&](const int4 &it, bool on) {
        if (on && train && (it.w >> 8) == 0 && p.P >= 2 && !(p.ablate & 8)) {
            const char *mv = reinterpret_cast<const char *>(p.table[s] + ((int64_t)it.x - id0) * p.P * S + S);
            for (int o = 128 * l; o < slot_bytes; o += 128 * G) prefetch_l2(mv + o);
        }
    };

    // L1 prefetch of one opposite snapshot row (whole 128-byte lines, one per lane): the gather proper then hits L1, so
    // the depth of the gather pipeline costs no registers
    const int row_bytes = S * 4;
    auto prefetch_opp = [&](int pos, bool on) {
        if (p.ablate & 16) return;
        if (on && 128 * l < row_bytes) prefetch_l1(reinterpret_cast<const char *>(p.snap[1 - s] + (int64_t)pos * S) + 128 * l);
        if (G * 128 < row_bytes && on && 128 * (l + G) < row_bytes)
            prefetch_l1(reinterpret_cast<const char *>(p.snap[1 - s] + (int64_t)pos * S) + 128 * (l + G));
    };

    int it = warp * NG + grp;
    if (it - grp >= nI) return;                          // warp-uniform
    // pipeline state at the top of an iteration: ir, x (own snapshot row) and rq0 / rq1 (records of the first two triples)
    // of the CURRENT item are loaded or in flight and the opposite rows of those two triples are on their way into L1;
    // ir_next is the item after it
    int4 ir = it < nI ? __ldg(irec + it) : zero4;
    int4 ir_next = it + stride < nI ? __ldg(irec + it + stride) : zero4;
    int4 rq0 = (ir.w & 0xff) ? __ldg(ps.rec + ir.z) : zero4;
    int4 rq1 = (ir.w & 0xff) > 1 ? __ldg(ps.rec + ir.z + 1) : zero4;
    float4 x[NV], acc[NV], bufA[NV], bufB[NV];
    g_load_row<G, NV>(x, p.snap[s] + (int64_t)ir.y * S, l, S4, (ir.w & 0xff) != 0);
    load_opp(bufA, rq0.x, (ir.w & 0xff) != 0);
    prefetch_opp(rq1.x, (ir.w & 0xff) > 1);
    prefetch_slots(ir, true);
#pragma unroll 1
    while (it - grp < nI) {
        const int n = ir.w & 0xff, part = ir.w >> 8, slot = ir.y, tok = ir.x, start = ir.z;   // n == 0: no item in this group
        const bool applies = train && part == 0 && n > 0;
        const int gap_own = (applies && closed) ? __ldcg(p.gap[s] + slot) : 0;
        const int n_next = ir_next.w & 0xff;
        prefetch_slots(ir_next, n_next != 0);
        const int4 rq0_next = n_next ? __ldg(ps.rec + ir_next.z) : zero4;
        const int4 rq1_next = n_next > 1 ? __ldg(ps.rec + ir_next.z + 1) : zero4;
#pragma unroll
        for (int r = 0; r < NV; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        float loss_d = 0.0f, sum_e = 0.0f;
        int n_eff = 0;
        auto triple = [&](const float4 (&y)[NV], const int4 rq, bool on) {
            float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                a2 = __ffma2_rn(make_float2(x[r].x, x[r].y), make_float2(y[r].x, y[r].y), a2);
                a2 = __ffma2_rn(make_float2(x[r].z, x[r].w), make_float2(y[r].z, y[r].w), a2);
            }
            const float z = group_sum<G>(a2.x + a2.y) + gbias;
            float e, lo;
            head_eval(HEAD, z, __int_as_float(rq.y), __int_as_float(rq.z), p.invB, p.nf, e, lo);
            bool mine = on;
            if (DP) mine = on && (rq.w / p.dp_block == p.dp_rank);
            e = mine ? e : 0.f; lo = mine ? lo : 0.f; n_eff += mine;
            loss_d += lo; sum_e += e;
            const float2 ee = make_float2(e, e);
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                const float2 lo2 = __ffma2_rn(ee, make_float2(y[r].x, y[r].y), make_float2(acc[r].x, acc[r].y));
                const float2 hi2 = __ffma2_rn(ee, make_float2(y[r].z, y[r].w), make_float2(acc[r].z, acc[r].w));
                acc[r] = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
            }
        };
        const int nmax = NG == 1 ? n : __reduce_max_sync(0xffffffffu, n);
        // two triples per trip, ping-pong register buffers: the row of triple q+1 is in flight while triple q is computed
#pragma unroll 1
        for (int q = 0; q < nmax; q += 2) {
            load_opp(bufB, rq1.x, q + 1 < n);
            const int4 rq2 = q + 2 < n ? __ldg(ps.rec + start + q + 2) : zero4;
            const int4 rq3 = q + 3 < n ? __ldg(ps.rec + start + q + 3) : zero4;
            triple(bufA, rq0, q < n);
            load_opp(bufA, rq2.x, q + 2 < n);
            triple(bufB, rq1, q + 1 < n);
            rq0 = rq2; rq1 = rq3;
        }
        // the first two gathers of the next item travel into L1 while this item's epilogue runs (the register buffers
        // hold the slot planes there)
        prefetch_opp(rq0_next.x, n_next != 0);
        prefetch_opp(rq1_next.x, n_next > 1);
        const int4 ir_nn = it + 2 * stride < nI ? __ldg(irec + it + 2 * stride) : zero4;

        // activity-L2: gradient n coef_c x_c and loss n (l2/d sum x^2 + l2 bias^2)
        const float fn = (float)n_eff;
        {
            float2 sq2 = make_float2(0.f, 0.f);
            float sq1 = 0.0f;
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                const int f = l + G * r;
                if (r < NV - 1 || f < S4) activity_chunk(acc[r], x[r], 4 * f, fn, p.ce, p.cbias, p.d, bcol, sq2, sq1);
            }
            const float sq = group_sum<G>(sq1 + (sq2.x + sq2.y));
            if (l == 0 && n > 0) {
                double *w = wa + 3 * grp;
                if (s == 0) { w[0] += (double)loss_d; w[1] += (double)sum_e; }
                w[2] += (double)(fn * p.reg_unscale * sq);
            }
        }
        const uint64_t pol_stream = p.l2_hints ? l2_policy_evict_first() : l2_policy_evict_normal();
        if (n > 0) {
            if (part) {
                // piece of a split segment: two-level, fixed-order combine inside the launch (see update_kernel)
                g_store_row<G, NV>(p.partial[s] + (int64_t)(part - 1) * S, acc, l, S4);
                const int4 lr = __ldg(ps.long_rec + ps.b_long[k] + tok);   // {token id, slot, first partial, pieces}
                const int piece = (part - 1) - lr.z, chunk = piece / kChunk;
                const int c_first = lr.z + chunk * kChunk;
                const int c_n = min(kChunk, lr.w - chunk * kChunk);
                __threadfence();
                int last = 0;
                if (l == 0) last = atomicAdd(p.chunk_cnt[s] + c_first, 1) == c_n - 1;
                if (__shfl_sync(gmask, last, grp * G)) {
                    __threadfence();
                    if (l == 0) p.chunk_cnt[s][c_first] = 0;
                    g_sum_partials<G, NV>(acc, p.partial[s] + (int64_t)c_first * S, c_n, 1, S, bufA, l, S4);
                    const int n_chunks = (lr.w + kChunk - 1) / kChunk;
                    if (n_chunks > 1) g_store_row<G, NV>(p.partial[s] + (int64_t)c_first * S, acc, l, S4);
                    __threadfence();
                    last = 0;
                    if (l == 0) last = atomicAdd(p.long_cnt[s] + tok, 1) == n_chunks - 1;
                    if (__shfl_sync(gmask, last, grp * G)) {
                        __threadfence();
                        if (l == 0) p.long_cnt[s][tok] = 0;     // ready for the next step
                        if (n_chunks > 1)
                            g_sum_partials<G, NV>(acc, p.partial[s] + (int64_t)lr.z * S, n_chunks, kChunk, S, bufA, l, S4);
                        if (!train) {
                            g_store_row<G, NV>(p.grad[s] + (int64_t)lr.y * S, acc, l, S4);
                        } else {
                            // x still holds the segment's snapshot row (all pieces share the slot)
                            g_apply_row<G, NV>(p, p.table[s] + ((int64_t)lr.x - id0) * p.P * S, x, acc, bufA, bufB, s, step, l,
                                               closed ? __ldcg(p.gap[s] + lr.y) : 0, pol_stream);
                        }
                    }
                }
            } else if (!train) {
                g_store_row<G, NV>(p.grad[s] + (int64_t)slot * S, acc, l, S4);
            } else {
                g_apply_row<G, NV>(p, (p.ablate & 1) ? p.partial[s] + (int64_t)((warp * NG + grp) & 63) * 3 * S : p.table[s] + ((int64_t)tok - id0) * p.P * S,
                                   x, acc, bufA, bufB, s, step, l, gap_own, pol_stream);
            }
        }
        __syncwarp();
        ir = ir_next; ir_next = ir_nn; rq0 = rq0_next; rq1 = rq1_next;
        it += stride;
        g_load_row<G, NV>(x, p.snap[s] + (int64_t)ir.y * S, l, S4, (ir.w & 0xff) != 0);
        load_opp(bufA, rq0.x, (ir.w & 0xff) != 0);
    }
}

#ifndef K5_MINB
#define K5_MINB 4
#endif
template <int G, int NV, int HEAD, bool DP>
__global__ void __launch_bounds__(128, K5_MINB) update_kernel5(const StepParams p) {
    constexpr int NG = 32 / G;
    __shared__ double w_acc[4][3 * NG];
    int k, step;
    if (!batch_index(p, k, step)) return;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    double *const wa = w_acc[wib];
    if (lane < 3 * NG) wa[lane] = 0.0;
    __syncwarp();
    k5_side<G, NV, HEAD, DP, 0>(p, k, step, lane, warp, nwarps, wa);
    __syncwarp();
    k5_side<G, NV, HEAD, DP, 1>(p, k, step, lane, warp, nwarps, wa);
    __syncwarp();
    // ---- end of step: per-group loss terms -> last CTA (ticket) adds them in (warp, group) order and finishes the step
    __shared__ double sh_red[3][128];
    __shared__ int is_last;
    if (lane < NG) {
        double *o = p.warp_out + 3 * (warp * NG + lane);
        o[0] = wa[3 * lane]; o[1] = wa[3 * lane + 1]; o[2] = wa[3 * lane + 2];
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&p.sc->ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (is_last) {
        __threadfence();
        const int tid = threadIdx.x;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (int w = tid; w < nwarps * NG; w += 128) {
            a0 += __ldcg(p.warp_out + 3 * w); a1 += __ldcg(p.warp_out + 3 * w + 1); a2 += __ldcg(p.warp_out + 3 * w + 2);
        }
        sh_red[0][tid] = a0; sh_red[1][tid] = a1; sh_red[2][tid] = a2;
        __syncthreads();
        for (int o = 64; o > 0; o >>= 1) {
            if (tid < o) { sh_red[0][tid] += sh_red[0][tid + o]; sh_red[1][tid] += sh_red[1][tid + o]; sh_red[2][tid] += sh_red[2][tid + o]; }
            __syncthreads();
        }
        if (tid == 0) finish_step(p, step, nullptr, sh_red[0][0], sh_red[1][0], sh_red[2][0]);
    }
}

}  // namespace glove
