// TRAIN step: model_fn(mode=TRAIN) of the reference [ref src/models/estimator.py:13-56] as two sm_100a launches.
//
//   stage_kernel   one thread per float4 of every distinct row / col id of the batch: (replay missed idle Adam steps,)
//                  publish the pre-step row into a compact snapshot snap[side][slot][S].  A snapshot row is
//                  [x_0..x_{d-1} | bias at column d+side | 1.0 at column d+1-side | 0..], so that
//                  dot(snap_row[i], snap_col[j]) over all S columns = sum_k R_ik C_jk + rb_i + cb_j with no masking,
//                  and sum_b e_b * snap_opposite lands the bias gradient sum_b e_b in the own bias column for free.
//                  stage_kernel<true> + commit_ls_kernel are the same replay run AHEAD of time on a side stream for the
//                  rows of the next batch that the step in flight cannot touch (glove_catchup_step).
//   update_kernel  one warp per work item (<= kItemMax triples of one id, both sides in one launch): gather the
//                  opposite rows from the snapshot with 128-bit loads (software-pipelined one triple ahead),
//                  warp-shuffle dot product, residual, loss, gradient accumulation in registers, and -- when the item
//                  is the whole segment -- the fused sparse optimizer update written in place to the packed table.
//                  All forward reads come from the snapshot, so the row side and the col side never race
//                  (SURVEY §7 hard part 2) and both sides run in one launch.  Segments longer than kItemMax are split
//                  into pieces whose partial sums are combined in a fixed order inside the same launch (two-level tree,
//                  the last warp to finish a chunk / a segment does the adding); the last CTA to finish (ticket) sums
//                  the per-warp loss terms in warp order, updates the scalar global bias, publishes the loss and
//                  increments the device step counter.
//
// Row-sharded tables split the step into glove_shard_stage / (pack, unpack | pull) / update / finish entry points around
// the exchange of snapshot rows; replicated data parallelism into glove_grad_step / glove_apply_step (apply_kernel).
//
// Everything is deterministic: no floating-point atomics, fixed summation order given (B, kItemMax, grid size).
#include <stdlib.h>

#include <new>

#include "glove_common.cuh"

namespace glove {

enum { MODE_TRAIN = 0, MODE_GRAD = 1, MODE_APPLY = 2, MODE_SHARD = 3 };  // SHARD: owner-computes update of own segments

struct StepParams {
    float *table[2];
    glove_scalars *sc;
    const PlanHeader *hdr;
    PlanSide side[2];
    float *snap[2];       // [B][S] pre-step snapshot rows
    float *partial[2];    // [max parts][S] partial gradient sums of split segments
    int32_t *long_cnt[2]; // [max long segments] chunks finished so far (self-resetting; workspace starts zeroed)
    int32_t *chunk_cnt[2];  // [max parts] pieces finished in the chunk that starts at this partial slot (self-resetting)
    unsigned long long *loss_acc;   // [6] fixed-point sums of the step's loss terms (self-resetting)
    int32_t *item_ctr;    // [2] next work item of each side handed out by the update kernel (self-resetting)
    int32_t *gap[2];      // [snapshot rows] closed-form replay: idle steps the staged row owed (0 = moments are current)
    float l2b1, l2b2;     // log2(beta1), log2(beta2) (host double, rounded once)
    float invB, ce, cbias, reg_unscale;   // 1/B; 2 s l2/(d B); 2 s l2/B; B/(2 s): host-computed with the same fp32 operations
    int32_t l2_hints;     // 1: table traffic evict_first, snapshot traffic evict_last (GLOVE_L2_HINTS=0 disables)
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
    int32_t n_shards, shard, v_loc;   // row-sharded tables (n_shards = 1: everything is owned by shard 0)
    int32_t run_stage, run_update;
    // peer gather (row-sharded tables, NVLink): peer_snap[side * kMaxShards + owner] = snapshot base of side `side` in the
    // step workspace of rank `owner` (peer-mapped device memory); nullptr = opposite rows come from the local snapshot
    const float *const *peer_snap;
    // device-side synchronisation of the row-sharded step over peer memory (peer_gather == 3): see sync_kernel
    int32_t *sync;               // local sync area: [0..7] stage-done epoch of rank r, [8..15] update-done epoch, [16] epoch
    float *peer_scal;            // local [kMaxShards][4]: loss scalars of rank r for the current step
    char *const *peer_base;      // [kMaxShards] workspace bases of the peers
    int64_t sync_off, scal_off;  // byte offsets of the two areas inside a workspace (same layout on every rank)
    int32_t dev_sync;
    float *const *push_snap;     // peer_gather == 4: [2][kMaxShards] snapshot bases of the peers the stage pushes rows to
    int32_t fused_sync;          // peer_gather == 4: the stage kernel announces its block itself, the update kernel waits for the owners
    int32_t *stage_ticket;       // CTAs of the stage kernel that have finished (self-resetting)
};

struct StepWs {
    float *snap[2];
    float *partial[2];
    int32_t *long_cnt[2];
    int32_t *chunk_cnt[2];
    const float **peer_tab;   // [2][kMaxShards] snapshot bases of the peers (glove_shard_set_peers)
    int32_t *gap[2];
    unsigned long long *loss_acc;
    int32_t *item_ctr;
    int32_t *sync;            // [32]
    float *peer_scal;         // [kMaxShards][4]
    char **peer_base;         // [kMaxShards]
    float *own_scal;          // [4] this rank's loss sums of the current step (one-call sharded step)
    size_t bytes;
};
__host__ __device__ static inline int64_t snapshot_rows(int32_t B) { return (int64_t)B + B / 4 + 64; }  // room for padded shard blocks
static inline int64_t max_parts_per_batch(int32_t B) { return 2 * (int64_t)B / kItemMax + 2; }
static StepWs step_ws_view(void *base, int32_t B, int32_t d) {
    StepWs w;
    char *p = (char *)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char *r = p ? p + off : nullptr; off += align_up(bytes); return r; };
    const int32_t S = table_stride(d);
    for (int s = 0; s < 2; ++s) w.snap[s] = (float *)take(sizeof(float) * (size_t)snapshot_rows(B) * S);
    for (int s = 0; s < 2; ++s) w.partial[s] = (float *)take(sizeof(float) * (size_t)max_parts_per_batch(B) * S);
    for (int s = 0; s < 2; ++s) w.long_cnt[s] = (int32_t *)take(sizeof(int32_t) * (size_t)(B / kItemMax + 2));
    for (int s = 0; s < 2; ++s) w.chunk_cnt[s] = (int32_t *)take(sizeof(int32_t) * (size_t)max_parts_per_batch(B));
    w.peer_tab = (const float **)take(sizeof(float *) * 2 * kMaxShards);
    for (int s = 0; s < 2; ++s) w.gap[s] = (int32_t *)take(sizeof(int32_t) * (size_t)snapshot_rows(B));
    w.loss_acc = (unsigned long long *)take(sizeof(unsigned long long) * 6);
    w.item_ctr = (int32_t *)take(sizeof(int32_t) * 4);   // [2] = ticket of the stage kernel (fused announcement)
    w.sync = (int32_t *)take(sizeof(int32_t) * 32);
    w.peer_scal = (float *)take(sizeof(float) * 4 * kMaxShards);
    w.peer_base = (char **)take(sizeof(char *) * kMaxShards);
    w.own_scal = (float *)take(sizeof(float) * 4);
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

template <int NV>
__device__ __forceinline__ void load_row(float4 (&x)[NV], const float *row, int lane, int S4) {
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        const int f = lane + 32 * r;
        x[r] = (r < NV - 1 || f < S4) ? ld4(row + 4 * f) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
template <int NV>
__device__ __forceinline__ void load_row_hint(float4 (&x)[NV], const float *row, int lane, int S4, uint64_t pol) {
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        const int f = lane + 32 * r;
        x[r] = (r < NV - 1 || f < S4) ? ld4_hint(row + 4 * f, pol) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
template <int NV>
__device__ __forceinline__ void load_row_nc(float4 (&x)[NV], const float *row, int lane, int S4) {
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        const int f = lane + 32 * r;
        x[r] = (r < NV - 1 || f < S4) ? ld4_nc(row + 4 * f) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
template <int NV>
__device__ __forceinline__ void store_row(float *row, const float4 (&x)[NV], int lane, int S4, uint64_t pol = 0) {
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        const int f = lane + 32 * r;
        if (r < NV - 1 || f < S4) {
            if (pol) st4_hint(row + 4 * f, x[r], pol); else st4(row + 4 * f, x[r]);
        }
    }
}

__device__ __forceinline__ bool batch_index(const StepParams &p, int &k, int &step) {
    step = p.sc->step;
    k = step - p.hdr->first_step;
    if (p.hdr->magic != kPlanMagic || k < 0 || k >= p.hdr->K || p.hdr->B != p.B ||
        (p.opt == GLOVE_OPT_ADAM && step >= p.alpha_len)) {
        if (threadIdx.x == 0 && blockIdx.x == 0) p.sc->error = 1;
        return false;
    }
    // row-sharded tables: the padded owner blocks of this batch must fit the snapshot (a batch whose ids pile up on one owner
    // does not); checked here, by every kernel of the step, so that the host need not read the block sizes back per chunk
    if (p.n_shards > 1 && (int64_t)p.n_shards * max(p.side[0].b_upad[k], p.side[1].b_upad[k]) > snapshot_rows(p.B)) {
        if (threadIdx.x == 0 && blockIdx.x == 0) p.sc->error = 3;
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
        const float br = b * r;
        l = br * r;
        e = (2.0f * invB) * br;
    } else {  // a = pos weight (value), b = neg weight
        const float sg = sigmoid_f(z);
        l = a * softplus_f(-z) + nf * (b * softplus_f(z));
        e = (a * (sg - 1.0f) + nf * b * sg) * invB;
    }
}

// device-side synchronisation over peer memory (see sync_staged_kernel / finish_sync_kernel below)
enum { SYNC_STAGED = 0, SYNC_UPDATED = 8, SYNC_EPOCH = 16 };
__device__ __forceinline__ int ld_acquire_sys(const int32_t *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int32_t *p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int32_t *peer_sync(const StepParams &p, int q) {
    return reinterpret_cast<int32_t *>(p.peer_base[q] + p.sync_off);
}

// ---- K1: stage -----------------------------------------------------------------------------------------------------
// Flat mapping: one thread per (slot, float4) of the snapshot, a warp per 32 consecutive float4s.  Rows that need no
// replay are a pure 128-bit copy; rows that do are replayed 4 columns per thread, so a long gap is spread over the
// ~S/4 threads of the row instead of serialising on one warp, and the many small chunks balance across the SMs.
//
// CATCHUP = true is the same arithmetic run AHEAD of time for the batch of step `step_arg` while the previous step is
// still executing on another stream: rows of that batch which are NOT in the previous batch (seg_prev == 0) cannot be
// touched by the step in flight, so their idle steps through step_arg-1 are replayed now and written back in place
// (x, m, v, last_step = step_arg).  The regular stage of step_arg then finds them current and degenerates to a copy;
// rows it could not pre-replay (seg_prev == 1, or no catch-up launched) are replayed there as before, so correctness
// never depends on the catch-up having run.
template <bool CATCHUP>
__global__ void __launch_bounds__(256) stage_kernel(const StepParams p, int step_arg) {
    int k, step;
    if (CATCHUP) {
        step = step_arg;
        k = step - p.hdr->first_step;
        if (p.hdr->magic != kPlanMagic || k < 1 || k >= p.hdr->K || p.hdr->B != p.B || step > p.alpha_len) return;
    } else if (!batch_index(p, k, step)) return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int seg0[2] = {p.side[0].b_seg[k], p.side[1].b_seg[k]};
    // this shard's block of slots on each side (everything when the tables are not sharded)
    const int own0[2] = {p.side[0].b_own[k * (kMaxShards + 1) + p.shard], p.side[1].b_own[k * (kMaxShards + 1) + p.shard]};
    const int U0 = p.side[0].b_own[k * (kMaxShards + 1) + p.shard + 1] - own0[0];
    const int U1 = p.side[1].b_own[k * (kMaxShards + 1) + p.shard + 1] - own0[1];
    const int pos0[2] = {p.shard * p.side[0].b_upad[k], p.shard * p.side[1].b_upad[k]};  // first snapshot row of the block
    const int64_t id0 = (int64_t)p.shard * p.v_loc;                                       // first (remapped) id owned
    const int S4 = p.S >> 2;
    const bool replay = (p.opt == GLOVE_OPT_ADAM && p.adam_mode == GLOVE_ADAM_REPLAY_EXACT);
    const int total = (U0 + U1) * S4;

    // ---- pass 1 (stage only): rows that are current are a pure 128-bit copy; 4 chunks per warp in flight
    if (!CATCHUP) {
        constexpr int UN = 4;
        for (int base0 = warp * 32 * UN; base0 < total; base0 += nwarps * 32 * UN) {
            float4 x[UN];
            int ls[UN], dsti[UN], fsel[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int idx = base0 + u * 32 + lane;
                dsti[u] = -1;
                if (idx < total) {
                    const int w = idx / S4, f = idx - w * S4;
                    const int s = w >= U0 ? 1 : 0;
                    const int j = s ? w - U0 : w;
                    const float *row = p.table[s] + ((int64_t)p.side[s].seg_id[seg0[s] + own0[s] + j] - id0) * p.P * p.S;
                    const int lcol = ls_col(p.d, s);
                    x[u] = ld4(row + 4 * f);
                    ls[u] = __float_as_int(row[lcol]);
                    dsti[u] = (s << 30) | ((pos0[s] + j) * S4 + f);  // side in bit 30, float4 index into snap[side]
                    fsel[u] = (lcol >> 2) == f ? (lcol & 3) : -1;
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                if (dsti[u] < 0) continue;
                if (replay && ls[u] > 0 && ls[u] < step) continue;  // needs replay: pass 2
                if (fsel[u] >= 0) f4c(x[u], fsel[u]) = 1.0f;        // 1.0 in the other side's bias column
                st4(p.snap[dsti[u] >> 30] + (int64_t)(dsti[u] & 0x3fffffff) * 4, x[u]);
            }
        }
        if (!replay) return;
    }

    // ---- pass 2: rows with missed idle Adam steps
    for (int base = warp * 32; base < total; base += nwarps * 32) {
        const int idx = base + lane;
        const bool active = idx < total;
        const int w = active ? idx / S4 : 0, f = active ? idx - w * S4 : 0;
        const int s = w >= U0 ? 1 : 0;
        const int j = s ? w - U0 : w;
        const int g = seg0[s] + own0[s] + j;
        const float *row = p.table[s] + ((int64_t)(active ? p.side[s].seg_id[g] : id0) - id0) * p.P * p.S;
        const int lcol = ls_col(p.d, s);
        int ls = active ? __float_as_int(row[lcol]) : 0;
        if (CATCHUP && active && p.side[s].seg_prev[g]) ls = 0;   // may be in flight: leave it to the stage
        const bool need = active && ls > 0 && ls < step;
        if (!__any_sync(0xffffffffu, need)) continue;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f), m = x, v = x;
        if (need) { x = ld4(row + 4 * f); m = ld4(row + p.S + 4 * f); v = ld4(row + 2 * p.S + 4 * f); }
        float2 xa = make_float2(x.x, x.y), xb = make_float2(x.z, x.w);
        float2 ma = make_float2(m.x, m.y), mb = make_float2(m.z, m.w);
        float2 va = make_float2(v.x, v.y), vb = make_float2(v.z, v.w);
        const int gap = need ? step - ls : 0;
        const int trips = __reduce_max_sync(0xffffffffu, gap);
        const float *al = p.alpha + (need ? ls : 0);
        bool moving = need;
        // |increment| shrinks monotonically (x0.9 per step from m, at most x1.012 from alpha and sqrt(v)): once an element
        // has not moved for a whole block of 4 steps it never moves again; from then on only the m, v decays remain
        // (2 packed multiplies per pair).  The test is made once per block of 4 steps.
#pragma unroll 1
        for (int i = 0; i < trips; i += 4) {
            if (moving) {
                const float2 oa = xa, ob = xb;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (i + j < gap) {
                        const float na = -__ldg(al + i + j);
                        adam_idle_step2(xa, ma, va, na, p.b1, p.b2, p.eps);
                        adam_idle_step2(xb, mb, vb, na, p.b1, p.b2, p.eps);
                    }
                }
                moving = (xa.x != oa.x) | (xa.y != oa.y) | (xb.x != ob.x) | (xb.y != ob.y);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (i + j < gap) {
                        ma = __fmul2_rn(ma, f2(p.b1)); mb = __fmul2_rn(mb, f2(p.b1));
                        va = __fmul2_rn(va, f2(p.b2)); vb = __fmul2_rn(vb, f2(p.b2));
                    }
                }
            }
        }
        if (need) {
            // the row is current through step-1 now: publish the decayed moments so that the update kernel reads them
            // as-is (the x plane of the table is rewritten by the update in this same step)
            x = make_float4(xa.x, xa.y, xb.x, xb.y);
            st4(const_cast<float *>(row) + p.S + 4 * f, make_float4(ma.x, ma.y, mb.x, mb.y));
            st4(const_cast<float *>(row) + 2 * p.S + 4 * f, make_float4(va.x, va.y, vb.x, vb.y));
            if (CATCHUP) {
                // catch-up: x goes back in place too; last_step is NOT touched here (other threads of the row may still
                // have to read it) -- commit_ls_kernel sets it once this kernel has finished
                st4(const_cast<float *>(row) + 4 * f, x);
            } else {
                if ((lcol >> 2) == f) f4c(x, lcol & 3) = 1.0f;
                st4(p.snap[s] + (int64_t)(pos0[s] + j) * p.S + 4 * f, x);
            }
        }
    }
}

// ---- K1 (GLOVE_ADAM_REPLAY, the default): stage with the closed-form replay -------------------------------------------
// One warp per distinct id of the batch, software-pipelined one row ahead.  A row that owes idle Adam steps
// (0 < last_step < step) needs its moments as well as its plane-0 row: whether it does is known only once the row has
// arrived, and a dependent second round trip to DRAM per row is what bounded the first version of this kernel.  The plan
// knows better: a row that is NOT in the previous batch (seg_prev == 0) has been idle for at least one step, so its
// three planes are fetched together, one row ahead (rows that were never updated pay two wasted plane reads once).  The
// run of idle steps is applied in closed form (replay_x4: one sqrt, one reciprocal and a cubic per element, independent
// of the gap) and the row is published to the snapshot.  NOTHING is written back to the table: the update kernel of the
// same step re-reads the moments anyway, scales them by b1^gap / b2^gap (gap is left in p.gap[side][position]) and
// writes x, m, v once.  Bound: HBM.
template <int NV>
__global__ void __launch_bounds__(256, 2) stage_closed_kernel(const StepParams p) {
    __shared__ ReplayTables tabs;
    extern __shared__ __align__(16) float push_smem[];   // peer-push only: [8 warps][2 buffers][S] rows on their way to the peers
    uint32_t n_pushed = 0;                               // rows this warp has handed to the copy engine
    int k, step;
    const bool ok = batch_index(p, k, step);
    replay_tables_init(tabs, p.b1, p.b2);
    if (!ok && !p.fused_sync) return;
    if (!ok) k = 0;     // fused announcement: a refused step stages nothing (total = 0 below) but still announces, or the peers would wait for ever
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int S4 = p.S >> 2;
    const int own_r = p.side[0].b_own[k * (kMaxShards + 1) + p.shard], own_c = p.side[1].b_own[k * (kMaxShards + 1) + p.shard];
    const int U0 = p.side[0].b_own[k * (kMaxShards + 1) + p.shard + 1] - own_r;
    const int U1 = p.side[1].b_own[k * (kMaxShards + 1) + p.shard + 1] - own_c;
    const int g_r = p.side[0].b_seg[k] + own_r, g_c = p.side[1].b_seg[k] + own_c;         // first owned segment of each side
    const int pos_r = p.shard * p.side[0].b_upad[k], pos_c = p.shard * p.side[1].b_upad[k]; // first snapshot row of the block
    const int64_t id0 = (int64_t)p.shard * p.v_loc;
    const int total = ok ? U0 + U1 : 0;
    const uint64_t pol_stream = p.l2_hints ? l2_policy_evict_first() : l2_policy_evict_normal();
    const uint64_t pol_keep = p.l2_hints ? l2_policy_evict_last() : l2_policy_evict_normal();
    float4 xn[NV], mn[NV], vn[NV];
    const float *row_n = nullptr;
    bool spec_n = false;
    // issue the loads of row w: plane 0 always, the moments when the plan says the row has been idle
    auto fetch = [&](int w) {
        const int sd = w >= U0 ? 1 : 0;
        const int g = sd ? g_c + (w - U0) : g_r + w;
        row_n = p.table[sd] + ((int64_t)__ldg(p.side[sd].seg_id + g) - id0) * p.P * p.S;
        spec_n = __ldg(p.side[sd].seg_prev + g) == 0;
        load_row_hint<NV>(xn, row_n, lane, S4, pol_stream);
        if (spec_n) { load_row_hint<NV>(mn, row_n + p.S, lane, S4, pol_stream); load_row_hint<NV>(vn, row_n + 2 * p.S, lane, S4, pol_stream); }
    };
    if (warp < total) fetch(warp);
#pragma unroll 1
    for (int w = warp; w < total; w += nwarps) {
        const int sd = w >= U0 ? 1 : 0;
        const int j = sd ? w - U0 : w;
        const float *row = row_n;
        const bool spec = spec_n;
        float4 x[NV], m[NV], v[NV];
#pragma unroll
        for (int r = 0; r < NV; ++r) { x[r] = xn[r]; m[r] = mn[r]; v[r] = vn[r]; }
        if (w + nwarps < total) fetch(w + nwarps);
        const int lcol = ls_col(p.d, sd);
        const int ls = __float_as_int(row_col<NV>(x, lcol, lane));
        const int gap = (ls > 0 && ls < step) ? step - ls : 0;
        if (gap) {
            if (!spec) { load_row_hint<NV>(m, row + p.S, lane, S4, pol_stream); load_row_hint<NV>(v, row + 2 * p.S, lane, S4, pol_stream); }
            const ReplayCoef c = replay_coef(tabs, p.alpha, ls, gap, p.l2b1, p.l2b2, lane);
#pragma unroll
            for (int r = 0; r < NV; ++r) x[r] = replay_x4(x[r], m[r], v[r], c, p.eps);
        }
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int f = lane + 32 * r;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (4 * f + c == lcol) f4c(x[r], c) = 1.0f;     // 1.0 in the other side's bias column
        }
        const int pos = (sd ? pos_c : pos_r) + j;
        store_row<NV>(p.snap[sd] + (int64_t)pos * p.S, x, lane, S4, pol_keep);
        if (lane == 0) p.gap[sd][pos] = gap;
        if (p.push_snap) {
            // fused exchange: the row goes straight from these registers into the snapshot of every shard whose work items
            // read it (request mask from the plan), as posted NVLink writes that overlap the staging of the next rows
            //
            // The row is parked in a per-warp shared-memory buffer and ONE bulk asynchronous copy per destination
            // (cp.async.bulk shared -> global: 1,216 bytes each) does the rest: the warp does not wait on NVLink, and the
            // link sees whole rows instead of 16-byte stores.  Two buffers per warp: a buffer is rewritten only after the
            // copies of the row before last have read it (wait_group.read 1).
            unsigned mask = (unsigned)__ldg(p.side[sd].seg_push + (sd ? g_c + (w - U0) : g_r + w));
            if (mask) {
                float *buf = push_smem + ((threadIdx.x >> 5) * 2 + (n_pushed & 1u)) * p.S;
                if (lane == 0) bulk_wait_read_1();
                __syncwarp();
                store_row<NV>(buf, x, lane, S4);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    const uint32_t src = smem_addr(buf), bytes = (uint32_t)p.S * 4u;
                    while (mask) {
                        const int r = __ffs(mask) - 1;
                        mask &= mask - 1;
                        bulk_s2g(p.push_snap[sd * kMaxShards + r] + (int64_t)pos * p.S, src, bytes);
                    }
                    bulk_commit();
                }
                ++n_pushed;
            }
        }
    }
    if (p.push_snap && lane == 0) bulk_wait_all();   // every pushed row has left before the kernel (and its announcement) ends
    if (p.fused_sync) {
        // fused announcement (instead of the sync_staged_kernel launch): the last CTA to finish tells every peer that this
        // rank's block -- and every row it pushed -- of the current step is complete
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (atomicAdd(p.stage_ticket, 1) == (int)gridDim.x - 1) {
                *p.stage_ticket = 0;
                __threadfence_system();
                const int epoch = p.sync[SYNC_EPOCH];
                for (int q = 0; q < p.n_shards; ++q) st_release_sys(peer_sync(p, q) + SYNC_STAGED + p.shard, epoch + 1);
            }
        }
    }
}

// second half of the catch-up: mark the pre-replayed rows current (same predicate as stage_kernel<true>)
__global__ void __launch_bounds__(256) commit_ls_kernel(const StepParams p, int step) {
    const int k = step - p.hdr->first_step;
    if (p.hdr->magic != kPlanMagic || k < 1 || k >= p.hdr->K || p.hdr->B != p.B || step > p.alpha_len) return;
    const int seg0[2] = {p.side[0].b_seg[k], p.side[1].b_seg[k]};
    const int own0[2] = {p.side[0].b_own[k * (kMaxShards + 1) + p.shard], p.side[1].b_own[k * (kMaxShards + 1) + p.shard]};
    const int U0 = p.side[0].b_own[k * (kMaxShards + 1) + p.shard + 1] - own0[0];
    const int U1 = p.side[1].b_own[k * (kMaxShards + 1) + p.shard + 1] - own0[1];
    const int64_t id0 = (int64_t)p.shard * p.v_loc;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < U0 + U1; w += gridDim.x * blockDim.x) {
        const int s = w >= U0 ? 1 : 0;
        const int g = seg0[s] + own0[s] + (s ? w - U0 : w);
        if (p.side[s].seg_prev[g]) continue;
        float *ls_word = p.table[s] + ((int64_t)p.side[s].seg_id[g] - id0) * p.P * p.S + ls_col(p.d, s);
        const int ls = __float_as_int(*ls_word);
        if (ls > 0 && ls < step) *ls_word = __int_as_float(step);
    }
}

// ---- optimizer epilogue on a row held in registers ---------------------------------------------------------------
// x = pre-step snapshot row (1.0 in the last_step column), G = de-duplicated gradient (0 outside columns 0..d-1 and
// the bias column).  Padding columns carry x = G = m = v = 0 and stay 0 through every formula below.  In replay mode the
// stage kernel has already decayed m, v through step-1, so the moments are read as-is.
template <int NV>
__device__ __forceinline__ void apply_row(const StepParams &p, float *row, float4 (&x)[NV], const float4 (&G)[NV],
                                          float4 (&s1)[NV], float4 (&s2)[NV], int s, int step, int lane, int gap,
                                          uint64_t store_pol = 0) {
    // s1 / s2 = optimizer slot planes 1 / 2 of the row, already loaded by the caller (prefetched at item start)
    // gap = idle steps the row owed when it was staged (closed-form replay: the moments in the table are `gap` steps old)
    const int S4 = p.S >> 2;
    if (p.opt == GLOVE_OPT_ADAM) {
        const float na = -__ldg(p.alpha + step);
        if (gap > 0) {
            float dm, dv;
            replay_decay(gap, p.l2b1, p.l2b2, dm, dv);
#pragma unroll
            for (int r = 0; r < NV; ++r) { s1[r] = scale4(s1[r], dm); s2[r] = scale4(s2[r], dv); }
        }
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            float2 xa = make_float2(x[r].x, x[r].y), xb = make_float2(x[r].z, x[r].w);
            float2 ma = make_float2(s1[r].x, s1[r].y), mb = make_float2(s1[r].z, s1[r].w);
            float2 va = make_float2(s2[r].x, s2[r].y), vb = make_float2(s2[r].z, s2[r].w);
            adam_update2(xa, ma, va, make_float2(G[r].x, G[r].y), na, p.b1, p.b2, p.eps);
            adam_update2(xb, mb, vb, make_float2(G[r].z, G[r].w), na, p.b1, p.b2, p.eps);
            x[r] = make_float4(xa.x, xa.y, xb.x, xb.y);
            s1[r] = make_float4(ma.x, ma.y, mb.x, mb.y);
            s2[r] = make_float4(va.x, va.y, vb.x, vb.y);
        }
        store_row<NV>(row + p.S, s1, lane, S4, store_pol);
        store_row<NV>(row + 2 * p.S, s2, lane, S4, store_pol);
    } else if (p.opt == GLOVE_OPT_ADAGRAD) {
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            float2 xa = make_float2(x[r].x, x[r].y), xb = make_float2(x[r].z, x[r].w);
            float2 aa = make_float2(s1[r].x, s1[r].y), ab = make_float2(s1[r].z, s1[r].w);
            adagrad_update2(xa, aa, make_float2(G[r].x, G[r].y), -p.lr, p.eps);
            adagrad_update2(xb, ab, make_float2(G[r].z, G[r].w), -p.lr, p.eps);
            x[r] = make_float4(xa.x, xa.y, xb.x, xb.y);
            s1[r] = make_float4(aa.x, aa.y, ab.x, ab.y);
        }
        store_row<NV>(row + p.S, s1, lane, S4, store_pol);
    } else {
#pragma unroll
        for (int r = 0; r < NV; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) sgd_update(f4c(x[r], c), f4v(G[r], c), p.lr);
    }
    const int lcol = ls_col(p.d, s);
    const int lc = lcol & 3;
    const float lsv = __int_as_float(step + 1);
#pragma unroll
    for (int r = 0; r < NV; ++r)
        if (lane + 32 * r == (lcol >> 2)) {
            x[r].x = lc == 0 ? lsv : x[r].x; x[r].y = lc == 1 ? lsv : x[r].y;
            x[r].z = lc == 2 ? lsv : x[r].z; x[r].w = lc == 3 ? lsv : x[r].w;
        }
    store_row<NV>(row, x, lane, S4, store_pol);
}

// ---- end of step: loss, global bias, step counter --------------------------------------------------------------------------------------
__device__ void finish_step(const StepParams &p, int step, const float *reduced /* MODE_APPLY */, double ld = 0.0,
                            double se = 0.0, double rg = 0.0) {
    if (reduced) { ld = reduced[0]; se = reduced[1]; rg = reduced[2]; }
    if (p.mode == MODE_GRAD || p.mode == MODE_SHARD) {
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
        g = __fsub_rn(g, div_pos(__fmul_rn(a, m), __fadd_rn(sqrt_pos(v), p.eps)));
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

// ---- K2: update ----------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ float dot_row(const float4 (&x)[NV], const float4 (&y)[NV]) {
    float2 a = make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        a = __ffma2_rn(make_float2(x[r].x, x[r].y), make_float2(y[r].x, y[r].y), a);
        a = __ffma2_rn(make_float2(x[r].z, x[r].w), make_float2(y[r].z, y[r].w), a);
    }
    return warp_sum(a.x + a.y);
}
template <int NV>
__device__ __forceinline__ void axpy_row(float4 (&acc)[NV], float e, const float4 (&y)[NV]) {
    const float2 ee = make_float2(e, e);
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        const float2 lo = __ffma2_rn(ee, make_float2(y[r].x, y[r].y), make_float2(acc[r].x, acc[r].y));
        const float2 hi = __ffma2_rn(ee, make_float2(y[r].z, y[r].w), make_float2(acc[r].z, acc[r].w));
        acc[r] = make_float4(lo.x, lo.y, hi.x, hi.y);
    }
}

// activity-L2 of one row: acc_c <- [coef_c != 0] acc_c + n coef_c x_c and the lane's share of sum_c coef_c x_c^2,
// coef = ce on the embedding columns, cbias on the bias column, 0 elsewhere (there acc holds the opposite bias that was
// multiplied in, and is cleared).  A float4 chunk that lies wholly inside the embedding columns -- all but one chunk of
// the row -- takes the packed path; only the chunk that straddles column d does the per-column select.
template <int NV>
__device__ __forceinline__ float activity_l2(float4 (&acc)[NV], const float4 (&x)[NV], float fn, float ce, float cbias, int d,
                                             int bcol, int lane) {
    float2 sq2 = make_float2(0.f, 0.f);
    float sq = 0.0f;
    const float2 ce2 = make_float2(ce, ce), fn2 = make_float2(fn, fn);
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        const int col0 = 4 * (lane + 32 * r);
        if (col0 + 3 < d) {
            const float2 xl = make_float2(x[r].x, x[r].y), xh = make_float2(x[r].z, x[r].w);
            const float2 cl = __fmul2_rn(ce2, xl), ch = __fmul2_rn(ce2, xh);
            const float2 al = __ffma2_rn(fn2, cl, make_float2(acc[r].x, acc[r].y));
            const float2 ah = __ffma2_rn(fn2, ch, make_float2(acc[r].z, acc[r].w));
            acc[r] = make_float4(al.x, al.y, ah.x, ah.y);
            sq2 = __ffma2_rn(cl, xl, sq2);
            sq2 = __ffma2_rn(ch, xh, sq2);
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int col = col0 + c;
                const float cf = col < d ? ce : (col == bcol ? cbias : 0.0f);
                const float xv = f4v(x[r], c);
                const float cx = cf * xv;
                f4c(acc[r], c) = (cf != 0.0f ? f4c(acc[r], c) : 0.0f) + fn * cx;
                sq += cx * xv;
            }
        }
    }
    return sq + (sq2.x + sq2.y);
}

// Order-independent accumulation of float terms: a term is split exactly into its integer part and a 2^-40 fixed-point
// fraction (|fraction| < 1 -> < 2^40 per term, so 2^22 terms fit an int64; the integer parts fit as long as the true sum
// is below 2^63); both limbs are summed with integer additions, which commute and associate exactly.  Resolution 9e-13.
__device__ __forceinline__ void fx_add(long long *acc, float x, long long *bad) {
    if (!(fabsf(x) < 4.0e18f)) *bad = 1;                          // NaN / inf / out of range: reported, never silently dropped
    const long long hi = __float2ll_rz(x);
    const float frac = __fsub_rn(x, __ll2float_rn(hi));          // exact: |x| < 2^24 has an exact difference, larger x no fraction
    acc[0] += hi;
    acc[1] += __float2ll_rn(__fmul_rn(frac, 1099511627776.0f));   // 2^40
}
__device__ __forceinline__ double fx_value(long long hi, long long lo) { return (double)hi + (double)lo * (1.0 / 1099511627776.0); }

constexpr int kChunk = 16;  // pieces per first-level combine of a split segment

// acc = sum_{i < n} rows[i * stride_rows] in index order; 3 rows in flight (the caller lends 3 dead row buffers).
// L2-coherent loads (the rows were written by other SMs in this launch).
template <int NV>
__device__ __forceinline__ void sum_partials(float4 (&acc)[NV], const float *base, int n, int stride_rows, int S,
                                             float4 (&b0)[NV], float4 (&b1)[NV], float4 (&b2)[NV], int lane, int S4) {
#pragma unroll
    for (int r = 0; r < NV; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int q = 0; q < n; q += 3) {
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int f = lane + 32 * r;
            const bool in = r < NV - 1 || f < S4;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float *p0 = base + (int64_t)q * stride_rows * S + 4 * f;
            b0[r] = in ? __ldcg(reinterpret_cast<const float4 *>(p0)) : z;
            b1[r] = (in && q + 1 < n) ? __ldcg(reinterpret_cast<const float4 *>(p0 + (int64_t)stride_rows * S)) : z;
            b2[r] = (in && q + 2 < n) ? __ldcg(reinterpret_cast<const float4 *>(p0 + 2 * (int64_t)stride_rows * S)) : z;
        }
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            acc[r].x += b0[r].x; acc[r].y += b0[r].y; acc[r].z += b0[r].z; acc[r].w += b0[r].w;
            acc[r].x += b1[r].x; acc[r].y += b1[r].y; acc[r].z += b1[r].z; acc[r].w += b1[r].w;
            acc[r].x += b2[r].x; acc[r].y += b2[r].y; acc[r].z += b2[r].z; acc[r].w += b2[r].w;
        }
    }
}

// One warp per work item.  ncu on the round-1 form of this kernel showed the warps (16 / SM at 128 registers) stalled on
// three chains of dependent loads per item (item record -> triple records -> gathered rows, and the optimizer planes,
// whose DRAM latency is longer than the ~1 us a typical 2.7-triple item lives) and the SMs idle for 18 % of the launch
// waiting for the slowest warps.  Hence
//   * software pipeline across items: the record of item i+2 and the triple records of item i+1 are fetched while item i
//     is gathered; the optimizer planes are requested at item start and consumed in the epilogue (L1 / L2 prefetches of
//     the following items' rows were measured to change nothing and are gone: profiles/r02_update_kernel.md);
//   * dynamic list scheduling: the work list is sorted by decreasing length at plan time (LPT); a warp takes its first
//     three items by position and every later one from a per-side counter (one atomic per item, issued three items
//     ahead), so all SMs finish together whatever the latency each of them sees;
//   * which warp processes which item is therefore timing-dependent, so the loss terms are accumulated in 128-bit
//     fixed point (two int64 limbs per term: integer part and 2^-40 fraction): integer addition is associative, the
//     sums do not depend on the order, and the step stays bit-reproducible.  (The tables never depended on it: an item's
//     arithmetic is self-contained and split segments are combined in piece order.)
template <int NV, int HEAD, bool DP>
__global__ void __launch_bounds__(128, (NV <= 3 ? 4 : 3)) update_kernel(const StepParams p) {
    int k, step;
    if (!batch_index(p, k, step)) return;
    if (p.fused_sync) {
        // fused wait (instead of the wait_staged_kernel launch): the rows this CTA gathers were pushed by their owners' stage
        // kernels; one thread polls the owners' announcements in this rank's own memory, the CTA waits at the barrier
        if (threadIdx.x == 0) {
            const int epoch = p.sync[SYNC_EPOCH];
            for (int q = 0; q < p.n_shards; ++q)
                while (ld_acquire_sys(p.sync + SYNC_STAGED + q) < epoch + 1) __nanosleep(64);
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int S4 = p.S >> 2;
    const float gbias = p.sc->g;
    const bool train = p.mode == MODE_TRAIN || p.mode == MODE_SHARD;
    const bool closed = p.opt == GLOVE_OPT_ADAM && p.adam_mode == GLOVE_ADAM_REPLAY;
    const int64_t id0 = (int64_t)p.shard * p.v_loc;
    const uint64_t pol_keep = p.l2_hints ? l2_policy_evict_last() : l2_policy_evict_normal();
    const uint64_t pol_stream = p.l2_hints ? l2_policy_evict_first() : l2_policy_evict_normal();
    // this warp's share of the step's loss terms as order-independent fixed-point sums (see fx_add), kept in shared
    // memory to save registers: {data loss, B * sum e, reg term} x {integer limb, 2^-40 fraction limb}
    __shared__ long long w_acc[4][7];    // [6]: a term was not finite
    long long *const wa = w_acc[threadIdx.x >> 5];
    if (lane < 7) wa[lane] = 0;
    __syncwarp();

#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
        const PlanSide &ps = p.side[s];
        const int oi0 = p.mode == MODE_SHARD ? ps.b_own_item[k * (kMaxShards + 1) + p.shard] : 0;
        const int4 *const irec = ps.item_rec + ps.b_item[k] + oi0;
        const int nI = (p.mode == MODE_SHARD ? ps.b_own_item[k * (kMaxShards + 1) + p.shard + 1] : ps.b_item[k + 1] - ps.b_item[k]) - oi0;
        const float *opp_base = p.snap[1 - s];
        const float *own_base = p.snap[s];
        const float *const *peer = p.peer_snap ? p.peer_snap + (1 - s) * kMaxShards : nullptr;
        const int upad_opp = peer ? max(p.side[1 - s].b_upad[k], 1) : 1;
        auto load_opp = [&](float4 (&buf)[NV], int pos) {
            if (peer) {
                const float *base = reinterpret_cast<const float *>(__ldg(reinterpret_cast<const unsigned long long *>(peer + pos / upad_opp))) + (int64_t)pos * p.S;
#pragma unroll
                for (int r = 0; r < NV; ++r) {
                    const int f = lane + 32 * r;
                    buf[r] = (r < NV - 1 || f < S4) ? __ldcg(reinterpret_cast<const float4 *>(base + 4 * f)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
                const float *base = opp_base + (int64_t)pos * p.S;
#pragma unroll
                for (int r = 0; r < NV; ++r) {
                    const int f = lane + 32 * r;
                    buf[r] = (r < NV - 1 || f < S4) ? ld4(base + 4 * f) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        auto load_own = [&](float4 (&buf)[NV], int slot) {
            const float *base = own_base + (int64_t)slot * p.S;
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                const int f = lane + 32 * r;
                buf[r] = (r < NV - 1 || f < S4) ? ld4(base + 4 * f) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        const int4 *rec = ps.rec;
        const int bcol = bias_col(p.d, s);
        int itl = warp;
        if (itl >= nI) continue;
        // a warp's first three items are dealt by position, the later ones come from the side's counter
        int it_n = warp + nwarps, it_nn = warp + 2 * nwarps;
        // pipeline state: ir / myrec / x / bufA (first gather) of the CURRENT item are in registers when an iteration starts
        int4 ir = __ldg(irec + itl);
        int4 ir_next = it_n < nI ? __ldg(irec + it_n) : ir;
        int4 myrec = lane < (ir.w & 0xff) ? __ldg(rec + ir.z + lane) : make_int4(0, 0, 0, 0);
        float4 x[NV], acc[NV], bufA[NV], bufB[NV], s1[NV], s2[NV];
        load_own(x, ir.y);
        load_opp(bufA, __shfl_sync(0xffffffffu, myrec.x, 0));
#pragma unroll 1
        while (itl < nI) {
            const int slot = ir.y, n = ir.w & 0xff, part = ir.w >> 8;
            float *row = p.table[s] + (part ? 0 : (int64_t)ir.x - id0) * p.P * p.S;
            const bool applies = train && part == 0;
            const int gap_own = (applies && closed) ? __ldcg(p.gap[s] + slot) : 0;
            if (applies && p.P >= 2) load_row<NV>(s1, row + p.S, lane, S4);
            if (applies && p.P >= 3) load_row<NV>(s2, row + 2 * p.S, lane, S4);
            const bool has_next = it_n < nI;
            const int itnn = it_nn;
            const int4 ir_nn = itnn < nI ? __ldg(irec + itnn) : ir_next;
            int it_nnn = nI;                   // the item after that: taken from the counter now, needed one item from now
            if (itnn < nI && lane == 0) it_nnn = 3 * nwarps + atomicAdd(p.item_ctr + s, 1);
            const int4 myrec_next = (has_next && lane < (ir_next.w & 0xff)) ? __ldg(rec + ir_next.z + lane) : make_int4(0, 0, 0, 0);
#pragma unroll
            for (int r = 0; r < NV; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            float loss_d = 0.0f, sum_e = 0.0f;
            int n_eff = 0;
#pragma unroll 1
            for (int q = 0; q < n; q += 2) {
                const bool hasB = q + 1 < n;
                if (hasB) load_opp(bufB, __shfl_sync(0xffffffffu, myrec.x, q + 1));
                {
                    const float a = __int_as_float(__shfl_sync(0xffffffffu, myrec.y, q));
                    const float b = __int_as_float(__shfl_sync(0xffffffffu, myrec.z, q));
                    float e, l;
                    head_eval(HEAD, dot_row<NV>(x, bufA) + gbias, a, b, p.invB, p.nf, e, l);
                    if (DP) {
                        const bool mine = __shfl_sync(0xffffffffu, myrec.w, q) / p.dp_block == p.dp_rank;
                        e = mine ? e : 0.f; l = mine ? l : 0.f; n_eff += mine;
                    }
                    loss_d += l; sum_e += e;
                    axpy_row<NV>(acc, e, bufA);
                }
                if (hasB) {
                    if (q + 2 < n) load_opp(bufA, __shfl_sync(0xffffffffu, myrec.x, q + 2));
                    const float a = __int_as_float(__shfl_sync(0xffffffffu, myrec.y, q + 1));
                    const float b = __int_as_float(__shfl_sync(0xffffffffu, myrec.z, q + 1));
                    float e, l;
                    head_eval(HEAD, dot_row<NV>(x, bufB) + gbias, a, b, p.invB, p.nf, e, l);
                    if (DP) {
                        const bool mine = __shfl_sync(0xffffffffu, myrec.w, q + 1) / p.dp_block == p.dp_rank;
                        e = mine ? e : 0.f; l = mine ? l : 0.f; n_eff += mine;
                    }
                    loss_d += l; sum_e += e;
                    axpy_row<NV>(acc, e, bufB);
                }
            }
            if (!DP) n_eff = n;
            const float fn = (float)n_eff;
            const float sq = warp_sum(activity_l2<NV>(acc, x, fn, p.ce, p.cbias, p.d, bcol, lane));
            if (lane == 0) {
                if (s == 0) { fx_add(wa, loss_d, wa + 6); fx_add(wa + 2, sum_e * (float)p.B, wa + 6); }
                fx_add(wa + 4, fn * p.reg_unscale * sq, wa + 6);
            }

            if (part) {
                store_row<NV>(p.partial[s] + (int64_t)(part - 1) * p.S, acc, lane, S4);
                const int4 lr = __ldg(ps.long_rec + ps.b_long[k] + ir.x);   // {token id, slot, first partial, pieces}
                const int piece = (part - 1) - lr.z, chunk = piece / kChunk;
                const int c_first = lr.z + chunk * kChunk;
                const int c_n = min(kChunk, lr.w - chunk * kChunk);
                __threadfence();
                int last = 0;
                if (lane == 0) last = atomicAdd(p.chunk_cnt[s] + c_first, 1) == c_n - 1;
                if (__shfl_sync(0xffffffffu, last, 0)) {
                    __threadfence();
                    if (lane == 0) p.chunk_cnt[s][c_first] = 0;
                    sum_partials<NV>(acc, p.partial[s] + (int64_t)c_first * p.S, c_n, 1, p.S, bufA, bufB, x, lane, S4);
                    const int n_chunks = (lr.w + kChunk - 1) / kChunk;
                    if (n_chunks > 1) store_row<NV>(p.partial[s] + (int64_t)c_first * p.S, acc, lane, S4);
                    __threadfence();
                    last = 0;
                    if (lane == 0) last = atomicAdd(p.long_cnt[s] + ir.x, 1) == n_chunks - 1;
                    if (__shfl_sync(0xffffffffu, last, 0)) {
                        __threadfence();
                        if (lane == 0) p.long_cnt[s][ir.x] = 0;     // ready for the next step
                        if (n_chunks > 1)
                            sum_partials<NV>(acc, p.partial[s] + (int64_t)lr.z * p.S, n_chunks, kChunk, p.S, bufA, bufB, x, lane, S4);
                        if (!train) {
                            store_row<NV>(p.grad[s] + (int64_t)lr.y * p.S, acc, lane, S4);
                        } else {
                            float *lrow = p.table[s] + ((int64_t)lr.x - id0) * p.P * p.S;
                            load_row<NV>(x, p.snap[s] + (int64_t)lr.y * p.S, lane, S4);
                            if (p.P >= 2) load_row<NV>(s1, lrow + p.S, lane, S4);
                            if (p.P >= 3) load_row<NV>(s2, lrow + 2 * p.S, lane, S4);
                            apply_row<NV>(p, lrow, x, acc, s1, s2, s, step, lane, closed ? __ldcg(p.gap[s] + lr.y) : 0, pol_stream);
                        }
                    }
                }
            } else if (!train) {
                store_row<NV>(p.grad[s] + (int64_t)slot * p.S, acc, lane, S4);
            } else {
                apply_row<NV>(p, row, x, acc, s1, s2, s, step, lane, gap_own, pol_stream);
            }
            ir = ir_next; ir_next = ir_nn; myrec = myrec_next;
            itl = it_n; it_n = it_nn; it_nn = __shfl_sync(0xffffffffu, it_nnn, 0);
            if (itl < nI) {
                load_own(x, ir.y);
                load_opp(bufA, __shfl_sync(0xffffffffu, myrec.x, 0));
            }
        }
    }
    // ---- end of step: per-warp fixed-point sums -> global accumulators (integer atomics: order-independent); the last CTA
    // to finish (ticket) converts them, finishes the step and clears the accumulators and the item counters for the next one
    __shared__ int is_last;
    if (lane < 6 && wa[lane] != 0) atomicAdd(p.loss_acc + lane, (unsigned long long)wa[lane]);
    if (lane == 6 && wa[6] != 0) p.sc->error = 2;   // non-finite loss term (diverged run): sticky, surfaced by the host
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&p.sc->ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        double v[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            v[i] = fx_value((long long)__ldcg(p.loss_acc + 2 * i), (long long)__ldcg(p.loss_acc + 2 * i + 1));
            p.loss_acc[2 * i] = 0; p.loss_acc[2 * i + 1] = 0;
        }
        p.item_ctr[0] = 0; p.item_ctr[1] = 0;
        finish_step(p, step, nullptr, v[0], v[1] / (double)p.B, v[2]);
    }
}

// ---- row-sharded exchange: pack the rows other shards requested / unpack the rows this shard received ------------------
// Layout of both buffers: for every peer (rank order) the rows of the side-0 request list followed by those of side 1.
// PACK: peer r gets the rows of MY block that r's work items need; UNPACK: rows from owner q land at their snapshot position.
template <bool PACK>
__global__ void __launch_bounds__(256) exchange_kernel(const StepParams p, float *buf) {
    int k, step;
    if (!batch_index(p, k, step)) return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int S4 = p.S >> 2;
    const int me = p.shard, N = p.n_shards;
    constexpr int W = kMaxShards + 1;
    int base = 0;   // first buffer row of the current (peer, side) block
    for (int peer = 0; peer < N; ++peer) {
        for (int s = 0; s < 2; ++s) {
            // PACK: requester = peer, owner = me;  UNPACK: requester = me, owner = peer
            const int r = PACK ? peer : me, q = PACK ? me : peer;
            const int32_t *off = p.side[s].need_off + ((int64_t)k * kMaxShards + r) * W;
            const int lo = off[q], n = off[q + 1] - lo;
            const int32_t *pos = p.side[s].need_pos + lo;
            float *snap = p.snap[1 - s];
            for (int i = warp; i < n; i += nwarps) {
                float *a = snap + (int64_t)pos[i] * p.S, *b = buf + (int64_t)(base + i) * p.S;
                for (int f = lane; f < S4; f += 32) {
                    if (PACK) st4(b + 4 * f, ld4(a + 4 * f)); else st4(a + 4 * f, ld4(b + 4 * f));
                }
            }
            base += n;
        }
    }
}

// PULL (peer-mapped workspaces): every row this shard's work items need from another owner is read ONCE from that owner's
// snapshot over NVLink and written to the same position of the local snapshot -- pack + all-to-all + unpack in one launch,
// with no send / receive buffers.  One warp per row, two rows in flight per warp.
__global__ void __launch_bounds__(256) pull_kernel(const StepParams p, const float *const *peer_tab) {
    int k, step;
    if (!batch_index(p, k, step)) return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int S4 = p.S >> 2;
    const int me = p.shard, N = p.n_shards;
    constexpr int W = kMaxShards + 1;
    // flattened list of (owner, side) blocks: first[b] = first global row index of block b
    int first[2 * kMaxShards + 1];
    first[0] = 0;
#pragma unroll
    for (int b = 0; b < 2 * kMaxShards; ++b) {
        const int q = b >> 1, s = b & 1;
        int n = 0;
        if (q < N && q != me) {
            const int32_t *off = p.side[s].need_off + ((int64_t)k * kMaxShards + me) * W;
            n = off[q + 1] - off[q];
        }
        first[b + 1] = first[b] + n;
    }
    const int total = first[2 * kMaxShards];
    auto locate = [&](int i, const float *&src, float *&dst) {
        int b = 0;
#pragma unroll
        for (int c = 1; c < 2 * kMaxShards; ++c) b += (i >= first[c]);
        const int q = b >> 1, s = b & 1;
        const int32_t *off = p.side[s].need_off + ((int64_t)k * kMaxShards + me) * W;
        const int pos = p.side[s].need_pos[off[q] + (i - first[b])];
        src = peer_tab[(1 - s) * kMaxShards + q] + (int64_t)pos * p.S;
        dst = p.snap[1 - s] + (int64_t)pos * p.S;
    };
    // device-side sync: rows are read only after their owners have announced their blocks of this step.  ONE thread per
    // CTA polls (its own memory, system-scope acquire) and the CTA waits at a barrier: thousands of warps polling the same
    // line delay the very store they are waiting for
    if (p.dev_sync) {
        if (threadIdx.x == 0) {
            const int epoch = p.sync[SYNC_EPOCH];
            for (int q = 0; q < N; ++q)
                if (q != me && first[2 * q + 2] > first[2 * q])
                    while (ld_acquire_sys(p.sync + SYNC_STAGED + q) < epoch + 1) __nanosleep(100);
        }
        __syncthreads();
    }
    for (int i = 2 * warp; i < total; i += 2 * nwarps) {
        const float *sa, *sb = nullptr;
        float *da, *db = nullptr;
        locate(i, sa, da);
        const bool two = i + 1 < total;
        if (two) locate(i + 1, sb, db);
        float4 va[4], vb[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int f = lane + 32 * r;
            if (f < S4) va[r] = __ldcg(reinterpret_cast<const float4 *>(sa + 4 * f));
            if (two && f < S4) vb[r] = __ldcg(reinterpret_cast<const float4 *>(sb + 4 * f));
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int f = lane + 32 * r;
            if (f < S4) st4(da + 4 * f, va[r]);
            if (two && f < S4) st4(db + 4 * f, vb[r]);
        }
    }
}

// ---- device-side synchronisation of the row-sharded step (peer_gather == 3) --------------------------------------------
// Every rank's step workspace is peer-mapped (NVLink / NVSwitch), so the two synchronisation points of a step need neither
// the host nor NCCL: a rank announces "my snapshot block of this step is staged" / "my update of this step is done, here
// are my three loss sums" by writing an epoch number (and the sums) straight into every peer's workspace with
// system-scope release stores, and waits by polling ITS OWN memory.  Epochs count steps since the peers were registered
// (monotonic, identical on every rank), so a fast rank's next announcement can never be mistaken for the current one.
//   stage(t) -> sync_kernel<STAGED> -> pull_kernel (waits, owner by owner, for the blocks it reads) -> update(t)
//            -> finish_sync_kernel (announce + wait for all + sum the loss scalars in rank order + finish the step)
// The "update done" wait also protects the snapshots: a rank restages (t+1) only after every peer has finished the
// update -- and hence the pull -- of step t.  Four + one tiny launches, no host round trip: graph-capturable.
// one thread per peer: announce that this rank's snapshot block of the current step is complete (launched behind the
// stage kernel on the same stream: the kernel boundary orders its writes before this one)
__global__ void sync_staged_kernel(const StepParams p) {
    const int q = threadIdx.x;
    if (q >= p.n_shards) return;
    const int epoch = p.sync[SYNC_EPOCH];
    __threadfence_system();
    st_release_sys(peer_sync(p, q) + SYNC_STAGED + p.shard, epoch + 1);
}

// peer_gather == 4 (rows are pushed by their owners' stage kernels): wait until every owner has announced
__global__ void wait_staged_kernel(const StepParams p) {
    const int q = threadIdx.x;
    if (q >= p.n_shards) return;
    const int epoch = p.sync[SYNC_EPOCH];
    while (ld_acquire_sys(p.sync + SYNC_STAGED + q) < epoch + 1) __nanosleep(64);
}

// end of a row-sharded step without the host: publish this rank's loss sums to every peer, wait for everybody's, add them
// in rank order (the same order on every rank: identical, deterministic totals) and finish the step
__global__ void finish_sync_kernel(const StepParams p, const float *mine) {
    int k, step;
    const bool ok = batch_index(p, k, step);
    const int q = threadIdx.x;
    const int epoch = p.sync[SYNC_EPOCH];
    __shared__ float tot[3];
    if (q < p.n_shards) {
        float *dst = reinterpret_cast<float *>(p.peer_base[q] + p.scal_off) + 4 * p.shard;
        dst[0] = mine[0]; dst[1] = mine[1]; dst[2] = mine[2];
        __threadfence_system();
        st_release_sys(peer_sync(p, q) + SYNC_UPDATED + p.shard, epoch + 1);
        while (ld_acquire_sys(p.sync + SYNC_UPDATED + q) < epoch + 1) __nanosleep(64);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int r = 0; r < p.n_shards; ++r) {
            const volatile float *src = p.peer_scal + 4 * r;
            a0 += src[0]; a1 += src[1]; a2 += src[2];
        }
        tot[0] = a0; tot[1] = a1; tot[2] = a2;
        if (ok) finish_step(p, step, tot);
        p.sync[SYNC_EPOCH] = epoch + 1;
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
    const int own0[2] = {p.side[0].b_own[k * (kMaxShards + 1) + p.shard], p.side[1].b_own[k * (kMaxShards + 1) + p.shard]};
    const int U0 = p.side[0].b_own[k * (kMaxShards + 1) + p.shard + 1] - own0[0];
    const int U1 = p.side[1].b_own[k * (kMaxShards + 1) + p.shard + 1] - own0[1];
    const int pos0[2] = {p.shard * p.side[0].b_upad[k], p.shard * p.side[1].b_upad[k]};
    const int64_t id0 = (int64_t)p.shard * p.v_loc;
    const int S4 = p.S >> 2;
    for (int w = warp; w < U0 + U1; w += nwarps) {
        const int s = w >= U0 ? 1 : 0;
        const int j = s ? w - U0 : w;     // index inside this shard's block: also the index into its gradient block
        float *row = p.table[s] + ((int64_t)p.side[s].seg_id[seg0[s] + own0[s] + j] - id0) * p.P * p.S;
        float4 x[NV], G[NV], s1[NV], s2[NV];
        load_row<NV>(x, p.snap[s] + (int64_t)(pos0[s] + j) * p.S, lane, S4);
        load_row<NV>(G, p.grad[s] + (int64_t)j * p.S, lane, S4);
        if (p.P >= 2) load_row<NV>(s1, row + p.S, lane, S4);
        if (p.P >= 3) load_row<NV>(s2, row + 2 * p.S, lane, S4);
        const int gap = (p.opt == GLOVE_OPT_ADAM && p.adam_mode == GLOVE_ADAM_REPLAY) ? __ldcg(p.gap[s] + pos0[s] + j) : 0;
        apply_row<NV>(p, row, x, G, s1, s2, s, step, lane, gap);
    }
}

__global__ void apply_finish_kernel(const StepParams p, const float *reduced) {
    int k, step;
    if (!batch_index(p, k, step)) return;
    if (threadIdx.x == 0) finish_step(p, step, reduced);
}

// ---- host side -------------------------------------------------------------------------------------------------------
// measurement switch, read once from the environment (default = the shipped configuration)
struct Tuning {
    int l2_hints;        // GLOVE_L2_HINTS (default 1): L2 eviction-priority hints on table / snapshot traffic
    int fused_sync;      // GLOVE_FUSED_SYNC (default 0): peer-push announcement / wait inside the stage / update kernels
};
static const Tuning &tuning() {
    static const Tuning t = [] {
        Tuning v{1, 0};
        if (const char *e = getenv("GLOVE_L2_HINTS")) v.l2_hints = atoi(e) != 0;
        if (const char *e = getenv("GLOVE_FUSED_SYNC")) v.fused_sync = atoi(e) != 0;
        return v;
    }();
    return t;
}

static int fill_params(const glove_step_args *a, StepParams &p, int mode) {
    GLOVE_REQUIRE(a, "step: null args");
    GLOVE_REQUIRE(a->row_table && a->col_table && a->scalars && a->plan && a->workspace, "step: null pointer in args");
    GLOVE_REQUIRE(a->V > 0 && a->d > 0 && a->B > 0 && a->plan_K > 0, "step: bad sizes");
    GLOVE_REQUIRE(a->head == GLOVE_HEAD_GLOVE || a->head == GLOVE_HEAD_LOGISTIC, "step: unsupported head %d", a->head);
    GLOVE_REQUIRE(a->optimizer >= 0 && a->optimizer <= 2, "step: unsupported optimizer %d", a->optimizer);
    GLOVE_REQUIRE(a->adam_mode >= 0 && a->adam_mode <= 2, "step: unsupported adam_mode %d", a->adam_mode);
    GLOVE_REQUIRE(a->struct_size == sizeof(glove_step_args), "step: glove_step_args.struct_size %u != %zu (ABI mismatch)",
                  (unsigned)a->struct_size, sizeof(glove_step_args));
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
        p.snap[s] = w.snap[s]; p.partial[s] = w.partial[s]; p.long_cnt[s] = w.long_cnt[s]; p.chunk_cnt[s] = w.chunk_cnt[s];
        p.grad[s] = nullptr;
    }
    p.gap[0] = w.gap[0]; p.gap[1] = w.gap[1];
    p.loss_acc = w.loss_acc; p.item_ctr = w.item_ctr;
    p.sync = w.sync; p.peer_scal = w.peer_scal; p.peer_base = w.peer_base;
    p.sync_off = (char *)w.sync - (char *)a->workspace; p.scal_off = (char *)w.peer_scal - (char *)a->workspace;
    p.dev_sync = (a->peer_gather >= 3 && a->n_shards > 1) ? 1 : 0;
    p.push_snap = (a->peer_gather == 4 && a->n_shards > 1) ? (float *const *)w.peer_tab : nullptr;
    p.fused_sync = 0;               // set by glove_shard_train_step only: callers that drive the phases themselves keep the launches
    p.stage_ticket = w.item_ctr + 2;
    p.l2b1 = replay_log2(a->beta1); p.l2b2 = replay_log2(a->beta2);
    p.l2_hints = tuning().l2_hints;
    p.grad_scalars = nullptr;
    p.alpha = a->alpha; p.alpha_len = a->alpha_len;
    p.loss_out = a->loss_cap > 0 ? a->loss_out : nullptr; p.loss_cap = a->loss_cap > 0 ? a->loss_cap : 1;
    p.V = a->V; p.d = a->d; p.S = S; p.P = table_planes(a->optimizer); p.B = a->B; p.K = a->plan_K;
    p.head = a->head; p.opt = a->optimizer; p.adam_mode = a->adam_mode; p.mode = mode;
    p.lr = a->learning_rate; p.l2 = a->l2_reg; p.rs = a->reg_scale; p.nf = a->neg_factor;
    p.b1 = a->beta1; p.b2 = a->beta2; p.eps = a->epsilon;
    p.invB = 1.0f / (float)p.B;
    p.ce = (2.0f * p.rs * p.l2) / ((float)p.d * (float)p.B);
    p.cbias = (2.0f * p.rs * p.l2) / (float)p.B;
    p.reg_unscale = (float)p.B / (2.0f * p.rs);
    p.dp_world = mode == MODE_TRAIN ? 1 : (a->dp_world > 1 ? a->dp_world : 1);
    p.dp_rank = a->dp_rank;
    if (p.dp_world > 1) GLOVE_REQUIRE(a->B % p.dp_world == 0 && a->dp_rank >= 0 && a->dp_rank < p.dp_world, "step: bad dp split");
    p.dp_block = a->B / p.dp_world;
    p.n_shards = a->n_shards > 1 ? a->n_shards : 1;
    p.shard = p.n_shards > 1 ? a->shard : 0;
    GLOVE_REQUIRE(p.n_shards <= kMaxShards && p.shard >= 0 && p.shard < p.n_shards, "step: bad shard %d of %d", a->shard, a->n_shards);
    GLOVE_REQUIRE(p.n_shards == 1 || mode != MODE_TRAIN, "step: row-sharded tables need the stage / update / finish split");
    p.v_loc = (int32_t)((a->V + p.n_shards - 1) / p.n_shards);
    p.run_stage = p.run_update = 1;
    p.peer_snap = (a->peer_gather == 1 && p.n_shards > 1) ? w.peer_tab : nullptr;   // 1: gather inside the update kernel
    return GLOVE_OK;
}

template <typename Kern>
static int occupancy_grid(Kern kern, int threads, size_t dyn_smem = 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, dyn_smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    int dev = 0, sms = kNumSMs;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return per_sm * sms;
}

template <int NV>
static int launch_step(const StepParams &p, cudaStream_t stream, cudaEvent_t *ev = nullptr) {
    static int g_stage = 0, g_stage_closed = 0, g_update = 0, g_apply = 0;
    if (!g_stage) g_stage = occupancy_grid(stage_kernel<false>, 256);
    if (!g_stage_closed) g_stage_closed = occupancy_grid(stage_closed_kernel<NV>, 256);
    if (!g_update) {
        g_update = occupancy_grid(update_kernel<NV, GLOVE_HEAD_GLOVE, false>, 128);
        if (const char *e = getenv("GLOVE_UPDATE_CTAS")) {   // measurement: cap on resident CTAs per SM
            int dev = 0, sms = kNumSMs;
            if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const int n = atoi(e);
            if (n > 0 && n * sms < g_update) g_update = n * sms;
        }
    }
    if (!g_apply) g_apply = occupancy_grid(apply_kernel<NV>, 128);
    if (p.mode == MODE_TRAIN || p.mode == MODE_GRAD || p.mode == MODE_SHARD) {
        if (ev) cudaEventRecord(ev[0], stream);
        if (p.run_stage) {
            if (p.opt == GLOVE_OPT_ADAM && p.adam_mode == GLOVE_ADAM_REPLAY) {
                // peer-push: per-warp row buffers for the bulk copies to the peers (2.4 KB per warp at d = 300)
                const size_t push_smem = p.push_snap ? (size_t)8 * 2 * p.S * sizeof(float) : 0;
                stage_closed_kernel<NV><<<g_stage_closed, 256, push_smem, stream>>>(p);
            }
            else stage_kernel<false><<<g_stage, 256, 0, stream>>>(p, 0);
        }
        if (ev) cudaEventRecord(ev[1], stream);
        const bool dp = p.dp_world > 1;
        if (p.run_update) {
            if (p.head == GLOVE_HEAD_GLOVE) {
                if (dp) update_kernel<NV, GLOVE_HEAD_GLOVE, true><<<g_update, 128, 0, stream>>>(p);
                else update_kernel<NV, GLOVE_HEAD_GLOVE, false><<<g_update, 128, 0, stream>>>(p);
            } else {
                if (dp) update_kernel<NV, GLOVE_HEAD_LOGISTIC, true><<<g_update, 128, 0, stream>>>(p);
                else update_kernel<NV, GLOVE_HEAD_LOGISTIC, false><<<g_update, 128, 0, stream>>>(p);
            }
        }
        if (ev) { cudaEventRecord(ev[2], stream); cudaEventRecord(ev[3], stream); }
    } else {
        apply_kernel<NV><<<g_apply, 128, 0, stream>>>(p);
        apply_finish_kernel<<<1, 32, 0, stream>>>(p, p.grad_scalars);
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
int64_t glove_step_snapshot_rows(int32_t B) { return B > 0 ? snapshot_rows(B) : 0; }
size_t glove_step_snapshot_offset(int32_t B, int32_t d, int32_t side) {
    if (B <= 0 || d <= 0 || side < 0 || side > 1) return 0;
    StepWs w = step_ws_view((void *)256, B, d);   // any non-null base: offsets are relative to it
    return (size_t)((char *)w.snap[side] - (char *)256);
}

// Row-sharded tables, peer gather: records where every rank's snapshot lives.  peer_workspaces[r] = base of rank r's step
// workspace as mapped in THIS process (symmetric / IPC memory; entry `shard` is the local one), same B and d everywhere.
int glove_shard_set_peers(const glove_step_args *args, const void *const *peer_workspaces, int32_t n_peers, void *stream) {
    GLOVE_REQUIRE(args && args->workspace && peer_workspaces && n_peers > 1 && n_peers <= kMaxShards,
                  "glove_shard_set_peers: bad arguments");
    StepWs w = step_ws_view(args->workspace, args->B, args->d);
    if (args->workspace_bytes < w.bytes)
        return set_error(GLOVE_EWORKSPACE, "glove_shard_set_peers: workspace %zu < required %zu", args->workspace_bytes, w.bytes);
    const float *tab[2 * kMaxShards] = {};
    for (int r = 0; r < n_peers; ++r) {
        GLOVE_REQUIRE(peer_workspaces[r], "glove_shard_set_peers: null workspace of rank %d", r);
        StepWs pw = step_ws_view(const_cast<void *>(peer_workspaces[r]), args->B, args->d);
        for (int s = 0; s < 2; ++s) tab[s * kMaxShards + r] = pw.snap[s];
    }
    GLOVE_CHECK_CUDA(cudaMemcpyAsync(w.peer_tab, tab, sizeof(tab), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    const char *bases[kMaxShards] = {};
    for (int r = 0; r < n_peers; ++r) bases[r] = (const char *)peer_workspaces[r];
    GLOVE_CHECK_CUDA(cudaMemcpyAsync(w.peer_base, bases, sizeof(bases), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    GLOVE_CHECK_CUDA(cudaMemsetAsync(w.sync, 0, sizeof(int32_t) * 32, (cudaStream_t)stream));   // epochs restart with the registration
    GLOVE_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return GLOVE_OK;
}

int glove_train_step(const glove_step_args *args, void *stream) {
    StepParams p;
    int rc = fill_params(args, p, MODE_TRAIN);
    if (rc != GLOVE_OK) return rc;
    return dispatch(p, (cudaStream_t)stream);
}

// ---- K TRAIN steps as one CUDA graph ------------------------------------------------------------------------------------
// The step index lives in device memory and every kernel derives its batch from it, so the launch sequence of a step does
// not depend on WHICH step it is: n_steps x (stage, update) captured once replay correctly for any run of n_steps
// consecutive steps served by the same plan buffer.  One graph launch replaces 2 n_steps kernel launches -- what keeps a
// small-batch run (B = 1,024: ~10 us of kernel time per step) from being bound by launch overhead.
struct glove_step_graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int32_t n_steps = 0;
};

int glove_step_graph_create(const glove_step_args *args, int32_t n_steps, glove_step_graph **out) {
    GLOVE_REQUIRE(out, "glove_step_graph_create: null output");
    *out = nullptr;
    GLOVE_REQUIRE(n_steps > 0 && n_steps <= 4096, "glove_step_graph_create: n_steps out of range");
    const bool shard = args && args->n_shards > 1;     // row-sharded tables: the one-call step with device-side synchronisation
    GLOVE_REQUIRE(!shard || args->peer_gather >= 3, "glove_step_graph_create: row-sharded steps are capturable only with peer_gather >= 3");
    StepParams p;
    int rc = fill_params(args, p, shard ? MODE_SHARD : MODE_TRAIN);
    if (rc != GLOVE_OK) return rc;
    {   // size the grids (occupancy queries, environment) outside the capture
        StepParams q = p;
        q.run_stage = q.run_update = 0;
        rc = dispatch(q, nullptr);
        if (rc != GLOVE_OK) return rc;
    }
    glove_step_graph *g = new (std::nothrow) glove_step_graph();
    GLOVE_REQUIRE(g, "glove_step_graph_create: out of memory");
    cudaStream_t cs = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
        for (int32_t i = 0; i < n_steps && rc == GLOVE_OK; ++i) rc = shard ? glove_shard_train_step(args, cs) : dispatch(p, cs);
        e = cudaStreamEndCapture(cs, &g->graph);
    }
    if (e == cudaSuccess && rc == GLOVE_OK) e = cudaGraphInstantiate(&g->exec, g->graph, 0);
    if (cs) cudaStreamDestroy(cs);
    if (e != cudaSuccess || rc != GLOVE_OK) {
        glove_step_graph_destroy(g);
        return rc != GLOVE_OK ? rc : set_error(GLOVE_ECUDA, "glove_step_graph_create: %s", cudaGetErrorString(e));
    }
    g->n_steps = n_steps;
    *out = g;
    return GLOVE_OK;
}

int glove_step_graph_launch(glove_step_graph *g, void *stream) {
    GLOVE_REQUIRE(g && g->exec, "glove_step_graph_launch: null graph");
    GLOVE_CHECK_CUDA(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
    return GLOVE_OK;
}

int glove_step_graph_destroy(glove_step_graph *g) {
    if (!g) return GLOVE_OK;
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
    return GLOVE_OK;
}

int glove_catchup_step(const glove_step_args *args, int32_t step_index, void *stream) {
    StepParams p;
    int rc = fill_params(args, p, args && args->n_shards > 1 ? MODE_GRAD : MODE_TRAIN);
    if (rc != GLOVE_OK) return rc;
    if (p.opt != GLOVE_OPT_ADAM || p.adam_mode != GLOVE_ADAM_REPLAY_EXACT) return GLOVE_OK;
    static int grid = 0;
    if (!grid) {
        grid = occupancy_grid(stage_kernel<true>, 256);
        // tuning knob: resident catch-up CTAs per SM (the catch-up shares the SMs with the step in flight)
        if (const char *e = getenv("GLOVE_CATCHUP_CTAS_PER_SM")) {
            int dev = 0, sms = kNumSMs;
            if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const int n = atoi(e);
            if (n > 0 && n * sms < grid) grid = n * sms;
        }
    }
    // leave room for the step in flight: the catch-up is arithmetic-bound and is meant to fill its memory stalls
    stage_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(p, step_index);
    commit_ls_kernel<<<kNumSMs, 256, 0, (cudaStream_t)stream>>>(p, step_index);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
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

// fused: the peer-push announcement / wait happen inside the stage / update kernels (one-call step only)
static int shard_stage(const glove_step_args *args, void *stream, int fused) {
    StepParams p;
    int rc = fill_params(args, p, MODE_SHARD);
    if (rc != GLOVE_OK) return rc;
    p.run_update = 0;
    p.fused_sync = fused;
    return dispatch(p, (cudaStream_t)stream);
}
static int shard_update(const glove_step_args *args, float *loss_scalars, void *stream, int fused) {
    StepParams p;
    int rc = fill_params(args, p, MODE_SHARD);
    if (rc != GLOVE_OK) return rc;
    GLOVE_REQUIRE(loss_scalars, "glove_shard_update_step: null scalar buffer");
    p.grad_scalars = loss_scalars;
    p.dp_world = 1;           // ownership is by segment, not by triple
    p.run_stage = 0;
    p.fused_sync = fused;
    return dispatch(p, (cudaStream_t)stream);
}
int glove_shard_stage_step(const glove_step_args *args, void *stream) { return shard_stage(args, stream, 0); }
int glove_shard_update_step(const glove_step_args *args, float *loss_scalars, void *stream) {
    return shard_update(args, loss_scalars, stream, 0);
}

int glove_shard_pack_step(const glove_step_args *args, float *send_buf, void *stream) {
    StepParams p;
    int rc = fill_params(args, p, MODE_SHARD);
    if (rc != GLOVE_OK) return rc;
    GLOVE_REQUIRE(send_buf, "glove_shard_pack_step: null buffer");
    exchange_kernel<true><<<kNumSMs * 4, 256, 0, (cudaStream_t)stream>>>(p, send_buf);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}

int glove_shard_pull_step(const glove_step_args *args, void *stream) {
    StepParams p;
    int rc = fill_params(args, p, MODE_SHARD);
    if (rc != GLOVE_OK) return rc;
    GLOVE_REQUIRE(p.n_shards > 1 && args->peer_gather, "glove_shard_pull_step: needs row-sharded tables and registered peers");
    StepWs w = step_ws_view(args->workspace, args->B, args->d);
    pull_kernel<<<kNumSMs * 4, 256, 0, (cudaStream_t)stream>>>(p, w.peer_tab);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}

// device-side synchronisation (peer_gather == 3): the two announcements of a step as separate tiny launches (what the
// one-call step below strings together; exposed for callers that drive the phases themselves, e.g. N shards emulated
// on one GPU, where every shard must have announced before any shard waits)
int glove_shard_signal_staged(const glove_step_args *args, void *stream) {
    StepParams p;
    int rc = fill_params(args, p, MODE_SHARD);
    if (rc != GLOVE_OK) return rc;
    GLOVE_REQUIRE(p.dev_sync, "glove_shard_signal_staged: needs peer_gather >= 3 and registered peers");
    sync_staged_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}

int glove_shard_wait_staged(const glove_step_args *args, void *stream) {
    StepParams p;
    int rc = fill_params(args, p, MODE_SHARD);
    if (rc != GLOVE_OK) return rc;
    GLOVE_REQUIRE(p.dev_sync, "glove_shard_wait_staged: needs peer_gather >= 3 and registered peers");
    wait_staged_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}

int glove_shard_finish_sync(const glove_step_args *args, const float *loss_scalars, void *stream) {
    StepParams p;
    int rc = fill_params(args, p, MODE_APPLY);
    if (rc != GLOVE_OK) return rc;
    GLOVE_REQUIRE(p.dev_sync && loss_scalars, "glove_shard_finish_sync: needs peer_gather >= 3, registered peers and the scalar buffer");
    finish_sync_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p, loss_scalars);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}

// One row-sharded TRAIN step with no host involvement between its phases: stage -> announce -> pull -> update ->
// announce + wait + finish (this rank's loss sums travel through a scratch slot of the workspace).
int glove_shard_train_step(const glove_step_args *args, void *stream) {
    GLOVE_REQUIRE(args && args->workspace, "glove_shard_train_step: null args");
    float *scal = step_ws_view(args->workspace, args->B, args->d).own_scal;
    // peer-push with the closed-form Adam stage: announcement and wait can live inside the two big kernels (3 launches per step)
    const int fused = tuning().fused_sync && args->peer_gather == 4 && args->n_shards > 1 && args->optimizer == GLOVE_OPT_ADAM &&
                      args->adam_mode == GLOVE_ADAM_REPLAY;
    int rc = shard_stage(args, stream, fused);
    if (rc == GLOVE_OK && !fused) rc = glove_shard_signal_staged(args, stream);
    if (rc == GLOVE_OK && !fused) rc = args->peer_gather == 4 ? glove_shard_wait_staged(args, stream) : glove_shard_pull_step(args, stream);
    if (rc == GLOVE_OK) rc = shard_update(args, scal, stream, fused);
    if (rc == GLOVE_OK) rc = glove_shard_finish_sync(args, scal, stream);
    return rc;
}

int glove_shard_unpack_step(const glove_step_args *args, const float *recv_buf, void *stream) {
    StepParams p;
    int rc = fill_params(args, p, MODE_SHARD);
    if (rc != GLOVE_OK) return rc;
    GLOVE_REQUIRE(recv_buf, "glove_shard_unpack_step: null buffer");
    exchange_kernel<false><<<kNumSMs * 4, 256, 0, (cudaStream_t)stream>>>(p, const_cast<float *>(recv_buf));
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}

int glove_shard_finish_step(const glove_step_args *args, const float *loss_scalars, void *stream) {
    StepParams p;
    int rc = fill_params(args, p, MODE_APPLY);
    if (rc != GLOVE_OK) return rc;
    GLOVE_REQUIRE(loss_scalars, "glove_shard_finish_step: null scalar buffer");
    apply_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p, loss_scalars);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
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
