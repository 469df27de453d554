/* CPU oracle (plain C port) for the GloVe TRAIN step.  TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Restates the same arithmetic as oracle/glove_oracle.py (which carries the per-function reference citations):
 *   forward   src/models/model_utils.py:41-54          loss  src/models/estimator.py:48-56
 *   L2        src/models/model_utils.py:8-21,51-52     logistic head  src/models/logistic_matrix_factorisation.py:50-54
 *   optimizer src/models/train_utils.py:13-16 (legacy Keras OptimizerV2 Adam / Adagrad / SGD semantics)
 * The arithmetic itself lives in tensorflow==2.11.0 / keras==2.11.0 / tensorflow-estimator==2.11.0, which are not
 * vendored and not installable here: PARITY UNPINNED (no reference tests/golden vectors exist for this path).
 *
 * Used only by tests/ (checker), and by bench.py's cpu_baseline / --impl reference legs (timed on host cores with
 * OpenMP over the dense Adam sweeps, which is where the reference's CPU time goes).  The product never links it.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).  The same source compiled with
 * -DORACLE_F64 (libglove_oracle64.so) carries every variable and every operation in double: the fp64 SHADOW of the
 * oracle, used by the parity tests to tell rounding noise of the fp32 oracle from errors of the CUDA path
 * (hyper-parameters b1, b2, eps and the alpha table keep their fp32 VALUES, so both precisions solve the same problem).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HEAD_GLOVE 0
#define HEAD_LOGISTIC 1
#define OPT_ADAM 0
#define OPT_ADAGRAD 1
#define OPT_SGD 2
#define ADAM_DENSE 0 /* legacy Keras: decay + apply on ALL rows every step */
#define ADAM_LAZY 1  /* LazyAdam: touched rows only (not the reference; divergence measurement) */

#ifdef ORACLE_F64
typedef double real;
#define RC(x) ((double)(x##f)) /* the fp32 VALUE of the constant, carried in double */
#define R_SQRT sqrt
#define R_FMAX fmax
#define R_LOG1P log1p
#define R_EXP exp
#define R_FABS fabs
#else
typedef float real;
#define RC(x) x##f
#define R_SQRT sqrtf
#define R_FMAX fmaxf
#define R_LOG1P log1pf
#define R_EXP expf
#define R_FABS fabsf
#endif

static const real B1 = RC(0.9), B2 = RC(0.999), EPS = RC(1e-7);

typedef struct {
    int32_t V, d;
    real *R, *C, *rb, *cb; /* [V,d] [V,d] [V] [V] */
    /* optimizer slots: Adam uses s0 = m, s1 = v; Adagrad uses s0 = accumulator; SGD none */
    real *R_s0, *R_s1, *C_s0, *C_s1, *rb_s0, *rb_s1, *cb_s0, *cb_s1;
    real g, g_s0, g_s1;
    int32_t step;
} oracle_state;

int glove_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static real softplusf_(real x) { return R_FMAX(x, RC(0.0)) + R_LOG1P(R_EXP(-R_FABS(x))); }
static real sigmoidf_(real x) { return RC(1.0) / (RC(1.0) + R_EXP(-x)); }

/* dense (all-rows) legacy Adam on one table of n elements, touched rows already scatter-added into m, v */
static void adam_dense_table(real *x, real *m, real *v, int64_t rows, int64_t width, const int32_t *slot_of,
                             const real *G, real alpha) {
    const real omb1 = RC(1.0) - B1, omb2 = RC(1.0) - B2;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        real *xr = x + r * width, *mr = m + r * width, *vr = v + r * width;
        int32_t s = slot_of[r];
        if (s >= 0) {
            const real *g = G + (int64_t)s * width;
            for (int64_t k = 0; k < width; ++k) {
                real gk = g[k];
                real mk = mr[k] * B1;
                mk = mk + gk * omb1;
                real vk = vr[k] * B2;
                vk = vk + (gk * gk) * omb2;
                mr[k] = mk; vr[k] = vk;
                xr[k] = xr[k] - (alpha * mk) / (R_SQRT(vk) + EPS);
            }
        } else {
            for (int64_t k = 0; k < width; ++k) {
                real mk = mr[k] * B1, vk = vr[k] * B2;
                mr[k] = mk; vr[k] = vk;
                xr[k] = xr[k] - (alpha * mk) / (R_SQRT(vk) + EPS);
            }
        }
    }
}

static void adam_lazy_table(real *x, real *m, real *v, int64_t width, const int32_t *uniq, int32_t n_uniq,
                            const real *G, real alpha) {
    const real omb1 = RC(1.0) - B1, omb2 = RC(1.0) - B2;
#pragma omp parallel for schedule(static)
    for (int32_t s = 0; s < n_uniq; ++s) {
        int64_t r = uniq[s];
        real *xr = x + r * width, *mr = m + r * width, *vr = v + r * width;
        const real *g = G + (int64_t)s * width;
        for (int64_t k = 0; k < width; ++k) {
            real gk = g[k];
            real mk = mr[k] * B1 + gk * omb1;
            real vk = vr[k] * B2 + (gk * gk) * omb2;
            mr[k] = mk; vr[k] = vk;
            xr[k] = xr[k] - (alpha * mk) / (R_SQRT(vk) + EPS);
        }
    }
}

static void adagrad_table(real *x, real *acc, int64_t width, const int32_t *uniq, int32_t n_uniq, const real *G,
                          real lr) {
#pragma omp parallel for schedule(static)
    for (int32_t s = 0; s < n_uniq; ++s) {
        int64_t r = uniq[s];
        const real *g = G + (int64_t)s * width;
        for (int64_t k = 0; k < width; ++k) {
            real a = acc[r * width + k] + g[k] * g[k];
            acc[r * width + k] = a;
            x[r * width + k] -= (lr * g[k]) / (R_SQRT(a) + EPS);
        }
    }
}

static void sgd_table(real *x, int64_t width, const int32_t *uniq, int32_t n_uniq, const real *G, real lr) {
#pragma omp parallel for schedule(static)
    for (int32_t s = 0; s < n_uniq; ++s) {
        int64_t r = uniq[s];
        for (int64_t k = 0; k < width; ++k) x[r * width + k] -= lr * G[(int64_t)s * width + k];
    }
}

/* Runs n_steps TRAIN steps over explicit batches.  batch_idx[n_steps*B] indexes the COO arrays.
 * colA/colB = (glove_value, glove_weight) for the glove head, (value, neg_weight) for the logistic head.
 * alpha[step] is the fp32 Adam step-size table shared with the CUDA path.  losses[n_steps] receives the pre-update
 * loss of every step.  Returns 0, or -1 on allocation failure / bad arguments. */
int glove_oracle_train(oracle_state *st, const int32_t *row, const int32_t *col, const real *colA,
                       const real *colB, const int64_t *batch_idx, int32_t n_steps, int32_t B, int32_t head,
                       int32_t optimizer, real lr, real l2, real reg_scale, real neg_factor, int32_t adam_mode,
                       const real *alpha, real *losses) {
    const int32_t V = st->V, d = st->d;
    if (B <= 0 || V <= 0 || d <= 0) return -1;
    int32_t *slot_r = (int32_t *)malloc(sizeof(int32_t) * V), *slot_c = (int32_t *)malloc(sizeof(int32_t) * V);
    int32_t *uniq_r = (int32_t *)malloc(sizeof(int32_t) * B), *uniq_c = (int32_t *)malloc(sizeof(int32_t) * B);
    real *GR = (real *)malloc(sizeof(real) * (size_t)B * d), *GC = (real *)malloc(sizeof(real) * (size_t)B * d);
    real *Grb = (real *)malloc(sizeof(real) * B), *Gcb = (real *)malloc(sizeof(real) * B);
    real *e = (real *)malloc(sizeof(real) * B), *z = (real *)malloc(sizeof(real) * B);
    if (!slot_r || !slot_c || !uniq_r || !uniq_c || !GR || !GC || !Grb || !Gcb || !e || !z) return -1;
    for (int32_t r = 0; r < V; ++r) slot_r[r] = slot_c[r] = -1;
    const real fB = (real)B, fd = (real)d;
    const real ce = (RC(2.0) * reg_scale * l2) / (fd * fB), cb_ = (RC(2.0) * reg_scale * l2) / fB;

    for (int32_t s = 0; s < n_steps; ++s) {
        const int64_t *idx = batch_idx + (int64_t)s * B;
        /* forward + per-example residual */
        double data = 0.0, sq_r = 0.0, sq_c = 0.0, sq_rb = 0.0, sq_cb = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : data, sq_r, sq_c, sq_rb, sq_cb)
        for (int32_t b = 0; b < B; ++b) {
            int64_t t = idx[b];
            const real *Ri = st->R + (int64_t)row[t] * d, *Cj = st->C + (int64_t)col[t] * d;
            real ep = RC(0.0), nr = RC(0.0), nc = RC(0.0);
            for (int32_t k = 0; k < d; ++k) { ep += Ri[k] * Cj[k]; nr += Ri[k] * Ri[k]; nc += Cj[k] * Cj[k]; }
            real rbv = st->rb[row[t]], cbv = st->cb[col[t]];
            real zb = ((ep + rbv) + cbv) + st->g;
            z[b] = zb;
            if (head == HEAD_GLOVE) {
                real r_ = zb - colA[t];
                data += (double)(colB[t] * r_ * r_);
                e[b] = (RC(2.0) / fB) * colB[t] * r_;
            } else {
                real sg = sigmoidf_(zb);
                data += (double)(colA[t] * softplusf_(-zb)) + (double)neg_factor * (double)(colB[t] * softplusf_(zb));
                e[b] = (colA[t] * (sg - RC(1.0)) + neg_factor * colB[t] * sg) / fB;
            }
            sq_r += nr; sq_c += nc; sq_rb += (double)(rbv * rbv); sq_cb += (double)(cbv * cbv);
        }
        double reg = (double)reg_scale * ((double)(l2 / fd) * sq_r / B + (double)(l2 / fd) * sq_c / B +
                                          (double)l2 * sq_rb / B + (double)l2 * sq_cb / B +
                                          (double)l2 * (double)st->g * (double)st->g);
        losses[s] = (real)(data / B + reg);

        /* de-duplicated gradients, duplicates summed in batch order */
        int32_t n_r = 0, n_c = 0;
        real se = RC(0.0);
        for (int32_t b = 0; b < B; ++b) {
            int64_t t = idx[b];
            int32_t i = row[t], j = col[t];
            const real *Ri = st->R + (int64_t)i * d, *Cj = st->C + (int64_t)j * d;
            int32_t sr = slot_r[i], sc = slot_c[j];
            if (sr < 0) { sr = slot_r[i] = n_r; uniq_r[n_r++] = i; memset(GR + (int64_t)sr * d, 0, sizeof(real) * d); Grb[sr] = RC(0.0); }
            if (sc < 0) { sc = slot_c[j] = n_c; uniq_c[n_c++] = j; memset(GC + (int64_t)sc * d, 0, sizeof(real) * d); Gcb[sc] = RC(0.0); }
            real eb = e[b];
            real *gr = GR + (int64_t)sr * d, *gc = GC + (int64_t)sc * d;
            for (int32_t k = 0; k < d; ++k) {
                gr[k] += eb * Cj[k] + ce * Ri[k];
                gc[k] += eb * Ri[k] + ce * Cj[k];
            }
            Grb[sr] += eb + cb_ * st->rb[i];
            Gcb[sc] += eb + cb_ * st->cb[j];
            se += eb;
        }
        real dg = se + (RC(2.0) * reg_scale * l2) * st->g;

        if (optimizer == OPT_ADAM) {
            real a = alpha[st->step];
            if (adam_mode == ADAM_DENSE) {
                adam_dense_table(st->R, st->R_s0, st->R_s1, V, d, slot_r, GR, a);
                adam_dense_table(st->C, st->C_s0, st->C_s1, V, d, slot_c, GC, a);
                adam_dense_table(st->rb, st->rb_s0, st->rb_s1, V, 1, slot_r, Grb, a);
                adam_dense_table(st->cb, st->cb_s0, st->cb_s1, V, 1, slot_c, Gcb, a);
            } else {
                adam_lazy_table(st->R, st->R_s0, st->R_s1, d, uniq_r, n_r, GR, a);
                adam_lazy_table(st->C, st->C_s0, st->C_s1, d, uniq_c, n_c, GC, a);
                adam_lazy_table(st->rb, st->rb_s0, st->rb_s1, 1, uniq_r, n_r, Grb, a);
                adam_lazy_table(st->cb, st->cb_s0, st->cb_s1, 1, uniq_c, n_c, Gcb, a);
            }
            real gm = st->g_s0 + (dg - st->g_s0) * (RC(1.0) - B1);
            real gv = st->g_s1 + (dg * dg - st->g_s1) * (RC(1.0) - B2);
            st->g = st->g - (a * gm) / (R_SQRT(gv) + EPS);
            st->g_s0 = gm; st->g_s1 = gv;
        } else if (optimizer == OPT_ADAGRAD) {
            adagrad_table(st->R, st->R_s0, d, uniq_r, n_r, GR, lr);
            adagrad_table(st->C, st->C_s0, d, uniq_c, n_c, GC, lr);
            adagrad_table(st->rb, st->rb_s0, 1, uniq_r, n_r, Grb, lr);
            adagrad_table(st->cb, st->cb_s0, 1, uniq_c, n_c, Gcb, lr);
            real ga = st->g_s0 + dg * dg;
            st->g = st->g - (lr * dg) / (R_SQRT(ga) + EPS);
            st->g_s0 = ga;
        } else {
            sgd_table(st->R, d, uniq_r, n_r, GR, lr);
            sgd_table(st->C, d, uniq_c, n_c, GC, lr);
            sgd_table(st->rb, 1, uniq_r, n_r, Grb, lr);
            sgd_table(st->cb, 1, uniq_c, n_c, Gcb, lr);
            st->g = st->g - lr * dg;
        }
        for (int32_t k = 0; k < n_r; ++k) slot_r[uniq_r[k]] = -1;
        for (int32_t k = 0; k < n_c; ++k) slot_c[uniq_c[k]] = -1;
        st->step += 1;
    }
    free(slot_r); free(slot_c); free(uniq_r); free(uniq_c); free(GR); free(GC); free(Grb); free(Gcb); free(e); free(z);
    return 0;
}
