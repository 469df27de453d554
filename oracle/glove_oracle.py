"""CPU oracle for the GloVe training / eval / top-k hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in NumPy fp32, the arithmetic that yxtay/glove-tensorflow delegates to
tensorflow==2.11.0 / keras==2.11.0 / tensorflow-estimator==2.11.0 (requirements.txt:15,38,40 -- not
vendored under /root/reference, not installable here).  It is the CHECKER for the CUDA path: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import it.  The product package (``glove_tensorflow_b200``) never imports anything from ``oracle/``.

PARITY UNPINNED: the reference ships no tests, fixtures, seeds or golden vectors for this path (SURVEY §4,
§8c).  What pins this oracle is (a) the README sample rows for the preprocessing transforms
(README.md:48-59 vs src/data/text8.py:129-139), (b) golden files produced by importing the reference's own
pure-pandas preprocessor (tests/golden/make_golden.py), (c) a torch-autograd restatement of the Keras layer
and the estimator heads that cross-checks every closed-form gradient here (tests/test_oracle.py).  TF-internal
semantics that could not be confirmed without a TF install are behind explicit switches
(``reg_scale``, ``adam_mode``).

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32

# Keras legacy OptimizerV2 defaults (tf.keras.optimizers.get({"class_name": ..., "config": {"learning_rate": lr}}),
# src/models/train_utils.py:13-16)
ADAM_BETA1 = 0.9
ADAM_BETA2 = 0.999
KERAS_EPSILON = 1e-7
ADAGRAD_INIT_ACC = 0.1

HEAD_GLOVE = "glove"
HEAD_LOGISTIC = "logistic"


# --------------------------------------------------------------------------------------------------------------
# Preprocessing transforms (the only true known-answer material the reference offers)
# --------------------------------------------------------------------------------------------------------------
def glove_weight(count, alpha=0.75, x_max=100):
    """src/data/text8.py:138-139 -- clip((count / x_max) ** alpha, 0, 1)."""
    return np.clip(np.power(np.asarray(count, dtype=np.float64) / x_max, alpha), 0, 1)


def glove_value(value):
    """src/data/text8.py:133 -- ln(value)."""
    return np.log(np.asarray(value, dtype=np.float64))


# --------------------------------------------------------------------------------------------------------------
# Model state
# --------------------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class State:
    """The five trainable variables of MatrixFactorisation (src/models/model_utils.py:31-39) plus optimizer slots.

    R, C: [V, d] f32 row / col embedding tables; rb, cb: [V] f32 biases (stored [V, 1] in Keras);
    g: f32 scalar global bias (zeros init, model_utils.py:39).  ``slots`` holds per-variable optimizer state.
    """
    R: np.ndarray
    C: np.ndarray
    rb: np.ndarray
    cb: np.ndarray
    g: np.float32
    slots: Dict[str, np.ndarray] = dataclasses.field(default_factory=dict)
    step: int = 0  # global_step (src/models/estimator.py:44-45: optimizer.iterations = global_step)

    def copy(self) -> "State":
        return State(self.R.copy(), self.C.copy(), self.rb.copy(), self.cb.copy(), f32(self.g),
                     {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in self.slots.items()}, self.step)


def init_state(vocab_size: int, embedding_size: int, seed: int = 0) -> State:
    """Keras ``Embedding`` default initialiser 'uniform' = U(-0.05, 0.05) for all four tables, global bias zeros
    (src/models/model_utils.py:7-15,39).  TF's RNG stream cannot be reproduced; parity is defined on injected
    tables, so any seeded draw from the same distribution is a valid initial state."""
    rng = np.random.default_rng(seed)
    u = lambda *s: rng.uniform(-0.05, 0.05, size=s).astype(f32)
    return State(u(vocab_size, embedding_size), u(vocab_size, embedding_size), u(vocab_size), u(vocab_size), f32(0))


# --------------------------------------------------------------------------------------------------------------
# Forward
# --------------------------------------------------------------------------------------------------------------
def logits(st: State, i: np.ndarray, j: np.ndarray) -> np.ndarray:
    """MatrixFactorisation.call (src/models/model_utils.py:41-54):
    z_b = sum_k R[i_b,k] C[j_b,k] + rb[i_b] + cb[j_b] + g, with Add([embed_product, row_bias, col_bias, global])
    evaluated left to right in fp32."""
    ep = np.sum(st.R[i] * st.C[j], axis=-1, dtype=f32)
    return (((ep + st.rb[i]).astype(f32) + st.cb[j]).astype(f32) + f32(st.g)).astype(f32)


def _softplus(x: np.ndarray) -> np.ndarray:
    """tf.nn.sigmoid_cross_entropy_with_logits form: max(x,0) + log1p(exp(-|x|))."""
    x = x.astype(f32)
    return (np.maximum(x, f32(0)) + np.log1p(np.exp(-np.abs(x)).astype(f32)).astype(f32)).astype(f32)


def _sigmoid(x: np.ndarray) -> np.ndarray:
    x = x.astype(f32)
    return (f32(1) / (f32(1) + np.exp(-x).astype(f32))).astype(f32)


def reg_loss(st: State, i, j, l2_reg: float, reg_scale: float) -> np.float32:
    """Activity L2 (src/models/model_utils.py:8-13,18-21,51-52; src/models/estimator.py:55):
    s * [ (l2/d)/B sum_b |R[i_b]|^2 + (l2/d)/B sum_b |C[j_b]|^2 + l2/B sum_b rb[i_b]^2 + l2/B sum_b cb[j_b]^2
          + l2 * g^2 ].
    ``reg_scale`` s = 2 for estimator.py under TF >= 2.4 (get_losses_for(None)+get_losses_for(features) counts
    every term twice), s = 1 for estimator_v1.py:21-26,151 and keras.py."""
    B = f32(len(i))
    d = f32(st.R.shape[1])
    lam = f32(l2_reg)
    t_r = (lam / d) * np.sum(np.square(st.R[i]), dtype=f32) / B
    t_c = (lam / d) * np.sum(np.square(st.C[j]), dtype=f32) / B
    t_rb = lam * np.sum(np.square(st.rb[i]), dtype=f32) / B
    t_cb = lam * np.sum(np.square(st.cb[j]), dtype=f32) / B
    t_g = lam * f32(st.g) * f32(st.g)
    return f32(f32(reg_scale) * (t_r + t_c + t_rb + t_cb + t_g))


def forward_loss(st: State, batch: Dict[str, np.ndarray], head: str = HEAD_GLOVE, l2_reg: float = 0.01,
                 reg_scale: float = 2.0, neg_factor: float = 1.0):
    """Returns (loss, z, e) where e_b = dL_data/dz_b.

    glove head: tf.estimator.RegressionHead(weight_column) (src/models/estimator.py:48-56):
        L_data = sum_b w_b (z_b - y_b)^2 / B  (SUM_OVER_BATCH_SIZE), e_b = (2/B) w_b (z_b - y_b).
    logistic head: BinaryClassHead(weight=value,'pos') + BinaryClassHead(weight=neg_weight,'neg') merged by
        MultiHead([pos, neg], [1, neg_factor]) with labels ones / zeros
        (src/models/logistic_matrix_factorisation.py:50-54):
        L_data = (1/B) sum p_b softplus(-z_b) + nu (1/B) sum n_b softplus(z_b),
        e_b = (1/B) [ p_b (sigma(z_b) - 1) + nu n_b sigma(z_b) ].
    """
    i, j = batch["row"], batch["col"]
    B = f32(len(i))
    z = logits(st, i, j)
    if head == HEAD_GLOVE:
        y, w = batch["target"].astype(f32), batch["weight"].astype(f32)
        r = (z - y).astype(f32)
        data = np.sum(w * r * r, dtype=f32) / B
        e = (f32(2) / B) * w * r
    elif head == HEAD_LOGISTIC:
        p, n = batch["pos"].astype(f32), batch["neg"].astype(f32)
        nu = f32(neg_factor)
        data = np.sum(p * _softplus(-z), dtype=f32) / B + nu * (np.sum(n * _softplus(z), dtype=f32) / B)
        s = _sigmoid(z)
        e = (p * (s - f32(1)) + nu * n * s) / B
    else:
        raise ValueError(head)
    loss = f32(data + reg_loss(st, i, j, l2_reg, reg_scale))
    return loss, z, e.astype(f32)


# --------------------------------------------------------------------------------------------------------------
# Backward: IndexedSlices gradients, deduplicated the way OptimizerV2 does it
# --------------------------------------------------------------------------------------------------------------
def sparse_grads(st: State, batch, e: np.ndarray, l2_reg: float, reg_scale: float):
    """Gradients of L = L_data + L_reg w.r.t. the five variables.

    Per occurrence (backward through ResourceGather + activity regulariser):
        dR_b = e_b C[j_b] + (2 s l2 / (d B)) R[i_b]        drb_b = e_b + (2 s l2 / B) rb[i_b]
        dC_b = e_b R[i_b] + (2 s l2 / (d B)) C[j_b]        dcb_b = e_b + (2 s l2 / B) cb[j_b]
        dg   = sum_b e_b + 2 s l2 g
    then OptimizerV2._deduplicate_indexed_slices: unique ids + unsorted_segment_sum, i.e. duplicates are summed in
    batch order (np.add.at is sequential).  Returns {name: (unique_ids, summed_grad)} and dg."""
    i, j = batch["row"], batch["col"]
    B = f32(len(i))
    d = f32(st.R.shape[1])
    ce = f32(2.0 * reg_scale * l2_reg) / (d * B)
    cbias = f32(2.0 * reg_scale * l2_reg) / B
    dR = (e[:, None] * st.C[j] + ce * st.R[i]).astype(f32)
    dC = (e[:, None] * st.R[i] + ce * st.C[j]).astype(f32)
    drb = (e + cbias * st.rb[i]).astype(f32)
    dcb = (e + cbias * st.cb[j]).astype(f32)
    dg = f32(np.sum(e, dtype=f32) + f32(2.0 * reg_scale * l2_reg) * f32(st.g))

    def dedupe(ids, vals):
        uniq, inv = np.unique(ids, return_inverse=True)
        out = np.zeros((len(uniq),) + vals.shape[1:], dtype=f32)
        np.add.at(out, inv, vals)
        return uniq, out

    return {"R": dedupe(i, dR), "C": dedupe(j, dC), "rb": dedupe(i, drb), "cb": dedupe(j, dcb)}, dg


# --------------------------------------------------------------------------------------------------------------
# Optimizers (legacy Keras OptimizerV2 semantics)
# --------------------------------------------------------------------------------------------------------------
def alpha_table(learning_rate: float, n_steps: int, beta1: float = ADAM_BETA1, beta2: float = ADAM_BETA2):
    """alpha[s] for 0-based step s (t = s + 1): lr * sqrt(1 - beta2^t) / (1 - beta1^t), evaluated in fp32 like
    Adam._prepare_local.  Shared verbatim by the CUDA path (uploaded as a device table) so that both sides use
    bit-identical step sizes."""
    t = np.arange(1, n_steps + 1, dtype=f32)
    b1p = np.power(f32(beta1), t).astype(f32)
    b2p = np.power(f32(beta2), t).astype(f32)
    return (f32(learning_rate) * np.sqrt(f32(1) - b2p).astype(f32) / (f32(1) - b1p)).astype(f32)


def _slots(st: State, name: str, like: np.ndarray, init: float = 0.0):
    if name not in st.slots:
        st.slots[name] = np.full_like(like, f32(init))
    return st.slots[name]


def _adam_untouched_step(x, m, v, alpha, b1, b2, eps):
    """One step of legacy-Keras Adam for rows with zero gradient: m*=b1, v*=b2, x -= (alpha*m)/(sqrt(v)+eps)."""
    m *= b1
    v *= b2
    x -= ((alpha * m).astype(f32) / (np.sqrt(v).astype(f32) + eps)).astype(f32)


def idle_run_sequential(x, m, v, alpha, ls: int, T: int, dtype=f32):
    """The zero-gradient steps s = ls .. T-1 of legacy-Keras Adam on untouched elements, one after the other in
    ``dtype`` (what the dense sweep of the reference does to a row between two touches).  Returns (x, m, v)."""
    x, m, v = (np.array(a, dtype) for a in (x, m, v))
    b1, b2, eps = dtype(f32(ADAM_BETA1)), dtype(f32(ADAM_BETA2)), dtype(f32(KERAS_EPSILON))
    for s in range(ls, T):
        _adam_untouched_step_t(x, m, v, dtype(alpha[s]), b1, b2, eps, dtype)
    return x, m, v


def _adam_untouched_step_t(x, m, v, alpha, b1, b2, eps, dtype):
    m *= b1
    v *= b2
    x -= ((alpha * m).astype(dtype) / (np.sqrt(v).astype(dtype) + eps)).astype(dtype)


REPLAY_WINDOW, REPLAY_TERMS = 192, 4


def idle_run_closed_form(x, m, v, alpha, ls: int, T: int):
    """fp32 restatement of the CUDA path's closed-form replay (csrc/glove_common.cuh: replay_coef / replay_x2), the
    default ``adam_mode='replay'``:  with r = sqrt(v0), D = r + eps, q = r / D, u_j = 1 - b2^(j/2),
        x_T = x_ls - (m0 / D) * sum_{k<4} q^k T_k,   T_k = sum_{j=1..min(gap,192)} alpha[ls+j-1] b1^j u_j^k,
        m_T = b1^gap m0,  v_T = b2^gap v0.
    Not part of the reference: it is the thing under test (against idle_run_sequential in fp64)."""
    x, m, v = (np.array(a, f32) for a in (x, m, v))
    gap = T - ls
    if gap <= 0:
        return x, m, v
    b1, b2, eps = float(f32(ADAM_BETA1)), float(f32(ADAM_BETA2)), f32(KERAS_EPSILON)
    n = min(gap, REPLAY_WINDOW)
    j = np.arange(1, n + 1, dtype=np.float64)
    pb1 = np.exp2(j * np.log2(b1)).astype(f32)
    u = (-np.expm1(0.5 * j * np.log(b2))).astype(f32)
    a = (np.asarray(alpha[ls:ls + n], f32) * pb1).astype(f32)
    Tk = []
    for _ in range(REPLAY_TERMS):
        Tk.append(f32(np.sum(a, dtype=f32)))
        a = (a * u).astype(f32)
    r = np.sqrt(np.maximum(v, f32(1e-30))).astype(f32)
    D = (r + eps).astype(f32)
    inv = (f32(1) / D).astype(f32)
    q = (r * inv).astype(f32)
    poly = np.full_like(x, Tk[3])
    for k in (2, 1, 0):
        poly = (poly * q + Tk[k]).astype(f32)
    x = (x - ((m * inv).astype(f32) * poly).astype(f32)).astype(f32)
    dm = f32(np.exp2(np.float64(f32(gap * f32(np.log2(b1))))))
    dv = f32(np.exp2(np.float64(f32(gap * f32(np.log2(b2))))))
    return x, (m * dm).astype(f32), (v * dv).astype(f32)


def apply_adam(st: State, grads, dg, alpha: np.float32, adam_mode: str = "keras_dense"):
    """Legacy tf.keras.optimizers.Adam._resource_apply_sparse (non-lazy): every step
        M <- b1 M (all rows); M[U] += (1-b1) G;  V <- b2 V (all rows); V[U] += (1-b2) G^2;
        X <- X - alpha_t M / (sqrt(V) + eps)  (ALL rows),
    with G the de-duplicated gradient.  ``adam_mode='lazy'`` restricts all three to the touched rows U (LazyAdam);
    it is NOT what the reference computes and exists only to measure the divergence.
    Scalar global bias: dense ResourceApplyAdam: m += (g'-m)(1-b1); v += (g'^2-v)(1-b2); g -= alpha m/(sqrt(v)+eps).
    """
    b1, b2, eps = f32(ADAM_BETA1), f32(ADAM_BETA2), f32(KERAS_EPSILON)
    omb1, omb2 = f32(1) - b1, f32(1) - b2
    for name in ("R", "C", "rb", "cb"):
        x = getattr(st, name)
        m = _slots(st, name + "/m", x)
        v = _slots(st, name + "/v", x)
        uniq, G = grads[name]
        if adam_mode == "keras_dense":
            m *= b1
            m[uniq] += (G * omb1).astype(f32)
            v *= b2
            v[uniq] += ((G * G).astype(f32) * omb2).astype(f32)
            x -= ((alpha * m).astype(f32) / (np.sqrt(v).astype(f32) + eps)).astype(f32)
        elif adam_mode == "lazy":
            mu = (m[uniq] * b1).astype(f32) + (G * omb1).astype(f32)
            vu = (v[uniq] * b2).astype(f32) + ((G * G).astype(f32) * omb2).astype(f32)
            m[uniq], v[uniq] = mu, vu
            x[uniq] -= ((alpha * mu).astype(f32) / (np.sqrt(vu).astype(f32) + eps)).astype(f32)
        else:
            raise ValueError(adam_mode)
    gm = f32(st.slots.get("g/m", f32(0)))
    gv = f32(st.slots.get("g/v", f32(0)))
    gm = f32(gm + (dg - gm) * omb1)
    gv = f32(gv + (dg * dg - gv) * omb2)
    st.g = f32(st.g - (alpha * gm) / (np.sqrt(gv) + eps))
    st.slots["g/m"], st.slots["g/v"] = gm, gv


def apply_adagrad(st: State, grads, dg, lr: float):
    """Legacy Keras Adagrad (initial_accumulator_value=0.1, epsilon=1e-7), truly sparse after de-duplication:
        acc[U] += G^2;  X[U] -= lr * G / (sqrt(acc[U]) + eps)."""
    lr, eps = f32(lr), f32(KERAS_EPSILON)
    for name in ("R", "C", "rb", "cb"):
        x = getattr(st, name)
        acc = _slots(st, name + "/acc", x, ADAGRAD_INIT_ACC)
        uniq, G = grads[name]
        a = (acc[uniq] + (G * G).astype(f32)).astype(f32)
        acc[uniq] = a
        x[uniq] -= ((lr * G).astype(f32) / (np.sqrt(a).astype(f32) + eps)).astype(f32)
    ga = f32(st.slots.get("g/acc", f32(ADAGRAD_INIT_ACC)))
    ga = f32(ga + dg * dg)
    st.g = f32(st.g - (lr * dg) / (np.sqrt(ga) + eps))
    st.slots["g/acc"] = ga


def apply_sgd(st: State, grads, dg, lr: float):
    """Legacy Keras SGD (momentum 0): X[U] -= lr * G."""
    lr = f32(lr)
    for name in ("R", "C", "rb", "cb"):
        x = getattr(st, name)
        uniq, G = grads[name]
        x[uniq] -= (lr * G).astype(f32)
    st.g = f32(st.g - lr * dg)


def train_step(st: State, batch, *, head=HEAD_GLOVE, optimizer="Adam", learning_rate=0.001, l2_reg=0.01,
               reg_scale=2.0, neg_factor=1.0, adam_mode="keras_dense", alpha: Optional[np.ndarray] = None):
    """One Estimator TRAIN step (src/models/estimator.py:13-56 through train_and_evaluate, :95): forward, loss,
    de-duplicated sparse gradients, optimizer apply, global_step += 1.  Returns the pre-update loss."""
    loss, _, e = forward_loss(st, batch, head, l2_reg, reg_scale, neg_factor)
    grads, dg = sparse_grads(st, batch, e, l2_reg, reg_scale)
    if optimizer == "Adam":
        a = alpha[st.step] if alpha is not None else alpha_table(learning_rate, st.step + 1)[st.step]
        apply_adam(st, grads, dg, f32(a), adam_mode)
    elif optimizer == "Adagrad":
        apply_adagrad(st, grads, dg, learning_rate)
    elif optimizer == "SGD":
        apply_sgd(st, grads, dg, learning_rate)
    else:
        raise ValueError("unsupported optimizer %r" % optimizer)
    st.step += 1
    return loss


def train(st: State, coo: Dict[str, np.ndarray], batches: Iterable[np.ndarray], **kw) -> List[np.float32]:
    """Run explicit batches (index lists into the COO) -- the injected batch order parity is defined on."""
    batches = list(batches)
    if kw.get("optimizer", "Adam") == "Adam" and kw.get("alpha") is None:
        kw["alpha"] = alpha_table(kw.get("learning_rate", 0.001), st.step + len(batches))
    losses = []
    for idx in batches:
        losses.append(train_step(st, {k: v[idx] for k, v in coo.items()}, **kw))
    return losses


# --------------------------------------------------------------------------------------------------------------
# Eval metrics (RegressionHead, mode=EVAL; src/models/estimator.py:87-92)
# --------------------------------------------------------------------------------------------------------------
def eval_metrics(st: State, coo, batch_size: int, l2_reg=0.01, reg_scale=2.0):
    """One pass over the csv in file order in batches of ``batch_size`` (last batch may be short):
    average_loss = sum w l / sum w;  loss = mean over batches of (sum w l / B_k + reg_k);
    label/mean, prediction/mean = weighted means;  regularization_loss = mean over batches of reg_k."""
    n = len(coo["row"])
    swl = sw = swy = swz = 0.0
    losses, regs = [], []
    for s in range(0, n, batch_size):
        b = {k: v[s:s + batch_size] for k, v in coo.items()}
        z = logits(st, b["row"], b["col"]).astype(np.float64)
        y, w = b["target"].astype(np.float64), b["weight"].astype(np.float64)
        l = (z - y) ** 2
        swl += float(np.sum(w * l)); sw += float(np.sum(w)); swy += float(np.sum(w * y)); swz += float(np.sum(w * z))
        reg = float(reg_loss(st, b["row"], b["col"], l2_reg, reg_scale))
        regs.append(reg)
        losses.append(float(np.sum(w * l)) / len(z) + reg)
    return {"average_loss": swl / sw, "loss": float(np.mean(losses)), "label/mean": swy / sw,
            "prediction/mean": swz / sw, "regularization_loss": float(np.mean(regs))}


# --------------------------------------------------------------------------------------------------------------
# Cosine top-k (PREDICT; src/models/utils.py:12-19, src/models/model_utils.py:81-110)
# --------------------------------------------------------------------------------------------------------------
def l2_normalize(x: np.ndarray) -> np.ndarray:
    """tf.math.l2_normalize(x, -1): x * rsqrt(max(sum(x^2), 1e-12))."""
    ss = np.sum(np.square(x.astype(f32)), axis=-1, keepdims=True, dtype=f32)
    return (x * (f32(1) / np.sqrt(np.maximum(ss, f32(1e-12))))).astype(f32)


def cosine_topk(table: np.ndarray, query_ids: np.ndarray, k: int):
    """get_predictions: cosine_sim = l2norm(R[input_id]) @ l2norm(R)^T (row table only), tf.math.top_k(k, sorted):
    values descending, ties -> lower index first.  Returns (sim [N,k] f32, idx [N,k] int32)."""
    tn = l2_normalize(table)
    sim = tn[query_ids] @ tn.T
    # stable argsort of -sim gives descending values with lower index first on ties
    idx = np.argsort(-sim, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(sim, idx, axis=1).astype(f32), idx.astype(np.int32)


def format_embeddings(table: np.ndarray, vocab: Sequence[str]) -> Dict[str, dict]:
    """export_embeddings.format_predictions (src/models/export_embeddings.py:13-26): row table only, '<UNK>'
    skipped, later duplicates of a token overwrite earlier ones."""
    out = {}
    for tok, row in zip(vocab, table):
        if tok != "<UNK>":
            out[tok] = {"item_id": tok, "item_embedding": [float(x) for x in row]}
    return out


# --------------------------------------------------------------------------------------------------------------
# Epoch shuffle: keyed bijection on [0, n) (new design; the reference only has tf.data's 10k-row shuffle window,
# src/models/data_utils.py:12-21).  Restated here so tests can check the CUDA kernel index for index.
# --------------------------------------------------------------------------------------------------------------
def _mix32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64) & np.uint64(0xFFFFFFFF)
    x = ((x ^ (x >> np.uint64(16))) * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    x = ((x ^ (x >> np.uint64(15))) * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
    return x ^ (x >> np.uint64(16))


def feistel_permute(pos: np.ndarray, n: int, key: int, rounds: int = 4) -> np.ndarray:
    """Cycle-walking balanced Feistel network over 2*h bits (2^(2h) >= n).  Bijective on [0, n)."""
    bits = max(2, int(n - 1).bit_length())
    h = (bits + 1) // 2
    mask = np.uint64((1 << h) - 1)
    x = pos.astype(np.uint64).copy()
    todo = np.ones(x.shape, dtype=bool)
    while todo.any():
        v = x[todo]
        l, r = v >> np.uint64(h), v & mask
        for rd in range(rounds):
            k = np.uint64((key * 0x9E3779B1 + rd * 0x85EBCA6B) & 0xFFFFFFFF)
            l, r = r, l ^ (_mix32(r ^ k) & mask)
        v = (l << np.uint64(h)) | r
        x[todo] = v
        todo[todo] = v >= np.uint64(n)
    return x.astype(np.int64)
