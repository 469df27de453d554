"""Train / eval harness: the part of the reference's src/models/train_utils.py that matters to the training path --
step loop with ``max_steps`` (:39-40), checkpoint every <= 300 s (:26-27) and after the last step, EVAL after each
checkpoint over one pass of the same csv (:47-54, src/models/estimator.py:86-92), resume from the latest checkpoint in
``job_dir`` (what tf.estimator.Estimator does on restart)."""
import glob
import json
import logging
import os
import time

import numpy as np

logger = logging.getLogger(__name__)
EVAL_INTERVAL = 300


def checkpoint_path(job_dir, step):
    return os.path.join(job_dir, "model.ckpt-%d.npz" % step)


KEEP_CHECKPOINT_MAX = 5   # tf.estimator.RunConfig default


def _checkpoint_steps(job_dir):
    out = []
    for path in glob.glob(os.path.join(job_dir, "model.ckpt-*.npz")):
        try:
            out.append((int(os.path.basename(path)[len("model.ckpt-"):-len(".npz")]), path))
        except ValueError:
            continue
    return sorted(out)


def latest_checkpoint(job_dir):
    """The checkpoint named by the ``checkpoint`` index (written AFTER the file it names was renamed into place, so it
    never points at a truncated file); without a usable index, the newest ``model.ckpt-N.npz`` that opens."""
    index = os.path.join(job_dir, "checkpoint")
    if os.path.exists(index):
        try:
            with open(index) as f:
                path = os.path.join(job_dir, json.load(f)["model_checkpoint_path"])
            if os.path.exists(path):
                return path
        except (ValueError, KeyError, OSError):
            pass
    for _, path in reversed(_checkpoint_steps(job_dir)):
        try:
            with np.load(path, allow_pickle=False) as z:
                z["step"]
            return path
        except Exception:   # truncated / foreign file: try the one before
            continue
    return None


def save_checkpoint(engine, job_dir, keep=KEEP_CHECKPOINT_MAX):
    """Variables in the reference's layout (4 tables + global bias) + optimizer slots + global_step + the per-row
    "ever updated" masks.  Written to a temporary file and renamed, index updated after the rename, older checkpoints
    pruned to the last ``keep`` (tf.estimator keeps 5)."""
    state = engine.get_state(slots=True)  # flushes lazy Adam state: tables are exactly the dense-Keras values
    sc = engine.read_scalars()
    step = int(state["step"])
    path = checkpoint_path(job_dir, step)
    touched = {"touched_%s" % side: (engine.get_last_step(side) > 0).cpu().numpy() for side in ("row", "col")}
    tmp = path + ".tmp.npz"
    np.savez(tmp, **{k.replace("/", "__"): v for k, v in state.items()}, g_s0=np.float32(sc["g_s0"]),
             g_s1=np.float32(sc["g_s1"]), shuffle_key=np.int64(engine.shuffle_key), **touched)
    os.replace(tmp, path)
    index_tmp = os.path.join(job_dir, "checkpoint.tmp")
    with open(index_tmp, "w") as f:
        json.dump({"model_checkpoint_path": os.path.basename(path), "global_step": step}, f)
    os.replace(index_tmp, os.path.join(job_dir, "checkpoint"))
    if keep and keep > 0:
        for _, old in _checkpoint_steps(job_dir)[:-keep]:
            try:
                os.remove(old)
            except OSError:
                pass
    return path


def load_checkpoint(engine, path):
    import torch
    z = np.load(path, allow_pickle=False)
    step = int(z["step"])
    if step > engine.max_steps:
        raise ValueError("checkpoint %s is at global_step %d, beyond this run's %d steps (--train-steps counts the steps "
                         "already in the checkpoint: raise it to continue training)" % (path, step, engine.max_steps))
    engine.load_state(z["R"], z["C"], z["rb"], z["cb"], float(z["g"]))
    dev = engine.device
    for side, (tab, bias) in (("row", ("R", "rb")), ("col", ("C", "cb"))):
        slots = []
        for p in range(1, engine.P):
            e, b = z["%s__s%d" % (tab, p - 1)], z["%s__s%d" % (bias, p - 1)]
            slots += [e, b.reshape(-1, 1)]
            engine.set_plane(side, p, torch.from_numpy(e).to(dev), torch.from_numpy(b).to(dev))
        # A flushed checkpoint holds every row current through `step`.  last_step = step on every row that was ever
        # updated, 0 ("never touched": the stage skips it) on the others.  The mask is stored; for checkpoints written
        # before it was, a row counts as touched when ANY of its slot values (embedding or bias, m or v) is non-zero --
        # m alone underflows to 0 after ~900 idle steps while v is still decaying.
        if "touched_%s" % side in z.files:
            touched = torch.from_numpy(z["touched_%s" % side]).to(dev)
        elif slots:
            touched = torch.from_numpy(np.any(np.concatenate(slots, 1) != 0, axis=1)).to(dev)
        else:
            touched = torch.zeros(engine.V, dtype=torch.bool, device=dev)
        ls = torch.where(touched, torch.full((engine.V,), step, dtype=torch.int32, device=dev),
                         torch.zeros(engine.V, dtype=torch.int32, device=dev))
        engine.set_last_step(side, ls)
    engine._write_scalars(g_s0=float(z["g_s0"]), g_s1=float(z["g_s1"]))
    engine.set_step(step)
    return step


def train_and_evaluate(engine, train_steps, job_dir, eval_fn=None, save_checkpoints_secs=EVAL_INTERVAL, log_every=100):
    """tf.estimator.train_and_evaluate for the local case (ref src/models/estimator.py:95): train to ``train_steps``
    (``max_steps`` semantics: counts steps already in the checkpoint), checkpoint + evaluate every <= 300 s and at the
    end.  Returns the list of (step, metrics)."""
    from .summary import EventWriter
    history = []
    last_ckpt = time.time()
    t0, s0 = time.time(), engine.host_step
    # TensorBoard summaries, as model_fn's add_summary + tf.estimator write them [ref src/models/model_utils.py:113-118]:
    # every `log_every` (save_summary_steps = 100) TRAIN steps `loss`, `global_step/sec`, `mf/global_bias` and the
    # histograms `mf/row_biases`, `mf/col_biases` into <job_dir>; the EVAL metrics into <job_dir>/eval
    train_writer, eval_writer = EventWriter(job_dir), None
    try:
        while engine.host_step < train_steps:
            n = min(log_every, train_steps - engine.host_step)
            t1 = time.time()
            losses = engine.train(n)
            if not np.all(np.isfinite(losses)):
                raise FloatingPointError("NaN loss at step %d" % engine.host_step)
            rate = n / max(time.time() - t1, 1e-9)
            logger.info("step %d loss %.6f (%.1f steps/s)", engine.host_step, float(losses[-1]),
                        (engine.host_step - s0) / max(time.time() - t0, 1e-9))
            rb, cb = engine.bias_vectors()
            train_writer.add(engine.host_step, scalars={"loss": float(losses[-1]), "global_step/sec": rate,
                                                        "mf/global_bias": float(engine.read_scalars()["g"])},
                             histograms={"mf/row_biases": rb, "mf/col_biases": cb})
            if time.time() - last_ckpt >= min(save_checkpoints_secs, 300) or engine.host_step >= train_steps:
                save_checkpoint(engine, job_dir)
                last_ckpt = time.time()
                if eval_fn is not None:
                    metrics = eval_fn(engine)
                    metrics["global_step"] = engine.host_step
                    history.append((engine.host_step, metrics))
                    logger.info("eval @%d: %s", engine.host_step, metrics)
                    if eval_writer is None:
                        eval_writer = EventWriter(os.path.join(job_dir, "eval"))
                    eval_writer.add(engine.host_step, scalars={k: v for k, v in metrics.items() if k != "global_step"})
    finally:
        train_writer.close()
        if eval_writer is not None:
            eval_writer.close()
    return history
