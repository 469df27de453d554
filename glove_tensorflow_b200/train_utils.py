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


def latest_checkpoint(job_dir):
    best, best_step = None, -1
    for path in glob.glob(os.path.join(job_dir, "model.ckpt-*.npz")):
        try:
            step = int(os.path.basename(path)[len("model.ckpt-"):-len(".npz")])
        except ValueError:
            continue
        if step > best_step:
            best, best_step = path, step
    return best


def save_checkpoint(engine, job_dir):
    """Variables in the reference's layout (4 tables + global bias) + optimizer slots + global_step."""
    state = engine.get_state(slots=True)  # flushes lazy Adam state: tables are exactly the dense-Keras values
    sc = engine.read_scalars()
    path = checkpoint_path(job_dir, int(state["step"]))
    np.savez(path, **{k.replace("/", "__"): v for k, v in state.items()}, g_s0=np.float32(sc["g_s0"]),
             g_s1=np.float32(sc["g_s1"]), shuffle_key=np.int64(engine.shuffle_key))
    with open(os.path.join(job_dir, "checkpoint"), "w") as f:
        json.dump({"model_checkpoint_path": os.path.basename(path), "global_step": int(state["step"])}, f)
    return path


def load_checkpoint(engine, path):
    import torch
    z = np.load(path, allow_pickle=False)
    engine.load_state(z["R"], z["C"], z["rb"], z["cb"], float(z["g"]))
    step = int(z["step"])
    dev = engine.device
    for side, (tab, bias) in (("row", ("R", "rb")), ("col", ("C", "cb"))):
        for p in range(1, engine.P):
            engine.set_plane(side, p, torch.from_numpy(z["%s__s%d" % (tab, p - 1)]).to(dev),
                             torch.from_numpy(z["%s__s%d" % (bias, p - 1)]).to(dev))
        # every row of a flushed checkpoint is current through `step` (0 = never touched is indistinguishable from
        # "touched, then flushed" only when m = v = 0, where the idle step is a no-op anyway)
        ls = torch.full((engine.V,), step, dtype=torch.int32, device=dev)
        if engine.optimizer == "Adam":
            m = torch.from_numpy(z["%s__s0" % tab]).to(dev)
            ls = torch.where((m != 0).any(dim=1), ls, torch.zeros_like(ls))
        engine.set_last_step(side, ls)
    engine._write_scalars(g_s0=float(z["g_s0"]), g_s1=float(z["g_s1"]))
    engine.set_step(step)
    return step


def train_and_evaluate(engine, train_steps, job_dir, eval_fn=None, save_checkpoints_secs=EVAL_INTERVAL, log_every=100):
    """tf.estimator.train_and_evaluate for the local case (ref src/models/estimator.py:95): train to ``train_steps``
    (``max_steps`` semantics: counts steps already in the checkpoint), checkpoint + evaluate every <= 300 s and at the
    end.  Returns the list of (step, metrics)."""
    history = []
    last_ckpt = time.time()
    t0, s0 = time.time(), engine.host_step
    while engine.host_step < train_steps:
        n = min(log_every, train_steps - engine.host_step)
        losses = engine.train(n)
        if not np.all(np.isfinite(losses)):
            raise FloatingPointError("NaN loss at step %d" % engine.host_step)
        logger.info("step %d loss %.6f (%.1f steps/s)", engine.host_step, float(losses[-1]),
                    (engine.host_step - s0) / max(time.time() - t0, 1e-9))
        if time.time() - last_ckpt >= min(save_checkpoints_secs, 300) or engine.host_step >= train_steps:
            save_checkpoint(engine, job_dir)
            last_ckpt = time.time()
            if eval_fn is not None:
                metrics = eval_fn(engine)
                metrics["global_step"] = engine.host_step
                history.append((engine.host_step, metrics))
                logger.info("eval @%d: %s", engine.host_step, metrics)
    return history
