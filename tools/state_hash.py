"""Trains a small Zipf problem for a few steps and prints a hash of the tables + the losses: run it under different
GLOVE_UPDATE_KERNEL / GLOVE_* settings to check that kernel variants are bit-identical.
usage: python tools/state_hash.py [V d B steps]"""
import hashlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glove_tensorflow_b200.engine import GloveEngine

V, d, B, steps = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (20000, 300, 8192, 40)))
opt = sys.argv[5] if len(sys.argv) > 5 else "Adam"
rng = np.random.default_rng(3)
n = 40 * B
p = 1.0 / np.arange(1, V + 1); p /= p.sum()
row = rng.choice(V, n, p=p).astype(np.int32); col = rng.choice(V, n, p=p).astype(np.int32)
tgt = rng.normal(2.0, 1.0, n).astype(np.float32); w = rng.uniform(0.1, 1, n).astype(np.float32)
eng = GloveEngine(V, d, optimizer=opt, learning_rate=0.01, batch_size=B, plan_steps=8, max_steps=steps + 16)
eng.init_uniform(1)
eng.set_coo(row, col, tgt, w, shuffle_key=7)
losses = eng.train(steps)
st = eng.get_state(slots=True)
h = hashlib.sha256()
for k in sorted(st):
    if isinstance(st[k], np.ndarray):
        h.update(np.ascontiguousarray(st[k]).tobytes())
print("kernel=%s V=%d d=%d B=%d steps=%d opt=%s tables=%s losses=%s last=%.9g" % (
    os.environ.get("GLOVE_UPDATE_KERNEL", "default"), V, d, B, steps, opt, h.hexdigest()[:16],
    hashlib.sha256(np.asarray(losses, np.float32).tobytes()).hexdigest()[:16], float(losses[-1])))
