"""Development probe: host-side enqueue time vs device time of one training step (is the step launch-bound?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vnet_tensorflow_b200.engine import VNetEngine
from vnet_tensorflow_b200.init import initialize
from vnet_tensorflow_b200.synthetic import synth_batch

P, B = 128, 2
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
eng = VNetEngine(num_classes=2, in_channels=1, patch_shape=(P, P, P), max_batch=B, precision=prec,
                 loss="weighted_sorensen", loss_weights=(0.1, 1.0))
initialize(eng, 42)
img, lab = synth_batch(0, B, P, 1, 2)
eng.upload_batch(img, lab)
for i in range(3):
    eng.train_step_resident(B, 0.01, seed=i)
eng.sync()
enq, tot = [], []
for i in range(10):
    t0 = time.perf_counter()
    eng.train_step_resident(B, 0.01, seed=10 + i)
    t1 = time.perf_counter()
    eng.sync()
    t2 = time.perf_counter()
    enq.append((t1 - t0) * 1e3)
    tot.append((t2 - t0) * 1e3)
print("STEP_PROBE %s env WG=%s: enqueue %.2f ms, enqueue+sync %.2f ms (median of 10 isolated steps)" % (
    prec, os.environ.get("VNB_WGRAD_STREAM", "1"), float(np.median(enq)), float(np.median(tot))))
eng.event_record(0)
for i in range(10):
    eng.train_step_resident(B, 0.01, seed=30 + i)
eng.event_record(1)
eng.sync()
print("STEP_PROBE back-to-back: %.2f ms/step" % (eng.event_elapsed_ms() / 10))
