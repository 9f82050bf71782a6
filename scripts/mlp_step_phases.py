"""Per-phase time of the single-launch MLP train step (csrc/mlp_step.cu): CTA 0's %globaltimer stamps after every grid barrier.
   python scripts/mlp_step_phases.py [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import neuralnetlib_b200 as nn
from neuralnetlib_b200 import _backend as B
from test_gpu_models import build_cfg

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
rng = np.random.default_rng(0)
m = build_cfg(1)
x = rng.random((batch, 784)).astype(np.float32)
y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, batch)]
for _ in range(5):
    m.train_on_batch(x, y)
plan = m._last_plan
prog = plan.program("train")
sync = plan._mlp_keep[2]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=B.device())
names = ["fwd1", "bar", "fwd2", "bar", "head", "bar", "bwd3", "bar", "bwd2", "bar", "bwd1", "bar", "adam"]
for cold in (True, False):
    acc = np.zeros(13)
    reps = 20
    for _ in range(reps):
        if cold:
            flush.zero_()
        prog.run()
        torch.cuda.synchronize()
        st = sync[2:32].cpu().numpy().view(np.uint64)[:14].astype(np.int64)
        acc += np.diff(st) / reps
    print("cold L2" if cold else "warm L2", " ".join(f"{n}={v / 1e3:.1f}us" for n, v in zip(names, acc)), f"total={acc.sum() / 1e3:.1f}us")

# launch floor: a captured graph of ONE trivial kernel (step_begin), timed the same way (events around the launch, L2 flushed)
from neuralnetlib_b200 import _ops as ops
ss = B.zeros((8,), dtype=torch.float64)
ops.graph_begin(); ops.step_begin(ss, None, None, 0); g = ops.graph_end()
st = B.stream_obj()
for cold in (True, False):
    tot = 0.0
    for _ in range(50):
        if cold:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); ops.graph_launch(g); b.record(st); b.synchronize()
        tot += a.elapsed_time(b)
    print("one-node graph floor", "after flush" if cold else "back to back", f"{tot / 50 * 1e3:.1f} us")
tot = 0.0
for _ in range(50):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st); prog.run(); b.record(st); b.synchronize()
    tot += a.elapsed_time(b)
print("mlp step (events, after flush)", f"{tot / 50 * 1e3:.1f} us")
