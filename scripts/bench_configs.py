"""Resident-step timing of the other BASELINE.json configs (1 GPU): K replays of the captured train-step graph with the batch already
in HBM, CUDA events, L2 flushed between steps. Not the driver's bench (that is bench.py on config 2); recorded in DESIGN.md."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import neuralnetlib_b200 as nn
from neuralnetlib_b200 import _backend as B
from test_gpu_models import build_cfg


def time_model(m, x, y, steps=20):
    for _ in range(4):
        m.train_on_batch(x, y)
    plan = m._last_plan
    prog = plan.program("train")
    st = B.stream_obj()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=B.device())
    tot = 0.0
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); prog.run(); b.record(st); b.synchronize()
        tot += a.elapsed_time(b)
    return tot / steps


out = {}
rng = np.random.default_rng(0)
for name, cfg, batch, math in (("cfg1 MLP 784-128-64-10 b32", 1, 32, "fp32"), ("cfg3 deep CNN 32x32x3 b1024", 3, 1024, "fp32"),
                               ("cfg3 deep CNN 32x32x3 b256", 3, 256, "fp32"), ("cfg3 deep CNN 32x32x3 b1024 tf32x3", 3, 1024, "tf32x3"),
                               ("cfg3 deep CNN 32x32x3 b256 tf32x3", 3, 256, "tf32x3"), ("cfg3 deep CNN 32x32x3 b1024 tf32", 3, 1024, "tf32"),
                               ("cfg4 wide MLP 4x4096 b8192 fp32", 4, 8192, "fp32"), ("cfg4 wide MLP 4x4096 b8192 tf32", 4, 8192, "tf32"),
                               ("cfg4 wide MLP 4x4096 b8192 tf32x3", 4, 8192, "tf32x3"), ("cfg4 wide MLP 4x4096 b256 fp32", 4, 256, "fp32"),
                               ("cfg4 wide MLP 4x4096 b256 tf32x3", 4, 256, "tf32x3")):
    if len(sys.argv) > 1 and not any(a in name for a in sys.argv[1:]):
        continue
    nn.set_math(math)
    m = build_cfg(cfg)
    shp = {1: (784,), 3: (32, 32, 3), 4: (4096,)}[cfg]
    x = rng.random((batch,) + shp).astype(np.float32)
    y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, batch)]
    ms = time_model(m, x, y, steps=10 if cfg == 4 else 20)
    out[name] = {"ms_per_step": round(ms, 4), "samples_per_s": round(batch / ms * 1e3)}
    print(name, out[name], flush=True)
    del m
    torch.cuda.empty_cache()
nn.set_math("fp32")
print(json.dumps(out))
