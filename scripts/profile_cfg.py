"""Run a few train steps of one BASELINE config (for an ncu launch list): python scripts/profile_cfg.py <cfg> <batch> [math]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import neuralnetlib_b200 as nn
from test_gpu_models import build_cfg

cfg, batch = int(sys.argv[1]), int(sys.argv[2])
nn.set_math(sys.argv[3] if len(sys.argv) > 3 else "fp32")
rng = np.random.default_rng(0)
m = build_cfg(cfg)
shp = {1: (784,), 2: (28, 28, 1), 3: (32, 32, 3), 4: (4096,)}[cfg]
x = rng.random((batch,) + shp).astype(np.float32)
y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, batch)]
for _ in range(3):
    print(m.train_on_batch(x, y))
torch.cuda.synchronize()
