"""Diagnostic (GPU): is the config-2 conv weight gradient reproducible across fresh models / after other models ran?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import nnl_oracle as O
import neuralnetlib_b200 as nn
from test_gpu_models import build_cfg

def nrel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a.reshape(b.shape) - b).max() / max(np.abs(b).max(), 1e-300)

x, y = O.synthetic_batch(2, 256)
specs, _ = O.config_specs(2)
ref = O.OracleSequential(specs, "cce", O.OracleAdam(), seed=42)
ref.train_on_batch(x, y)
rconv = [r for r in ref.layers if "w" in r][0]
prev = None
def run(tag):
    global prev
    m = build_cfg(2)
    m.train_on_batch(x, y)
    conv = [l for l in m.layers if hasattr(l, "weights")][0]
    dw = conv.d_weights
    print(f"{tag}: conv dw vs oracle {nrel(dw, rconv['dw']):.3e} db {nrel(conv.d_bias, rconv['db']):.3e}"
          f"  vs previous run {0.0 if prev is None else nrel(dw, prev):.3e}  |dw|max {np.abs(dw).max():.3e}")
    prev = dw
run("fresh 1"); run("fresh 2")
nn.set_math("tf32")
m4 = build_cfg(4, 1024)
rng = np.random.default_rng(0)
m4.train_on_batch(rng.random((256, 1024)).astype(np.float32), np.eye(10, dtype=np.float32)[rng.integers(0, 10, 256)])
nn.set_math("fp32")
run("after tf32 cfg4"); run("again")
m1 = build_cfg(1)
x1, y1 = O.synthetic_batch(1, 32)
m1.train_on_batch(x1, y1)
run("after cfg1"); run("again")
