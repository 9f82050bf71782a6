"""cfg3 b1024: gradients under tf32x3 vs fp32 (both on the GPU): are the deviations isolated (pool / ReLU flips) or broad?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import neuralnetlib_b200 as nn
import nnl_oracle as O
from test_gpu_models import build_cfg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x, y = O.synthetic_batch(3, B)
out = {}
for math in ("fp32", "tf32x3", "tf32"):
    nn.set_math(math)
    m = build_cfg(3)
    loss = m.train_on_batch(x, y)
    out[math] = (loss, [np.array(L.d_weights) for L in m.layers if hasattr(L, "weights")], np.array(m.predictions))
for math in ("tf32x3", "tf32"):
    print(f"== {math} vs fp32: loss {out[math][0]:.9f} vs {out['fp32'][0]:.9f}; pred nrel {np.abs(out[math][2]-out['fp32'][2]).max()/np.abs(out['fp32'][2]).max():.2e}")
    for i, (a, b) in enumerate(zip(out[math][1], out["fp32"][1])):
        d = np.abs(a - b); sc = np.abs(b).max()
        bad = d > 2e-5 * sc
        F = a.shape[-1]
        per_f = bad.reshape(-1, F).sum(0)
        top = np.argsort(-per_f)[:6]
        print(f"  layer {i}: shape {a.shape} worst {d.max()/sc:.2e} median {np.median(d)/sc:.2e} beyond 2e-5: {bad.sum()} of {d.size}; "
              f"worst filters {[(int(t), int(per_f[t])) for t in top]}")
