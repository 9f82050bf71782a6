"""cfg4 at several sizes: L2-relative deviation of every Dense gradient from the fp64 oracle, per math mode."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import neuralnetlib_b200 as nn
import nnl_oracle as O
from test_gpu_models import build_cfg
for width, batch in ((1024, 256), (4096, 256), (4096, 2048)):
    specs, _ = O.config_specs(4)
    for sp in specs:
        if sp.get("units") == 4096:
            sp["units"] = width
    rng = np.random.default_rng(0)
    x = rng.random((batch, width)).astype(np.float32)
    y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, batch)]
    ref = O.OracleSequential(specs, "cce", O.OracleAdam(), seed=42)
    rl = ref.train_on_batch(x, y)
    for math in ("fp32", "tf32x3"):
        nn.set_math(math)
        m = build_cfg(4, width)
        l = m.train_on_batch(x, y)
        devs = []
        for layer, r in zip([L for L in m.layers if hasattr(L, "weights")], [r for r in ref.layers if "w" in r]):
            a, b = np.asarray(layer.d_weights, np.float64), r["dw"]
            devs.append(float(np.sqrt(np.sum((a - b) ** 2) / np.sum(b ** 2))))
        print(f"width {width} batch {batch} {math}: loss {l:.8f} vs {rl:.8f}; L2-rel dw per layer (first..last): " + " ".join(f"{d:.1e}" for d in devs), flush=True)
