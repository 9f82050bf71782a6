"""Diagnostic (GPU): per-layer error of one train step in math mode tf32 vs the fp64 oracle, next to numpy emulations of
TF32 operand truncation / round-to-nearest applied to every Dense contraction of the oracle. Tells numerics from bugs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import nnl_oracle as O
import neuralnetlib_b200 as nn
from test_gpu_models import build_cfg


def trunc(a):
    return (np.asarray(a, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)


def rnd(a):
    u = np.asarray(a, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32).astype(np.float64)


def nrel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a.reshape(b.shape) - b).max() / max(np.abs(b).max(), 1e-300)


def oracle_step(cfg, width, x, y, q=None):
    specs, _ = O.config_specs(cfg)
    if width:
        for sp in specs:
            if sp.get("units") == 4096:
                sp["units"] = width
    f0, b0 = O.dense_fwd, O.dense_bwd
    if q is not None:
        O.dense_fwd = lambda x_, w_, b_: q(x_) @ q(w_) + b_
        O.dense_bwd = lambda x_, w_, e_: (q(e_) @ q(w_).T, q(x_).T @ q(e_), e_.sum(axis=0, keepdims=True))
    try:
        ref = O.OracleSequential(specs, "cce", O.OracleAdam(), seed=42)
        loss = ref.train_on_batch(x, y)
    finally:
        O.dense_fwd, O.dense_bwd = f0, b0
    return ref, loss


for cfg, batch, width in [(4, 256, 1024), (1, 32, None), (2, 256, None)]:
    if width:
        rng = np.random.default_rng(0)
        x = rng.random((batch, width)).astype(np.float32)
        y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, batch)]
    else:
        x, y = O.synthetic_batch(cfg, batch)
    ref, rloss = oracle_step(cfg, width, x, y)
    rt, tloss = oracle_step(cfg, width, x, y, trunc)
    rr, nloss = oracle_step(cfg, width, x, y, rnd)
    print(f"== cfg{cfg} batch {batch} width {width}: oracle loss {rloss:.8f} trunc-emul {tloss:.8f} round-emul {nloss:.8f}")
    for mode in ("fp32", "tf32"):
        nn.set_math(mode)
        m = build_cfg(cfg, width)
        loss = m.train_on_batch(x, y)
        print(f"  [{mode}] loss {loss:.8f}  pred nrel {nrel(m.predictions, ref.predictions):.2e}")
        mine = [l for l in m.layers if hasattr(l, "weights")]
        for k, (layer, rl, tl, nl) in enumerate(zip(mine, [r for r in ref.layers if "w" in r], [r for r in rt.layers if "w" in r],
                                                    [r for r in rr.layers if "w" in r])):
            print(f"    layer {k}: dw vs oracle {nrel(layer.d_weights, rl['dw']):.2e}  vs trunc-emul {nrel(layer.d_weights, tl['dw']):.2e}"
                  f"  vs round-emul {nrel(layer.d_weights, nl['dw']):.2e} | emul-vs-oracle trunc {nrel(tl['dw'], rl['dw']):.2e} round {nrel(nl['dw'], rl['dw']):.2e}"
                  f" | db {nrel(layer.d_bias, rl['db']):.2e}")
    nn.set_math("fp32")
