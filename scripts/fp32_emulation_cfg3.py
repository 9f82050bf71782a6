"""How far does a float32 evaluation of the reference's OWN algorithm sit from its float64 evaluation, for one step of config 3 at
batch 1024? Runs the NumPy oracle twice (float64; float32 parameters / activations, NumPy's pairwise float32 sums) and prints,
per weighted layer, the norm-relative deviation of the gradients — the floor any fp32 implementation (ours included) shares.
CPU only (~1 min):   python scripts/fp32_emulation_cfg3.py [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import nnl_oracle as O

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x, y = O.synthetic_batch(3, batch)
specs, _ = O.config_specs(3)
ref = O.OracleSequential(specs, "cce", O.OracleAdam(), seed=42)
ref.train_on_batch(x, y)

emu = O.OracleSequential([dict(s) for s in specs], "cce", O.OracleAdam(), seed=42)
_asarray = np.asarray


def asarray32(a, dtype=None, **kw):     # OracleSequential.forward / train_on_batch cast their inputs to float64: keep float32
    return _asarray(a, dtype=np.float32 if dtype is np.float64 else dtype, **kw)


# initialise lazily-created parameters with one tiny float64 forward, then cast everything to float32
emu.forward(x[:2].astype(np.float64), training=False)
for L in emu.layers:
    for k, v in list(L.items()):
        if isinstance(v, np.ndarray) and v.dtype == np.float64:
            L[k] = v.astype(np.float32)
O.np.asarray = asarray32
try:
    emu.train_on_batch(x.astype(np.float32), y.astype(np.float32))
finally:
    O.np.asarray = _asarray
for i, (a, b) in enumerate(zip(emu.layers, ref.layers)):
    if "dw" in b:
        for k in ("dw", "db"):
            d = np.abs(np.asarray(a[k], np.float64).reshape(b[k].shape) - b[k])
            sc = np.abs(b[k]).max()
            print(f"layer {i:2d} {b['type']:7s} {k}: dtype {np.asarray(a[k]).dtype}, max |d| / max|ref| = {d.max() / sc:.2e}, "
                  f"elements beyond 2e-5: {int(np.sum(d > 2e-5 * sc))} of {d.size}")
