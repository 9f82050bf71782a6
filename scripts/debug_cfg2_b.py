import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import nnl_oracle as O
import test_gpu_models as T
from helpers import nrel
T.test_train_on_batch_oracle_full_size(2, 256, 3, 2e-5)
print("test function passed when called from a script")
for order in ("gpu_first", "oracle_first"):
    x, y = O.synthetic_batch(2, 256)
    specs, _ = O.config_specs(2)
    if order == "gpu_first":
        m = T.build_cfg(2)
        ref = O.OracleSequential(specs, "cce", O.OracleAdam(), seed=42)
        l1 = m.train_on_batch(x, y); l2 = ref.train_on_batch(x, y)
    else:
        ref = O.OracleSequential(specs, "cce", O.OracleAdam(), seed=42)
        l2 = ref.train_on_batch(x, y)
        m = T.build_cfg(2)
        l1 = m.train_on_batch(x, y)
    conv = m.layers[1]; rc = ref.layers[1]
    print(order, l1, l2, "dw nrel", nrel(conv.d_weights, rc["dw"]), "w0 equal", nrel(m.layers[5].weights, ref.layers[6]["w"]) if False else "")
    print("  gpu dw[0,0,0,:4]", conv.d_weights[0, 0, 0, :4], " oracle", rc["dw"][0, 0, 0, :4])
