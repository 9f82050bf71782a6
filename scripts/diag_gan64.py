import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import neuralnetlib_b200 as nn
from bench import build_gan
from helpers import check_packed
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "gan_b64.npz")))
for math in sys.argv[1:] or ["tf32x3", "fp32"]:
    nn.set_math(math)
    gan = build_gan()
    real = g["real"].astype(np.float32)
    dl, gl = gan.train_on_batch(real, batch_size=int(g["batch"]), n_critic=1)
    print(math, "d_loss", dl, float(g["s0_d_loss"]), "g_loss", gl, float(g["s0_g_loss"]))
    for nm, M in (("G", gan.generator), ("D", gan.discriminator)):
        for li, L in enumerate(M.layers):
            if hasattr(L, "weights"):
                for k, arr in (("dw", L.d_weights), ("db", L.d_bias)):
                    try:
                        e = check_packed(g, f"s0_{nm}{li}_{k}", arr, 1e9)
                    except AssertionError as ex:
                        e = str(ex)[:80]
                    print(f"   {nm}{li} {type(L).__name__:16s} {k}: {e}")
