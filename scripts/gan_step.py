"""A few GAN.train_on_batch steps at batch 512 (for the ncu launch list)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import bench
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 512
gan = bench.build_gan()
real = np.random.default_rng(0).random((bs, 28, 28, 1)).astype(np.float32)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 4):
    out = gan.train_on_batch(real, None, bs, 1)
print(out)
