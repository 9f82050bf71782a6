"""Where does one GAN.train_on_batch (cfg5, batch 512) spend its time? Per-program, per-op CUDA-event intervals."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import torch.distributed as dist
import bench
from neuralnetlib_b200 import _backend as B
B.load_library()
T = bench.Timer(torch, dist, 1, B)
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 512
gan = bench.build_gan()
real = np.random.default_rng(0).random((bs, 28, 28, 1)).astype(np.float32)
for _ in range(4):
    out = gan.train_on_batch(real, None, bs, 1)
ms, out = T.time_steps(lambda: gan.train_on_batch(real, None, bs, 1), 10)
t0 = time.perf_counter()
for _ in range(10):
    gan.train_on_batch(real, None, bs, 1)
torch.cuda.synchronize()
print(f"GAN b{bs}: {ms:.3f} ms/step (events), {(time.perf_counter() - t0) * 100:.3f} ms/step (wall); losses {out}")
for nm, net in (("G", gan.generator), ("D", gan.discriminator)):
    for key, plan in net._plans.items():
        for kind, prog in plan.programs.items():
            per = T.per_op(prog, 3)
            tot = sum(per) * 1e3
            print(f"  {nm} plan {key} program {kind}: {len(prog.ops)} ops, {tot:.1f} us eager; graph={'yes' if prog.graph is not None else 'no'}")
            rows = sorted(zip(per, [m[0] for m in prog.meta]), reverse=True)[:6]
            print("      " + ", ".join(f"{n}={t * 1e3:.0f}" for t, n in rows))
