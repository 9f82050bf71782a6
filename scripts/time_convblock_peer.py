"""Timing (1 GPU): first-block kernels mode 0 vs mode 3 (peer-fused, world of one) -- the cost of the in-kernel exchange protocol."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from neuralnetlib_b200 import _backend as B, _ops as ops, parallel as P
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29534")
dist.init_process_group("gloo", rank=0, world_size=1)
comm = P.Communicator(use_peer="force")
Bn, H, W, C, F = 256, 28, 28, 1, 32
rng = np.random.default_rng(0)
dev = lambda a: B.to_device(np.asarray(a, dtype=np.float32))
xd, wd, bd = dev(rng.random((Bn, H, W, C))), dev(rng.normal(size=(3, 3, C, F)) * 0.3), dev(rng.normal(size=(F,)) * 0.1)
Pn = H * W * F
gd, btd = dev(np.ones(Pn)), dev(np.zeros(Pn))
g = B.ConvGeom(Bn, H, W, C, F, 3, 3, 1, 1, 1, 1, H, W)
ed = dev(rng.normal(size=(Bn, H // 2, W // 2, F)))
ws = B.empty((max(ops.convblock_ws_bytes(g) // 4, 1),))
nb = ops.convblock_region_bytes(g, 1)
pf, pb = (comm.handle,) + comm.region(nb), (comm.handle,) + comm.region(nb)
rm, rv, sm, si = B.zeros((Pn,)), B.to_device(np.ones(Pn)), B.empty((Pn,)), B.empty((Pn,))
yp = B.empty((Bn, H // 2, W // 2, F))
ss = B.zeros((16,), dtype=torch.float64)
dg, dbt, dw, db = B.empty((Pn,)), B.empty((Pn,)), B.empty((3, 3, C, F)), B.empty((F,))
st = B.stream_obj()
def fwd(mode, peer): ops.convblock_fwd(g, xd, wd, bd, 1, 0.0, gd, btd, rm, rv, sm, si, None, mode, Bn, yp, 0.9, 1e-5, peer)
def bwd(mode, peer): ops.convblock_bwd(g, xd, wd, bd, 1, 0.0, gd, btd, sm, si, ed, ss, (), dg.data_ptr(), dbt.data_ptr(), None, None, None, mode, Bn, 1, 2, 3, (), ws, peer)
def timeit(fn, n=30):
    for _ in range(5): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(n): fn()
    b.record(st); b.synchronize()
    return a.elapsed_time(b) / n * 1e3
print("fwd mode0 %.1f us   mode3(world 1) %.1f us" % (timeit(lambda: fwd(0, None)), timeit(lambda: fwd(3, pf))))
print("bwd mode0 %.1f us   mode3(world 1) %.1f us" % (timeit(lambda: bwd(0, None)), timeit(lambda: bwd(3, pb))))
print("timeouts", comm.timeouts())
os._exit(0)
