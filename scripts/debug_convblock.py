import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import nnl_oracle as O
from neuralnetlib_b200 import _backend as B, _ops as ops
Bn,H,W,C,F,act = 16,6,4,1,8,"linear"
rng = np.random.default_rng(Bn * 7 + H + C)
x = rng.random((Bn, H, W, C)).astype(np.float32)
w = (rng.normal(size=(3, 3, C, F)) * 0.3).astype(np.float32)
b = (rng.normal(size=(1, F)) * 0.1).astype(np.float32)
P = H*W*F
gamma = (1 + 0.1 * rng.normal(size=P)).astype(np.float32); beta = (0.1 * rng.normal(size=P)).astype(np.float32)
g = B.ConvGeom(Bn, H, W, C, F, 3, 3, 1, 1, 1, 1, H, W)
dev = lambda a: B.to_device(np.asarray(a, dtype=np.float32))
xd, wd, bd, gd, btd = dev(x), dev(w), dev(b.reshape(-1)), dev(gamma), dev(beta)
rm, rv = B.zeros((P,)), B.to_device(np.ones(P)); sm, si = B.empty((P,)), B.empty((P,)); yp = B.empty((Bn, H//2, W//2, F))
ops.convblock_fwd(g, xd, wd, bd, 0, 0.01, gd, btd, rm, rv, sm, si, None, 0, Bn, yp, 0.9, 1e-5)
e = rng.normal(size=(Bn, H//2, W//2, F)).astype(np.float32)
ed = dev(e); ss = B.zeros((16,), dtype=torch.float64); ops.sumsq(ed, ss, 0)
ws = B.zeros((max(ops.convblock_ws_bytes(g)//4,1),))
dg, dbt, dw, db, r = B.empty((P,)), B.empty((P,)), B.zeros((3,3,C,F)), B.zeros((F,)), B.empty((Bn,H,W,F))
ops.convblock_bwd(g, xd, wd, bd, 0, 0.01, gd, btd, sm, si, ed, ss, (0,), dg.data_ptr(), dbt.data_ptr(), dw, db, r, 0, Bn, 1, 2, None, (1,2), ws)
B.synchronize()
print("ss", ss.cpu().numpy()[:5]); print("ws", ws.cpu().numpy().reshape(-1, 10, F)[0]); print("dw", dw.cpu().numpy().reshape(9, F)); print("db", db.cpu().numpy())
