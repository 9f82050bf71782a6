import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neuralnetlib_b200 import _backend as B, _ops as ops
def trunc(a):
    return (a.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)
def rnd(a):
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32).astype(np.float64)
def nrel(a, b): return np.abs(a - b).max() / np.abs(b).max()
rng = np.random.default_rng(0)
M, N, K = 256, 1024, 1024
x = np.maximum(rng.normal(size=(M, K)), 0).astype(np.float32)
w = (rng.normal(size=(K, N)) / 32).astype(np.float32)
e = rng.normal(size=(M, N)).astype(np.float32)
xd, wd, ed, bd = B.to_device(x), B.to_device(w), B.to_device(e), B.zeros((N,))
ws = B.empty((ops.dense_ws_bytes(M, N, K) // 4 + 1,))
y = B.empty((M, N)); dx = B.empty((M, K)); dw = B.empty((K, N)); db = B.empty((N,))
ops.dense_fwd(1, xd, wd, bd, y, 0, 0.0, ws)
ops.dense_bwd(1, xd, wd, ed, None, (), dx, dw, db, 0, 0.0, None, None, ws)
B.synchronize()
y, dx, dw = (t.cpu().numpy().astype(np.float64) for t in (y, dx, dw))
x64, w64, e64 = x.astype(np.float64), w.astype(np.float64), e.astype(np.float64)
print("fwd: vs exact %.2e  vs trunc-emul %.2e  vs round-emul %.2e" % (nrel(y, x64 @ w64), nrel(y, trunc(x) @ trunc(w)), nrel(y, rnd(x) @ rnd(w))))
print("dw : vs exact %.2e  vs trunc-emul %.2e  vs round-emul %.2e" % (nrel(dw, x64.T @ e64), nrel(dw, trunc(x).T @ trunc(e)), nrel(dw, rnd(x).T @ rnd(e))))
print("dx : vs exact %.2e  vs trunc-emul %.2e  vs round-emul %.2e" % (nrel(dx, e64 @ w64.T), nrel(dx, trunc(e) @ trunc(w).T), nrel(dx, rnd(e) @ rnd(w).T)))
print("round-emul vs exact: fwd %.2e dw %.2e dx %.2e" % (nrel(rnd(x) @ rnd(w), x64 @ w64), nrel(rnd(x).T @ rnd(e), x64.T @ e64), nrel(rnd(e) @ rnd(w).T, e64 @ w64.T)))
