import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neuralnetlib_b200 import _backend as B, _ops as ops
np.set_printoptions(linewidth=250, precision=1, suppress=True)
M, N, K = 128, 64, 32
def run(x, w):
    xd, wd = B.to_device(x), B.to_device(w)
    bd = B.zeros((N,)); y = B.empty((M, N))
    ws = B.empty((max(ops.dense_ws_bytes(M, N, K) // 4, 1),))
    ops.dense_fwd(1, xd, wd, bd, y, 0, 0.0, ws)
    B.synchronize()
    return y.cpu().numpy()
# test 1: D[m,n] = m
x = np.zeros((M, K), np.float32); x[:, 0] = np.arange(M)
w = np.zeros((K, N), np.float32); w[0, :] = 1
d = run(x, w)
print("T1 col0 (expect 0..127):", d[:, 0][:40]); print("T1 row5 (expect all 5):", d[5, :16])
# test 2: D[m,n] = (m%32)*100 + n
x = np.zeros((M, K), np.float32); x[np.arange(M), np.arange(M) % 32] = 1
w = (np.arange(K)[:, None] * 100 + np.arange(N)[None, :]).astype(np.float32)
d = run(x, w)
print("T2 row0 (expect 0..63):", d[0, :40]); print("T2 row3 (expect 300..):", d[3, :12]); print("T2 col1 (expect k*100+1):", d[:36, 1])
ref = x @ w
print("max err", np.abs(d - ref).max())
