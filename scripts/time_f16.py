"""Dense 8192x4096x4096 forward / dgrad / wgrad in math mode tf32x3 with and without the fp16-plane path (B200NN_F16X3=0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neuralnetlib_b200 import _backend as B, _ops as ops
dev = B.device()
M, N, K = 8192, 4096, 4096
g = torch.Generator(device="cpu"); g.manual_seed(0)
rnd = lambda *s: (torch.rand(*s, generator=g) - 0.5).to(dev)
x, w, b, y, err = rnd(M, K), rnd(K, N), rnd(1, N), torch.empty(M, N, device=dev), rnd(M, N)
dx, dw, db = torch.empty(M, K, device=dev), torch.empty(K, N, device=dev), torch.empty(1, N, device=dev)
mc = ops.MATH_TF32X3
ws = torch.empty(max(ops.dense_ws_bytes(M, N, K, mc) // 4, 1), device=dev)
st = B.stream_obj()
def t(fn, reps=5):
    fn(); fn()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(reps): fn()
    e.record(st); e.synchronize()
    return a.elapsed_time(e) / reps
fl = 2.0 * M * N * K
for name, fn in (("fwd", lambda: ops.dense_fwd(mc, x, w, b, y, ops.ACT_RELU, 0.0, ws)),
                 ("dgrad", lambda: ops.dense_bwd(mc, x, w, err, None, (), dx, None, None, ops.ACT_RELU, 0.0, None, None, ws)),
                 ("wgrad", lambda: ops.dense_bwd(mc, x, w, err, None, (), None, dw, db, ops.ACT_NONE, 0.0, None, None, ws))):
    ms = t(fn)
    print(f"{name}: {ms*1e3:.0f} us  {fl/ms/1e9:.0f} TFLOP/s-equivalent  (ws {ws.numel()*4/2**20:.0f} MiB)")
ref = (x.double() @ w.double() + b.double()).clamp_min(0)
ops.dense_fwd(mc, x, w, b, y, ops.ACT_RELU, 0.0, ws); torch.cuda.synchronize()
print("fwd rel err", float((y.double() - ref).abs().max() / ref.abs().max()))
