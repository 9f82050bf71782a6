"""Time the Dense GEMMs (forward + backward) of config 4's hidden layer in the three math modes (CUDA events, L2 flushed)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuralnetlib_b200 import _backend as B, _ops as ops

M, N, K = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (256, 4096, 4096)))
B.stream_obj()
dev = B.device()
x = torch.rand(M, K, device=dev); w = torch.randn(K, N, device=dev) / K ** 0.5; b = torch.zeros(N, device=dev)
err = torch.randn(M, N, device=dev); y = torch.empty(M, N, device=dev)
dx, dw, db = torch.empty(M, K, device=dev), torch.empty(K, N, device=dev), torch.empty(N, device=dev)
ws = torch.empty(max(ops.dense_ws_bytes(M, N, K, 2) // 4, 1), device=dev)
ss = torch.zeros(16, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flops = 2.0 * M * N * K
for math, name in ((0, "fp32"), (1, "tf32"), (2, "tf32x3")):
    with torch.cuda.stream(B.stream_obj()):
        for what in ("fwd", "bwd"):
            ts = []
            for it in range(8):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if what == "fwd":
                    ops.dense_fwd(math, x, w, b, y, 1, 0.0, ws)
                else:
                    ops.dense_bwd(math, x, w, err, ss, (), dx, dw, db, 1, 0.0, 1, 2, ws)
                e1.record(); e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = sorted(ts[2:])[len(ts[2:]) // 2]
            print(f"{name:7s} {what} M={M} N={N} K={K}: {t*1e3:8.1f} us  {flops * (1 if what == 'fwd' else 2) / t / 1e9:7.1f} TFLOP/s")
