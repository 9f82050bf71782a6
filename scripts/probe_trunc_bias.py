"""Is the tensor core's truncating accumulate a (nearly) uniform relative shrink of the result? Fit y_gpu ~ (1 + a) y_ref per GEMM
and report a, the error before and after removing it (tf32x3 in-kernel split vs fp16-plane path)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neuralnetlib_b200 import _backend as B, _ops as ops
dev = B.device()
g = torch.Generator(device="cpu"); g.manual_seed(0)
mc = ops.MATH_TF32X3
for kind in ("relu_x", "signed_x"):
    for (M, N, K) in ((2048, 2048, 1024), (2048, 2048, 2048), (2048, 2048, 4096), (1024, 4096, 8192)):
        x = torch.randn(M, K, generator=g)
        if kind == "relu_x": x = x.clamp_min(0)
        w = torch.randn(K, N, generator=g) / K ** 0.5
        xd, wd = x.to(dev), w.to(dev)
        y = torch.empty(M, N, device=dev)
        ws = torch.empty(max(ops.dense_ws_bytes(M, N, K, mc) // 4, 1), device=dev)
        ops.dense_fwd(mc, xd, wd, None, y, ops.ACT_NONE, 0.0, ws); torch.cuda.synchronize()
        ref = (xd.double() @ wd.double())
        yd = y.double()
        a = float((yd * ref).sum() / (ref * ref).sum()) - 1.0
        e0 = float((yd - ref).abs().max() / ref.abs().max())
        e1 = float((yd / (1 + a) - ref).abs().max() / ref.abs().max())
        print(f"{kind} M{M} N{N} K{K} f16={os.environ.get('B200NN_F16X3','1')}: a={a:.3e} a/K={a/K:.3e} err={e0:.2e} after_unscale={e1:.2e}", flush=True)
