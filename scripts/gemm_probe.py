"""Runs the tensor-core contractions of bench.py's `gemm` table once each (for ncu / quick timing).
    python scripts/gemm_probe.py [tf32|tf32x3] [reps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import torch.distributed as dist
import bench
from neuralnetlib_b200 import _backend as B
B.load_library()
T = bench.Timer(torch, dist, 1, B)
maths = tuple(sys.argv[1].split(",")) if len(sys.argv) > 1 else ("tf32",)
out = bench.gemm_microbench(T, float(sys.argv[2]) if len(sys.argv) > 2 else 700.0, maths)
for m, rows in out.items():
    for k, v in rows.items():
        print(f"{m:7s} {k:36s} {v['us']:9.1f} us {v['tflops']:7.1f} TFLOP/s")
