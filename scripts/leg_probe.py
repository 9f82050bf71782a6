"""One bench leg with the full per-op table:  python scripts/leg_probe.py <cfg> <math> [steps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
os.environ["B200NN_BENCH_ALLOPS"] = "1"
import torch
import torch.distributed as dist
import bench
from neuralnetlib_b200 import _backend as B
B.load_library()
T = bench.Timer(torch, dist, 1, B)
cfg, math = int(sys.argv[1]), sys.argv[2]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
peaks, _ = bench.load_peaks()
r = bench.run_leg(cfg, math, T, steps, 3, 1, 0, peaks, 700.0, None)
print(f"cfg{cfg} {math}: {r['ms_per_step']:.4f} ms/step resident, {r['e2e']['ms_per_step']:.4f} ms e2e, {r['value']:.0f} samples/s, kernels {r.get('kernels_per_step')}")
if "kernel_breakdown" in r:
    tot = sum(x[1] for x in r["kernel_breakdown"]["rows"])
    for row in r["kernel_breakdown"]["rows"]:
        print(f"   {row[0]:28s} {row[1]:9.1f} us  {row[2] / 1e6:9.1f} MB  {row[3] if row[3] is not None else 0:8.1f} GB/s")
    print(f"   sum of eager op intervals {tot:.1f} us")
