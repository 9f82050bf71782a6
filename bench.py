#!/usr/bin/env python
"""bench.py — training samples/s (fwd + bwd + clips + Adam) of BASELINE config 2 through neuralnetlib_b200.

    python bench.py --gpus N --steps K --warmup W              # this repo's arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

A "step" is one `Sequential.train_on_batch` of the CNN  Input(28,28,1)-Conv2D(32,3,same)+ReLU-BatchNorm-MaxPool(2)-
Flatten-Dense(64)+ReLU-Dense(10)+Softmax, CCE, Adam, batch 256 per GPU, synthetic MNIST-shaped data
(BASELINE.json configs[1]). One JSON line is printed by rank 0:

  value     samples/s with the batch already resident in HBM (the captured step graph, CUDA events per step,
            L2 flushed between steps by a 256 MiB write outside the timed events, max over ranks);
  e2e       the same metric through the public API with HOST buffers: pinned-host -> device copy of x and y and the
            device -> host read of the loss inside the timed region of every step;
  roofline  the dominant kernel of the step: algorithmic bytes per launch / its mean CUDA-event duration, against the
            measured HBM copy bandwidth in MEASURED_PEAKS.json;
  cpu_baseline  the numpy fp64 oracle port of the reference path (oracle/nnl_oracle.py) timed on this box's host cores
            (rank 0, N = 1 only, a bounded number of steps).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"   # NCCL prints its version banner to STDOUT; rank 0 must print exactly one JSON line

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

BATCH_PER_GPU = 256
CFG = 2
WORKLOAD = "cfg2 CNN Conv2D(32,3,same)+ReLU+BatchNorm+MaxPool2+Flatten+Dense64+ReLU+Dense10+Softmax, 28x28x1, CCE, Adam, batch 256/GPU"
METRIC = "training samples/sec (fwd+bwd+Adam)"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons every 20 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
def oracle_steps_per_second(steps, warmup, batch, budget_s=45.0):
    """The reference's algorithm for this path (numpy fp64 oracle port) on the host cores. At most `steps` train steps are timed,
    fewer when they would not fit `budget_s` seconds of CPU work (the sample is bounded, the rate is what is reported)."""
    import nnl_oracle as O
    x, y = O.synthetic_batch(CFG, batch)
    specs, _ = O.config_specs(CFG)
    m = O.OracleSequential(specs, "cce", O.OracleAdam(), seed=42)
    t_w = time.perf_counter()
    n_w = max(1, min(warmup, 2))
    for _ in range(n_w):
        m.train_on_batch(x, y)
    est = (time.perf_counter() - t_w) / n_w
    steps = max(3, min(steps, int(budget_s / max(est, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(steps):
        m.train_on_batch(x, y)
    dt = time.perf_counter() - t0
    return steps / dt, dt, steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count()
    sps, dt, n_timed = oracle_steps_per_second(args.steps, args.warmup, BATCH_PER_GPU)
    value = sps * BATCH_PER_GPU
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / sps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": BATCH_PER_GPU, "note": "CPU path does not shard; one process on the host"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{n_timed} train_on_batch steps at batch {BATCH_PER_GPU} ({dt:.1f} s of CPU work, bounded to ~45 s) "
                                   f"after warm-up (numpy {np.__version__} fp64 oracle port of the reference path, OPENBLAS threads = all cores)"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------
def build_model():
    from neuralnetlib_b200.layers import BatchNormalization, Conv2D, Dense, Flatten, Input, MaxPooling2D
    from neuralnetlib_b200.models import Sequential
    from neuralnetlib_b200.optimizers import Adam
    m = Sequential(random_state=42)
    m.add(Input((28, 28, 1)))
    m.add(Conv2D(32, 3, padding="same", activation="relu"))
    m.add(BatchNormalization())
    m.add(MaxPooling2D(2))
    m.add(Flatten())
    m.add(Dense(64, activation="relu"))
    m.add(Dense(10, activation="softmax"))
    m.compile(loss_function="categorical_crossentropy", optimizer=Adam())
    return m


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)

    import neuralnetlib_b200 as nn
    from neuralnetlib_b200 import _backend as B
    import nnl_oracle as O  # synthetic data recipe only; the oracle is never on the timed path

    nn.set_math(args.math)
    lib = B.load_library()
    stream = B.stream_obj()
    model = build_model()
    if world > 1:
        from neuralnetlib_b200 import parallel
        parallel.make_data_parallel(model)
    x, y = O.synthetic_batch(CFG, BATCH_PER_GPU, seed=rank)
    x_pin = torch.from_numpy(x).pin_memory()
    y_pin = torch.from_numpy(y).pin_memory()
    x_np, y_np = x_pin.numpy(), y_pin.numpy()

    W, K = max(args.warmup, 3), args.steps
    for _ in range(W):
        model.train_on_batch(x_np, y_np)
    plan = model._last_plan
    prog = plan.program("train")
    assert prog.graph is not None, "train step was not captured into a CUDA graph"

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=B.device())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=B.device())
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- (1) resident: K replays of the captured step graph, inputs already in HBM -------------------------------
    launches0 = lib.b200nn_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    for s, e in ev:
        flush.zero_()                     # evict L2 (126 MB) between steps; outside the timed events
        s.record(stream)
        prog.run()
        e.record(stream)
    barrier()
    step_ms = [s.elapsed_time(e) for s, e in ev]
    resident_ms = reduce_max(sum(step_ms)) / K
    kernels_per_step = getattr(plan, "kernels_per_step", None)

    # ---- (2) end to end through the public API: pinned host batch -> device, step, loss -> host ------------------
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    for s, e in ev2:
        flush.zero_()
        s.record(stream)
        loss = model.train_on_batch(x_np, y_np)
        e.record(stream)
    barrier()
    e2e_ms = reduce_max(sum(s.elapsed_time(e) for s, e in ev2)) / K
    clocks = sampler.stop() if rank == 0 else None

    # ---- (3) per-kernel times: the same step run op by op with an event pair around every launch ------------------
    names = [n for n, _ in prog.meta]
    per_op = [0.0] * len(prog.ops)
    reps = max(3, min(K, 20))
    for _ in range(reps):
        for _ in range(12):
            flush.zero_()   # ~0.5 ms of device work queued first, so that the host enqueues every op (and its event pair) before the
                            # device reaches it: the event intervals then measure kernels, not Python launch latency
        evs = []
        for fn in prog.ops:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream)
            evs.append((a, b))
        torch.cuda.synchronize()
        for i, (a, b) in enumerate(evs):
            per_op[i] += a.elapsed_time(b) / reps
    launches_eager_step = (lib.b200nn_launch_count() - launches0) // reps if reps else 0
    top = max(range(len(per_op)), key=lambda i: per_op[i])
    top_name, top_bytes = prog.meta[top]
    peak, peak_src = load_peaks()
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(top_name)
    except Exception:
        pass
    # ALGORITHMIC bytes per SURVEY.md 8(d) (fp32, perfect fusion inside a layer, tensors materialised when saved for backward or
    # separated by a batch-wide reduction; S = conv-output, P = pooled elements per sample): conv block forward 2S + 2P, backward of the
    # first block 4S + 3P. The fused first-block kernels cover exactly those layer groups; every other kernel's figure is the
    # tensors it reads / writes once (Plan.emit io lists), which is the survey's definition for Dense / optimizer kernels.
    S_el, P_el = 28 * 28 * 32, 14 * 14 * 32
    survey_bytes = {"convblock_fwd": (2 * S_el + 2 * P_el) * 4 * BATCH_PER_GPU, "convblock_fwd_peer": (2 * S_el + 2 * P_el) * 4 * BATCH_PER_GPU,
                    "convblock_bwd": (4 * S_el + 3 * P_el) * 4 * BATCH_PER_GPU, "convblock_bwd_peer": (4 * S_el + 3 * P_el) * 4 * BATCH_PER_GPU}
    io_bytes = top_bytes
    top_bytes = survey_bytes.get(top_name, top_bytes)
    achieved = top_bytes / (per_op[top] * 1e-3) / 1e9 if per_op[top] > 0 else 0.0
    step_bytes = (6 * S_el + 5 * P_el) * 4 * BATCH_PER_GPU + sum(b for n_, b in prog.meta if n_.startswith("optimizer"))
    io_step_bytes = sum(b for _, b in prog.meta)
    breakdown = sorted(((names[i], round(per_op[i] * 1e3, 2), prog.meta[i][1]) for i in range(len(per_op))), key=lambda t: -t[1])
    if os.environ.get("B200NN_BENCH_ALLOPS") and rank == 0:
        print("[bench] per-op us (program order): " + ", ".join(f"{names[i]}={per_op[i] * 1e3:.1f}" for i in range(len(per_op)))
              + f" | sum={sum(per_op) * 1e3:.1f}", file=sys.stderr)

    if world > 1:
        dist.barrier()
    if rank != 0:
        sys.stdout.flush()
        os._exit(0)   # the captured step graphs still hold the NCCL communicator: let process exit reclaim it

    total_batch = BATCH_PER_GPU * world
    line = {
        "metric": METRIC, "value": total_batch / (resident_ms * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": resident_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": total_batch, "parallelism": f"dp{world}",
                   "l2": "flushed between steps (256 MiB device write outside the timed events)",
                   "math": {"tf32": "tf32: Dense GEMMs on tcgen05 kind::tf32 (rel 2e-3), conv/BN/pool/Adam fp32",
                            "tf32x3": "tf32x3: Dense GEMMs on tcgen05 with hi/lo operand split (rel 1e-5), conv/BN/pool/Adam fp32",
                            }.get(args.math, "fp32 FFMA everywhere (parity path, rel 1e-5)"),
                   "step": "one captured CUDA graph per step",
                   **({"data_parallel": "exact global-batch semantics: NCCL gradient sum + "
                       + ("one-shot NVLink peer-memory exchanges" if model._dp.comm.peer else "NCCL")
                       + " for clip norms / BatchNorm moments, all inside the step graph"} if world > 1 else {})},
        "e2e": {"value": total_batch / (e2e_ms * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(x.nbytes + y.nbytes), "d2h_bytes_per_step": 8, "last_loss": loss},
        "gpu_launches": int(launches_eager_step * K),
        "kernels_per_step": int(launches_eager_step),
        "roofline": {"bound": "hbm", "kernel": top_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": top_bytes, "kernel_us": per_op[top] * 1e3,
                     "algorithmic_bytes_definition": "SURVEY.md 8(d) per-sample figure x 256 samples per launch (conv block bwd: (4S+3P)*4 B, "
                                                     "fwd: (2S+2P)*4 B); the kernel itself touches only io_bytes_per_launch because the conv "
                                                     "output is recomputed from the input image instead of being stored (traffic = ncu DRAM bytes)",
                     "io_bytes_per_launch": io_bytes,
                     "step_algorithmic_bytes": step_bytes, "step_io_bytes": io_step_bytes,
                     "step_frac_of_hbm_roofline": (step_bytes / (resident_ms * 1e-3) / 1e9) / peak},
        "kernel_breakdown_us": breakdown[:8],
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        budget_steps = 24
        sps, dt, budget_steps = oracle_steps_per_second(budget_steps, 1, BATCH_PER_GPU, 20.0)
        line["cpu_baseline"] = {"value": sps * BATCH_PER_GPU, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{budget_steps} train_on_batch steps of the same workload at batch {BATCH_PER_GPU} "
                                          f"after 1 warm-up ({dt:.1f} s of CPU work; numpy fp64 oracle port of the reference path)"}
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        os._exit(0)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--math", default=os.environ.get("B200NN_MATH", "fp32"), choices=["fp32", "tf32", "tf32x3"],
                    help="Dense contractions: fp32 = FFMA parity path (rel 1e-5); tf32 = tcgen05 tensor cores (rel 2e-3); "
                         "tf32x3 = tcgen05 with error compensation (rel 1e-5)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
