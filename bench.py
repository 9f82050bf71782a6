#!/usr/bin/env python
"""bench.py — training samples/s (fwd + bwd + clips + Adam) through neuralnetlib_b200, and GEMM TFLOP/s vs peak.

    python bench.py --gpus N --steps K --warmup W                    # this repo's arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

Headline (`value`, `e2e`, `roofline`): BASELINE.json configs[1] — one `Sequential.train_on_batch` of the CNN
Input(28,28,1)-Conv2D(32,3,same)+ReLU-BatchNorm-MaxPool(2)-Flatten-Dense(64)+ReLU-Dense(10)+Softmax, CCE, Adam, batch 256
per GPU (weak scaling), synthetic MNIST-shaped data. One JSON line is printed by rank 0:

  value     samples/s with the batch already resident in HBM (the captured step graph, CUDA events per step,
            L2 flushed between steps by a 256 MiB write outside the timed events, max over ranks);
  e2e       the same metric through the public API with HOST buffers: pinned-host -> device copy of x and y and the
            device -> host read of the loss inside the timed region of every step;
  roofline  the dominant kernel of the step (bound named per kernel), against MEASURED_PEAKS.json;
  cpu_baseline  the numpy fp64 oracle port of the reference path (oracle/nnl_oracle.py) timed on this box's host cores
            (rank 0, N = 1 only, a bounded number of steps);
  tf32_peak the yardstick for the tensor-pipe fractions, measured IN THIS RUN: cuBLAS TF32 (torch.matmul, allow_tf32) at
            8192^3, best of 10 (burst) and a back-to-back loop (sustained); MEASURED_PEAKS.json's bf16 figures beside it;
  configs   the other BASELINE.json configs with the same timing rules — cfg1 (MLP b32), cfg3 (deep CNN b1024), cfg4 (wide MLP
            b8192), cfg5 (GAN b512) — each with ms_per_step, samples/s, e2e, math mode, a roofline (algorithmic FLOPs of
            SURVEY.md §8d ÷ step time ÷ tf32_peak, algorithmic bytes ÷ step time ÷ HBM peak) and its kernel breakdown. Under
            --gpus N > 1 cfg3 / cfg4 / cfg5 run as STRONG-scaling legs (global batch 1024 / 8192 / 512 sharded over the ranks);
  gemm      the Dense 8192x4096x4096 and cfg3 conv2 / conv3 contractions (fwd, dgrad, wgrad) timed one by one through the C ABI:
            us, TFLOP/s, fraction of tf32_peak, per math mode.
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

# algorithmic FLOPs per sample (2MNK per GEMM, fwd + dW + dX, first layer's dX excluded) and algorithmic bytes
# per sample (fp32, perfect fusion inside a layer) of the BASELINE configs; Adam = 28 B / parameter / step
FLOPS_PER_SAMPLE = {1: 454_400, 2: 3_315_456, 3: 233_816_064, 4: 369_344_512, 5: 82_272_640}
PARAMS = {1: 109_386, 2: 402_442, 3: 411_786, 4: 67_166_218, 5: 299_521 + 420_481}
BYTES_PER_SAMPLE = {3: 908_288 * 4}
LEG_NAMES = {1: "cfg1 MLP 784-128-64-10 relu/softmax, Adam, batch 32",
             3: "cfg3 deep CNN 3x[Conv2D(64/128/256,3,same)+ReLU+BN+MaxPool2]+Flatten+Dense10, 32x32x3, Adam, global batch 1024",
             4: "cfg4 wide MLP 4xDense(4096)+ReLU+Dense10, Adam, global batch 8192",
             5: "cfg5 conv GAN (G: Dense6272-Reshape-ConvT64s2-ConvT32s2-Conv1 sigmoid; D: Conv32s2-Conv64s2-Dense128-Dense1), "
                "28x28x1, BCE, Adam x2, batch 512, n_critic 1"}
LEG_BATCH = {1: 32, 3: 1024, 4: 8192, 5: 512}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons of one GPU, read in-process through NVML by a background thread: every 2 ms while a timed
    region runs (Timer.time_steps switches it to `fast`), every 20 ms otherwise. The timing loop itself is never delayed (an inline
    NVML read on one rank would skew the ranks of a data-parallel step against each other). Falls back to `nvidia-smi -lms 20`."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nvml = self.handle = None
        self.fast = False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def _sample(self):
        n, h = self.nvml, self.handle
        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        flag = lambda bit: "Active" if r & bit else "Not Active"   # noqa: E731
        self.rows.append([str(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)), str(self.max_sm),
                          str(n.nvmlDeviceGetPowerUsage(h) / 1000.0), flag(0x8), flag(0x40), flag(0x20), flag(0x4)])

    def sample_inline(self):
        if self.nvml is not None:
            try:
                self._sample()
            except Exception:
                pass

    def _poll(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._stop.wait(0.002 if self.fast else 0.02)

    def start(self):
        if self.nvml is not None:
            threading.Thread(target=self._poll, daemon=True).start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        return len(self.rows)

    def summary(self, lo=0, hi=None):
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[lo:hi]:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        self._stop.set()
        if self.proc is None:
            return
        time.sleep(0.25)
        self.proc.terminate()


# ---------------------------------------------------------------------------------------------------------------
# reference arm: the oracle port of the reference path on the host cores
def _oracle_model(cfg):
    import nnl_oracle as O
    specs, _ = O.config_specs(cfg)
    return O.OracleSequential(specs, "cce", O.OracleAdam(), seed=42)


def oracle_steps_per_second(cfg, steps, warmup, batch, budget_s=45.0):
    """The reference's algorithm for this path (numpy fp64 oracle port) on the host cores. At most `steps` train steps are timed,
    fewer when they would not fit `budget_s` seconds of CPU work (the sample is bounded, the rate is what is reported)."""
    import nnl_oracle as O
    x, y = O.synthetic_batch(cfg, batch)
    m = _oracle_model(cfg)
    t_w = time.perf_counter()
    n_w = max(1, min(warmup, 2))
    for _ in range(n_w):
        m.train_on_batch(x, y)
    est = (time.perf_counter() - t_w) / n_w
    steps = max(1, min(steps, int(budget_s / max(est, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(steps):
        m.train_on_batch(x, y)
    dt = time.perf_counter() - t0
    return steps / dt, dt, steps


def oracle_gan_steps_per_second(batch, budget_s=30.0):
    """One GAN.train_on_batch of the oracle port at a REDUCED batch (the reference's Conv2DTranspose loops make B=512 take
    minutes per step); the per-sample rate is what is reported, with the batch stated."""
    import nnl_oracle as O
    gs, ds = O.gan_specs()
    G = O.OracleSequential(gs, "bce", O.OracleAdam(), seed=7)
    D = O.OracleSequential(ds, "bce", O.OracleAdam(), seed=7)
    real = np.random.default_rng(0).random((batch, 28, 28, 1))
    O.gan_train_on_batch(G, D, real, batch, 32, 7)
    t0 = time.perf_counter()
    n = 0
    while n < 1 or time.perf_counter() - t0 < budget_s * 0.5:
        O.gan_train_on_batch(G, D, real, batch, 32, 7)
        n += 1
        if n >= 3:
            break
    dt = time.perf_counter() - t0
    return n / dt, dt, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count()
    sps, dt, n_timed = oracle_steps_per_second(CFG, args.steps, args.warmup, BATCH_PER_GPU)
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
    if not args.headline_only:
        # the other BASELINE configs on the same host cores, each a bounded sample (the CPU path needs 15-20 s per step at
        # cfg3 / cfg4 size): 1 warm-up + as many steps as fit the budget
        legs = {}
        for cfg, budget in ((1, 3.0), (3, 25.0), (4, 20.0)):
            try:
                s, d, n = oracle_steps_per_second(cfg, 50 if cfg == 1 else 2, 1, LEG_BATCH[cfg], budget)
                legs[f"cfg{cfg}"] = {"workload": LEG_NAMES[cfg], "value": s * LEG_BATCH[cfg], "unit": "samples/s", "ms_per_step": 1e3 / s,
                                     "sample": f"{n} step(s) at batch {LEG_BATCH[cfg]} after 1 warm-up ({d:.1f} s)"}
            except Exception as e:   # noqa: BLE001
                legs[f"cfg{cfg}"] = {"error": repr(e)}
        try:
            gb = 32
            s, d, n = oracle_gan_steps_per_second(gb)
            legs["cfg5"] = {"workload": LEG_NAMES[5], "value": s * gb, "unit": "samples/s", "ms_per_step": 1e3 / s,
                            "sample": f"{n} GAN.train_on_batch step(s) at batch {gb} (NOT 512: the port's per-sample cost is flat in the "
                                      f"batch; B=512 takes minutes per step) ({d:.1f} s)"}
        except Exception as e:   # noqa: BLE001
            legs["cfg5"] = {"error": repr(e)}
        line["configs"] = legs
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------
def build_model(cfg=2):
    from neuralnetlib_b200.layers import BatchNormalization, Conv2D, Dense, Flatten, Input, MaxPooling2D
    from neuralnetlib_b200.models import Sequential
    from neuralnetlib_b200.optimizers import Adam
    m = Sequential(random_state=42)
    if cfg == 1:
        m.add(Input(784)); m.add(Dense(128, activation="relu")); m.add(Dense(64, activation="relu"))
        m.add(Dense(10, activation="softmax"))
    elif cfg == 2:
        m.add(Input((28, 28, 1)))
        m.add(Conv2D(32, 3, padding="same", activation="relu"))
        m.add(BatchNormalization())
        m.add(MaxPooling2D(2))
        m.add(Flatten())
        m.add(Dense(64, activation="relu"))
        m.add(Dense(10, activation="softmax"))
    elif cfg == 3:
        m.add(Input((32, 32, 3)))
        for f in (64, 128, 256):
            m.add(Conv2D(f, 3, padding="same", activation="relu")); m.add(BatchNormalization()); m.add(MaxPooling2D(2))
        m.add(Flatten()); m.add(Dense(10, activation="softmax"))
    elif cfg == 4:
        m.add(Input(4096))
        for _ in range(4):
            m.add(Dense(4096, activation="relu"))
        m.add(Dense(10, activation="softmax"))
    else:
        raise ValueError(cfg)
    m.compile(loss_function="categorical_crossentropy", optimizer=Adam())
    return m


def build_gan():
    """The notebook GAN of SURVEY.md §8(d) (gan-mnist-convolutional.ipynb cells 3-5)."""
    from neuralnetlib_b200.layers import Conv2D, Conv2DTranspose, Dense, Flatten, Input, Reshape
    from neuralnetlib_b200.models import GAN, Sequential
    from neuralnetlib_b200.optimizers import Adam
    G = Sequential(random_state=7)
    G.add(Input(32)); G.add(Dense(7 * 7 * 128)); G.add(Reshape((7, 7, 128)))
    G.add(Conv2DTranspose(64, 3, strides=2, padding="same", activation="relu"))
    G.add(Conv2DTranspose(32, 3, strides=2, padding="same", activation="relu"))
    G.add(Conv2D(1, 3, padding="same", activation="sigmoid"))
    D = Sequential(random_state=7)
    D.add(Input((28, 28, 1)))
    D.add(Conv2D(32, 3, strides=2, padding="same", activation="relu"))
    D.add(Conv2D(64, 3, strides=2, padding="same", activation="relu"))
    D.add(Flatten()); D.add(Dense(128, activation="relu")); D.add(Dense(1, activation="sigmoid"))
    gan = GAN(latent_dim=32, random_state=7)
    gan.compile(G, D, Adam(), Adam(), "bce")
    return gan


class Timer:
    """CUDA-event timing of K steps on the launching stream, L2 flushed between steps (outside the events), max over ranks."""

    def __init__(self, torch, dist, world, B):
        self.torch, self.dist, self.world, self.B = torch, dist, world, B
        self.sampler = None                   # ClockSampler on rank 0: sampled between the timed steps
        self.stream = B.stream_obj()
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=B.device())

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce_max(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.B.device())
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def time_steps(self, fn, K):
        """mean ms per call of fn over K calls."""
        torch = self.torch
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        self.barrier()
        out = None
        if self.sampler is not None:
            self.sampler.fast = True          # 2 ms NVML polling (background thread) while the timed steps run
        for s, e in ev:
            self.flush.zero_()                # evict L2 (126 MB) between steps; outside the timed events
            s.record(self.stream)
            out = fn()
            e.record(self.stream)
        if self.sampler is not None:
            self.sampler.sample_inline()      # every step is enqueued, the device is still running them: one guaranteed sample under load
        self.barrier()
        if self.sampler is not None:
            self.sampler.fast = False
        return self.reduce_max(sum(s.elapsed_time(e) for s, e in ev)) / K, out

    def per_op(self, prog, reps):
        """The same step run op by op with an event pair around every launch, behind ~0.5 ms of queued device work so that the
        host has enqueued every op before the device reaches it. These intervals include the gaps between dependent launches,
        so their sum exceeds the graph step: each is an UPPER bound of the kernel's in-graph time."""
        torch = self.torch
        acc = [0.0] * len(prog.ops)
        for _ in range(reps):
            for _ in range(12):
                self.flush.zero_()
            evs = []
            for fn in prog.ops:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(self.stream); fn(); b.record(self.stream)
                evs.append((a, b))
            torch.cuda.synchronize()
            for i, (a, b) in enumerate(evs):
                acc[i] += a.elapsed_time(b) / reps
        return acc


def measure_tf32_peak(torch):
    """cuBLAS TF32 GEMM 8192^3 (torch.matmul with allow_tf32): the yardstick the tensor-pipe fractions are quoted against. A library
    call used ONLY as a measuring stick, never on the product path."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
        c = torch.empty(n, n, device="cuda")
        for _ in range(3):
            torch.matmul(a, b, out=c)
        best = 1e9
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); torch.matmul(a, b, out=c); e.record(); e.synchronize()
            best = min(best, s.elapsed_time(e))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 60
        s.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e.record(); e.synchronize()
        sus = s.elapsed_time(e) / reps
        fl = 2.0 * n ** 3
        del a, b, c
        return {"burst_tflops": fl / (best * 1e-3) / 1e12, "sustained_tflops": fl / (sus * 1e-3) / 1e12,
                "how": "torch.matmul fp32 with allow_tf32 (cuBLAS TF32) 8192^3: best of 10 / 60 back to back; yardstick only"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def run_leg(cfg, math, T, K, W, world, rank, peaks, tf32_peak, clock_sampler):
    """One BASELINE config through the public API: resident (captured step graph) and end-to-end timings + kernel breakdown."""
    import torch
    import neuralnetlib_b200 as nn
    from neuralnetlib_b200 import _backend as B
    import nnl_oracle as O
    lib = B.lib()
    nn.set_math(math)
    gb = LEG_BATCH[cfg]
    if world > 1 and gb % world:
        return {"skipped": f"global batch {gb} does not divide over {world} ranks"}
    lb = gb // world
    mark = clock_sampler.mark() if clock_sampler is not None else 0
    res = {"workload": LEG_NAMES[cfg], "math": math, "global_batch": gb, "per_gpu_batch": lb,
           "scaling": "strong" if world > 1 else "single GPU"}
    if cfg == 5:
        gan = build_gan()
        if world > 1:
            from neuralnetlib_b200 import parallel
            if not hasattr(parallel, "make_gan_data_parallel"):
                return {"skipped": "GAN under data parallelism is not wired"}
            parallel.make_gan_data_parallel(gan)
        real = np.random.default_rng(0).random((gb, 28, 28, 1)).astype(np.float32)
        real_l = real   # every rank passes the GLOBAL batch; GAN.train_on_batch keeps this rank's rows of the seeded index choice
        for _ in range(W):
            out = gan.train_on_batch(real, None, gb, 1)
        l0 = lib.b200nn_launch_count()
        e2e_ms, out = T.time_steps(lambda: gan.train_on_batch(real, None, gb, 1), K)
        launches = (lib.b200nn_launch_count() - l0) // K
        res.update({"ms_per_step": e2e_ms, "value": gb / (e2e_ms * 1e-3), "unit": "samples/s", "kernels_per_step": int(launches),
                    "e2e": {"value": gb / (e2e_ms * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms,
                            "h2d_bytes_per_step": int(lb * 784 * 4 + lb * 32 * 4 * 2 + 3 * lb * 4), "d2h_bytes_per_step": 16,
                            "last_losses": [float(out[0]), float(out[1])]},
                    "note": "GAN.train_on_batch through the public API (host noise + real batch uploaded, two losses read back every "
                            "step): value == e2e, there is no separate resident replay for the alternating G/D step"})
        fl = FLOPS_PER_SAMPLE[5] * gb
        res["roofline"] = {"bound": "launch/latency", "step_gflop": fl / 1e9, "achieved_tflops": fl / (e2e_ms * 1e-3) / 1e12,
                           "frac_of_tf32_peak": fl / (e2e_ms * 1e-3) / 1e12 / tf32_peak, "us_per_kernel": e2e_ms * 1e3 / max(launches, 1)}
        gan.generator._drop_plans(); gan.discriminator._drop_plans()
        return res
    model = build_model(cfg)
    if world > 1:
        from neuralnetlib_b200 import parallel
        parallel.make_data_parallel(model)
    x, y = O.synthetic_batch(cfg, gb, seed=0)
    x_pin = torch.from_numpy(np.ascontiguousarray(x[rank * lb:(rank + 1) * lb])).pin_memory()
    y_pin = torch.from_numpy(np.ascontiguousarray(y[rank * lb:(rank + 1) * lb])).pin_memory()
    x_np, y_np = x_pin.numpy(), y_pin.numpy()
    for _ in range(W):
        model.train_on_batch(x_np, y_np)
    plan = model._last_plan
    prog = plan.program("train")
    assert prog.graph is not None, "train step was not captured into a CUDA graph"
    resident_ms, _ = T.time_steps(prog.run, K)
    e2e_ms, loss = T.time_steps(lambda: model.train_on_batch(x_np, y_np), K)
    l0 = lib.b200nn_launch_count()
    per_op = T.per_op(prog, 3)
    launches = (lib.b200nn_launch_count() - l0) // 3
    names = [n for n, _ in prog.meta]
    breakdown = sorted(((names[i], round(per_op[i] * 1e3, 1), prog.meta[i][1],
                         round(prog.meta[i][1] / (per_op[i] * 1e-3) / 1e9, 1) if per_op[i] > 0 else None) for i in range(len(per_op))),
                       key=lambda t: -t[1])
    fl = FLOPS_PER_SAMPLE[cfg] * gb
    tfl = fl / (resident_ms * 1e-3) / 1e12
    res.update({"ms_per_step": resident_ms, "value": gb / (resident_ms * 1e-3), "unit": "samples/s", "kernels_per_step": int(launches),
                "e2e": {"value": gb / (e2e_ms * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(x_pin.numel() * 4 + y_pin.numel() * 4), "d2h_bytes_per_step": 8, "last_loss": loss},
                "kernel_breakdown": {"columns": ["op", "us (eager upper bound)", "io bytes", "GB/s"],
                                     "rows": breakdown if os.environ.get("B200NN_BENCH_ALLOPS") else breakdown[:10]}})
    roof = {"step_gflop": fl / 1e9, "achieved_tflops": tfl / world, "tf32_peak_tflops": tf32_peak,
            "frac_of_tf32_peak": (tfl / world / tf32_peak) if tf32_peak else None}
    if cfg in BYTES_PER_SAMPLE:
        by = BYTES_PER_SAMPLE[cfg] * gb + PARAMS[cfg] * 28 * world
        roof.update({"step_algorithmic_bytes": by, "achieved_gbs": by / world / (resident_ms * 1e-3) / 1e9,
                     "frac_of_hbm_peak": by / world / (resident_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]})
    if cfg == 1:
        roof["bound"] = "launch/latency"
        roof["us_per_graph_node"] = resident_ms * 1e3 / max(launches, 1)
        roof["note_latency"] = ("one cooperative kernel per step (csrc/mlp_step.cu): 7 phases separated by grid barriers; a captured graph of "
                                "ONE trivial kernel measures 6.2 us the same way (scripts/mlp_step_phases.py), the kernel's own phases sum to ~33 us")
    elif cfg == 3:
        # floors: 239 GFLOP at the TF32 yardstick vs 3.72 GB at the HBM copy rate -> the step is HBM-bound
        roof["bound"] = "hbm"
        roof["frac"] = roof["frac_of_hbm_peak"]
    else:
        roof["bound"] = "tensor"
        roof["frac"] = roof["frac_of_tf32_peak"]
    roof["note"] = "per-GPU rates: FLOPs (bytes) of the global step / ranks / step time"
    res["roofline"] = roof
    if clock_sampler is not None:
        res["clocks"] = clock_sampler.summary(mark)
    model._drop_plans()
    del model, plan, prog
    torch.cuda.empty_cache()
    return res


def gemm_microbench(T, tf32_peak, maths=("tf32", "tf32x3")):
    """The tensor-core contractions one by one through the C ABI (ops.dense_* / ops.conv2d_*): CUDA events around each call
    (a call = the GEMM launch plus, where the path has them, its split-K reduce / bias-gradient launches), operands larger
    than L2 or L2 flushed between calls."""
    import torch
    from neuralnetlib_b200 import _backend as B
    from neuralnetlib_b200 import _ops as ops
    dev = B.device()
    out = {}
    g = torch.Generator(device="cpu"); g.manual_seed(0)

    def rnd(*shape):
        return (torch.rand(*shape, generator=g) - 0.5).to(dev)

    def timeit(fn, reps=int(os.environ.get("B200NN_PROBE_REPS", "5"))):
        if reps > 1:
            fn(); fn()
        ms, _ = T.time_steps(fn, reps)
        return ms

    for math in maths:
        mc = {"fp32": ops.MATH_FP32, "tf32": ops.MATH_TF32, "tf32x3": ops.MATH_TF32X3}[math]
        rows = {}
        M, N, K = 8192, 4096, 4096
        x, w, b, y, err = rnd(M, K), rnd(K, N), rnd(1, N), torch.empty(M, N, device=dev), rnd(M, N)
        dx, dw, db = torch.empty(M, K, device=dev), torch.empty(K, N, device=dev), torch.empty(1, N, device=dev)
        ws = torch.empty(max(ops.dense_ws_bytes(M, N, K, mc) // 4, 1), device=dev)
        fl = 2.0 * M * N * K
        for name, fn in (("dense_fwd", lambda: ops.dense_fwd(mc, x, w, b, y, ops.ACT_RELU, 0.0, ws)),
                         ("dense_dgrad", lambda: ops.dense_bwd(mc, x, w, err, None, (), dx, None, None, ops.ACT_RELU, 0.0, None, None, ws)),
                         ("dense_wgrad", lambda: ops.dense_bwd(mc, x, w, err, None, (), None, dw, db, ops.ACT_NONE, 0.0, None, None, ws))):
            ms = timeit(fn)
            rows[f"{name} 8192x4096x4096"] = {"us": ms * 1e3, "tflops": fl / (ms * 1e-3) / 1e12, "frac_of_tf32_peak": fl / (ms * 1e-3) / 1e12 / tf32_peak}
        del x, w, b, y, err, dx, dw, db, ws
        for cname, (Hc, Cc, Fc) in (("conv2 16x16x64->128", (16, 64, 128)), ("conv3 8x8x128->256", (8, 128, 256))):
            Bn = 1024
            geo = B.ConvGeom(Bn, Hc, Hc, Cc, Fc, 3, 3, 1, 1, 1, 1, Hc, Hc)
            x, w, b = rnd(Bn, Hc, Hc, Cc), rnd(3, 3, Cc, Fc), rnd(1, Fc)
            y, err = torch.empty(Bn, Hc, Hc, Fc, device=dev), rnd(Bn, Hc, Hc, Fc)
            dx, dw, db = torch.empty_like(x), torch.empty_like(w), torch.empty(Fc, device=dev)
            ws = torch.empty(max(ops.conv2d_ws_bytes(geo, mc) // 4, 1), device=dev)
            fl = 2.0 * Bn * Hc * Hc * Fc * 9 * Cc
            for name, fn in (("fwd", lambda: ops.conv2d_fwd(mc, geo, x, w, b, y, ops.ACT_RELU, 0.0, ws)),
                             ("dgrad", lambda: ops.conv2d_bwd(mc, geo, x, w, err, None, (), dx, None, None, ops.ACT_NONE, 0.0, None, None, ws)),
                             ("wgrad", lambda: ops.conv2d_bwd(mc, geo, x, w, err, None, (), None, dw, db, ops.ACT_NONE, 0.0, None, None, ws))):
                ms = timeit(fn)
                rows[f"{cname} b1024 {name}"] = {"us": ms * 1e3, "tflops": fl / (ms * 1e-3) / 1e12,
                                                 "frac_of_tf32_peak": fl / (ms * 1e-3) / 1e12 / tf32_peak}
            del x, w, b, y, err, dx, dw, db, ws
        out[math] = rows
        torch.cuda.empty_cache()
    return out


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

    math = args.math or nn.get_math()
    nn.set_math(math)
    lib = B.load_library()
    stream = B.stream_obj()
    T = Timer(torch, dist, world, B)
    model = build_model(2)
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

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        T.sampler = sampler

    # ---- (1) resident: K replays of the captured step graph, inputs already in HBM -------------------------------
    launches0 = lib.b200nn_launch_count()
    resident_ms, _ = T.time_steps(prog.run, K)
    launches_resident = lib.b200nn_launch_count() - launches0   # graph replays add their kernel nodes (b200nn_graph_launch)
    # ---- (2) end to end through the public API: pinned host batch -> device, step, loss -> host ------------------
    e2e_ms, loss = T.time_steps(lambda: model.train_on_batch(x_np, y_np), K)
    clocks = sampler.summary() if rank == 0 else None
    launches_timed = lib.b200nn_launch_count() - launches0   # graph replays are counted per kernel node by the C ABI

    # ---- (3) per-kernel times: the same step run op by op with an event pair around every launch ------------------
    names = [n for n, _ in prog.meta]
    reps = max(3, min(K, 20))
    l1 = lib.b200nn_launch_count()
    per_op = T.per_op(prog, reps)
    launches_eager_step = (lib.b200nn_launch_count() - l1) // reps if reps else 0
    top = max(range(len(per_op)), key=lambda i: per_op[i])
    top_name, top_bytes = prog.meta[top]
    peaks, peak_src = load_peaks()
    peak = float(peaks["hbm_gbs"])
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
    # what bounds the dominant kernel: the first-block kernels recompute the convolution from the input image instead of storing
    # it (DRAM traffic 16x below the algorithmic bytes), so they are fp32-issue bound, not HBM bound; `frac` compares them with a
    # perfectly fused per-layer implementation streaming the algorithmic bytes at the HBM copy rate
    issue_bound = top_name.startswith("convblock")
    dp_note = None
    if world > 1:
        dp_note = ("exact global-batch semantics: NCCL gradient sum + "
                   + ("one-shot NVLink peer-memory exchanges" if model._dp.comm.peer else "NCCL")
                   + " for clip norms / BatchNorm moments, all inside the step graph")

    total_batch = BATCH_PER_GPU * world
    line = {
        "metric": METRIC, "value": total_batch / (resident_ms * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": resident_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": total_batch, "parallelism": f"dp{world}",
                   "l2": "flushed between steps (256 MiB device write outside the timed events)",
                   "math": {"tf32": "tf32: eligible Dense / Conv2D contractions on tcgen05 kind::tf32 (rel 2e-3 per contraction), rest fp32",
                            "tf32x3": "tf32x3: eligible Dense / Conv2D contractions on tcgen05 with hi/lo operand split (rel 1e-5), rest fp32 FFMA",
                            }.get(math, "fp32 FFMA everywhere (parity path, rel 1e-5)"),
                   "step": "one captured CUDA graph per step",
                   **({"data_parallel": dp_note} if world > 1 else {})},
        "e2e": {"value": total_batch / (e2e_ms * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(x.nbytes + y.nbytes), "d2h_bytes_per_step": 8, "last_loss": loss},
        "gpu_launches": int(launches_resident if launches_resident > 0 else launches_eager_step * K),
        "kernels_per_step": int(launches_eager_step),
        "roofline": {"bound": "fp32 issue (compared against the HBM roofline of a per-layer implementation)" if issue_bound else "hbm",
                     "kernel": top_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": top_bytes, "kernel_us": per_op[top] * 1e3,
                     "kernel_us_note": "eager event interval behind queued work: an upper bound of the in-graph kernel time, so frac is a lower bound",
                     "algorithmic_bytes_definition": "SURVEY.md 8(d) per-sample figure x 256 samples per launch (conv block bwd: (4S+3P)*4 B, "
                                                     "fwd: (2S+2P)*4 B); the kernel itself touches only io_bytes_per_launch because the conv "
                                                     "output is recomputed from the input image instead of being stored (traffic = ncu DRAM bytes)",
                     "io_bytes_per_launch": io_bytes,
                     "step_algorithmic_bytes": step_bytes, "step_io_bytes": io_step_bytes,
                     "step_frac_of_hbm_roofline": (step_bytes / (resident_ms * 1e-3) / 1e9) / peak},
        "kernel_breakdown_us": breakdown[:8],
        "clocks": clocks,
    }
    _ = launches_timed
    model._drop_plans()
    del model, plan, prog
    torch.cuda.empty_cache()

    cfg1_leg = None
    if not args.headline_only and world == 1:
        # config 1 (a 45 us latency-bound step) is timed BEFORE the 8192^3 yardstick GEMMs: those leave the board in its power-capped
        # clock state for a while, which would be charged to the smallest workload
        try:
            cfg1_leg = run_leg(1, math, T, max(3, min(K, 20)), max(3, min(W, 5)), world, rank, load_peaks()[0], None, sampler if rank == 0 else None)
        except Exception as e:   # noqa: BLE001
            import traceback
            cfg1_leg = {"error": repr(e), "trace": traceback.format_exc()[-600:]}
    if not args.headline_only:
        try:
            tfp = measure_tf32_peak(torch)
        except Exception as e:   # noqa: BLE001
            tfp = {"burst_tflops": float(peaks.get("bf16_tflops", 1590.0)) / 2, "sustained_tflops": float(peaks.get("bf16_tflops_sustained", 1400.0)) / 2,
                   "how": f"cuBLAS TF32 measurement failed ({e!r}); half of the bf16 figures"}
        tfp.update({"bf16_tflops_measured_peaks": peaks.get("bf16_tflops"), "bf16_tflops_sustained_measured_peaks": peaks.get("bf16_tflops_sustained")})
        line["tf32_peak"] = tfp
        tf32_peak = float(tfp["burst_tflops"])
        Kl, Wl = max(3, min(K, 20)), max(3, min(W, 5))
        legs = {}
        plan_legs = [(1, math)] if world == 1 else []
        plan_legs += [(3, math), (3, "tf32"), (4, math), (4, "tf32"), (5, math)]
        seen = set()
        for cfg, m_ in plan_legs:
            key = f"cfg{cfg}_{m_}"
            if key in seen:
                continue
            seen.add(key)
            if cfg == 1 and cfg1_leg is not None:
                r = cfg1_leg
                if "roofline" in r:
                    r["roofline"]["tf32_peak_tflops"] = tf32_peak
                    r["roofline"]["frac_of_tf32_peak"] = r["roofline"]["achieved_tflops"] / tf32_peak
                r["steps"], r["warmup"] = Kl, Wl
                legs[key] = r
                continue
            try:
                r = run_leg(cfg, m_, T, Kl if cfg != 4 else min(Kl, 10), Wl, world, rank, peaks, tf32_peak, sampler if rank == 0 else None)
            except Exception as e:   # noqa: BLE001  (a failing leg must not lose the headline)
                import traceback
                r = {"error": repr(e), "trace": traceback.format_exc()[-600:]}
            r["steps"], r["warmup"] = (Kl if cfg != 4 else min(Kl, 10)), Wl
            legs[key] = r
        nn.set_math(math)
        line["configs"] = legs
        if world == 1:
            try:
                line["gemm"] = gemm_microbench(T, tf32_peak)
                line["gemm"]["note"] = "fractions against tf32_peak.burst_tflops (each call is timed alone)"
            except Exception as e:   # noqa: BLE001
                line["gemm"] = {"error": repr(e)}
    if rank == 0:
        sampler.stop()

    if world > 1:
        dist.barrier()
    if rank != 0:
        sys.stdout.flush()
        _dp_teardown(dist)
        return 0

    if world == 1 and not args.no_cpu_baseline:
        budget_steps = 24
        sps, dt, budget_steps = oracle_steps_per_second(CFG, budget_steps, 1, BATCH_PER_GPU, 20.0)
        line["cpu_baseline"] = {"value": sps * BATCH_PER_GPU, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{budget_steps} train_on_batch steps of the same workload at batch {BATCH_PER_GPU} "
                                          f"after 1 warm-up ({dt:.1f} s of CPU work; numpy fp64 oracle port of the reference path)"}
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        _dp_teardown(dist)
    return 0


def _dp_teardown(dist):
    """Orderly exit of a data-parallel rank: plans / graphs were dropped per leg, so the communicators can be destroyed and the process
    group closed. A watchdog ends the process if NCCL's destroy does not return within 20 s (the line is already printed)."""
    watchdog = threading.Timer(20.0, lambda: os._exit(0))
    watchdog.daemon = True
    watchdog.start()
    try:
        from neuralnetlib_b200 import parallel
        parallel.shutdown()
        dist.destroy_process_group()
    except Exception as e:   # noqa: BLE001
        print(f"[bench] teardown: {e!r}", file=sys.stderr)
    watchdog.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="only the cfg2 headline leg (no configs / gemm / tf32_peak objects)")
    ap.add_argument("--math", default=os.environ.get("B200NN_MATH"), choices=["fp32", "tf32", "tf32x3"],
                    help="contractions: fp32 = FFMA everywhere (rel 1e-5); tf32 = tcgen05 tensor cores (rel 2e-3 per contraction); "
                         "tf32x3 = tcgen05 with error compensation (rel 1e-5) — the package default")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
