"""Data parallelism, host side, on CPU: world-size-2 `gloo` process groups exercise (a) the device-independent pieces of
neuralnetlib_b200/parallel.py (row sharding, flat gradient layout, buckets, the byte-string bootstrap the NCCL id and the
CUDA IPC handles travel through) and (b) the data-parallel SCHEDULE itself: the oracle run on row shards with exactly the
sums the CUDA path exchanges (tests/dp_reference.py) must reproduce the single-process oracle on the global batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import nnl_oracle as O
from dp_reference import ShardedOracle
from neuralnetlib_b200 import parallel as P


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(fn, world, *args):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in out:
        assert r[1] == "ok", r
    return out


def _entry(fn, rank, world, port, q, *args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    try:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        fn(rank, world, *args)
        q.put((rank, "ok"))
    except BaseException as e:  # noqa: BLE001
        import traceback
        q.put((rank, "fail", repr(e), traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def test_shard_rows_and_layout():
    assert [P.shard_rows(256, 4, r) for r in range(4)] == [(0, 64), (64, 128), (128, 192), (192, 256)]
    with pytest.raises(ValueError):
        P.shard_rows(10, 4, 0)
    with pytest.raises(ValueError):
        P.shard_rows(8, 2, 2)
    offs, total = P.flat_layout([10, 640, 6272 * 64, 64, 288, 32])
    assert offs[0] == 0 and all(o % P.ALIGN_ELEMS == 0 for o in offs) and total % P.ALIGN_ELEMS == 0
    assert all(offs[i + 1] >= offs[i] + n for i, n in enumerate([10, 640, 6272 * 64, 64, 288]))
    sizes = [10, 640, 6272 * 64, 64, 288, 32]
    b = P.make_buckets(offs, sizes, total, 1 << 18)
    assert b[0][0] == 0 and b[-1][1] == total and all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
    assert [k for _, _, k in b] == sorted(k for _, _, k in b) and b[-1][2] == len(sizes) - 1
    one = P.make_buckets(offs, sizes, total, 1 << 30)
    assert one == [(0, total, len(sizes) - 1)]


def _bootstrap(rank, world):
    uid = os.urandom(128) if rank == 0 else None
    got = P.broadcast_bytes(uid, 0)
    assert isinstance(got, bytes) and len(got) == 128
    all_ids = P.all_gather_bytes(got)
    assert len(all_ids) == world and all(a == all_ids[0] for a in all_ids)       # every rank received rank 0's id
    handles = P.all_gather_bytes(bytes([rank]) * 64)
    assert [h[0] for h in handles] == list(range(world)) and all(len(h) == 64 for h in handles)


def test_bootstrap_bytes_over_gloo():
    _spawn(_bootstrap, 2)


def _allreduce_np(a):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy().reshape(np.shape(a))


def _sharded_step(rank, world, cfg, batch, steps, loss="cce"):
    x, y = O.synthetic_batch(cfg, batch)
    specs, _ = O.config_specs(cfg)
    if loss == "mse":   # a non-fused (activation, loss) pair: the error entering backward is LossFunction.derivative itself
        specs = specs[:-1] + [{"type": "act", "fn": "tanh"}]
    full = O.OracleSequential(specs, loss, O.OracleAdam(), seed=42)
    lo, hi = P.shard_rows(batch, world, rank)
    mine = ShardedOracle(specs, loss, O.OracleAdam(), 42, _allreduce_np, world)
    for s in range(steps):
        lf = full.train_on_batch(x, y)
        lm = mine.train_on_batch(x[lo:hi], y[lo:hi])
        assert abs(lf - lm) <= 1e-12 * max(1.0, abs(lf)), (s, lf, lm)
        for a, b in zip(full.layers, mine.layers):
            if "w" in a:
                for k in ("dw", "db", "w", "b"):
                    assert np.abs(a[k] - b[k]).max() <= 1e-9 * max(np.abs(a[k]).max(), 1e-30), (s, a["type"], k)
            if a["type"] == "bn":
                for k in ("running_mean", "running_var"):
                    assert np.abs(a[k] - b[k]).max() <= 1e-12, (s, k)
    assert mine.optimizer.t == full.optimizer.t


@pytest.mark.parametrize("cfg,batch,steps", [(1, 32, 3), (2, 16, 2)])
def test_sharded_schedule_reproduces_global_batch(cfg, batch, steps):
    """SUM of gradients (Q4), global clip norms (Q2), global BatchNorm moments (Q5) and backward reductions across two
    ranks == the single-process oracle on the whole batch, step after step (Adam state included)."""
    _spawn(_sharded_step, 2, cfg, batch, steps)


def test_sharded_schedule_mse_uses_global_size():
    """MeanSquaredError.derivative divides by the element count of the WHOLE batch (losses.py:84): a rank that divides by its
    shard's size doubles every gradient on two ranks."""
    _spawn(_sharded_step, 2, 1, 32, 2, "mse")
