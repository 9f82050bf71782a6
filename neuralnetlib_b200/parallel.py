"""Data-parallel `Sequential.fit` over the GPUs of one NVLink/NVSwitch box: one process per GPU (torchrun), the batch
sharded by rows, weights / optimizer state / BatchNorm running statistics replicated.

The reference is single-process, and several of its quantities are global over the batch. For the sharded run to
reproduce the single-process result on the GLOBAL batch, the following are summed over the ranks inside every step
(csrc/comm.cu; all captured into the step graph, no host involvement):

  * weight / bias gradients — SUM, not mean: models.py:200-205 never divides the error by the batch (Q4). All gradient
    tensors of a model are views into one flat buffer, reduced with NCCL; the per-tensor clip (models.py:240-242) and
    the optimizer then run on identical data on every rank;
  * the float64 sum-of-squares slots behind every error clip (models.py:214) and the loss / accuracy accumulators;
  * BatchNormalization's per-position moments (layers.py:1237-1238) and its two backward reductions (:1261-1262).

The last two are latency-bound (a few doubles, or a few hundred KB of statistics, on the critical path of the step), so
they go through the one-shot peer-memory exchange when CUDA IPC is available; NCCL otherwise.

Host-side pieces (sharding, flat-buffer layout, bootstrap over torch.distributed) are device-independent and are
covered by world-size-2 `gloo` tests on CPU (tests/test_dp_gloo.py).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _backend as B

ALIGN_ELEMS = 64   # every tensor of the flat gradient buffer starts on a 256-byte boundary (TMA / float4 alignment)


# ---- device-independent host logic ----------------------------------------------------------------------------------
def shard_rows(n_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous row range [lo, hi) of `rank` in a global batch of `n_rows` rows. The global batch must divide evenly:
    every rank runs the same captured graph on the same local shape."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    if n_rows % world:
        raise ValueError(f"global batch of {n_rows} rows is not divisible by the {world} data-parallel ranks")
    per = n_rows // world
    return rank * per, (rank + 1) * per


def flat_layout(sizes, align: int = ALIGN_ELEMS):
    """Offsets (in elements) of tensors of `sizes` elements packed into one flat buffer, each aligned to `align`
    elements; returns (offsets, total)."""
    offs, cur = [], 0
    for n in sizes:
        offs.append(cur)
        cur += (int(n) + align - 1) // align * align
    return offs, cur


def make_buckets(offsets, sizes, total, bucket_elems: int):
    """Split the flat buffer [0, total) into contiguous buckets of about `bucket_elems` elements on tensor boundaries,
    in buffer order (= the order in which backward produces the gradients: last layer first). Returns
    [(lo, hi, last_tensor_index)]: bucket k may be reduced as soon as tensor `last_tensor_index` has been written."""
    buckets, lo = [], 0
    for i, (o, n) in enumerate(zip(offsets, sizes)):
        end = offsets[i + 1] if i + 1 < len(offsets) else total
        if end - lo >= bucket_elems or i == len(offsets) - 1:
            buckets.append((lo, end, i))
            lo = end
    return buckets


def broadcast_bytes(payload: bytes | None, src: int = 0, group=None) -> bytes:
    """Broadcast a small byte string (the NCCL unique id) from `src` through torch.distributed (any backend)."""
    import torch.distributed as dist
    box = [payload]
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]


def all_gather_bytes(payload: bytes, group=None) -> list[bytes]:
    import torch.distributed as dist
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, payload, group=group)
    return out


# ---- communicator -----------------------------------------------------------------------------------------------------
import weakref

_LIVE_COMMS = []     # every Communicator of this process, for shutdown()
_LIVE_MODELS = weakref.WeakSet()   # every model made data parallel (their plans hold graphs that reference a communicator)


def shutdown(models=()):
    """Orderly teardown of the data-parallel state of this process: drop the plans (and with them the captured step graphs, which
    reference the NCCL communicator) of `models`, drain the device, destroy every communicator. Call on every rank, then
    torch.distributed.destroy_process_group()."""
    for m in list(models) + list(_LIVE_MODELS):
        if hasattr(m, "_drop_plans"):
            m._drop_plans()
    torch.cuda.synchronize()
    while _LIVE_COMMS:
        _LIVE_COMMS.pop().destroy()


class Communicator:
    """b200nn_comm (include/b200nn.h): NCCL communicator + optional CUDA-IPC symmetric buffer for one-shot exchanges.
    Bootstrapped over an initialised torch.distributed process group (torchrun)."""

    def __init__(self, group=None, sym_bytes: int | None = None, use_peer: bool | None = None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun and call init_process_group)")
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        lib = B.lib()
        B.stream_obj()
        nccl_path = None
        try:
            import nvidia.nccl
            cand = os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so.2")
            nccl_path = cand if os.path.exists(cand) else None
        except Exception:
            pass
        B.check(lib.b200nn_comm_load_nccl(nccl_path.encode() if nccl_path else None))
        uid = None
        if self.rank == 0:
            buf = C.create_string_buffer(128)
            B.check(lib.b200nn_comm_unique_id(buf))
            uid = buf.raw
        uid = broadcast_bytes(uid, 0, group)
        h = C.c_void_p()
        B.check(lib.b200nn_comm_create(uid, self.rank, self.world, C.byref(h)))
        self.handle = h.value
        self.nccl_version = int(lib.b200nn_comm_nccl_version())
        if use_peer is None:
            use_peer = os.environ.get("B200NN_PEER", "1") != "0"
        self.peer = False
        self._n_sites = 0
        self.peer_max_bytes = int(os.environ.get("B200NN_PEER_MAX_BYTES", str(2 << 20)))
        if use_peer and (self.world > 1 or use_peer == "force"):   # "force": map the symmetric buffer even for a world of one (tests)
            self._open_peer(sym_bytes or int(os.environ.get("B200NN_SYM_BYTES", str(256 << 20))))
        dist.barrier(group)
        _LIVE_COMMS.append(self)

    def _open_peer(self, sym_bytes):
        lib = B.lib()
        hb = int(lib.b200nn_comm_ipc_handle_bytes())
        buf = C.create_string_buffer(hb)
        ok = lib.b200nn_comm_peer_alloc(self.handle, sym_bytes, buf) == 0
        handles = all_gather_bytes(buf.raw if ok else b"", self.group)
        if not all(len(hh) == hb for hh in handles):
            return
        ok = lib.b200nn_comm_peer_open(self.handle, b"".join(handles)) == 0
        oks = all_gather_bytes(b"1" if ok else b"0", self.group)
        self.peer = all(o == b"1" for o in oks)   # all ranks or none: the two paths must never be mixed

    def site(self, nbytes: int) -> int:
        """Reserve a one-shot exchange site for payloads of up to `nbytes`; -1 = use NCCL (no peer memory, or the payload
        is bandwidth-bound). Every rank must reserve the same sites in the same order."""
        if not self.peer or nbytes > self.peer_max_bytes or self._n_sites >= 96:
            return -1
        self._n_sites += 1
        s = C.c_int(-1)
        B.check(B.lib().b200nn_comm_site_create(self.handle, int(nbytes), C.byref(s)))
        return s.value

    def region(self, nbytes: int):
        """A raw region of the symmetric buffer for kernels that exchange over peer memory themselves: (site, offset) or
        None when peer memory is unavailable. Every rank must create the same regions in the same order."""
        if not self.peer or self._n_sites >= 96:
            return None
        s, off = C.c_int(-1), C.c_size_t(0)
        if B.lib().b200nn_comm_region_create(self.handle, int(nbytes), C.byref(s), C.byref(off)) != 0:
            return None
        self._n_sites += 1
        return s.value, off.value

    def allreduce(self, t: torch.Tensor, site: int = -1, stream: int | None = None):
        """In-place sum of a contiguous float32 / float64 device tensor over the ranks, asynchronous on `stream` (a raw
        cudaStream_t; default: the package stream)."""
        if self.world == 1:
            return
        if not t.is_contiguous() or t.dtype not in (torch.float32, torch.float64):
            raise TypeError("allreduce needs a contiguous float32 / float64 device tensor")
        B.check(B.lib().b200nn_comm_allreduce_sum(self.handle, int(site), t.data_ptr(), t.numel(),
                                                  0 if t.dtype == torch.float32 else 1, B.stream() if stream is None else stream))

    def allreduce2(self, a: torch.Tensor, b: torch.Tensor, site: int, stream: int | None = None):
        """One latency exchange for a float32 tensor and a float64 tensor (b200nn_comm_allreduce_sum2)."""
        if self.world == 1:
            return
        B.check(B.lib().b200nn_comm_allreduce_sum2(self.handle, int(site), a.data_ptr(), a.numel(), b.data_ptr(), b.numel(),
                                                   B.stream() if stream is None else stream))

    def timeouts(self) -> int:
        return int(B.lib().b200nn_comm_timeouts(self.handle))

    def destroy(self):
        if getattr(self, "handle", None):
            B.lib().b200nn_comm_destroy(self.handle)
            self.handle = None


# ---- model-level state ---------------------------------------------------------------------------------------------------
class DataParallel:
    """Attached to a Sequential as `model._dp`: the communicator, the flat gradient buffer and its buckets."""

    def __init__(self, comm: Communicator):
        self.comm = comm
        self.rank, self.world = comm.rank, comm.world
        self.flat = None
        self.buckets = None
        self._flat_key = None
        # a bucket is closed once it holds this many bytes: it is then reduced on a side stream while backward continues
        self.bucket_elems = int(os.environ.get("B200NN_BUCKET_BYTES", str(256 << 10))) // 4
        self.bucket_layers = None   # per bucket: the layer whose backward completes it
        self.side = torch.cuda.Stream(device=B.device())

    def flatten_grads(self, layers):
        """Re-point d_weights / d_bias of every parametrised layer (update order: last layer first) at views of one flat
        buffer, so that the gradient exchange is a few large NCCL calls instead of two small ones per layer."""
        params = []
        for L in reversed(layers):
            if hasattr(L, "weights") and L._weights is not None:
                params.append((L, "_d_weights", L._weights.numel()))
                if L._bias is not None:
                    params.append((L, "_d_bias", L._bias.numel()))
        key = tuple((id(L), name, n) for L, name, n in params)
        if key == self._flat_key:
            return
        sizes = [n for _, _, n in params]
        offs, total = flat_layout(sizes)
        self.flat = B.zeros((max(total, 1),))
        for (L, name, n), o in zip(params, offs):
            old = getattr(L, name)
            view = self.flat[o:o + n]
            setattr(L, name, view.view(old.shape) if old is not None and old.numel() == n else view)
        self.buckets = make_buckets(offs, sizes, total, self.bucket_elems)
        self.bucket_layers = [params[k][0] for _, _, k in self.buckets]
        self._flat_key = key


def make_data_parallel(model, comm: Communicator | None = None, group=None):
    """Turn `model` (a compiled Sequential) into a data-parallel replica. Call on every rank after the process group is
    up. `train_on_batch(x, y)` then takes the LOCAL shard of the global batch; `fit(x, y, batch_size=G)` takes the global
    dataset and shards every batch of G rows by `shard_rows`. All ranks must build identical models (same seeds)."""
    if comm is None:
        comm = Communicator(group)
    model._dp = DataParallel(comm) if comm.world > 1 else None
    model._drop_plans()
    _LIVE_MODELS.add(model)
    return model


def make_gan_data_parallel(gan, comm: Communicator | None = None, group=None):
    """Data-parallel `GAN.train_on_batch` / `fit` (models.py:2609-2667 on a sharded batch): generator and discriminator become
    data-parallel replicas sharing ONE communicator, and `train_on_batch(real_samples, None, batch_size=G)` — called with the same
    GLOBAL arguments on every rank — trains on this rank's rows of the globally seeded real-index choice and latent noise. The
    discriminator sees 2 * G / world rows per rank; gradient sums, clip norms and both losses are those of the global batch."""
    if gan.generator is None or gan.discriminator is None:
        raise ValueError("compile the GAN before make_gan_data_parallel")
    if comm is None:
        comm = Communicator(group)
    for net in (gan.generator, gan.discriminator):
        net._dp = DataParallel(comm) if comm.world > 1 else None
        net._drop_plans()
        _LIVE_MODELS.add(net)
    gan._dp = gan.discriminator._dp
    return gan
