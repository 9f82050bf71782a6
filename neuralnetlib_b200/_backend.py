"""ctypes binding of libb200nn.so (see include/b200nn.h) and the device plumbing around it.

PyTorch is used for device memory, streams and torch.distributed only; every numeric op goes through the
C ABI. There is NO CPU fallback: without the shared library or without a CUDA device the package raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np
import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "_lib", "libb200nn.so")

_lib = None
_lock = threading.Lock()
_stream = None
_device = None

c_void_p, c_int, c_float, c_ll, c_u64, c_size_t = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_uint64, C.c_size_t


class ConvGeom(C.Structure):
    _fields_ = [(n, c_int) for n in ("N", "H", "W", "C", "F", "kh", "kw", "sh", "sw", "ph", "pw", "OH", "OW")]


class ConvTGeom(C.Structure):
    _fields_ = [(n, c_int) for n in ("N", "H", "W", "C", "F", "kh", "kw", "sh", "sw", "pt", "pl", "OH", "OW")]


class PoolGeom(C.Structure):
    _fields_ = [(n, c_int) for n in ("N", "H", "W", "C", "ph", "pw", "sh", "sw", "pad_h", "pad_w", "OH", "OW")]


class MlpLayer(C.Structure):
    """b200nn_mlp_layer (include/b200nn.h)."""
    _fields_ = [(n, c_void_p) for n in ("w", "b", "dw", "db", "out", "err")] + \
               [("K", c_int), ("N", c_int), ("act", c_int), ("alpha", c_float), ("chain", c_u64),
                ("s_out0", c_int), ("s_out1", c_int), ("gi_w", c_int), ("gi_b", c_int)]


MLP_MAX_LAYERS, MLP_MAX_BATCH = 8, 64


class MlpStep(C.Structure):
    """b200nn_mlp_step (include/b200nn.h)."""
    _fields_ = [(n, c_int) for n in ("L", "B", "n_slots", "n_params", "n_layers_opt", "total_blocks")] + \
               [(n, c_void_p) for n in ("x", "y", "ss", "acc", "gss", "hyper", "t_counter", "params", "block_map", "sync")] + \
               [("s0", c_int), ("layer", MlpLayer * MLP_MAX_LAYERS)]


# numpy mirror of b200nn_param (include/b200nn.h)
PARAM_DTYPE = np.dtype([("p", np.uint64), ("g", np.uint64), ("m", np.uint64), ("v", np.uint64), ("n", np.int64),
                        ("t_index", np.int32), ("clip", np.int32), ("chain", np.uint64)], align=True)
OPT_CHUNK = 4096

_P = c_void_p
_SIGNATURES = {
    "b200nn_version": (c_int, []),
    "b200nn_last_error": (c_size_t, [C.c_char_p, c_size_t]),
    "b200nn_has_tcgen05": (c_int, []),
    "b200nn_launch_count": (c_ll, []),
    "b200nn_step_begin": (c_int, [_P, c_int, _P, c_int, _P, c_ll, _P]),
    "b200nn_zero_f64": (c_int, [_P, c_int, _P]),
    "b200nn_gather_rows": (c_int, [_P, _P, _P, c_ll, c_ll, _P]),
    "b200nn_upload": (c_int, [_P, _P, c_size_t, _P]),
    "b200nn_read_sync": (c_int, [_P, _P, c_size_t, _P]),
    "b200nn_graph_begin": (c_int, [_P]),
    "b200nn_graph_end": (c_int, [_P, C.POINTER(c_void_p)]),
    "b200nn_graph_launch": (c_int, [_P, _P]),
    "b200nn_graph_destroy": (c_int, [_P]),
    "b200nn_act_fwd": (c_int, [c_int, c_float, _P, _P, c_ll, _P]),
    "b200nn_act_bwd": (c_int, [c_int, c_float, _P, _P, _P, c_u64, _P, _P, c_ll, _P]),
    "b200nn_scale_apply": (c_int, [_P, _P, c_u64, _P, c_ll, _P]),
    "b200nn_sumsq": (c_int, [_P, c_ll, _P, _P]),
    "b200nn_softmax_fwd": (c_int, [_P, _P, c_int, c_int, _P]),
    "b200nn_loss_fwd_bwd": (c_int, [c_int, c_int, _P, _P, _P, _P, _P, c_int, c_int, _P]),
    "b200nn_loss_fwd_bwd_ex": (c_int, [c_int, c_int, _P, _P, _P, _P, _P, c_int, c_int, c_ll, c_float, _P]),
    "b200nn_dense_ws_bytes": (c_size_t, [c_int, c_int, c_int]),
    "b200nn_dense_ws_bytes_math": (c_size_t, [c_int, c_int, c_int, c_int]),
    "b200nn_dense_fwd": (c_int, [c_int, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, _P, _P]),
    "b200nn_dense_softmax_cce": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "b200nn_dense_head_partials": (c_int, [c_int, c_int]),
    "b200nn_dense_bwd_head": (c_int, [_P, _P, _P, _P, c_u64, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, _P, _P, _P, _P]),
    "b200nn_dense_bwd": (c_int, [c_int, _P, _P, _P, _P, c_u64, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, _P, _P,
                                  _P, _P]),
    "b200nn_conv2d_ws_bytes": (c_size_t, [C.POINTER(ConvGeom)]),
    "b200nn_conv2d_ws_bytes_math": (c_size_t, [c_int, C.POINTER(ConvGeom)]),
    "b200nn_conv2d_fwd": (c_int, [c_int, C.POINTER(ConvGeom), _P, _P, _P, _P, c_int, c_float, _P, _P]),
    "b200nn_conv2d_bwd": (c_int, [c_int, C.POINTER(ConvGeom), _P, _P, _P, _P, c_u64, _P, _P, _P, c_int, c_float, _P, _P,
                                   _P, _P]),
    "b200nn_convt_ws_bytes": (c_size_t, [C.POINTER(ConvTGeom)]),
    "b200nn_convt_ws_bytes_math": (c_size_t, [c_int, C.POINTER(ConvTGeom)]),
    "b200nn_convt_fwd": (c_int, [c_int, C.POINTER(ConvTGeom), _P, _P, _P, _P, c_int, c_float, _P, _P]),
    "b200nn_convt_bwd": (c_int, [c_int, C.POINTER(ConvTGeom), _P, _P, _P, _P, c_u64, _P, _P, _P, c_int, c_float, _P, _P,
                                  _P, _P]),
    "b200nn_bn_fwd_train": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_ll, c_float, c_float, _P]),
    "b200nn_bn_fwd_infer": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_ll, c_float, _P]),
    "b200nn_bn_bwd": (c_int, [_P, _P, _P, c_u64, _P, _P, _P, _P, _P, _P, c_int, c_ll, c_int, c_float, _P, _P, _P]),
    "b200nn_maxpool_fwd": (c_int, [C.POINTER(PoolGeom), _P, _P, _P]),
    "b200nn_maxpool_bwd": (c_int, [C.POINTER(PoolGeom), _P, _P, _P, _P, c_u64, _P, _P, _P]),
    "b200nn_fused_block_max_batch": (c_int, []),
    "b200nn_bn_pool_fwd_train": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, c_float, _P]),
    "b200nn_pool_bn_act_bwd": (c_int, [_P, _P, _P, c_u64, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int,
                                        c_float, _P, _P, _P, _P]),
    "b200nn_convblock_supported": (c_int, [C.POINTER(ConvGeom), c_int, c_float]),
    "b200nn_convblock_ws_bytes": (c_size_t, [C.POINTER(ConvGeom)]),
    "b200nn_convblock_fwd": (c_int, [C.POINTER(ConvGeom), _P, _P, _P, c_int, c_float, _P, _P, _P, _P, _P, _P, _P, c_int, c_ll, _P,
                                      c_float, c_float, _P, c_int, c_size_t, _P]),
    "b200nn_convblock_bwd": (c_int, [C.POINTER(ConvGeom), _P, _P, _P, c_int, c_float, _P, _P, _P, _P, _P, _P, c_u64, _P, _P, _P, _P,
                                      _P, c_int, c_ll, _P, _P, _P, c_u64, _P, _P, c_int, c_size_t, _P]),
    "b200nn_convblock_region_bytes": (c_size_t, [C.POINTER(ConvGeom), c_int]),
    "b200nn_comm_region_create": (c_int, [_P, c_size_t, C.POINTER(c_int), C.POINTER(c_size_t)]),
    "b200nn_convblock_wgrad_reduce": (c_int, [C.POINTER(ConvGeom), _P, _P, c_u64, _P, _P, _P]),
    "b200nn_adam_multi": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, c_int, _P, _P]),
    "b200nn_sgd_multi": (c_int, [_P, c_int, _P, _P, _P, c_int, _P, _P]),
    "b200nn_opt_multi": (c_int, [c_int, _P, c_int, c_int, _P, _P, _P, _P, c_int, _P, _P]),
    "b200nn_mlp_train_step": (c_int, [C.POINTER(MlpStep), _P]),
    # data parallelism (csrc/comm.cu) and the split BatchNorm kernels
    "b200nn_comm_load_nccl": (c_int, [C.c_char_p]),
    "b200nn_comm_nccl_version": (c_int, []),
    "b200nn_comm_unique_id": (c_int, [_P]),
    "b200nn_comm_create": (c_int, [_P, c_int, c_int, C.POINTER(c_void_p)]),
    "b200nn_comm_destroy": (c_int, [_P]),
    "b200nn_comm_ipc_handle_bytes": (c_size_t, []),
    "b200nn_comm_peer_alloc": (c_int, [_P, c_size_t, _P]),
    "b200nn_comm_peer_open": (c_int, [_P, _P]),
    "b200nn_comm_has_peer": (c_int, [_P]),
    "b200nn_comm_timeouts": (c_int, [_P]),
    "b200nn_comm_site_create": (c_int, [_P, c_size_t, C.POINTER(c_int)]),
    "b200nn_comm_allreduce_sum": (c_int, [_P, c_int, _P, c_ll, c_int, _P]),
    "b200nn_comm_allreduce_sum2": (c_int, [_P, c_int, _P, c_int, _P, c_int, _P]),
    "b200nn_bn_stats_partial": (c_int, [_P, _P, c_int, c_ll, _P]),
    "b200nn_bn_fwd_apply": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_ll, c_ll, c_float, c_float, _P]),
    "b200nn_bn_pool_fwd_apply": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_ll, c_float, c_float, _P]),
    "b200nn_bn_bwd_partial": (c_int, [_P, _P, _P, c_u64, _P, _P, _P, _P, c_int, c_ll, _P]),
    "b200nn_bn_bwd_apply": (c_int, [_P, _P, _P, c_u64, _P, _P, _P, _P, _P, _P, c_int, c_ll, c_ll, c_int, c_float, _P, _P, _P]),
    "b200nn_pool_bn_bwd_partial": (c_int, [_P, _P, _P, c_u64, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "b200nn_mul": (c_int, [_P, _P, _P, c_u64, _P, _P, c_ll, _P]),
    "b200nn_ae_clip": (c_int, [_P, _P, c_ll, c_float, _P, _P]),
    "b200nn_axpy": (c_int, [_P, _P, c_float, _P, c_ll, _P]),
    "b200nn_avgpool2d_fwd": (c_int, [C.POINTER(PoolGeom), _P, _P, _P]),
    "b200nn_avgpool2d_bwd": (c_int, [C.POINTER(PoolGeom), _P, _P, c_u64, _P, _P, _P]),
    "b200nn_gap2d_fwd": (c_int, [_P, _P, c_int, c_int, c_int, _P]),
    "b200nn_gap2d_bwd": (c_int, [_P, _P, c_u64, _P, _P, c_int, c_int, c_int, _P]),
    "b200nn_upsample2d_fwd": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "b200nn_upsample2d_bwd": (c_int, [_P, _P, c_u64, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "b200nn_im2col": (c_int, [C.POINTER(ConvGeom), _P, _P, _P]),
    "b200nn_col2im": (c_int, [C.POINTER(ConvGeom), _P, _P, c_u64, _P, _P, _P]),
    "b200nn_bn_stats_stream_ws_bytes": (c_size_t, [c_int, c_ll]),
    "b200nn_bn_stats_stream": (c_int, [_P, _P, _P, c_int, c_ll, _P]),
    "b200nn_pool_bn_bwd_stream_ws_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "b200nn_pool_bn_bwd_stream_partial": (c_int, [_P, _P, _P, c_u64, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "b200nn_pool_bn_act_bwd_stream_apply": (c_int, [_P, _P, _P, c_u64, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_ll,
                                                     c_int, c_float, _P, _P, _P]),
    "b200nn_pool_bn_act_bwd_apply": (c_int, [_P, _P, _P, c_u64, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_ll,
                                              c_int, c_float, _P, _P, _P]),
}


def exported_symbols():
    """Every entry point include/b200nn.h declares (checked against the .so by tests/test_abi.py)."""
    return sorted(_SIGNATURES)


def load_library(path: str | None = None):
    """dlopen libb200nn.so and attach argtypes. Building is `python -m neuralnetlib_b200._build`
    (or __graft_entry__.build()); a missing library is an error, never a fallback."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = path or LIB_PATH
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `python -m neuralnetlib_b200._build` (needs nvcc). "
                "neuralnetlib_b200 has no CPU fallback.")
        lib = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def lib():
    return _lib if _lib is not None else load_library()


class B200NNError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        buf = C.create_string_buffer(512)
        lib().b200nn_last_error(buf, 512)
        msg = buf.value.decode(errors="replace")
        if rc == -1:
            raise ValueError(f"b200nn: {msg}")
        raise B200NNError(f"b200nn error {rc}: {msg}")


def device() -> torch.device:
    """The CUDA device of this process (one process per GPU: LOCAL_RANK selects it)."""
    global _device
    if _device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("neuralnetlib_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        idx = int(os.environ.get("LOCAL_RANK", "0")) % max(torch.cuda.device_count(), 1)
        torch.cuda.set_device(idx)
        _device = torch.device("cuda", idx)
    return _device


def stream_obj() -> "torch.cuda.Stream":
    """One non-default stream per process: every kernel, copy and graph of the package runs on it
    (the legacy default stream cannot be captured into a CUDA graph)."""
    global _stream
    if _stream is None:
        dev = device()
        _stream = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(_stream)
    return _stream


def stream() -> int:
    return stream_obj().cuda_stream


def synchronize():
    stream_obj().synchronize()


def empty(shape, dtype=torch.float32) -> torch.Tensor:
    stream_obj()
    return torch.empty(shape, dtype=dtype, device=device())


def zeros(shape, dtype=torch.float32) -> torch.Tensor:
    stream_obj()
    return torch.zeros(shape, dtype=dtype, device=device())


def to_device(a, dtype=torch.float32) -> torch.Tensor:
    """Host array -> fresh device tensor (contiguous, `dtype`)."""
    stream_obj()
    if isinstance(a, torch.Tensor):
        return a.to(device=device(), dtype=dtype).contiguous()
    arr = np.ascontiguousarray(np.asarray(a), dtype={torch.float32: np.float32, torch.float64: np.float64,
                                                      torch.int64: np.int64, torch.int32: np.int32}[dtype])
    return torch.from_numpy(arr).to(device())


def to_host(t: torch.Tensor) -> np.ndarray:
    """Device tensor -> float64 ndarray (the reference's dtype). Synchronises."""
    return t.detach().to("cpu").numpy().astype(np.float64)


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()
