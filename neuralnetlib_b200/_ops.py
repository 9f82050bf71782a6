"""Tensor-level wrappers over the C ABI (include/b200nn.h): torch tensors in, raw pointers out.

Nothing here computes: each function validates shapes, packs the lazy-clip chain and calls libb200nn.
`ss` is the float64 slot tensor of the calling plan; slots are ints, chains are tuples of ints.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _backend as B
from ._backend import ConvGeom, ConvTGeom, PoolGeom, check, lib, ptr, stream

ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, ACT_LEAKYRELU, ACT_ELU, ACT_SELU, ACT_GELU = 0, 1, 2, 3, 4, 5, 6, 7
LOSS_MSE, LOSS_BCE, LOSS_CCE, LOSS_MAE, LOSS_HUBER = 0, 1, 2, 3, 4
MATH_FP32, MATH_TF32, MATH_TF32X3 = 0, 1, 2


def pack_chain(chain) -> int:
    chain = tuple(chain or ())
    if chain and chain[0] == "range":   # ("range", lo, count): every slot of [lo, lo + count), in order (deferred clips)
        _, lo, cnt = chain
        if not (0 <= lo < 65536 and 0 <= cnt < 65536):
            raise ValueError("clip slot range out of bounds")
        return (1 << 63) | int(lo) | (int(cnt) << 16) if cnt else 0
    if len(chain) > 4:
        raise ValueError("clip chain longer than 4 slots")
    v = len(chain)
    for i, s in enumerate(chain):
        if not 0 <= s < 4096:
            raise ValueError("clip slot out of range")
        v |= int(s) << (8 + 12 * i)
    return v


def _slot_ptr(ss, slot):
    return None if (ss is None or slot is None) else ss.data_ptr() + 8 * int(slot)


def _f32(*ts):
    for t in ts:
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
            raise TypeError("b200nn ops need contiguous float32 tensors")


# ---- bookkeeping ------------------------------------------------------------------------------------
def step_begin(ss, acc, counter=None, inc=0):
    check(lib().b200nn_step_begin(ptr(ss), 0 if ss is None else ss.numel(), ptr(acc), 0 if acc is None else acc.numel(),
                                  ptr(counter), int(inc), stream()))


def zero_f64(t):
    check(lib().b200nn_zero_f64(ptr(t), t.numel(), stream()))


def gather_rows(src, idx, dst):
    _f32(src, dst)
    rows = idx.numel()
    row_elems = src.numel() // src.shape[0]
    assert idx.dtype == torch.int64 and dst.numel() == rows * row_elems
    check(lib().b200nn_gather_rows(ptr(src), ptr(idx), ptr(dst), rows, row_elems, stream()))


def graph_begin():
    check(lib().b200nn_graph_begin(stream()))


def graph_end():
    h = C.c_void_p()
    check(lib().b200nn_graph_end(stream(), C.byref(h)))
    return h.value


def graph_launch(h):
    check(lib().b200nn_graph_launch(h, stream()))


def graph_destroy(h):
    if h:
        check(lib().b200nn_graph_destroy(h))


# ---- elementwise -------------------------------------------------------------------------------------
def act_fwd(act, alpha, x, y):
    _f32(x, y)
    check(lib().b200nn_act_fwd(act, alpha, ptr(x), ptr(y), x.numel(), stream()))


def act_bwd(act, alpha, y, err, ss, chain, dx, slot_out):
    _f32(y, err, dx)
    check(lib().b200nn_act_bwd(act, alpha, ptr(y), ptr(err), ptr(ss), pack_chain(chain), ptr(dx), _slot_ptr(ss, slot_out),
                               err.numel(), stream()))


def scale_apply(inp, ss, chain, out):
    _f32(inp, out)
    check(lib().b200nn_scale_apply(ptr(inp), ptr(ss), pack_chain(chain), ptr(out), inp.numel(), stream()))


def sumsq(x, ss, slot):
    _f32(x)
    check(lib().b200nn_sumsq(ptr(x), x.numel(), _slot_ptr(ss, slot), stream()))


def softmax_fwd(x, p):
    _f32(x, p)
    check(lib().b200nn_softmax_fwd(ptr(x), ptr(p), x.shape[0], x.numel() // x.shape[0], stream()))


def loss_fwd_bwd(loss, fused_pair, p, y, err, acc, ss, slot_out, world=1, param=1.0):
    """`world`: data-parallel ranks sharing the batch (MSE's derivative divides by the GLOBAL element count); `param`: Huber delta."""
    _f32(p, y, err)
    rows = p.shape[0]
    check(lib().b200nn_loss_fwd_bwd_ex(loss, int(fused_pair), ptr(p), ptr(y), ptr(err), ptr(acc), _slot_ptr(ss, slot_out),
                                       rows, p.numel() // rows, rows * int(world), float(param), stream()))


# ---- Dense -------------------------------------------------------------------------------------------
def dense_ws_bytes(M, N, K, math=MATH_FP32):
    return int(lib().b200nn_dense_ws_bytes_math(math, M, N, K))


def dense_fwd(math, x, w, b, y, act, alpha, ws):
    _f32(x, w, b, y)
    K, N = w.shape
    M = x.numel() // K
    check(lib().b200nn_dense_fwd(math, ptr(x), ptr(w), ptr(b), ptr(y), M, N, K, act, alpha, ptr(ws), stream()))


def dense_softmax_cce(x, w, b, y_true, p, err, acc, ss, slot_out, dw_partial=None):
    _f32(x, w, b, y_true, p, err, dw_partial)
    K, N = w.shape
    M = x.numel() // K
    check(lib().b200nn_dense_softmax_cce(ptr(x), ptr(w), ptr(b), ptr(y_true), ptr(p), ptr(err), ptr(acc), _slot_ptr(ss, slot_out),
                                         M, N, K, ptr(dw_partial), stream()))


def dense_head_partials(M, K):
    return int(lib().b200nn_dense_head_partials(int(M), int(K)))


def dense_bwd_head(x, w, err, ss, chain, dx, dw, db, prev_act, prev_alpha, slot0, slot1, dw_partial):
    _f32(x, w, err, dx, dw, db, dw_partial)
    K, N = w.shape
    M = x.numel() // K
    check(lib().b200nn_dense_bwd_head(ptr(x), ptr(w), ptr(err), ptr(ss), pack_chain(chain), ptr(dx), ptr(dw), ptr(db), M, N, K,
                                      prev_act, prev_alpha, _slot_ptr(ss, slot0), _slot_ptr(ss, slot1), ptr(dw_partial), stream()))


def dense_bwd(math, x, w, err, ss, chain, dx, dw, db, prev_act, prev_alpha, slot0, slot1, ws):
    _f32(x, w, err, dx, dw, db)
    K, N = w.shape
    M = x.numel() // K
    check(lib().b200nn_dense_bwd(math, ptr(x), ptr(w), ptr(err), ptr(ss), pack_chain(chain), ptr(dx), ptr(dw), ptr(db),
                                 M, N, K, prev_act, prev_alpha, _slot_ptr(ss, slot0), _slot_ptr(ss, slot1), ptr(ws),
                                 stream()))


# ---- Conv2D / Conv2DTranspose ------------------------------------------------------------------------
def conv2d_ws_bytes(g: ConvGeom, math=MATH_FP32):
    return int(lib().b200nn_conv2d_ws_bytes_math(math, C.byref(g)))


def conv2d_fwd(math, g, x, w, b, y, act, alpha, ws):
    _f32(x, w, b, y)
    check(lib().b200nn_conv2d_fwd(math, C.byref(g), ptr(x), ptr(w), ptr(b), ptr(y), act, alpha, ptr(ws), stream()))


def conv2d_bwd(math, g, x, w, err, ss, chain, dx, dw, db, prev_act, prev_alpha, slot0, slot1, ws):
    _f32(x, w, err, dx, dw, db)
    check(lib().b200nn_conv2d_bwd(math, C.byref(g), ptr(x), ptr(w), ptr(err), ptr(ss), pack_chain(chain), ptr(dx), ptr(dw),
                                  ptr(db), prev_act, prev_alpha, _slot_ptr(ss, slot0), _slot_ptr(ss, slot1), ptr(ws),
                                  stream()))


def convt_ws_bytes(g: ConvTGeom, math=MATH_FP32):
    return int(lib().b200nn_convt_ws_bytes_math(int(math), C.byref(g)))


def convt_fwd(math, g, x, w, b, y, act, alpha, ws):
    _f32(x, w, b, y)
    check(lib().b200nn_convt_fwd(math, C.byref(g), ptr(x), ptr(w), ptr(b), ptr(y), act, alpha, ptr(ws), stream()))


def convt_bwd(math, g, x, w, err, ss, chain, dx, dw, db, prev_act, prev_alpha, slot0, slot1, ws):
    _f32(x, w, err, dx, dw, db)
    check(lib().b200nn_convt_bwd(math, C.byref(g), ptr(x), ptr(w), ptr(err), ptr(ss), pack_chain(chain), ptr(dx), ptr(dw),
                                 ptr(db), prev_act, prev_alpha, _slot_ptr(ss, slot0), _slot_ptr(ss, slot1), ptr(ws),
                                 stream()))


# ---- fused first block (csrc/conv_block.cu) ------------------------------------------------------------
def convblock_supported(g: ConvGeom, act, alpha):
    return bool(lib().b200nn_convblock_supported(C.byref(g), int(act), float(alpha)))


def convblock_ws_bytes(g: ConvGeom):
    return int(lib().b200nn_convblock_ws_bytes(C.byref(g)))


def convblock_fwd(g, x, w, b, act, alpha, gamma, beta, rmean, rvar, save_mean, save_invstd, stats, mode, b_total, y_pool,
                  momentum, eps, peer=None):
    """peer: (comm handle, site, region offset) for mode 3."""
    _f32(x, w, b, gamma, beta, rmean, rvar, save_mean, save_invstd, y_pool)
    ch, site, off = peer if peer is not None else (None, -1, 0)
    check(lib().b200nn_convblock_fwd(C.byref(g), ptr(x), ptr(w), ptr(b), act, alpha, ptr(gamma), ptr(beta), ptr(rmean), ptr(rvar),
                                     ptr(save_mean), ptr(save_invstd), ptr(stats), int(mode), int(b_total), ptr(y_pool),
                                     momentum, eps, ch, int(site), int(off), stream()))


def convblock_bwd(g, x, w, b, act, alpha, gamma, beta, save_mean, save_invstd, err, ss, chain, sums_or_dg, dbeta, dw, db, r_out,
                  mode, b_total, slot0, slot1, slot2, chain_out, ws, peer=None):
    """sums_or_dg / dbeta: device pointers (ints) of d_gamma and d_beta."""
    _f32(x, w, b, gamma, beta, save_mean, save_invstd, err, dw, db, r_out)
    ch, site, off = peer if peer is not None else (None, -1, 0)
    check(lib().b200nn_convblock_bwd(C.byref(g), ptr(x), ptr(w), ptr(b), act, alpha, ptr(gamma), ptr(beta), ptr(save_mean),
                                     ptr(save_invstd), ptr(err), ptr(ss), pack_chain(chain), sums_or_dg, dbeta, ptr(dw), ptr(db),
                                     ptr(r_out), int(mode), int(b_total), _slot_ptr(ss, slot0), _slot_ptr(ss, slot1),
                                     _slot_ptr(ss, slot2), pack_chain(chain_out), ptr(ws), ch, int(site), int(off), stream()))


def convblock_region_bytes(g, world):
    return int(lib().b200nn_convblock_region_bytes(C.byref(g), int(world)))


def convblock_wgrad_reduce(g, ws, ss, chain_out, dw, db):
    _f32(ws, dw, db)
    check(lib().b200nn_convblock_wgrad_reduce(C.byref(g), ptr(ws), ptr(ss), pack_chain(chain_out), ptr(dw), ptr(db), stream()))


# ---- BatchNorm / MaxPool -----------------------------------------------------------------------------
def bn_fwd_train(x, gamma, beta, rmean, rvar, save_mean, save_invstd, y, momentum, eps):
    _f32(x, gamma, beta, rmean, rvar, save_mean, save_invstd, y)
    Bn = x.shape[0]
    check(lib().b200nn_bn_fwd_train(ptr(x), ptr(gamma), ptr(beta), ptr(rmean), ptr(rvar), ptr(save_mean), ptr(save_invstd),
                                    ptr(y), Bn, x.numel() // Bn, momentum, eps, stream()))


def bn_fwd_infer(x, gamma, beta, rmean, rvar, y, eps):
    _f32(x, gamma, beta, rmean, rvar, y)
    Bn = x.shape[0]
    check(lib().b200nn_bn_fwd_infer(ptr(x), ptr(gamma), ptr(beta), ptr(rmean), ptr(rvar), ptr(y), Bn, x.numel() // Bn, eps,
                                    stream()))


def bn_bwd(x, err, ss, chain, gamma, save_mean, save_invstd, dx, dgamma, dbeta, prev_act, prev_alpha, slot0, slot1):
    _f32(x, err, gamma, save_mean, save_invstd, dx, dgamma, dbeta)
    Bn = x.shape[0]
    check(lib().b200nn_bn_bwd(ptr(x), ptr(err), ptr(ss), pack_chain(chain), ptr(gamma), ptr(save_mean), ptr(save_invstd),
                              ptr(dx), ptr(dgamma), ptr(dbeta), Bn, x.numel() // Bn, prev_act, prev_alpha,
                              _slot_ptr(ss, slot0), _slot_ptr(ss, slot1), stream()))


def fused_block_max_batch():
    return int(lib().b200nn_fused_block_max_batch())


def bn_pool_fwd_train(x, gamma, beta, rmean, rvar, save_mean, save_invstd, y_pool, momentum, eps):
    _f32(x, gamma, beta, rmean, rvar, save_mean, save_invstd, y_pool)
    Bn, H, W, Cc = (int(v) for v in x.shape)
    check(lib().b200nn_bn_pool_fwd_train(ptr(x), ptr(gamma), ptr(beta), ptr(rmean), ptr(rvar), ptr(save_mean),
                                         ptr(save_invstd), ptr(y_pool), Bn, H, W, Cc, momentum, eps, stream()))


def pool_bn_act_bwd(x, err, ss, chain, gamma, beta, save_mean, save_invstd, r0, dgamma, dbeta, prev_act, prev_alpha, slot0,
                    slot1, slot2):
    _f32(x, err, gamma, beta, save_mean, save_invstd, r0, dgamma, dbeta)
    Bn, H, W, Cc = (int(v) for v in x.shape)
    check(lib().b200nn_pool_bn_act_bwd(ptr(x), ptr(err), ptr(ss), pack_chain(chain), ptr(gamma), ptr(beta), ptr(save_mean),
                                       ptr(save_invstd), ptr(r0), ptr(dgamma), ptr(dbeta), Bn, H, W, Cc, prev_act, prev_alpha,
                                       _slot_ptr(ss, slot0), _slot_ptr(ss, slot1), _slot_ptr(ss, slot2), stream()))


# data-parallel halves of the BatchNorm kernels (include/b200nn.h); `stats` is float64 [2P], `sums` float32 [2P] = dbeta | dgamma
def bn_stats_partial(x, stats):
    _f32(x)
    Bn = x.shape[0]
    assert stats.dtype == torch.float64 and stats.numel() == 2 * (x.numel() // Bn)
    check(lib().b200nn_bn_stats_partial(ptr(x), ptr(stats), Bn, x.numel() // Bn, stream()))


def bn_fwd_apply(x, gamma, beta, rmean, rvar, save_mean, save_invstd, stats, y, b_total, momentum, eps):
    _f32(x, gamma, beta, rmean, rvar, save_mean, save_invstd, y)
    Bn = x.shape[0]
    check(lib().b200nn_bn_fwd_apply(ptr(x), ptr(gamma), ptr(beta), ptr(rmean), ptr(rvar), ptr(save_mean), ptr(save_invstd),
                                    ptr(stats), ptr(y), Bn, x.numel() // Bn, int(b_total), momentum, eps, stream()))


def bn_pool_fwd_apply(x, gamma, beta, rmean, rvar, save_mean, save_invstd, stats, y_pool, b_total, momentum, eps):
    _f32(x, gamma, beta, rmean, rvar, save_mean, save_invstd, y_pool)
    Bn, H, W, Cc = (int(v) for v in x.shape)
    check(lib().b200nn_bn_pool_fwd_apply(ptr(x), ptr(gamma), ptr(beta), ptr(rmean), ptr(rvar), ptr(save_mean),
                                         ptr(save_invstd), ptr(stats), ptr(y_pool), Bn, H, W, Cc, int(b_total), momentum, eps,
                                         stream()))


def bn_bwd_partial(x, err, ss, chain, save_mean, save_invstd, sums):
    _f32(x, err, save_mean, save_invstd, sums)
    Bn = x.shape[0]
    P = x.numel() // Bn
    check(lib().b200nn_bn_bwd_partial(ptr(x), ptr(err), ptr(ss), pack_chain(chain), ptr(save_mean), ptr(save_invstd),
                                      sums.data_ptr() + 4 * P, sums.data_ptr(), Bn, P, stream()))


def bn_bwd_apply(x, err, ss, chain, gamma, save_mean, save_invstd, dx, sums, b_total, prev_act, prev_alpha, slot0, slot1):
    _f32(x, err, gamma, save_mean, save_invstd, dx, sums)
    Bn = x.shape[0]
    P = x.numel() // Bn
    check(lib().b200nn_bn_bwd_apply(ptr(x), ptr(err), ptr(ss), pack_chain(chain), ptr(gamma), ptr(save_mean), ptr(save_invstd),
                                    ptr(dx), sums.data_ptr() + 4 * P, sums.data_ptr(), Bn, P, int(b_total), prev_act, prev_alpha,
                                    _slot_ptr(ss, slot0), _slot_ptr(ss, slot1), stream()))


def pool_bn_bwd_partial(x, err, ss, chain, gamma, beta, save_mean, save_invstd, sums, slot0):
    _f32(x, err, gamma, beta, save_mean, save_invstd, sums)
    Bn, H, W, Cc = (int(v) for v in x.shape)
    P = H * W * Cc
    check(lib().b200nn_pool_bn_bwd_partial(ptr(x), ptr(err), ptr(ss), pack_chain(chain), ptr(gamma), ptr(beta), ptr(save_mean),
                                           ptr(save_invstd), sums.data_ptr() + 4 * P, sums.data_ptr(), Bn, H, W, Cc,
                                           _slot_ptr(ss, slot0), stream()))


def pool_bn_act_bwd_apply(x, err, ss, chain, gamma, beta, save_mean, save_invstd, r0, sums, b_total, prev_act, prev_alpha,
                          slot1, slot2):
    _f32(x, err, gamma, beta, save_mean, save_invstd, r0, sums)
    Bn, H, W, Cc = (int(v) for v in x.shape)
    P = H * W * Cc
    check(lib().b200nn_pool_bn_act_bwd_apply(ptr(x), ptr(err), ptr(ss), pack_chain(chain), ptr(gamma), ptr(beta),
                                             ptr(save_mean), ptr(save_invstd), ptr(r0), sums.data_ptr() + 4 * P, sums.data_ptr(),
                                             Bn, H, W, Cc, int(b_total), prev_act, prev_alpha, _slot_ptr(ss, slot1),
                                             _slot_ptr(ss, slot2), stream()))


# ---- sibling layers (csrc/layers_extra.cu) -------------------------------------------------------------------------------
def mul(a, b, ss, chain, out, slot_out):
    _f32(a, b, out)
    check(lib().b200nn_mul(ptr(a), ptr(b), ptr(ss), pack_chain(chain), ptr(out), _slot_ptr(ss, slot_out), a.numel(), stream()))


def avgpool2d_fwd(g: PoolGeom, x, y):
    _f32(x, y)
    check(lib().b200nn_avgpool2d_fwd(C.byref(g), ptr(x), ptr(y), stream()))


def avgpool2d_bwd(g: PoolGeom, err, ss, chain, dx, slot_out):
    _f32(err, dx)
    check(lib().b200nn_avgpool2d_bwd(C.byref(g), ptr(err), ptr(ss), pack_chain(chain), ptr(dx), _slot_ptr(ss, slot_out), stream()))


def gap2d_fwd(x, y):
    _f32(x, y)
    N, H, W, Cc = (int(v) for v in x.shape)
    check(lib().b200nn_gap2d_fwd(ptr(x), ptr(y), N, H * W, Cc, stream()))


def gap2d_bwd(err, ss, chain, dx, slot_out):
    _f32(err, dx)
    N, H, W, Cc = (int(v) for v in dx.shape)
    check(lib().b200nn_gap2d_bwd(ptr(err), ptr(ss), pack_chain(chain), ptr(dx), _slot_ptr(ss, slot_out), N, H * W, Cc, stream()))


def upsample2d_fwd(x, y, fh, fw):
    _f32(x, y)
    N, H, W, Cc = (int(v) for v in x.shape)
    check(lib().b200nn_upsample2d_fwd(ptr(x), ptr(y), N, H, W, Cc, int(fh), int(fw), stream()))


def upsample2d_bwd(err, ss, chain, dx, slot_out, fh, fw):
    _f32(err, dx)
    N, H, W, Cc = (int(v) for v in dx.shape)
    check(lib().b200nn_upsample2d_bwd(ptr(err), ptr(ss), pack_chain(chain), ptr(dx), _slot_ptr(ss, slot_out), N, H, W, Cc, int(fh), int(fw),
                                      stream()))


def im2col(g: ConvGeom, x, col):
    _f32(x, col)
    check(lib().b200nn_im2col(C.byref(g), ptr(x), ptr(col), stream()))


def col2im(g: ConvGeom, dcol, ss, chain, dx, slot_out):
    _f32(dcol, dx)
    check(lib().b200nn_col2im(C.byref(g), ptr(dcol), ptr(ss), pack_chain(chain), ptr(dx), _slot_ptr(ss, slot_out), stream()))


# streaming forms for large batches (csrc/bn_pool.cu); `ws` is the plan's float workspace
def bn_stats_stream_ws_bytes(Bn, P):
    return int(lib().b200nn_bn_stats_stream_ws_bytes(int(Bn), int(P)))


def bn_stats_stream(x, stats, ws):
    _f32(x, ws)
    Bn = x.shape[0]
    P = x.numel() // Bn
    assert stats.dtype == torch.float64 and stats.numel() == 2 * P and ws.numel() * 4 >= bn_stats_stream_ws_bytes(Bn, P)
    check(lib().b200nn_bn_stats_stream(ptr(x), ptr(stats), ptr(ws), Bn, P, stream()))


def pool_bn_bwd_stream_ws_bytes(Bn, H, W, Cc):
    return int(lib().b200nn_pool_bn_bwd_stream_ws_bytes(int(Bn), int(H), int(W), int(Cc)))


def pool_bn_bwd_stream_partial(x, err, ss, chain, gamma, beta, save_mean, save_invstd, sums, ws, slot0):
    """sums: float32 [2P] = dbeta | dgamma (receives the LOCAL sums)."""
    _f32(x, err, gamma, beta, save_mean, save_invstd, sums, ws)
    Bn, H, W, Cc = (int(v) for v in x.shape)
    P = H * W * Cc
    assert ws.numel() * 4 >= pool_bn_bwd_stream_ws_bytes(Bn, H, W, Cc)
    check(lib().b200nn_pool_bn_bwd_stream_partial(ptr(x), ptr(err), ptr(ss), pack_chain(chain), ptr(gamma), ptr(beta), ptr(save_mean),
                                                  ptr(save_invstd), sums.data_ptr() + 4 * P, sums.data_ptr(), ptr(ws), Bn, H, W, Cc,
                                                  _slot_ptr(ss, slot0), stream()))


def pool_bn_act_bwd_stream_apply(x, err, ss, chain, gamma, beta, save_mean, save_invstd, r0, sums, b_total, prev_act, prev_alpha,
                                 slot1, slot2):
    _f32(x, err, gamma, beta, save_mean, save_invstd, r0, sums)
    Bn, H, W, Cc = (int(v) for v in x.shape)
    P = H * W * Cc
    check(lib().b200nn_pool_bn_act_bwd_stream_apply(ptr(x), ptr(err), ptr(ss), pack_chain(chain), ptr(gamma), ptr(beta),
                                                    ptr(save_mean), ptr(save_invstd), ptr(r0), sums.data_ptr() + 4 * P, sums.data_ptr(),
                                                    Bn, H, W, Cc, int(b_total), prev_act, prev_alpha, _slot_ptr(ss, slot1),
                                                    _slot_ptr(ss, slot2), stream()))


def maxpool_fwd(g: PoolGeom, x, y):
    _f32(x, y)
    check(lib().b200nn_maxpool_fwd(C.byref(g), ptr(x), ptr(y), stream()))


def maxpool_bwd(g: PoolGeom, x, y, err, ss, chain, dx, slot_out):
    _f32(x, y, err, dx)
    check(lib().b200nn_maxpool_bwd(C.byref(g), ptr(x), ptr(y), ptr(err), ptr(ss), pack_chain(chain), ptr(dx),
                                   _slot_ptr(ss, slot_out), stream()))


# ---- optimizers --------------------------------------------------------------------------------------
class OptPlan:
    """Device-resident descriptor table for the multi-tensor optimizer kernels.
    entries: list of (p, g, m, v, t_index, clip[, chain]) tensors (m, v may be None for SGD); `ss`: the slot tensor the
    optional per-tensor chains (deferred error clips) refer to."""

    def __init__(self, entries, n_layers, ss=None):
        self.ss = ss
        self.keep = entries  # keep tensors alive
        self.n_layers = int(n_layers)
        desc = np.zeros(len(entries), dtype=B.PARAM_DTYPE)
        pairs = []
        for i, entry in enumerate(entries):
            p, g, m, v, t_index, clip = entry[:6]
            chain = entry[6] if len(entry) > 6 else ()
            _f32(p, g, m, v)
            assert p.numel() == g.numel()
            desc[i] = (p.data_ptr(), g.data_ptr(), 0 if m is None else m.data_ptr(), 0 if v is None else v.data_ptr(),
                       p.numel(), t_index, int(clip), pack_chain(chain))
            for c in range((p.numel() + B.OPT_CHUNK - 1) // B.OPT_CHUNK):
                pairs.append((i, c))
        self.n_params = len(entries)
        self.total_blocks = len(pairs)
        self.desc = torch.from_numpy(desc.view(np.uint8).copy()).to(B.device())
        self.block_map = torch.from_numpy(np.asarray(pairs, dtype=np.int32).reshape(-1).copy()).to(B.device())
        self.gss = B.zeros((self.n_params + 1,), dtype=torch.float64)   # + the two counters of the single-launch Adam


def adam_multi(plan: OptPlan, hyper, t_counter, on_stream=None):
    check(lib().b200nn_adam_multi(ptr(plan.desc), plan.n_params, plan.n_layers, ptr(plan.gss), ptr(hyper), ptr(t_counter),
                                  ptr(plan.block_map), plan.total_blocks, ptr(plan.ss), stream() if on_stream is None else on_stream))


OPT_ADAM, OPT_MOMENTUM, OPT_RMSPROP, OPT_ADABELIEF, OPT_RADAM = 0, 1, 2, 3, 4


def opt_multi(kind, plan: OptPlan, hyper, t_counter, on_stream=None):
    """hyper: device float64[6] {lr, beta1 | momentum | rho, beta2, eps, clip_norm, clip_value}."""
    check(lib().b200nn_opt_multi(int(kind), ptr(plan.desc), plan.n_params, plan.n_layers, ptr(plan.gss), ptr(hyper), ptr(t_counter),
                                 ptr(plan.block_map), plan.total_blocks, ptr(plan.ss), stream() if on_stream is None else on_stream))


def sgd_multi(plan: OptPlan, hyper, on_stream=None):
    check(lib().b200nn_sgd_multi(ptr(plan.desc), plan.n_params, ptr(plan.gss), ptr(hyper), ptr(plan.block_map),
                                 plan.total_blocks, ptr(plan.ss), stream() if on_stream is None else on_stream))


def mlp_train_step(desc):
    """desc: a filled _backend.MlpStep (kept alive by the caller). One launch = one whole train step of a small MLP."""
    check(lib().b200nn_mlp_train_step(C.byref(desc), stream()))


def ae_clip(g, out, threshold, scratch3):
    """Autoencoder clip rule (models.py:1033-1048) on one tensor; scratch3: float64[3] device tensor."""
    _f32(g, out)
    check(lib().b200nn_ae_clip(ptr(g), ptr(out), g.numel(), float(threshold), ptr(scratch3), stream()))


def axpy(a, b, s, out):
    _f32(a, b, out)
    check(lib().b200nn_axpy(ptr(a), ptr(b), float(s), ptr(out), a.numel(), stream()))
