"""GPU parity, op level: every C-ABI kernel family against (a) the golden vectors produced by executing
the upstream reference (tests/golden/ops.npz) and (b) the numpy oracle on seeded inputs at larger sizes.

Tolerances (BASELINE.json north_star): fp32 FFMA path within rel 1e-5 (norm-relative: max|a-b| / max|b|),
max-pool outputs and argmax bit-exact.
"""
import numpy as np
import pytest

import nnl_oracle as O
from helpers import nrel

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def K():
    import torch
    from neuralnetlib_b200 import _backend as B, _ops as ops

    class Ctx:
        pass

    k = Ctx()
    k.B, k.ops, k.torch = B, ops, torch
    k.dev = lambda a: B.to_device(np.asarray(a, dtype=np.float32))
    k.host = lambda t: (B.synchronize(), t.detach().cpu().numpy().astype(np.float64))[1]
    k.empty = B.empty
    k.ss = lambda n=16: B.zeros((n,), dtype=torch.float64)
    return k


def _f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


class Guarded:
    """Output buffers carved out of a larger allocation with sentinel-filled guard bands on both sides (compute-sanitizer is
    not available on the GPU pool): `check()` asserts that a kernel wrote nothing outside the tensors it was given."""
    GUARD, SENTINEL = 256, 12345.0

    def __init__(self, K):
        self.K, self.bufs = K, []

    def __call__(self, shape):
        n = int(np.prod(shape))
        flat = self.K.empty((n + 2 * self.GUARD,))
        flat.fill_(self.SENTINEL)
        self.bufs.append((flat, n))
        return flat[self.GUARD:self.GUARD + n].view(*shape)     # 1 KB guard keeps the 16-byte alignment

    def check(self):
        self.K.B.synchronize()
        for flat, n in self.bufs:
            lo, hi = flat[:self.GUARD].cpu().numpy(), flat[self.GUARD + n:].cpu().numpy()
            assert np.all(lo == self.SENTINEL) and np.all(hi == self.SENTINEL), "write outside an output buffer"


ACTS = {"linear": 0, "relu": 1, "sigmoid": 2, "tanh": 3, "leakyrelu": 4}


@pytest.mark.parametrize("name", list(ACTS))
def test_activation_fwd_bwd_golden(K, golden, name):
    g = golden("ops")
    alpha = 0.2 if name == "leakyrelu" else 0.0
    x = K.dev(g["act_x"]); y = K.empty(x.shape)
    K.ops.act_fwd(ACTS[name], alpha, x, y)
    assert nrel(K.host(y), g[f"act_{name}_y"]) < TOL
    err = K.dev(g["act_err"]); dx = K.empty(x.shape); ss = K.ss()
    K.ops.act_bwd(ACTS[name], alpha, y, err, ss, (), dx, 3)
    assert nrel(K.host(dx), g[f"act_{name}_dx"]) < TOL
    assert abs(K.host(ss)[3] - np.sum(g[f"act_{name}_dx"] ** 2)) < 1e-5 * np.sum(g[f"act_{name}_dx"] ** 2)


def test_activation_large_unaligned(K):
    rng = np.random.default_rng(1)
    n = 1_000_003
    x = rng.normal(size=n).astype(np.float32) * 4
    xd = K.dev(x); yd = K.empty((n,))
    for name in ("relu", "sigmoid", "tanh"):
        K.ops.act_fwd(ACTS[name], 0.0, xd, yd)
        assert nrel(K.host(yd), O.act_fwd(name, x.astype(np.float64))) < TOL


def test_softmax_and_losses_golden(K, golden):
    g = golden("ops")
    x = K.dev(g["act_x"]); p = K.empty(x.shape)
    K.ops.softmax_fwd(x, p)
    assert nrel(K.host(p), g["act_softmax_y"]) < TOL
    for kind, code, pk, yk in (("cce", 2, "loss_p", "loss_y"), ("mse", 0, "loss_p", "loss_y"), ("bce", 1, "loss_pb", "loss_yb")):
        pd, yd = K.dev(g[pk]), K.dev(g[yk])
        err = K.empty(pd.shape); acc = K.B.zeros((4,), dtype=K.torch.float64); ss = K.ss()
        K.ops.loss_fwd_bwd(code, 0, pd, yd, err, acc, ss, 0)
        rows, cols = g[pk].shape
        denom = rows if kind == "cce" else rows * cols
        assert abs(K.host(acc)[0] / denom - float(g[f"loss_{kind}"])) < 1e-5 * abs(float(g[f"loss_{kind}"]))
        assert nrel(K.host(err), g[f"loss_{kind}_d"]) < TOL
        # fused pair: err = p - y, not divided by the batch (models.py:200-205)
        K.ops.zero_f64(ss); K.ops.zero_f64(acc)
        K.ops.loss_fwd_bwd(code, 1, pd, yd, err, acc, ss, 1)
        assert nrel(K.host(err), _f32(g[pk]) - _f32(g[yk])) < TOL
        assert abs(K.host(ss)[1] - np.sum((g[pk] - g[yk]) ** 2)) < 1e-5
    # argmax count incl. first-index tie break (metrics.py:84-89)
    p = np.array([[0.5, 0.5, 0.0], [0.1, 0.2, 0.7], [0.3, 0.3, 0.4]], dtype=np.float32)
    y = np.array([[1, 0, 0], [0, 0, 1], [0, 1, 0]], dtype=np.float32)
    acc = K.B.zeros((4,), dtype=K.torch.float64)
    K.ops.loss_fwd_bwd(2, 1, K.dev(p), K.dev(y), K.empty(p.shape), acc, None, None)
    assert K.host(acc)[1] == np.sum(p.argmax(1) == y.argmax(1)) == 2


def test_clip_chain(K, golden):
    g = golden("ops")
    x = K.dev(g["clip_in"]); ss = K.ss(); out = K.empty(x.shape)
    K.ops.sumsq(x, ss, 5)
    K.ops.scale_apply(x, ss, (5,), out)
    assert nrel(K.host(out), g["clip_big"]) < TOL
    xs = K.dev(g["clip_in"] * 1e-2); K.ops.zero_f64(ss)
    K.ops.sumsq(xs, ss, 2)
    K.ops.scale_apply(xs, ss, (2,), out)
    assert nrel(K.host(out), g["clip_small"]) < 1e-7  # norm <= 1: untouched
    # two-slot chain: acc = s0 * clip(s0^2 * ss1)
    ss2 = K.B.to_device(np.array([4.0, 9.0, 0.25]), dtype=K.torch.float64)
    K.ops.scale_apply(x, ss2, (0, 1), out)   # s0 = 1/2; n2 = 1/4*9 -> s1 = 2/3 ; total 1/3
    assert nrel(K.host(out), _f32(g["clip_in"]) / 3.0) < 1e-6
    K.ops.scale_apply(x, ss2, (2, 0), out)   # slot2 <= 1: no clip; then 1/2
    assert nrel(K.host(out), _f32(g["clip_in"]) / 2.0) < 1e-6


def test_dense_golden(K, golden):
    g = golden("ops")
    x, w, b, err = (K.dev(g[k]) for k in ("dense_x", "dense_w", "dense_b", "dense_err"))
    y = K.empty((5, 3))
    K.ops.dense_fwd(0, x, w, b.reshape(-1), y, 0, 0.0, None)
    assert nrel(K.host(y), g["dense_y"]) < TOL
    dx, dw, db = K.empty((5, 7)), K.empty((7, 3)), K.empty((3,))
    K.ops.dense_bwd(0, x, w, err, None, (), dx, dw, db, 0, 0.0, None, None, None)
    assert nrel(K.host(dx), g["dense_dx"]) < TOL
    assert nrel(K.host(dw), g["dense_dw"]) < TOL
    assert nrel(K.host(db), g["dense_db"].reshape(-1)) < TOL


@pytest.mark.parametrize("M,N,Kd,act", [(32, 128, 784, "relu"), (256, 64, 6272, "relu"), (256, 10, 64, "linear"),
                                        (100, 70, 33, "sigmoid"), (1, 5, 3, "tanh"), (300, 300, 300, "leakyrelu"),
                                        (100, 48, 2048, "relu"), (256, 64, 1028, "tanh"), (7, 4, 1060, "linear"),   # skinny slab kernels
                                        (64, 10, 4096, "linear"), (512, 10, 64, "relu"), (600, 3, 50, "relu")])     # small-N kernels
def test_dense_oracle(K, M, N, Kd, act):
    rng = np.random.default_rng(M * 7 + N)
    x = rng.normal(size=(M, Kd)).astype(np.float32)
    w = (rng.normal(size=(Kd, N)) / np.sqrt(Kd)).astype(np.float32)
    b = rng.normal(size=(N,)).astype(np.float32)
    err = rng.normal(size=(M, N)).astype(np.float32)
    alpha = 0.1 if act == "leakyrelu" else 0.0
    xd, wd, bd, ed = K.dev(x), K.dev(w), K.dev(b), K.dev(err)
    wsb = K.ops.dense_ws_bytes(M, N, Kd)
    ws = K.empty((max(wsb // 4, 1),))
    G = Guarded(K)                      # the skinny-slab and small-N kernels index shared / global memory by hand
    y = G((M, N))
    K.ops.dense_fwd(0, xd, wd, bd, y, ACTS[act], alpha, ws)
    x64, w64 = x.astype(np.float64), w.astype(np.float64)
    ref = O.act_fwd(act, O.dense_fwd(x64, w64, b.astype(np.float64)), alpha)
    assert nrel(K.host(y), ref) < TOL
    # backward with a clip chain and the fused backward of a preceding activation (mask from x)
    ss = K.ss()
    K.ops.sumsq(ed, ss, 0)
    dx, dw, db = G((M, Kd)), G((Kd, N)), G((N,))
    xpos = np.maximum(x, 0)  # pretend x is a ReLU output
    xpd = K.dev(xpos)
    K.ops.dense_bwd(0, xpd, wd, ed, ss, (0,), dx, dw, db, ACTS["relu"], 0.0, 1, 2, ws)
    G.check()
    e64 = O.clip_l2(err.astype(np.float64))
    rdx, rdw, rdb = O.dense_bwd(xpos.astype(np.float64), w64, e64)
    hss = K.host(ss)
    assert abs(hss[1] - np.sum(rdx ** 2)) < 1e-4 * np.sum(rdx ** 2)
    masked = rdx * (xpos > 0)
    assert nrel(K.host(dx), masked) < TOL
    assert abs(hss[2] - np.sum(masked ** 2)) < 1e-4 * max(np.sum(masked ** 2), 1e-30)
    assert nrel(K.host(dw), rdw) < TOL
    assert nrel(K.host(db), rdb.reshape(-1)) < TOL
    # dx == None is allowed (first parametrised layer in fit)
    K.ops.dense_bwd(0, xpd, wd, ed, ss, (0,), None, dw, db, 0, 0.0, None, None, ws)
    assert nrel(K.host(dw), rdw) < TOL


def _conv_geom(K, N, H, W, C, F, kh, kw, sh, sw, pad):
    oh, ow, ph, pw = O.conv2d_geometry(H, W, kh, kw, sh, sw, pad)
    return K.B.ConvGeom(N, H, W, C, F, kh, kw, sh, sw, ph, pw, oh, ow), oh, ow


@pytest.mark.parametrize("gi", range(6))
def test_conv2d_golden(K, golden, gi):
    g = golden("ops")
    H, W, C, F, kh, kw, sh, sw, same = [int(v) for v in g[f"conv{gi}_cfg"]]
    geom, oh, ow = _conv_geom(K, 3, H, W, C, F, kh, kw, sh, sw, "same" if same else "valid")
    x, w, b, err = (K.dev(g[f"conv{gi}_{k}"]) for k in ("x", "w", "b", "err"))
    ws = K.empty((max(K.ops.conv2d_ws_bytes(geom) // 4, 1),))
    y = K.empty((3, oh, ow, F))
    K.ops.conv2d_fwd(0, geom, x, w, b.reshape(-1), y, 0, 0.0, ws)
    assert nrel(K.host(y), g[f"conv{gi}_y"]) < TOL
    dx, dw, db = K.empty((3, H, W, C)), K.empty((kh, kw, C, F)), K.empty((F,))
    K.ops.conv2d_bwd(0, geom, x, w, err, None, (), dx, dw, db, 0, 0.0, None, None, ws)
    assert nrel(K.host(dx), g[f"conv{gi}_dx"]) < TOL
    assert nrel(K.host(dw), g[f"conv{gi}_dw"]) < TOL
    assert nrel(K.host(db), g[f"conv{gi}_db"]) < TOL


@pytest.mark.parametrize("N,H,W,C,F,k,s,pad", [(16, 28, 28, 1, 32, 3, 1, "same"), (8, 32, 32, 3, 64, 3, 1, "same"),
                                               (4, 16, 16, 64, 128, 3, 1, "same"), (8, 28, 28, 1, 32, 3, 2, "same"),
                                               (8, 13, 13, 32, 64, 3, 2, "same"), (4, 28, 28, 32, 1, 3, 1, "same"),
                                               # large-batch small-K kernels (conv_smallk3.cuh): ragged last tile, generic K, strides
                                               (7, 30, 30, 2, 64, 3, 1, "valid"), (6, 31, 29, 1, 32, 5, 1, "same"),
                                               (33, 28, 28, 1, 64, 3, 2, "same"), (12, 32, 32, 3, 32, 3, 1, "same"),
                                               # few-filter kernels with lane <-> channel quad (conv_fewf_fwd2)
                                               (8, 28, 28, 32, 1, 3, 1, "same"), (24, 28, 28, 16, 2, 3, 2, "same"),
                                               # band kernels (3 x 3 / stride 1 / 'same' / F <= 2): ragged last band, two filters
                                               (5, 12, 20, 16, 2, 3, 1, "same"), (3, 10, 9, 8, 1, 3, 1, "same")])
def test_conv2d_oracle(K, N, H, W, C, F, k, s, pad):
    rng = np.random.default_rng(H * 31 + C)
    x = rng.normal(size=(N, H, W, C)).astype(np.float32)
    w = (rng.normal(size=(k, k, C, F)) / np.sqrt(k * k * C)).astype(np.float32)
    b = rng.normal(size=(F,)).astype(np.float32)
    geom, oh, ow = _conv_geom(K, N, H, W, C, F, k, k, s, s, pad)
    err = rng.normal(size=(N, oh, ow, F)).astype(np.float32)
    xd, wd, bd, ed = K.dev(x), K.dev(w), K.dev(b), K.dev(err)
    ws = K.empty((max(K.ops.conv2d_ws_bytes(geom) // 4, 1),))
    G = Guarded(K)                      # small-K tile GEMMs, few-filter / band kernels: hand-indexed shared memory and tails
    y = G((N, oh, ow, F))
    K.ops.conv2d_fwd(0, geom, xd, wd, bd, y, 1, 0.0, ws)
    ref = np.maximum(O.conv2d_fwd(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), (s, s), pad), 0)
    assert nrel(K.host(y), ref) < TOL
    ss = K.ss(); K.ops.sumsq(ed, ss, 0)
    dx, dw, db = G(x.shape), G(w.shape), G((F,))
    K.ops.conv2d_bwd(0, geom, xd, wd, ed, ss, (0,), dx, dw, db, 0, 0.0, 1, None, ws)
    G.check()
    rdx, rdw, rdb = O.conv2d_bwd(x.astype(np.float64), w.astype(np.float64), O.clip_l2(err.astype(np.float64)), (s, s), pad)
    assert nrel(K.host(dx), rdx) < TOL
    assert nrel(K.host(dw), rdw) < TOL
    assert nrel(K.host(db), rdb) < TOL
    assert abs(K.host(ss)[1] - np.sum(rdx ** 2)) < 1e-4 * np.sum(rdx ** 2)


@pytest.mark.parametrize("math,tol", [(2, 1e-5), (1, 2e-3)])
@pytest.mark.parametrize("N,H,W,C,F,k,s,pad", [(7, 30, 30, 2, 64, 3, 1, "valid"), (6, 31, 29, 1, 32, 5, 1, "same"),
                                               (33, 28, 28, 1, 64, 3, 2, "same"), (12, 32, 32, 3, 32, 3, 1, "same"),
                                               (9, 32, 32, 3, 64, 3, 1, "same")])
def test_conv2d_smallk_tensor_modes(K, monkeypatch, N, H, W, C, F, k, s, pad, math, tol):
    """Small-K (first-layer) convolutions in the tensor-core math modes: the mma.sync TF32 / 3xTF32 tile kernels of
    conv_smallk3.cuh (ragged last tile, K = 9 / 18 / 25 / 27 padded to 32, strides, both filter counts) against the fp64 oracle;
    tf32x3 must meet the FFMA path's 1e-5, and differ in bits from the FFMA result (the tensor cores really ran). The 3xTF32 kernels
    are off by default (no faster than the FFMA kernels, see conv_smallk3.cuh) and forced here with B200NN_SMALLK_MMA=2."""
    monkeypatch.setenv("B200NN_SMALLK_MMA", "2")
    rng = np.random.default_rng(H * 13 + C + F)
    x = rng.normal(size=(N, H, W, C)).astype(np.float32)
    w = (rng.normal(size=(k, k, C, F)) / np.sqrt(k * k * C)).astype(np.float32)
    b = rng.normal(size=(F,)).astype(np.float32)
    geom, oh, ow = _conv_geom(K, N, H, W, C, F, k, k, s, s, pad)
    err = rng.normal(size=(N, oh, ow, F)).astype(np.float32)
    xd, wd, bd, ed = K.dev(x), K.dev(w), K.dev(b), K.dev(err)
    G = Guarded(K)
    ws = G((max(K.ops.conv2d_ws_bytes(geom, math) // 4, 1),))
    y = G((N, oh, ow, F))
    K.ops.conv2d_fwd(math, geom, xd, wd, bd, y, 1, 0.0, ws)
    ref = np.maximum(O.conv2d_fwd(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), (s, s), pad), 0)
    assert nrel(K.host(y), ref) < tol
    ss = K.ss(); K.ops.sumsq(ed, ss, 0)
    dx, dw, db = G(x.shape), G(w.shape), G((F,))
    K.ops.conv2d_bwd(math, geom, xd, wd, ed, ss, (0,), dx, dw, db, 0, 0.0, 1, None, ws)
    G.check()
    rdx, rdw, rdb = O.conv2d_bwd(x.astype(np.float64), w.astype(np.float64), O.clip_l2(err.astype(np.float64)), (s, s), pad)
    assert nrel(K.host(dx), rdx) < tol
    assert nrel(K.host(dw), rdw) < tol
    assert nrel(K.host(db), rdb) < max(tol, 1e-5)
    y0 = K.empty((N, oh, ow, F)); K.ops.conv2d_fwd(0, geom, xd, wd, bd, y0, 1, 0.0, ws)
    assert not np.array_equal(K.host(y0), K.host(y))
    if math == 1:
        assert nrel(K.host(y), ref) > 20 * nrel(K.host(y0), ref)


@pytest.mark.parametrize("math,tol", [(2, 1e-5), (1, 2e-3)])
@pytest.mark.parametrize("N,H,W,C,F,k,act", [(8, 16, 16, 64, 128, 3, 1), (37, 8, 8, 128, 256, 3, 1), (45, 4, 4, 32, 32, 3, 0),
                                             (3, 32, 32, 32, 96, 5, 1), (2, 4, 128, 32, 64, 3, 0), (7, 64, 8, 64, 32, 1, 1)])
def test_conv2d_tcgen05_implicit_gemm(K, N, H, W, C, F, k, act, math, tol):
    """Tensor-core convolution (math modes tf32 / tf32x3): implicit GEMM on the tcgen05 kernel with the im2col rows fetched
    by 4-D TMA boxes from the NHWC tensor (zero fill = 'same' padding) and the weights by a 3-D map in the reference's
    (c, ky, kx) K order. Forward, dgrad (mirrored taps, activation-backward mask, clip multiplier, sums of squares) and
    wgrad (row permutation to (c, ky, kx), db) against the fp64 oracle; tf32x3 must meet the FFMA path's 1e-5."""
    rng = np.random.default_rng(H * 31 + C + k)
    x = np.maximum(rng.normal(size=(N, H, W, C)), 0).astype(np.float32)     # x is a ReLU output: used as the dgrad mask
    w = (rng.normal(size=(k, k, C, F)) / np.sqrt(k * k * C)).astype(np.float32)
    b = rng.normal(size=(F,)).astype(np.float32)
    geom, oh, ow = _conv_geom(K, N, H, W, C, F, k, k, 1, 1, "same")
    assert (oh, ow) == (H, W)
    assert K.ops.conv2d_ws_bytes(geom, math) >= K.ops.conv2d_ws_bytes(geom, 0)
    err = rng.normal(size=(N, oh, ow, F)).astype(np.float32)
    xd, wd, bd, ed = K.dev(x), K.dev(w), K.dev(b), K.dev(err)
    G = Guarded(K)
    ws = G((max(K.ops.conv2d_ws_bytes(geom, math) // 4, 1),))     # also: the workspace the sizing call promises is enough
    y = G((N, oh, ow, F))
    K.ops.conv2d_fwd(math, geom, xd, wd, bd, y, act, 0.0, ws)
    x64, w64 = x.astype(np.float64), w.astype(np.float64)
    ref = O.conv2d_fwd(x64, w64, b.astype(np.float64), (1, 1), "same")
    ref = np.maximum(ref, 0) if act else ref
    assert nrel(K.host(y), ref) < tol
    ss = K.ss(); K.ops.sumsq(ed, ss, 0)
    dx, dw, db = G(x.shape), G(w.shape), G((F,))
    K.ops.conv2d_bwd(math, geom, xd, wd, ed, ss, (0,), dx, dw, db, 1, 0.0, 1, 2, ws)
    G.check()
    rdx, rdw, rdb = O.conv2d_bwd(x64, w64, O.clip_l2(err.astype(np.float64)), (1, 1), "same")
    assert nrel(K.host(dx), rdx * (x > 0)) < tol
    assert nrel(K.host(dw), rdw) < tol
    assert nrel(K.host(db), rdb) < 1e-5
    h = K.host(ss)
    assert abs(h[1] - np.sum(rdx ** 2)) < max(1e-4, 4 * tol) * np.sum(rdx ** 2)
    assert abs(h[2] - np.sum((rdx * (x > 0)) ** 2)) < max(1e-4, 4 * tol) * np.sum((rdx * (x > 0)) ** 2)
    if math == 2:   # the tensor cores really ran: single-pass TF32 on the same inputs is far from this accuracy
        y1 = K.empty((N, oh, ow, F)); K.ops.conv2d_fwd(1, geom, xd, wd, bd, y1, act, 0.0, ws)
        assert nrel(K.host(y1), ref) > 20 * nrel(K.host(y), ref)


@pytest.mark.parametrize("N,H,W,C,F,k,act", [(8, 16, 16, 64, 128, 3, 1), (37, 8, 8, 128, 256, 3, 1), (7, 64, 8, 64, 64, 1, 0),
                                             (5, 4, 4, 192, 64, 3, 1), (3, 32, 32, 64, 64, 5, 1)])
def test_conv2d_tcgen05_f16_planes(K, monkeypatch, N, H, W, C, F, k, act):
    """tf32x3 forward / data-gradient convolution from fp16 hi / lo operand planes (f16_split.cu conv2_f16x3: the image operand as
    hi / lo IMAGES behind one 4-D TMA map, the weights as a K-major plane matrix in the producer's (tap, channel) order), forced on
    small geometries with B200NN_F16CONV=2: the FFMA path's 1e-5 against the fp64 oracle, guard bands around outputs and the
    workspace the sizing call promises, and a different bit pattern from the in-kernel 3xTF32 split (so the planes really ran)."""
    rng = np.random.default_rng(H * 17 + C + k)
    x = np.maximum(rng.normal(size=(N, H, W, C)), 0).astype(np.float32)
    w = (rng.normal(size=(k, k, C, F)) / np.sqrt(k * k * C)).astype(np.float32)
    b = rng.normal(size=(F,)).astype(np.float32)
    geom, oh, ow = _conv_geom(K, N, H, W, C, F, k, k, 1, 1, "same")
    err = (1e-3 * rng.normal(size=(N, oh, ow, F))).astype(np.float32)      # a small-magnitude operand: the range scale matters
    xd, wd, bd, ed = K.dev(x), K.dev(w), K.dev(b), K.dev(err)
    x64, w64 = x.astype(np.float64), w.astype(np.float64)
    ref = O.conv2d_fwd(x64, w64, b.astype(np.float64), (1, 1), "same")
    ref = np.maximum(ref, 0) if act else ref
    rdx, rdw, rdb = O.conv2d_bwd(x64, w64, O.clip_l2(err.astype(np.float64)), (1, 1), "same")
    out = {}
    for mode in ("2", "0"):
        monkeypatch.setenv("B200NN_F16CONV", mode)
        G = Guarded(K)
        ws = G((max(K.ops.conv2d_ws_bytes(geom, 2) // 4, 1),))
        y = G((N, oh, ow, F))
        K.ops.conv2d_fwd(2, geom, xd, wd, bd, y, act, 0.0, ws)
        ss = K.ss(); K.ops.sumsq(ed, ss, 0)
        dx, dw, db = G(x.shape), G(w.shape), G((F,))
        K.ops.conv2d_bwd(2, geom, xd, wd, ed, ss, (0,), dx, dw, db, 1, 0.0, 1, 2, ws)
        G.check()
        assert nrel(K.host(y), ref) < 1e-5, mode
        assert nrel(K.host(dx), rdx * (x > 0)) < 1e-5, mode
        assert nrel(K.host(dw), rdw) < 1e-5, mode
        h = K.host(ss)
        assert abs(h[1] - np.sum(rdx ** 2)) < 1e-4 * np.sum(rdx ** 2)
        out[mode] = (K.host(y).copy(), K.host(dx).copy())
    assert not np.array_equal(out["2"][0], out["0"][0]) and not np.array_equal(out["2"][1], out["0"][1])


def _convt_geom(K, N, H, W, C, F, kh, kw, sh, sw, pad):
    oh, ow, pt, pb, pl, pr = O.convT_geometry(H, W, kh, kw, sh, sw, pad)
    return K.B.ConvTGeom(N, H, W, C, F, kh, kw, sh, sw, pt, pl, oh, ow), oh, ow


@pytest.mark.parametrize("gi", range(6))
def test_convt_golden(K, golden, gi):
    g = golden("ops")
    H, W, C, F, kh, kw, sh, sw, same = [int(v) for v in g[f"convT{gi}_cfg"]]
    geom, oh, ow = _convt_geom(K, 3, H, W, C, F, kh, kw, sh, sw, "same" if same else "valid")
    x, w, b, err = (K.dev(g[f"convT{gi}_{k}"]) for k in ("x", "w", "b", "err"))
    ws = K.empty((max(K.ops.convt_ws_bytes(geom) // 4, 1),))
    y = K.empty((3, oh, ow, F))
    K.ops.convt_fwd(0, geom, x, w, b.reshape(-1), y, 0, 0.0, ws)
    assert nrel(K.host(y), g[f"convT{gi}_y"]) < TOL
    dx, dw, db = K.empty((3, H, W, C)), K.empty((kh, kw, F, C)), K.empty((F,))
    K.ops.convt_bwd(0, geom, x, w, err, None, (), dx, dw, db, 0, 0.0, None, None, ws)
    assert nrel(K.host(dx), g[f"convT{gi}_dx"]) < TOL
    assert nrel(K.host(dw), g[f"convT{gi}_dw"]) < TOL
    assert nrel(K.host(db), g[f"convT{gi}_db"].reshape(-1)) < TOL


@pytest.mark.parametrize("N,H,W,C,F", [(8, 7, 7, 128, 64), (8, 14, 14, 64, 32)])
def test_convt_oracle(K, N, H, W, C, F):
    rng = np.random.default_rng(H + C)
    x = np.maximum(rng.normal(size=(N, H, W, C)), 0).astype(np.float32)
    w = (rng.normal(size=(3, 3, F, C)) / np.sqrt(9 * C)).astype(np.float32)
    b = rng.normal(size=(F,)).astype(np.float32)
    geom, oh, ow = _convt_geom(K, N, H, W, C, F, 3, 3, 2, 2, "same")
    err = rng.normal(size=(N, oh, ow, F)).astype(np.float32)
    xd, wd, bd, ed = K.dev(x), K.dev(w), K.dev(b), K.dev(err)
    ws = K.empty((max(K.ops.convt_ws_bytes(geom) // 4, 1),))
    y = K.empty((N, oh, ow, F))
    K.ops.convt_fwd(0, geom, xd, wd, bd, y, 1, 0.0, ws)
    ref = np.maximum(O.convT_fwd(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), (2, 2), "same"), 0)
    assert nrel(K.host(y), ref) < TOL
    ss = K.ss(); K.ops.sumsq(ed, ss, 0)
    dx, dw, db = K.empty(x.shape), K.empty(w.shape), K.empty((F,))
    K.ops.convt_bwd(0, geom, xd, wd, ed, ss, (0,), dx, dw, db, 1, 0.0, 1, 2, ws)
    rdx, rdw, rdb = O.convT_bwd(x.astype(np.float64), w.astype(np.float64), O.clip_l2(err.astype(np.float64)), (2, 2))
    assert nrel(K.host(dx), rdx * (x > 0)) < TOL
    assert nrel(K.host(dw), rdw) < TOL
    assert nrel(K.host(db), rdb.reshape(-1)) < TOL
    assert np.all(K.host(dx)[:, -1] == 0) and np.all(K.host(dx)[:, :, -1] == 0)  # Q12 edge skip


def test_batchnorm_golden(K, golden):
    g = golden("ops")
    x1 = K.dev(g["bn_x1"]); shp = g["bn_x1"].shape[1:]
    P = int(np.prod(shp))
    gamma, beta = K.B.to_device(np.ones(P)), K.B.zeros((P,))
    rm, rv = K.B.zeros((P,)), K.B.to_device(np.ones(P))
    sm, si, y = K.empty((P,)), K.empty((P,)), K.empty(x1.shape)
    K.ops.bn_fwd_train(x1, gamma, beta, rm, rv, sm, si, y, 0.9, 1e-5)
    assert nrel(K.host(y), g["bn_y1"]) < TOL
    assert nrel(K.host(rm).reshape(shp), g["bn_rm1"]) < TOL and nrel(K.host(rv).reshape(shp), g["bn_rv1"]) < TOL
    e1 = K.dev(g["bn_e1"]); dx, dg, dbt = K.empty(x1.shape), K.empty((P,)), K.empty((P,))
    K.ops.bn_bwd(x1, e1, None, (), gamma, sm, si, dx, dg, dbt, 0, 0.0, None, None)
    assert nrel(K.host(dx), g["bn_dx1"]) < 2e-5
    assert nrel(K.host(dg).reshape(shp), g["bn_dgamma1"]) < TOL and nrel(K.host(dbt).reshape(shp), g["bn_dbeta1"]) < TOL
    x2 = K.dev(g["bn_x2"])
    K.ops.bn_fwd_train(x2, gamma, beta, rm, rv, sm, si, y, 0.9, 1e-5)
    assert nrel(K.host(y), g["bn_y2"]) < TOL
    assert nrel(K.host(rm).reshape(shp), g["bn_rm2"]) < TOL and nrel(K.host(rv).reshape(shp), g["bn_rv2"]) < TOL
    K.ops.bn_fwd_infer(x1, gamma, beta, rm, rv, y, 1e-5)
    assert nrel(K.host(y), g["bn_yinf"]) < TOL


@pytest.mark.parametrize("Bn,shape", [(256, (28, 28, 32)), (64, (8, 8, 5)), (7, (33,))])
def test_batchnorm_oracle(K, Bn, shape):
    rng = np.random.default_rng(Bn)
    x = np.maximum(rng.normal(size=(Bn,) + shape) * 0.05, 0).astype(np.float32)  # post-ReLU conv-like input
    e = rng.normal(size=x.shape).astype(np.float32)
    P = int(np.prod(shape))
    gamma = (1 + 0.1 * rng.normal(size=P)).astype(np.float32); beta = (0.1 * rng.normal(size=P)).astype(np.float32)
    xd, ed, gd, bd = K.dev(x), K.dev(e), K.dev(gamma), K.dev(beta)
    rm, rv = K.B.zeros((P,)), K.B.to_device(np.ones(P))
    sm, si, y = K.empty((P,)), K.empty((P,)), K.empty(x.shape)
    K.ops.bn_fwd_train(xd, gd, bd, rm, rv, sm, si, y, 0.9, 1e-5)
    g64, b64 = gamma.astype(np.float64).reshape(shape), beta.astype(np.float64).reshape(shape)
    ry, rrm, rrv, cache = O.bn_fwd(x.astype(np.float64), g64, b64, np.zeros(shape), np.ones(shape), 0.9, 1e-5, True)
    assert nrel(K.host(y), ry) < TOL
    assert nrel(K.host(rv).reshape(shape), rrv) < TOL
    ss = K.ss(); K.ops.sumsq(ed, ss, 0)
    dx, dg, dbt = K.empty(x.shape), K.empty((P,)), K.empty((P,))
    K.ops.bn_bwd(xd, ed, ss, (0,), gd, sm, si, dx, dg, dbt, 1, 0.0, 1, 2)
    rdx, rdg, rdb = O.bn_bwd(O.clip_l2(e.astype(np.float64)), g64, cache, 1e-5)
    assert nrel(K.host(dx), rdx * (x > 0)) < 2e-5
    assert nrel(K.host(dg).reshape(shape), rdg) < 2e-5
    h = K.host(ss)
    assert abs(h[1] - np.sum(rdx ** 2)) < 1e-4 * np.sum(rdx ** 2)
    assert abs(h[2] - np.sum((rdx * (x > 0)) ** 2)) < 1e-4 * np.sum(rdx ** 2)


@pytest.mark.parametrize("gi", range(4))
def test_maxpool_golden(K, golden, gi):
    g = golden("ops")
    H, W, C, ph, pw, sh, sw = [int(v) for v in g[f"pool{gi}_cfg"]]
    oh, ow = (H - ph) // sh + 1, (W - pw) // sw + 1
    geom = K.B.PoolGeom(3, H, W, C, ph, pw, sh, sw, 0, 0, oh, ow)
    x, err = K.dev(g[f"pool{gi}_x"]), K.dev(g[f"pool{gi}_err"])
    y = K.empty((3, oh, ow, C)); dx = K.empty(x.shape)
    K.ops.maxpool_fwd(geom, x, y)
    assert np.array_equal(K.host(y), _f32(g[f"pool{gi}_y"]))  # bit-exact selection
    K.ops.maxpool_bwd(geom, x, y, err, None, (), dx, None)
    assert nrel(K.host(dx), g[f"pool{gi}_dx"]) < TOL


def test_maxpool_ties_all_zero(K):
    geom = K.B.PoolGeom(1, 2, 2, 1, 2, 2, 2, 2, 0, 0, 1, 1)
    x = K.B.zeros((1, 2, 2, 1)); y = K.empty((1, 1, 1, 1)); dx = K.empty((1, 2, 2, 1))
    K.ops.maxpool_fwd(geom, x, y)
    K.ops.maxpool_bwd(geom, x, y, K.dev(np.ones((1, 1, 1, 1))), None, (), dx, None)
    assert np.array_equal(K.host(dx).ravel(), [1, 1, 1, 1])  # Q6: every tied maximum receives the gradient


def test_adam_sgd_golden(K, golden):
    g = golden("ops")
    torch = K.torch
    p = [K.dev(g[f"adam_p{i}_init"]) for i in range(4)]
    gs = [K.dev(g[f"adam_g{i}"]) for i in range(6)]
    m = [K.B.zeros(t.shape) for t in p]; v = [K.B.zeros(t.shape) for t in p]
    hyper = K.B.to_device(np.array([1e-3, 0.9, 0.999, 1e-8]), dtype=torch.float64)
    t = K.B.zeros((1,), dtype=torch.int64)
    # reference: opt.update(5, w0, g0, b0, g1); update(3, w1, g2, b1, g3); update(5, w0, g4, b0, g5) -> t = 1, 2, 3
    for (pi, gi) in (((0, 1), (0, 1)), ((2, 3), (2, 3)), ((0, 1), (4, 5))):
        plan = K.ops.OptPlan([(p[pi[0]], gs[gi[0]], m[pi[0]], v[pi[0]], 1, 0), (p[pi[1]], gs[gi[1]], m[pi[1]], v[pi[1]], 1, 0)], 1)
        K.ops.step_begin(None, None, t, 1)
        K.ops.adam_multi(plan, hyper, t)
    for i in range(4):
        assert nrel(K.host(p[i]), g[f"adam_p{i}_final"]) < TOL
    assert nrel(K.host(m[0]), g["adam_m_w5"]) < TOL and nrel(K.host(v[0]), g["adam_v_w5"]) < TOL
    assert int(K.host(t.double())[0]) == int(g["adam_t"])
    w = K.dev(g["sgd_w_init"]); gg = K.dev(g["sgd_g"])
    plan = K.ops.OptPlan([(w, gg, None, None, 1, 0)], 1)
    K.ops.sgd_multi(plan, K.B.to_device(np.array([0.05, 0, 0, 0]), dtype=torch.float64))
    assert nrel(K.host(w), g["sgd_w_final"]) < TOL


def test_adam_clip_and_large(K):
    rng = np.random.default_rng(3)
    n = 100_003
    p0 = rng.normal(size=n).astype(np.float32); g0 = rng.normal(size=n).astype(np.float32)
    pd, gd = K.dev(p0), K.dev(g0)
    md, vd = K.B.zeros((n,)), K.B.zeros((n,))
    hyper = K.B.to_device(np.array([1e-3, 0.9, 0.999, 1e-8]), dtype=K.torch.float64)
    t = K.B.zeros((1,), dtype=K.torch.int64)
    plan = K.ops.OptPlan([(pd, gd, md, vd, 2, 1)], 3)
    ref = O.OracleAdam(); ref.t = 0
    p64 = p0.astype(np.float64)
    for step in range(2):
        K.ops.step_begin(None, None, t, 3)
        K.ops.adam_multi(plan, hyper, t)
        ref.t = step * 3 + 1  # oracle increments to t = step*3 + 2 inside update
        ref.update(0, p64, O.clip_l2(g0.astype(np.float64)))
    assert nrel(K.host(pd), p64) < TOL
    assert nrel(K.host(md), ref.m_w[0]) < TOL and nrel(K.host(vd), ref.v_w[0]) < TOL


def test_gather_rows(K):
    rng = np.random.default_rng(0)
    src = rng.normal(size=(50, 7, 9)).astype(np.float32)
    idx = rng.permutation(50)[:20]
    d = K.empty((20, 7, 9))
    K.ops.gather_rows(K.dev(src), K.B.to_device(idx, dtype=K.torch.int64), d)
    assert np.array_equal(K.host(d), src[idx].astype(np.float64))


@pytest.mark.parametrize("Bn,H,W,C", [(256, 28, 28, 32), (64, 8, 8, 64), (130, 4, 6, 32), (300, 4, 4, 32), (1024, 4, 4, 32), (5, 2, 2, 32)])
def test_fused_bn_pool_block(K, Bn, H, W, C):
    """BatchNorm(train)+MaxPool forward and MaxPool+BatchNorm+ReLU backward as single kernels (cluster/DSMEM batch
    reduction for B > 128) against the unfused oracle chain incl. the three clip points between the layers."""
    rng = np.random.default_rng(Bn + H)
    x = np.maximum(rng.normal(size=(Bn, H, W, C)) * 0.05 - 0.01, 0).astype(np.float32)
    x[:, 0, 0, 0] = 0.0  # an all-dead window column: 4-way ties in the pool (Q6)
    x[:, 0, 1, 0] = 0.0; x[:, 1, 0, 0] = 0.0; x[:, 1, 1, 0] = 0.0
    P = H * W * C
    gamma = (1 + 0.1 * rng.normal(size=P)).astype(np.float32); beta = (0.1 * rng.normal(size=P)).astype(np.float32)
    beta.reshape(H, W, C)[:2, :2, 0] = 0.25; gamma.reshape(H, W, C)[:2, :2, 0] = 1.0
    xd, gd, bd = K.dev(x), K.dev(gamma), K.dev(beta)
    rm, rv = K.B.zeros((P,)), K.B.to_device(np.ones(P))
    G = Guarded(K)                      # sentinel guard bands around every output: the kernels index shared / global memory by hand
    sm, si = G((P,)), G((P,))
    yp = G((Bn, H // 2, W // 2, C))
    K.ops.bn_pool_fwd_train(xd, gd, bd, rm, rv, sm, si, yp, 0.9, 1e-5)
    G.check()
    shp = (H, W, C)
    g64, b64 = gamma.astype(np.float64).reshape(shp), beta.astype(np.float64).reshape(shp)
    x64 = x.astype(np.float64)
    ybn, rrm, rrv, cache = O.bn_fwd(x64, g64, b64, np.zeros(shp), np.ones(shp), 0.9, 1e-5, True)
    ryp = O.maxpool_fwd(ybn, (2, 2), (2, 2))
    assert nrel(K.host(yp), ryp) < TOL
    assert nrel(K.host(rm).reshape(shp), rrm) < TOL and nrel(K.host(rv).reshape(shp), rrv) < TOL
    e = rng.normal(size=ryp.shape).astype(np.float32)
    ed = K.dev(e); ss = K.ss(); K.ops.sumsq(ed, ss, 0)
    r0, dg, dbt = G(x.shape), G((P,)), G((P,))
    K.ops.pool_bn_act_bwd(xd, ed, ss, (0,), gd, bd, sm, si, r0, dg, dbt, 1, 0.0, 1, 2, 3)
    G.check()
    # reference chain: clip -> pool bwd -> clip -> bn bwd -> clip -> relu' ; the kernel emits the unclipped
    # intermediates plus their sums of squares, the chain (1, 2, 3) reproduces the clips
    U = O.maxpool_bwd(ybn, O.clip_l2(e.astype(np.float64)), (2, 2), (2, 2))
    D, rdg, rdb = O.bn_bwd(U, g64, cache, 1e-5)             # bn_bwd is linear: D(clip(U)) = s_U * D(U)
    R = D * (x64 > 0)
    h = K.host(ss)
    assert abs(h[1] - np.sum(U ** 2)) < 1e-4 * np.sum(U ** 2)
    assert abs(h[2] - np.sum(D ** 2)) < 1e-4 * np.sum(D ** 2)
    assert abs(h[3] - np.sum(R ** 2)) < 1e-4 * np.sum(R ** 2)
    assert nrel(K.host(r0), R) < 2e-5
    assert nrel(K.host(dg).reshape(shp), rdg) < 2e-5 and nrel(K.host(dbt).reshape(shp), rdb) < 2e-5
    out = K.empty(x.shape)
    K.ops.scale_apply(r0, ss, (1, 2, 3), out)
    ref = O.clip_l2(O.bn_bwd(O.clip_l2(U), g64, cache, 1e-5)[0]) * (x64 > 0)
    assert nrel(K.host(out), O.clip_l2(ref)) < 2e-5
    assert np.all(U[:, :2, :2, 0] == O.clip_l2(e.astype(np.float64))[:, :1, :1, 0][:, :, :, None].reshape(Bn, 1, 1))  # 4-way tie


@pytest.mark.parametrize("M,N,Kd", [(256, 64, 6272), (512, 512, 512), (300, 260, 200), (128, 32, 64), (64, 4096, 512),
                                    (1000, 128, 96), (256, 256, 2048)])
def test_dense_tf32_tcgen05(K, M, N, Kd):
    """tcgen05 kind::tf32 path (TMA + TMEM): all three GEMMs of a Dense layer (K-major x MN-major, K x K, MN x MN),
    split-K and M/N/K tails, within the north star's rel 2e-3 of the fp64 oracle."""
    assert K.B.lib().b200nn_has_tcgen05() == 1
    rng = np.random.default_rng(M + N + Kd)
    x = np.maximum(rng.normal(size=(M, Kd)), 0).astype(np.float32)
    w = (rng.normal(size=(Kd, N)) / np.sqrt(Kd)).astype(np.float32)
    b = rng.normal(size=(N,)).astype(np.float32)
    err = rng.normal(size=(M, N)).astype(np.float32)
    xd, wd, bd, ed = K.dev(x), K.dev(w), K.dev(b), K.dev(err)
    ws = K.empty((max(K.ops.dense_ws_bytes(M, N, Kd) // 4, 1),))
    y = K.empty((M, N))
    K.ops.dense_fwd(1, xd, wd, bd, y, 1, 0.0, ws)
    x64, w64 = x.astype(np.float64), w.astype(np.float64)
    ref = np.maximum(x64 @ w64 + b, 0)
    assert nrel(K.host(y), ref) < 2e-3
    ss = K.ss(); K.ops.sumsq(ed, ss, 0)
    dx, dw, db = K.empty((M, Kd)), K.empty((Kd, N)), K.empty((N,))
    K.ops.dense_bwd(1, xd, wd, ed, ss, (0,), dx, dw, db, 1, 0.0, 1, 2, ws)
    e64 = O.clip_l2(err.astype(np.float64))
    rdx, rdw, rdb = O.dense_bwd(x64, w64, e64)
    assert nrel(K.host(dx), rdx * (x > 0)) < 2e-3
    assert nrel(K.host(dw), rdw) < 2e-3
    assert nrel(K.host(db), rdb.reshape(-1)) < 1e-5
    h = K.host(ss)
    assert abs(h[1] - np.sum(rdx ** 2)) < 5e-3 * np.sum(rdx ** 2)


@pytest.mark.parametrize("M,N,Kd", [(256, 64, 6272), (512, 512, 512), (300, 260, 200), (128, 32, 64), (64, 4096, 512),
                                    (1000, 128, 96), (256, 256, 2048), (256, 4096, 4096),
                                    # large contractions take the fp16 hi/lo-plane path (f16_split.cu + gemm_tc2.cuh F16 mode)
                                    (2048, 2048, 2048), (2052, 1028, 4100)])
def test_dense_tf32x3_tcgen05(K, M, N, Kd):
    """Error-compensated tensor-core path (math mode 2): raw + lo operand tiles, three tcgen05 MMAs per k-step. Must
    reach the FFMA path's accuracy (rel 1e-5 of the fp64 oracle), not single-pass TF32's 2e-3."""
    rng = np.random.default_rng(M + N + Kd + 1)
    x = np.maximum(rng.normal(size=(M, Kd)), 0).astype(np.float32)
    w = (rng.normal(size=(Kd, N)) / np.sqrt(Kd)).astype(np.float32)
    b = rng.normal(size=(N,)).astype(np.float32)
    err = rng.normal(size=(M, N)).astype(np.float32)
    xd, wd, bd, ed = K.dev(x), K.dev(w), K.dev(b), K.dev(err)
    G = Guarded(K)
    ws = G((max(K.ops.dense_ws_bytes(M, N, Kd, 2) // 4, 1),))
    y = G((M, N))
    K.ops.dense_fwd(2, xd, wd, bd, y, 1, 0.0, ws)
    x64, w64 = x.astype(np.float64), w.astype(np.float64)
    assert nrel(K.host(y), np.maximum(x64 @ w64 + b, 0)) < 1e-5
    ss = K.ss(); K.ops.sumsq(ed, ss, 0)
    dx, dw, db = G((M, Kd)), G((Kd, N)), G((N,))
    K.ops.dense_bwd(2, xd, wd, ed, ss, (0,), dx, dw, db, 1, 0.0, 1, 2, ws)
    G.check()
    rdx, rdw, rdb = O.dense_bwd(x64, w64, O.clip_l2(err.astype(np.float64)))
    assert nrel(K.host(dx), rdx * (x > 0)) < 1e-5
    assert nrel(K.host(dw), rdw) < 1e-5
    assert nrel(K.host(db), rdb.reshape(-1)) < 1e-5
    # and it is far closer to the oracle than single-pass TF32 on the same inputs
    y1 = K.empty((M, N)); K.ops.dense_fwd(1, xd, wd, bd, y1, 1, 0.0, ws)
    if N > 32:    # N <= 32 is the FFMA small-N kernel in every math mode
        assert nrel(K.host(y), np.maximum(x64 @ w64 + b, 0)) * 20 < nrel(K.host(y1), np.maximum(x64 @ w64 + b, 0))


@pytest.mark.parametrize("Bn,H,W,C,F,act", [(256, 28, 28, 1, 32, "relu"), (64, 32, 32, 3, 64, "relu"), (37, 8, 6, 2, 8, "leakyrelu"),
                                            (200, 4, 4, 1, 16, "relu"), (150, 4, 6, 3, 8, "tanh"), (16, 6, 4, 1, 8, "linear")])
def test_convblock_first_block(K, Bn, H, W, C, F, act):
    """Fused first block (csrc/conv_block.cu): Conv2D(3x3 'same') + activation + BatchNorm(train) + MaxPool 2x2 with the conv
    output never materialised, forward and backward (weight / bias gradient, d_gamma / d_beta, the three clip slots, the
    optional error w.r.t. the conv output), against the unfused oracle chain. Also the two data-parallel halves with a
    world of one (mode 1 -> mode 2 must reproduce mode 0)."""
    rng = np.random.default_rng(Bn * 7 + H + C)
    x = rng.random((Bn, H, W, C)).astype(np.float32)
    w = (rng.normal(size=(3, 3, C, F)) * 0.3).astype(np.float32)
    b = (rng.normal(size=(1, F)) * 0.1).astype(np.float32)
    P = H * W * F
    shp = (H, W, F)
    gamma = (1 + 0.1 * rng.normal(size=P)).astype(np.float32); beta = (0.1 * rng.normal(size=P)).astype(np.float32)
    g = K.B.ConvGeom(Bn, H, W, C, F, 3, 3, 1, 1, 1, 1, H, W)
    kind, alpha = ACTS[act], 0.01
    assert K.ops.convblock_supported(g, kind, alpha)
    xd, wd, bd, gd, btd = K.dev(x), K.dev(w), K.dev(b.reshape(-1)), K.dev(gamma), K.dev(beta)
    x64, w64, b64 = x.astype(np.float64), w.astype(np.float64), b.astype(np.float64)
    g64, bt64 = gamma.astype(np.float64).reshape(shp), beta.astype(np.float64).reshape(shp)
    pre = O.conv2d_fwd(x64, w64, b64, (1, 1), "same")
    a = O.act_fwd(act, pre, alpha)
    ybn, rrm, rrv, cache = O.bn_fwd(a, g64, bt64, np.zeros(shp), np.ones(shp), 0.9, 1e-5, True)
    ryp = O.maxpool_fwd(ybn, (2, 2), (2, 2))
    e = rng.normal(size=ryp.shape).astype(np.float32)
    U = O.maxpool_bwd(ybn, O.clip_l2(e.astype(np.float64)), (2, 2), (2, 2))
    D, rdg, rdb = O.bn_bwd(U, g64, cache, 1e-5)
    R = O.act_bwd(act, pre, D, alpha)
    # reference: clip(U) -> bn bwd -> clip -> act bwd -> clip -> conv bwd
    Rc = O.clip_l2(O.act_bwd(act, pre, O.clip_l2(O.bn_bwd(O.clip_l2(U), g64, cache, 1e-5)[0]), alpha))
    _, rdw, rdbias = O.conv2d_bwd(x64, w64, Rc, (1, 1), "same")
    G = Guarded(K)                      # guard bands around every output and the workspace of the first-block kernels
    ws = G((max(K.ops.convblock_ws_bytes(g) // 4, 1),))
    for modes in ((0,), (1, 2)):
        rm, rv = K.B.zeros((P,)), K.B.to_device(np.ones(P))
        sm, si = G((P,)), G((P,))
        yp = G((Bn, H // 2, W // 2, F))
        stats = K.B.zeros((2 * P,), dtype=K.torch.float64)
        for md in modes:
            K.ops.convblock_fwd(g, xd, wd, bd, kind, alpha, gd, btd, rm, rv, sm, si, stats, md, Bn, yp if md != 1 else None, 0.9, 1e-5)
        assert nrel(K.host(yp), ryp) < 2e-5, modes
        assert nrel(K.host(rm).reshape(shp), rrm) < TOL and nrel(K.host(rv).reshape(shp), rrv) < TOL
        ed = K.dev(e); ss = K.ss(); K.ops.sumsq(ed, ss, 0)
        dg, dbt, dw, db, r = G((P,)), G((P,)), G((3, 3, C, F)), G((F,)), G((Bn, H, W, F))
        s2 = 3 if act != "linear" else None
        cout = (1, 2, 3) if act != "linear" else (1, 2)
        if modes == (0,):
            K.ops.convblock_bwd(g, xd, wd, bd, kind, alpha, gd, btd, sm, si, ed, ss, (0,), dg.data_ptr(), dbt.data_ptr(), dw, db, r, 0, Bn,
                                1, 2, s2, cout, ws)
        else:
            K.ops.convblock_bwd(g, xd, wd, bd, kind, alpha, gd, btd, sm, si, ed, ss, (0,), dg.data_ptr(), dbt.data_ptr(), None, None, None,
                                1, Bn, 1, None, None, (), None)
            K.ops.convblock_bwd(g, xd, wd, bd, kind, alpha, gd, btd, sm, si, ed, ss, (0,), dg.data_ptr(), dbt.data_ptr(), dw, db, r, 2, Bn,
                                None, 2, s2, cout, ws)
        h = K.host(ss)
        assert abs(h[1] - np.sum(U ** 2)) < 1e-4 * np.sum(U ** 2)
        assert abs(h[2] - np.sum(D ** 2)) < 1e-4 * np.sum(D ** 2)
        if s2 is not None:
            assert abs(h[3] - np.sum(R ** 2)) < 1e-4 * np.sum(R ** 2)
        assert nrel(K.host(r), R) < 3e-5
        assert nrel(K.host(dg).reshape(shp), rdg) < 3e-5 and nrel(K.host(dbt).reshape(shp), rdb) < 3e-5
        assert nrel(K.host(dw), rdw) < 3e-5, modes
        # with a linear activation the bias gradient is sum over the batch of a BatchNorm backward output == 0 identically
        assert np.abs(K.host(db) - rdbias.reshape(-1)).max() < 3e-5 * max(np.abs(rdbias).max(), np.abs(rdw).max())
        G.check()


def test_convblock_peer_fused_world_of_one(K):
    """Mode 3 of the first-block kernels (batch reductions merged across GPUs through peer memory INSIDE the kernel, persistent
    launch) with a world of one: the rank exchanges with itself through the CUDA-IPC symmetric buffer, so flags, parity
    double-buffering, round counters and the merge arithmetic all run; results must equal mode 0. Two rounds (both parities)."""
    import os
    import torch.distributed as dist
    from neuralnetlib_b200 import parallel as P
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("gloo", rank=0, world_size=1)
    comm = P.Communicator(use_peer="force")
    assert comm.peer
    Bn, H, W, C, F = 96, 8, 8, 1, 16
    rng = np.random.default_rng(5)
    x = rng.random((Bn, H, W, C)).astype(np.float32)
    w = (rng.normal(size=(3, 3, C, F)) * 0.3).astype(np.float32)
    b = (rng.normal(size=(F,)) * 0.1).astype(np.float32)
    Pn = H * W * F
    gamma = (1 + 0.1 * rng.normal(size=Pn)).astype(np.float32); beta = (0.1 * rng.normal(size=Pn)).astype(np.float32)
    g = K.B.ConvGeom(Bn, H, W, C, F, 3, 3, 1, 1, 1, 1, H, W)
    xd, wd, bd, gd, btd = K.dev(x), K.dev(w), K.dev(b), K.dev(gamma), K.dev(beta)
    e = rng.normal(size=(Bn, H // 2, W // 2, F)).astype(np.float32)
    ed = K.dev(e)
    ws = K.empty((max(K.ops.convblock_ws_bytes(g) // 4, 1),))
    nbytes = K.ops.convblock_region_bytes(g, 1)
    rf, rb = comm.region(nbytes), comm.region(nbytes)
    pf, pb = (comm.handle,) + rf, (comm.handle,) + rb

    def run(mode, peer_f, peer_b):
        rm, rv = K.B.zeros((Pn,)), K.B.to_device(np.ones(Pn))
        sm, si = K.empty((Pn,)), K.empty((Pn,))
        yp = K.empty((Bn, H // 2, W // 2, F))
        K.ops.convblock_fwd(g, xd, wd, bd, 1, 0.0, gd, btd, rm, rv, sm, si, None, mode, Bn, yp, 0.9, 1e-5, peer_f)
        ss = K.ss(); K.ops.sumsq(ed, ss, 0)
        dg, dbt, dw, db = K.empty((Pn,)), K.empty((Pn,)), K.empty((3, 3, C, F)), K.empty((F,))
        K.ops.convblock_bwd(g, xd, wd, bd, 1, 0.0, gd, btd, sm, si, ed, ss, (0,), dg.data_ptr(), dbt.data_ptr(), dw, db, None, mode, Bn,
                            1, 2, 3, (1, 2, 3), ws, peer_b)
        return [K.host(t) for t in (yp, rm, rv, dg, dbt, dw, db, ss)]
    ref = run(0, None, None)
    for rnd in range(3):
        got = run(3, pf, pb)
        for a_, b_ in zip(got, ref):
            assert nrel(a_, b_) < 2e-6, rnd
    assert comm.timeouts() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("n", [8192, 200704, 200705, 12288 + 3, 5_000_000])
def test_upload_pinned_and_pageable(K, n):
    """b200nn_upload (the batch upload of Sequential.train_on_batch) from pinned and from pageable host memory, odd byte counts
    included: the same bytes on the device, nothing written outside the destination."""
    rng = np.random.default_rng(n)
    a = rng.normal(size=(n,)).astype(np.float32)
    pinned = K.torch.empty(n, dtype=K.torch.float32).pin_memory()
    pinned.numpy()[:] = a
    for src in (pinned.numpy(), a):
        G = Guarded(K)
        dst = G((n,))
        K.B.check(K.B.lib().b200nn_upload(dst.data_ptr(), src.__array_interface__["data"][0], src.nbytes, K.B.stream()))
        G.check()
        assert np.array_equal(dst.cpu().numpy(), a)
