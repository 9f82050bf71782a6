"""CPU: the arithmetic model behind math mode `tf32x3` (csrc/gemm_tcgen05.cuh), restated in NumPy.

The tensor core reads an fp32 word as TF32 by truncating the 13 low mantissa bits. The kernel leaves the staged operand
untouched (so hi = trunc(x)), writes lo = rn_tf32(x - hi) next to it and issues  A.B + A_lo.B + A.B_lo.  This test pins what
that buys and what it does not:
  * with an exact accumulate the compensated product is within ~1e-6 of the fp64 product of the fp32 inputs, against ~3e-4
    for single-pass TF32 -- also for all-positive (ReLU) inputs, where the dropped lo.lo term is same-signed;
  * a truncating fp32 accumulate (what the tensor core does; modelled here per k-step of 8) costs more than everything the
    split leaves behind once the chain is long, which is why the kernel keeps the lo products in a second accumulator and
    cuts chains at 2048 k (X3_CHAIN_KB).
The GPU side of the claim is tests/test_gpu_ops.py::test_dense_tf32x3_tcgen05 / test_conv2d_tcgen05_implicit_gemm."""
import numpy as np


def trunc_tf32(a):
    return (np.asarray(a, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def rn_tf32(a):   # cvt.rna.tf32.f32: round to nearest, ties away from zero (on the magnitude bits)
    u = np.asarray(a, np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def nrel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def split(a):
    hi = trunc_tf32(a)
    lo = rn_tf32(a - hi)          # a - hi is exact in fp32 (13 significant bits)
    return hi.astype(np.float64), lo.astype(np.float64)


def test_compensated_product_recovers_fp32_accuracy():
    rng = np.random.default_rng(0)
    M, N, K = 64, 48, 4096
    x = np.maximum(rng.normal(size=(M, K)), 0).astype(np.float32)          # ReLU output: all products' lo parts same-signed
    w = (rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)
    exact = x.astype(np.float64) @ w.astype(np.float64)
    xh, xl = split(x)
    wh, wl = split(w)
    single = xh @ wh
    comp = xh @ wh + xl @ wh + xh @ wl
    assert 5e-5 < nrel(single, exact) < 2e-3          # single-pass TF32: the north star's 2e-3 per contraction, no better
    assert nrel(comp, exact) < 2e-6                   # compensated: fp32-level
    # what is dropped is lo.lo, a same-signed ~2^-22 fraction of every product: a tiny relative scaling of the result
    # (the rest of the residual is the rounding of lo itself to TF32, ~2^-23 per operand)
    dropped = xl @ wl
    assert nrel(comp + dropped, exact) < 5e-7
    pos = np.abs(exact) > 0.1 * np.abs(exact).max()
    assert np.all(np.abs(dropped[pos] / exact[pos]) < 2.0 ** -19)


def _chained(xh, wh, step, chain):
    """sum over k in steps of `step`, added into an fp32 accumulator with truncation toward zero; a fresh accumulator every
    `chain` k, the partial sums added in round-to-nearest fp32 (the split-K reduce)."""
    M, K = xh.shape
    total = np.zeros((M, wh.shape[1]), np.float32)
    for c0 in range(0, K, chain):
        acc = np.zeros_like(total, dtype=np.float64)
        for k0 in range(c0, min(K, c0 + chain), step):
            s = acc + xh[:, k0:k0 + step] @ wh[k0:k0 + step]
            f = s.astype(np.float32)                                          # nearest ...
            over = np.abs(f.astype(np.float64)) > np.abs(s)                   # ... pulled back toward zero where it overshot
            f = np.where(over, np.nextafter(f, np.float32(0)), f)
            acc = f.astype(np.float64)
        total = (total.astype(np.float64) + acc).astype(np.float32)
    return total.astype(np.float64)


def test_truncating_accumulate_is_the_dominant_residual_on_long_chains():
    rng = np.random.default_rng(1)
    M, N, K = 16, 16, 4096
    x = np.maximum(rng.normal(size=(M, K)), 0).astype(np.float32)
    w = np.abs(rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)      # all-positive sums: the truncation bias adds up
    xh, xl = split(x)
    wh, wl = split(w)
    exact = x.astype(np.float64) @ w.astype(np.float64)
    lo_terms = xl @ wh + xh @ wl                                             # second accumulator: 2^-11 of the main one
    one_chain = _chained(xh, wh, 8, K) + lo_terms
    cut = _chained(xh, wh, 8, 2048) + lo_terms
    assert nrel(one_chain, exact) > 4 * nrel(xh @ wh + lo_terms, exact)      # the accumulate, not the split, sets the error
    assert nrel(cut, exact) < nrel(one_chain, exact)                         # shorter chains help ...
    assert nrel(cut, exact) < 1e-5                                           # ... and 2048-k chains meet the parity bar
