"""CPU: the C-ABI shared library loads and exports every symbol include/b200nn.h declares
(no compute calls: there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200nn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200nn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from neuralnetlib_b200 import _backend as B
    lib = ctypes.CDLL(B.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200nn.h but not exported by libb200nn.so"
    assert sorted(B.exported_symbols()) == declared, set(B.exported_symbols()) ^ set(declared)
    assert lib.b200nn_version() >= 100


def test_no_cpu_fallback_without_cuda():
    import pytest
    import torch
    from neuralnetlib_b200 import _backend as B
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B.device()


def test_chain_packing():
    from neuralnetlib_b200 import _ops
    assert _ops.pack_chain(()) == 0
    assert _ops.pack_chain((5,)) == 1 | (5 << 8)
    assert _ops.pack_chain((1, 2, 3)) == 3 | (1 << 8) | (2 << 20) | (3 << 32)


def test_workspace_sizing_per_math_mode():
    """Host-only entry points (no device work): the tensor-core modes never need less workspace than the FFMA path, the
    compensated mode (chain-limited split-K) needs room for its partial sums, and geometries the tensor-core convolution
    does not take keep the FFMA sizing in fp32 and get room for their column buffers in the tensor-core modes."""
    from neuralnetlib_b200 import _backend as B
    lib = B.lib()
    for M, N, K in ((256, 4096, 4096), (8192, 4096, 4096), (256, 64, 6272), (32, 128, 784)):
        base = lib.b200nn_dense_ws_bytes(M, N, K)
        assert lib.b200nn_dense_ws_bytes_math(0, M, N, K) == base
        assert lib.b200nn_dense_ws_bytes_math(1, M, N, K) == base
        assert lib.b200nn_dense_ws_bytes_math(2, M, N, K) >= base
    # K = 4096 exceeds one 2048-k accumulation chain: at least two partial tiles of [M, N] floats
    assert lib.b200nn_dense_ws_bytes_math(2, 8192, 4096, 4096) >= 2 * 8192 * 4096 * 4
    assert lib.b200nn_dense_ws_bytes_math(2, 8192, 4096, 4096) <= (512 << 20) + 8192 * 4096 * 4   # bounded workspace
    ok = B.ConvGeom(64, 16, 16, 64, 128, 3, 3, 1, 1, 1, 1, 16, 16)          # stride 1, 'same', power-of-two image, C, F % 32 == 0
    odd = B.ConvGeom(64, 28, 28, 64, 128, 3, 3, 1, 1, 1, 1, 28, 28)         # 28 x 28: im2col + Dense kernels in the tensor-core modes
    strided = B.ConvGeom(64, 16, 16, 64, 128, 3, 3, 2, 2, 1, 1, 8, 8)
    for g in (odd, strided):
        assert lib.b200nn_conv2d_ws_bytes_math(0, ctypes.byref(g)) == lib.b200nn_conv2d_ws_bytes(ctypes.byref(g))
        col = g.N * g.OH * g.OW * g.kh * g.kw * g.C * 4
        for math in (1, 2):   # column buffer (forward) + its gradient (backward) + the permuted weights and their gradient
            assert lib.b200nn_conv2d_ws_bytes_math(math, ctypes.byref(g)) >= 2 * col
    assert lib.b200nn_conv2d_ws_bytes_math(0, ctypes.byref(ok)) == lib.b200nn_conv2d_ws_bytes(ctypes.byref(ok))
    for math in (1, 2):
        assert lib.b200nn_conv2d_ws_bytes_math(math, ctypes.byref(ok)) >= lib.b200nn_conv2d_ws_bytes(ctypes.byref(ok))


def test_workspace_sizing_fp16_plane_convolution(monkeypatch):
    """Host-only: a tf32x3 convolution that takes the fp16-plane form (config 3's conv2 at batch 1024) gets room for the hi / lo
    image planes of its larger operand and both weight plane matrices, plus the 256-byte tail that carries max |err| from the
    column-sum pass to the data gradient; B200NN_F16CONV=0 switches the planes off, and a geometry below the worth threshold
    keeps the in-kernel split's sizing unless the planes are forced (B200NN_F16CONV=2)."""
    from neuralnetlib_b200 import _backend as B
    lib = B.lib()
    big = B.ConvGeom(1024, 16, 16, 64, 128, 3, 3, 1, 1, 1, 1, 16, 16)
    small = B.ConvGeom(8, 16, 16, 64, 128, 3, 3, 1, 1, 1, 1, 16, 16)
    P, taps = 1024 * 16 * 16, 9
    planes = (2 * P * max(64, 128) + 2 * taps * 64 * 128) * 2      # one image operand at a time (x forward, err backward)
    monkeypatch.delenv("B200NN_F16CONV", raising=False)
    with_planes = lib.b200nn_conv2d_ws_bytes_math(2, ctypes.byref(big))
    assert with_planes >= planes + 256
    assert with_planes % 256 == 0
    monkeypatch.setenv("B200NN_F16CONV", "0")
    without = lib.b200nn_conv2d_ws_bytes_math(2, ctypes.byref(big))
    assert without < planes and without >= lib.b200nn_conv2d_ws_bytes_math(1, ctypes.byref(big))
    small_off = lib.b200nn_conv2d_ws_bytes_math(2, ctypes.byref(small))
    monkeypatch.delenv("B200NN_F16CONV")
    assert lib.b200nn_conv2d_ws_bytes_math(2, ctypes.byref(small)) == small_off          # below the threshold: no planes by default
    monkeypatch.setenv("B200NN_F16CONV", "2")
    assert lib.b200nn_conv2d_ws_bytes_math(2, ctypes.byref(small)) >= (2 * 8 * 256 * 128 + 2 * taps * 64 * 128) * 2
    # Dense: the tail word is part of every tf32x3 sizing
    assert lib.b200nn_dense_ws_bytes_math(2, 8192, 4096, 4096) % 256 == 0
    assert lib.b200nn_dense_ws_bytes_math(2, 8192, 4096, 4096) >= 2 * (8192 + 4096) * 4096 * 2 + 256
