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
