"""TEST INFRASTRUCTURE ONLY — loads the unmodified upstream reference (pure Python/NumPy)
from /root/reference so that golden vectors can be generated and the numpy oracle in
`oracle/nnl_oracle.py` can be validated against it.

The reference imports matplotlib at `neuralnetlib/models.py:6-7`; matplotlib is not installed in
this image and is only reached for plotting, so two stub modules are injected into `sys.modules`.
`/root/reference` does not exist on the GPU box: nothing that runs there may import this module
(the committed fixtures under `tests/golden/` are what travels).
"""
import sys
import types

REFERENCE_ROOT = "/root/reference"


def _install_matplotlib_stub():
    if "matplotlib" in sys.modules:
        return
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    mpl.get_backend = lambda: "agg"
    plt = types.ModuleType("matplotlib.pyplot")

    def _missing(name):
        def _f(*a, **k):
            raise RuntimeError(f"matplotlib.pyplot.{name} is stubbed (plotting is out of scope)")
        return _f

    plt.__getattr__ = _missing  # type: ignore[attr-defined]
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


def load_reference():
    """Return the imported upstream `neuralnetlib` package (raises if /root/reference is absent)."""
    import os
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("upstream reference not present (expected only in the build container)")
    _install_matplotlib_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import neuralnetlib  # noqa: F401
    import neuralnetlib.models  # noqa: F401
    return sys.modules["neuralnetlib"]
