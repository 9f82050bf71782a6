"""Drop-in aliasing: `install_alias()` registers neuralnetlib_b200 (and its sub-modules) in `sys.modules` under the upstream
package name, so that code written against `neuralnetlib.{activations,layers,losses,metrics,models,optimizers,callbacks,utils,
preprocessing}` runs on the B200 backend unchanged (INTEGRATION.md route 1)."""
import importlib
import sys

SUBMODULES = ("activations", "layers", "losses", "metrics", "models", "optimizers", "callbacks", "utils", "preprocessing")


def install_alias(name: str = "neuralnetlib", force: bool = False) -> None:
    if name in sys.modules and not force and getattr(sys.modules[name], "__backend__", None) is None:
        raise RuntimeError(f"a different '{name}' package is already imported; pass force=True to shadow it")
    pkg = importlib.import_module("neuralnetlib_b200")
    sys.modules[name] = pkg
    for sub in SUBMODULES:
        sys.modules[f"{name}.{sub}"] = importlib.import_module(f"neuralnetlib_b200.{sub}")
