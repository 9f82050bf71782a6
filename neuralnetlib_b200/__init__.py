"""neuralnetlib_b200 — B200-native (sm_100a) training backend behind neuralnetlib's Python API.

Drop-in for `neuralnetlib.{activations,layers,losses,optimizers,models,metrics,callbacks,utils}` on the
`Sequential.fit -> train_on_batch -> forward/backward -> Optimizer.update` path (and `GAN.train_on_batch`):
same class names, constructor signatures, string aliases and error behaviour; the numeric work runs in
hand-written CUDA kernels reached through the C ABI in include/b200nn.h, with parameters, activations and
optimizer state resident in HBM across steps. There is no CPU fallback.
"""
import os

__version__ = "0.1.0"
__backend__ = "libb200nn (sm_100a CUDA via C ABI)"

_MATH = os.environ.get("B200NN_MATH", "fp32").lower()


def set_math(mode: str) -> None:
    """'fp32': FFMA contractions, parity with the reference within rel 1e-5 (default).
    'tf32': tcgen05 tensor-core contractions (kind::tf32, fp32 accumulate), within rel 2e-3 per contraction.
    'tf32x3': tcgen05 with error compensation (operands split hi/lo in shared memory, three MMAs per k-step): Dense
    contractions on the tensor cores at fp32 accuracy."""
    global _MATH
    mode = mode.lower()
    if mode not in ("fp32", "tf32", "tf32x3"):
        raise ValueError("math mode must be 'fp32', 'tf32' or 'tf32x3'")
    _MATH = mode


def get_math() -> str:
    return _MATH
