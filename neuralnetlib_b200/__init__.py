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

_MATH = os.environ.get("B200NN_MATH", "tf32x3").lower()
_MATH_VERSION = 0   # bumped by set_math: plans built under another mode are rebuilt (models.Sequential._plan)


def set_math(mode: str) -> None:
    """'tf32x3' (default): every eligible Dense / Conv2D / Conv2DTranspose contraction runs on the tcgen05 tensor cores with
    error compensation (hi/lo operand split, three MMAs per k-step) and stays within the FFMA path's rel 1e-5 of the
    reference; contractions the tensor-core path does not cover (tiny or unaligned shapes) use the FFMA kernels.
    'tf32': single-pass tcgen05 kind::tf32 (fp32 accumulate), within rel 2e-3 per contraction.
    'fp32': FFMA contractions everywhere."""
    global _MATH, _MATH_VERSION
    mode = mode.lower()
    if mode not in ("fp32", "tf32", "tf32x3"):
        raise ValueError("math mode must be 'fp32', 'tf32' or 'tf32x3'")
    if mode != _MATH:
        _MATH_VERSION += 1
    _MATH = mode


def math_version() -> int:
    return _MATH_VERSION


def invalidate_plans() -> None:
    """A layer swapped a parameter TENSOR (assignment with a different element count): plans, captured graphs and optimizer tables
    built on the old pointer are rebuilt at the next step (the same version counter as set_math)."""
    global _MATH_VERSION
    _MATH_VERSION += 1


def get_math() -> str:
    return _MATH
