"""Hot-path helpers of neuralnetlib/preprocessing.py: one_hot_encode (:11-26) and clip_gradients (:220-224).
im2col_2d / col2im_2d (:34-126) have no equivalent here: the convolution kernels gather straight from NHWC
(implicit GEMM), no column buffer exists."""
from __future__ import annotations

import numpy as np
import torch

from . import _backend as B
from . import _ops as ops


def one_hot_encode(indices: np.ndarray, num_classes: int) -> np.ndarray:
    indices = np.asarray(indices)
    if indices.ndim not in (1, 2):
        raise ValueError("Unsupported input shape. Expected 1D or 2D indices.")
    return np.eye(num_classes)[indices]


def clip_gradients(gradients, threshold: float = 1.0):
    """Whole-tensor L2 clip. The plan never materialises this (the multiplier is applied lazily by the consumer);
    this function is the standalone, materialising form for API users."""
    if threshold != 1.0:
        scale = float(threshold)
        return clip_gradients(np.asarray(gradients) / scale) * scale if not isinstance(gradients, torch.Tensor) \
            else clip_gradients(gradients / scale) * scale
    host = not isinstance(gradients, torch.Tensor)
    g = B.to_device(np.asarray(gradients)) if host else gradients.contiguous()
    ss = B.zeros((1,), dtype=torch.float64)
    out = B.empty(g.shape)
    ops.sumsq(g, ss, 0)
    ops.scale_apply(g, ss, (0,), out)
    return B.to_host(out) if host else out
