"""Host-side helpers of the fit() loop (mirror of neuralnetlib/utils.py:11-15, 34-50, 95-143, 347-356).
Permutations come from the same NumPy generator calls as the reference so that seeds reproduce."""
from __future__ import annotations

import shutil
import sys
import time

import numpy as np


class History(dict):
    """dict that prints as nothing when it is the value of a notebook cell (utils.py:11-15)."""

    def __repr__(self):
        return ""


def shuffle_indices(n_samples: int, random_state: int | None = None) -> np.ndarray:
    """The permutation `shuffle` applies (utils.py:36-40): default_rng(seed).permutation(n)."""
    rng = np.random.default_rng(random_state if random_state is not None else int(time.time_ns()))
    return rng.permutation(n_samples)


def shuffle(x, y=None, random_state: int | None = None):
    x = np.array(x)
    idx = shuffle_indices(x.shape[0], random_state)
    if y is None:
        return x[idx]
    return x[idx], np.array(y)[idx]


def progress_bar(current: int, total: int, width: int = 30, message: str = "") -> None:
    frac = current / total
    filled = int(width * frac)
    sys.stdout.write("\r" + " " * shutil.get_terminal_size().columns)
    sys.stdout.write(f"\r[{'=' * filled}{'-' * (width - filled)}] {int(100 * frac)}% {message}")
    sys.stdout.flush()


def format_number(number) -> str:
    if number == 0:
        return "0.0"
    if abs(number) < 1e-3:
        exponent = int(f"{number:.1e}".split("e")[1])
        return f"{number:.{abs(exponent) + 1}e}"
    return f"{number:.4f}"


def train_test_split(x, y=None, test_size: float = 0.2, random_state: int | None = None, shuffle: bool = True):
    x = np.array(x)
    rng = np.random.default_rng(random_state if random_state is not None else int(time.time_ns()))
    idx = np.arange(len(x))
    if shuffle:
        rng.shuffle(idx)
    cut = int(len(x) * (1 - test_size))
    if y is None:
        return x[idx[:cut]], x[idx[cut:]]
    y = np.array(y)
    return x[idx[:cut]], x[idx[cut:]], y[idx[:cut]], y[idx[cut:]]
