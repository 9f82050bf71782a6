"""GPU parity of the single-launch MLP train step (csrc/mlp_step.cu, b200nn_mlp_train_step): against the fp64 oracle and against
the layer-by-layer program of the same model (B200NN_MLP_STEP=0), over several steps, shapes that are not multiples of any
tile, every fusable activation and a layer without activation. Same tolerances as the rest of the model-level suite: losses,
predictions and gradients rel 1e-5 (norm-relative), post-Adam weights by the quantile criterion of test_gpu_models.py."""
import numpy as np
import pytest

import nnl_oracle as O
from helpers import nrel

pytestmark = pytest.mark.gpu


def _build(dims, acts, seed=3):
    from neuralnetlib_b200 import layers as L, models as M, optimizers as Opt
    m = M.Sequential(random_state=seed)
    m.add(L.Input(dims[0]))
    for units, act in zip(dims[1:-1], acts):
        if act is None:
            m.add(L.Dense(units))
        else:
            m.add(L.Dense(units, activation=act))
    m.add(L.Dense(dims[-1], activation="softmax"))
    m.compile(loss_function="categorical_crossentropy", optimizer=Opt.Adam())
    return m


def _oracle_specs(dims, acts):
    specs = [{"type": "input"}]
    for units, act in zip(dims[1:-1], acts):
        specs.append({"type": "dense", "units": units})
        if act is not None:
            specs.append({"type": "act", "fn": act})
    specs += [{"type": "dense", "units": dims[-1]}, {"type": "act", "fn": "softmax"}]
    return specs


def _names(model):
    return [n for n, _ in model._last_plan.program("train").meta]


CASES = [(32, (784, 128, 64, 10), ("relu", "relu")),
         (7, (50, 70, 33, 5), ("tanh", "sigmoid")),
         (64, (100, 40, 10), ("leakyrelu",)),
         (33, (20, 16, 8, 12, 3), ("relu", None, "relu")),
         (1, (9, 5, 2), ("relu",))]


@pytest.mark.parametrize("Bn,dims,acts", CASES)
def test_mlp_step_matches_layer_kernels(monkeypatch, Bn, dims, acts):
    rng = np.random.default_rng(Bn + len(dims))
    x = rng.normal(size=(Bn, dims[0])).astype(np.float32)
    y = np.eye(dims[-1], dtype=np.float32)[rng.integers(0, dims[-1], Bn)]
    fused = _build(dims, acts)
    monkeypatch.setenv("B200NN_MLP_STEP", "0")
    plain = _build(dims, acts)
    for s in range(3):
        monkeypatch.setenv("B200NN_MLP_STEP", "0")
        lp = plain.train_on_batch(x, y)
        monkeypatch.setenv("B200NN_MLP_STEP", "1")
        lf = fused.train_on_batch(x, y)
        assert abs(lf - lp) <= 1e-5 * max(1.0, abs(lp)), (s, lf, lp)
        assert nrel(fused.predictions, plain.predictions) < 2e-5
        for a, b in zip(fused.layers, plain.layers):
            if hasattr(a, "weights"):
                assert nrel(a.d_weights, b.d_weights) < 3e-5, (s, a)
                assert nrel(a.d_bias, b.d_bias) < 3e-5, (s, a)
                d = np.abs(np.asarray(a.weights) - np.asarray(b.weights))
                scale = np.abs(np.asarray(b.weights)).max()
                assert int(np.sum(d > 2e-5 * scale + 1e-9)) <= max(2, int(2e-3 * d.size)) and d.max() <= 2e-3, (s, a, d.max())
        # teacher forcing: both continue from the layer-by-layer weights, so that Adam's conditioning at |g| ~ eps cannot accumulate
        for a, b in zip(fused.layers, plain.layers):
            if hasattr(a, "weights"):
                a.weights = np.asarray(b.weights); a.bias = np.asarray(b.bias)
        fused.optimizer.m_w, fused.optimizer.v_w = plain.optimizer.m_w, plain.optimizer.v_w
        fused.optimizer.m_b, fused.optimizer.v_b = plain.optimizer.m_b, plain.optimizer.v_b
    assert _names(fused) == ["mlp_train_step"]
    assert "mlp_train_step" not in _names(plain)
    assert fused.optimizer.t == plain.optimizer.t == 3 * (len(dims) - 1)          # Q3: t counts layers, not steps


@pytest.mark.parametrize("Bn,dims,acts", [c for c in CASES if "leakyrelu" not in c[2]])
def test_mlp_step_oracle(Bn, dims, acts):
    """One-step parity from identical states against the fp64 oracle (teacher forcing as in test_gpu_models.py), 3 steps."""
    rng = np.random.default_rng(Bn * 3 + len(dims))
    x = rng.normal(size=(Bn, dims[0])).astype(np.float32)
    y = np.eye(dims[-1], dtype=np.float32)[rng.integers(0, dims[-1], Bn)]
    m = _build(dims, acts)
    ref = O.OracleSequential(_oracle_specs(dims, acts), "cce", O.OracleAdam(), seed=3)
    mine = [(i, l) for i, l in enumerate(m.layers) if hasattr(l, "weights")]
    for s in range(3):
        loss, rloss = m.train_on_batch(x, y), ref.train_on_batch(x, y)
        theirs = [r for r in ref.layers if "w" in r]
        assert len(theirs) == len(mine)
        assert abs(loss - rloss) <= 1e-5 * max(1.0, abs(rloss)), (s, loss, rloss)
        assert nrel(m.predictions, ref.predictions) < 2e-5
        for (li, layer), rl in zip(mine, theirs):
            assert nrel(layer.d_weights, rl["dw"]) < 3e-5, (s, li)
            assert nrel(layer.d_bias.reshape(-1), rl["db"].reshape(-1)) < 3e-5, (s, li)
            d = np.abs(layer.weights - rl["w"])
            assert np.mean(d <= 1e-5 * np.abs(rl["w"]).max() + 1e-9) >= 0.995 and d.max() <= 2e-3, (s, li, d.max())
        assert m.optimizer.t == ref.optimizer.t
        mw, vw, mb, vb = m.optimizer.m_w, m.optimizer.v_w, m.optimizer.m_b, m.optimizer.v_b
        for (li, layer), rl in zip(mine, theirs):
            rl["w"], rl["b"] = layer.weights, layer.bias
            ref.optimizer.m_w[li], ref.optimizer.v_w[li] = mw[li], vw[li]
            ref.optimizer.m_b[li], ref.optimizer.v_b[li] = mb[li], vb[li]
    assert _names(m) == ["mlp_train_step"]


def test_mlp_step_not_used_outside_its_limits():
    """Batch > 64, a non-Adam optimizer or a convolution keep the layer-by-layer program."""
    from neuralnetlib_b200 import optimizers as Opt
    rng = np.random.default_rng(0)
    m = _build((30, 20, 4), ("relu",))
    x = rng.normal(size=(65, 30)).astype(np.float32)
    y = np.eye(4, dtype=np.float32)[rng.integers(0, 4, 65)]
    m.train_on_batch(x, y)
    assert "mlp_train_step" not in _names(m)
    m2 = _build((30, 20, 4), ("relu",))
    m2.compile(loss_function="categorical_crossentropy", optimizer=Opt.SGD())
    m2.train_on_batch(x[:16], y[:16])
    assert "mlp_train_step" not in _names(m2)
