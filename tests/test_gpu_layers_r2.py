"""GPU parity of the sibling layers and the ELU / SELU / GELU activations against vectors recorded from the
unmodified reference (oracle/make_golden_r2.py: layers_r2.npz), through the reference's own eager API (forward_pass / backward_pass
on NumPy arrays), plus a Sequential that strings them together under a plan. Tolerance: norm-relative 1e-5."""
import numpy as np
import pytest

from helpers import nrel

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _L():
    from neuralnetlib_b200 import layers
    return layers


@pytest.mark.parametrize("name", ["elu", "selu", "gelu"])
def test_activation_golden(golden, name):
    from neuralnetlib_b200 import activations as A
    g = golden("layers_r2")
    f = {"elu": A.ELU, "selu": A.SELU, "gelu": A.GELU}[name]()
    assert nrel(f(g["act_x"]), g[f"act_{name}_y"]) < TOL
    assert nrel(f.derivative(g["act_x"]), g[f"act_{name}_d"]) < TOL
    layer = _L().Activation(f)
    y = layer.forward_pass(g["act_x"])
    assert nrel(y, g[f"act_{name}_y"]) < TOL
    assert nrel(layer.backward_pass(g["act_err"]), g["act_err"] * g[f"act_{name}_d"]) < TOL      # layers.py:237
    assert isinstance(A.ActivationFunction.from_name(name), type(f))


@pytest.mark.parametrize("gi", [0, 1])
def test_dropout_golden(golden, gi):
    g = golden("layers_r2")
    L = _L().Dropout(float(g[f"drop{gi}_rate"]), random_state=int(g[f"drop{gi}_seed"]))
    y = L.forward_pass(g[f"drop{gi}_x"], training=True)
    assert nrel(y, g[f"drop{gi}_y"]) < TOL                      # same host generator call => same mask
    assert np.array_equal(y == 0, g[f"drop{gi}_y"] == 0)
    assert nrel(L.backward_pass(g[f"drop{gi}_err"]), g[f"drop{gi}_dx"]) < TOL
    assert nrel(L.forward_pass(g[f"drop{gi}_x"], training=False), g[f"drop{gi}_y_eval"]) < 1e-6
    assert type(_L().Layer.from_config(L.get_config())).__name__ == "Dropout"


@pytest.mark.parametrize("gi", [0, 1, 2])
def test_average_pooling_golden(golden, gi):
    g = golden("layers_r2")
    ph, pw, sh, sw = (int(v) for v in g[f"avg{gi}_cfg"])
    L = _L().AveragePooling2D((ph, pw), (sh, sw), "valid")
    assert nrel(L.forward_pass(g[f"avg{gi}_x"]), g[f"avg{gi}_y"]) < TOL
    assert nrel(L.backward_pass(g[f"avg{gi}_err"]), g[f"avg{gi}_dx"]) < TOL


def test_global_average_pooling_golden(golden):
    g = golden("layers_r2")
    L = _L().GlobalAveragePooling2D()
    assert nrel(L.forward_pass(g["gap_x"]), g["gap_y"]) < TOL
    assert nrel(L.backward_pass(g["gap_err"]), g["gap_dx"]) < TOL


@pytest.mark.parametrize("gi", [0, 1])
def test_upsampling_golden(golden, gi):
    g = golden("layers_r2")
    L = _L().UpSampling2D(tuple(int(v) for v in g[f"up{gi}_size"]))
    assert nrel(L.forward_pass(g[f"up{gi}_x"]), g[f"up{gi}_y"]) < 1e-6
    assert nrel(L.backward_pass(g[f"up{gi}_err"]), g[f"up{gi}_dx"]) < TOL
    with pytest.raises(NotImplementedError):
        _L().UpSampling2D(2, interpolation="bilinear")


@pytest.mark.parametrize("math", ["fp32", "tf32x3"])
@pytest.mark.parametrize("gi", [0, 1, 2])
def test_conv1d_golden(golden, gi, math):
    import neuralnetlib_b200 as nn
    nn.set_math(math)
    g = golden("layers_r2")
    filt, k, s, same = (int(v) for v in g[f"c1d{gi}_cfg"])
    L = _L().Conv1D(filt, k, strides=s, padding="same" if same else "valid", random_state=3)
    y = L.forward_pass(g[f"c1d{gi}_x"])
    assert nrel(L.weights, g[f"c1d{gi}_w"]) < 1e-7            # same generator calls => same initial weights (float32 storage)
    assert nrel(y, g[f"c1d{gi}_y"]) < TOL
    assert nrel(L.backward_pass(g[f"c1d{gi}_err"]), g[f"c1d{gi}_dx"]) < TOL
    assert nrel(L.d_weights, g[f"c1d{gi}_dw"]) < TOL
    assert L.d_bias.shape == g[f"c1d{gi}_db"].shape and nrel(L.d_bias, g[f"c1d{gi}_db"]) < TOL


@pytest.mark.parametrize("gi", [0, 1])
def test_maxpooling1d_golden(golden, gi):
    g = golden("layers_r2")
    pool, s = (int(v) for v in g[f"mp1d{gi}_cfg"])
    L = _L().MaxPooling1D(pool, s)
    assert nrel(L.forward_pass(g[f"mp1d{gi}_x"]), g[f"mp1d{gi}_y"]) < 1e-6
    assert nrel(L.backward_pass(g[f"mp1d{gi}_err"]), g[f"mp1d{gi}_dx"]) < TOL


def test_sibling_layers_in_a_planned_model():
    """Conv2D -> AveragePooling2D -> Dropout -> UpSampling2D -> Conv2D(ELU) -> GlobalAveragePooling2D -> Dense(GELU) -> Dense(softmax)
    trains under a captured plan: the loss falls, the graph replays, save / load round-trips the layer list."""
    import os
    import tempfile
    from neuralnetlib_b200 import layers as L, models as M, optimizers as Opt
    m = M.Sequential(random_state=3)
    m.add(L.Input((8, 8, 4)))
    m.add(L.Conv2D(8, 3, padding="same", activation="relu"))
    m.add(L.AveragePooling2D(2))
    m.add(L.Dropout(0.25))
    m.add(L.UpSampling2D(2))
    m.add(L.Conv2D(8, 3, padding="same", activation="elu"))
    m.add(L.GlobalAveragePooling2D())
    m.add(L.Dense(16, activation="gelu"))
    m.add(L.Dense(3, activation="softmax"))
    m.compile(loss_function="categorical_crossentropy", optimizer=Opt.Adam(learning_rate=0.01))
    rng = np.random.default_rng(0)
    x = rng.random((32, 8, 8, 4)).astype(np.float32)
    y = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 32)]
    losses = [m.train_on_batch(x, y) for _ in range(30)]
    assert losses[-1] < losses[0] and np.all(np.isfinite(losses))
    assert m._last_plan.program("train").graph is not None
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "m.json")
        m.save(path)
        m2 = M.Sequential.load(path)
    assert [type(a).__name__ for a in m2.layers] == [type(a).__name__ for a in m.layers]
    assert nrel(m2.predict(x), m.predict(x)) < 1e-5
