"""The reference's own unit-test suite, restated against the drop-in ("the 29 reference tests must pass against
it"). Each test below makes the same assertions as the cited reference test, on the same inputs, through the same public names —
importable as `neuralnetlib` via `neuralnetlib_b200.compat.install_alias()` exactly as a user switching packages would."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nl():
    from neuralnetlib_b200 import compat
    compat.install_alias()
    import neuralnetlib   # noqa: F401  (the alias)
    import neuralnetlib.activations, neuralnetlib.layers, neuralnetlib.losses, neuralnetlib.metrics  # noqa: F401,E401
    import neuralnetlib.models, neuralnetlib.optimizers  # noqa: F401,E401
    return neuralnetlib


# ---- /root/reference/tests/test_activations.py:10-66 ----------------------------------------------------------------------
X3 = np.array([-1.0, 0.0, 1.0])


@pytest.mark.parametrize("name,expected", [
    ("Sigmoid", lambda x: 1 / (1 + np.exp(-x))), ("ReLU", lambda x: np.maximum(0, x)), ("Tanh", np.tanh), ("Linear", lambda x: x),
    ("LeakyReLU", lambda x: np.where(x > 0, x, x * 0.01)), ("ELU", lambda x: np.where(x > 0, x, np.exp(x) - 1))])
def test_activation_formulas(nl, name, expected):
    f = getattr(nl.activations, name)()
    np.testing.assert_array_almost_equal(f(X3), expected(X3))


def test_selu(nl):
    selu = nl.activations.SELU()
    want = nl.activations.SELU.DEFAULT_SCALE * np.where(X3 > 0, X3, nl.activations.SELU.DEFAULT_ALPHA * (np.exp(X3) - 1))
    np.testing.assert_array_almost_equal(selu(X3), want)


def test_softmax_and_its_missing_derivative(nl):
    x = np.array([[1.0, 2.0, 3.0], [1.0, 2.0, 3.0]])
    e = np.exp(x - np.max(x, axis=1, keepdims=True))
    np.testing.assert_array_almost_equal(nl.activations.Softmax()(x), e / np.sum(e, axis=1, keepdims=True))
    with pytest.raises(NotImplementedError):
        nl.activations.Softmax().derivative(np.array([[1.0, 2.0, 3.0]]))


# ---- /root/reference/tests/test_layers.py:10-43 -----------------------------------------------------------------------------
def test_layer_base_class_raises(nl):
    layer = nl.layers.Layer()
    with pytest.raises(NotImplementedError):
        layer.forward_pass(np.array([1, 2, 3]))
    with pytest.raises(NotImplementedError):
        layer.backward_pass(np.array([1, 2, 3]))


def test_dense_layer_shapes(nl):
    x = np.array([[1, 2, 3], [4, 5, 6]])
    dense = nl.layers.Dense(2)
    out = dense.forward_pass(x)
    assert out.shape == (2, 2)
    assert dense.backward_pass(np.random.default_rng(0).random(out.shape)).shape == x.shape


def test_activation_layer(nl):
    x = np.array([[-1, 0, 1]])
    f = nl.activations.Sigmoid()
    act = nl.layers.Activation(f)
    out = act.forward_pass(x)
    np.testing.assert_array_almost_equal(out, f(x))
    assert act.backward_pass(np.random.default_rng(0).random(out.shape)).shape == x.shape


# ---- /root/reference/tests/test_losses.py:11-53 -----------------------------------------------------------------------------
def test_losses(nl):
    Ls = nl.losses
    y, p = np.array([1, 2, 3]), np.array([1, 2, 4])
    assert Ls.MeanSquaredError()(y, p) == pytest.approx(np.mean(np.power(y - p, 2)), abs=1e-7)
    assert Ls.MeanAbsoluteError()(y, p) == pytest.approx(np.mean(np.abs(y - p)), abs=1e-7)
    err = y - p
    want = np.mean(np.where(np.abs(err) <= 1.0, 0.5 * np.square(err), 1.0 * (np.abs(err) - 0.5)))
    assert Ls.Huber(delta=1.0)(y, p) == pytest.approx(want, abs=1e-7)
    yb, pb = np.array([0, 1, 1]), np.array([0.1, 0.8, 0.99])
    assert Ls.BinaryCrossentropy()(yb, pb) == pytest.approx(-np.mean(yb * np.log(pb) + (1 - yb) * np.log(1 - pb)), abs=1e-6)
    yc = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1]])
    pc = np.array([[0.7, 0.2, 0.1], [0.1, 0.8, 0.1], [0.2, 0.2, 0.6]])
    assert Ls.CategoricalCrossentropy()(yc, pc) == pytest.approx(-np.sum(yc * np.log(pc)) / 3, abs=1e-6)


# ---- /root/reference/tests/test_metrics.py:10-33 ----------------------------------------------------------------------------
def test_metrics(nl):
    y_true = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1]])
    y_pred = np.array([[0.7, 0.2, 0.1], [0.1, 0.8, 0.1], [0.1, 0.1, 0.8]])
    Mt = nl.metrics
    assert Mt.accuracy_score(y_pred, y_true) == pytest.approx(1.0)
    assert Mt.f1_score(y_pred, y_true) == pytest.approx(1.0)
    assert Mt.recall_score(y_pred, y_true) == pytest.approx(1.0)
    assert np.array_equal(Mt.confusion_matrix(y_pred, y_true), np.eye(3, dtype=int))


# ---- /root/reference/tests/test_models.py:12-39 -----------------------------------------------------------------------------
@pytest.fixture()
def small_model(nl):
    from neuralnetlib.activations import Sigmoid
    from neuralnetlib.layers import Dense, Input
    from neuralnetlib.models import Activation, CategoricalCrossentropy, Sequential     # re-exported by models, as upstream
    from neuralnetlib.optimizers import SGD
    m = Sequential()
    m.add(Input(10)); m.add(Dense(20)); m.add(Activation(Sigmoid()))
    m.compile(loss_function=CategoricalCrossentropy(), optimizer=SGD())
    rng = np.random.default_rng(0)
    return m, rng.random((100, 10)), rng.random((100, 20)), rng.random((10, 10)), rng.random((10, 20))


def test_model_train_on_batch_returns_float(small_model):
    m, x, y, _, _ = small_model
    assert isinstance(m.train_on_batch(x[:10], y[:10]), float)


def test_model_fit(small_model):
    m, x, y, _, _ = small_model
    m.fit(x, y, epochs=1, batch_size=10, verbose=False)


def test_model_evaluate(small_model):
    m, _, _, xt, yt = small_model
    loss, preds = m.evaluate(xt, yt)
    assert isinstance(loss, float)


def test_model_predict(small_model):
    m, _, _, xt, yt = small_model
    assert m.predict(xt).shape == yt.shape


# ---- /root/reference/tests/test_optimizers.py:10-68 -------------------------------------------------------------------------
@pytest.fixture()
def params():
    return (np.array([[0.1, -0.2], [0.4, 0.5]]), np.array([0.1, -0.3]), np.array([[0.01, -0.02], [0.04, 0.05]]), np.array([0.01, -0.03]))


def test_sgd_and_momentum_first_step(nl, params):
    for make in (lambda: nl.optimizers.SGD(learning_rate=0.01), lambda: nl.optimizers.Momentum(learning_rate=0.01, momentum=0.9)):
        w, b, gw, gb = (a.copy() for a in params)
        w0, b0 = w.copy(), b.copy()
        make().update(0, w, gw, b, gb)
        np.testing.assert_array_almost_equal(w, w0 - 0.01 * gw)
        np.testing.assert_array_almost_equal(b, b0 - 0.01 * gb)


def test_adam_first_step(nl, params):
    w, b, gw, gb = (a.copy() for a in params)
    w0, b0 = w.copy(), b.copy()
    nl.optimizers.Adam(learning_rate=0.01, beta_1=0.9, beta_2=0.999).update(0, w, gw, b, gb)
    for p, p0, g in ((w, w0, gw), (b, b0, gb)):
        mh, vh = (0.1 * g) / (1 - 0.9), (0.001 * np.square(g)) / (1 - 0.999)
        np.testing.assert_array_almost_equal(p, p0 - 0.01 * mh / (np.sqrt(vh) + 1e-8), decimal=2)


def test_rmsprop_first_step(nl, params):
    w, b, gw, gb = (a.copy() for a in params)
    w0, b0 = w.copy(), b.copy()
    nl.optimizers.RMSprop(learning_rate=0.01, rho=0.9).update(0, w, gw, b, gb)
    for p, p0, g in ((w, w0, gw), (b, b0, gb)):
        np.testing.assert_array_almost_equal(p, p0 - 0.01 * g / (np.sqrt(0.1 * np.square(g)) + 1e-8))
