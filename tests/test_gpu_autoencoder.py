"""GPU parity of the Autoencoder training loop (neuralnetlib/models.py:778-1565) against traces recorded from
the unmodified reference (oracle/make_golden_r2.py `ae` -> tests/golden/autoencoder.npz): its own clip rule (norm clip, clamp,
divide by np.std), the per-layer updates right after each layer's backward, the variational loss terms, skip connections,
the Sigmoid + BinaryCrossentropy special branch, evaluate / predict. Losses 1e-5, predictions / gradients 5e-5 (norm-relative;
the division by the tensor's own standard deviation is part of the compared quantity), SGD weights 2e-5, Adam weights by the
quantile criterion of test_gpu_models.py."""
import numpy as np
import pytest

from helpers import nrel

pytestmark = pytest.mark.gpu


def _build(loss, opt, skip):
    from neuralnetlib_b200.layers import Dense, Input
    from neuralnetlib_b200.models import Autoencoder
    ae = Autoencoder(gradient_clip_threshold=5.0, random_state=3, l2_reg=1e-4, variational=True, skip_connections=skip)
    ae.add_encoder_layer(Input(20))
    ae.add_encoder_layer(Dense(12, activation="relu", random_state=3))
    ae.add_encoder_layer(Dense(10, random_state=4))
    ae.add_decoder_layer(Dense(12, activation="relu", random_state=5))
    ae.add_decoder_layer(Dense(20, activation="sigmoid", random_state=6))
    ae.compile(encoder_loss=loss, decoder_loss=loss, encoder_optimizer=opt(), decoder_optimizer=opt())
    return ae


def test_ae_clip_rule(golden):
    import torch
    from neuralnetlib_b200 import _backend as B, _ops as ops
    g = golden("autoencoder")
    x = B.to_device(g["clip_in"])
    out, scratch = B.empty(tuple(x.shape)), B.zeros((3,), dtype=torch.float64)
    ops.ae_clip(x, out, 5.0, scratch)
    assert nrel(B.to_host(out), g["clip_out"]) < 1e-6
    ops.ae_clip(x, out, 0.0, scratch)                        # threshold <= 0: untouched (models.py:1037)
    assert nrel(B.to_host(out), g["clip_in"]) < 1e-7


@pytest.mark.parametrize("name,loss,skip", [("mse_adam", "mse", False), ("bce_sgd_skip", "bce", True)])
def test_autoencoder_train_on_batch_golden(golden, name, loss, skip):
    from neuralnetlib_b200 import optimizers as Opt
    g = golden("autoencoder")
    adam = name == "mse_adam"
    ae = _build(loss, (lambda: Opt.Adam(learning_rate=0.01)) if adam else (lambda: Opt.SGD(learning_rate=0.05)), skip)
    x = g["x"]
    for s in range(3):
        lv = ae.train_on_batch(x)
        ref = float(g[f"{name}_s{s}_loss"])
        assert abs(lv - ref) <= 2e-5 * max(1.0, abs(ref)), (s, lv, ref)
        assert nrel(ae.predictions, g[f"{name}_s{s}_pred"]) < 5e-5, s
        for side, layers in (("enc", ae.encoder_layers), ("dec", ae.decoder_layers)):
            for li, L in enumerate(layers):
                if hasattr(L, "weights"):
                    assert nrel(L.d_weights, g[f"{name}_s{s}_{side}{li}_dw"]) < 5e-5, (s, side, li)
                    assert nrel(L.d_bias, g[f"{name}_s{s}_{side}{li}_db"]) < 5e-5, (s, side, li)
                    wref = g[f"{name}_s{s}_{side}{li}_w"]
                    d = np.abs(np.asarray(L.weights) - wref)
                    if adam:    # lr / eps conditioning of Adam at |g| ~ 0
                        assert np.mean(d <= 2e-5 * np.abs(wref).max() + 1e-9) >= 0.98 and d.max() <= 2 * 0.01, (s, side, li, d.max())
                    else:
                        assert d.max() <= 2e-5 * np.abs(wref).max(), (s, side, li, d.max())
                    # teacher forcing, as in the Sequential tests: continue from the reference's state
                    L.weights = wref
                    L.bias = g[f"{name}_s{s}_{side}{li}_b"]
    if not adam:
        ev_loss, ev_pred = ae.evaluate(x, batch_size=10)
        assert abs(ev_loss - float(g[f"{name}_eval_loss"])) <= 2e-5 * max(1.0, abs(float(g[f"{name}_eval_loss"])))
        assert nrel(ev_pred, g[f"{name}_eval_pred"]) < 5e-5
        assert nrel(ae.predict(x, output_latent=True), g[f"{name}_latent"]) < 5e-5


def test_autoencoder_fit_runs_and_learns():
    from neuralnetlib_b200 import optimizers as Opt
    rng = np.random.default_rng(0)
    x = rng.random((64, 20))
    ae = _build("mse", lambda: Opt.Adam(learning_rate=0.01), False)
    h = ae.fit(x, epochs=6, batch_size=16, verbose=False, random_state=1, validation_data=(x[:16], x[:16]))
    assert len(h["loss"]) == 6 and len(h["val_loss"]) == 6
    assert h["loss"][-1] < h["loss"][0]
