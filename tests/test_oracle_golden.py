"""CPU: pins the numpy oracle (oracle/nnl_oracle.py) against fixtures produced by executing the
unmodified upstream reference (oracle/make_golden.py), and against the reference's own closed-form
unit tests. fp64 vs fp64 => tolerance 1e-11 (BLAS summation order only)."""
import numpy as np
import pytest

import nnl_oracle as O
from helpers import check_packed, nrel

TOL = 1e-11


def test_dense_act_loss_clip(golden):
    g = golden("ops")
    assert nrel(O.dense_fwd(g["dense_x"], g["dense_w"], g["dense_b"]), g["dense_y"]) < TOL
    dx, dw, db = O.dense_bwd(g["dense_x"], g["dense_w"], g["dense_err"])
    assert nrel(dx, g["dense_dx"]) < TOL and nrel(dw, g["dense_dw"]) < TOL and nrel(db, g["dense_db"]) < TOL
    for nm in ("sigmoid", "relu", "tanh", "linear", "leakyrelu", "softmax"):
        a = 0.2 if nm == "leakyrelu" else 0.01
        assert nrel(O.act_fwd(nm, g["act_x"], a), g[f"act_{nm}_y"]) < TOL
        if nm != "softmax":
            assert nrel(O.act_bwd(nm, g["act_x"], g["act_err"], a), g[f"act_{nm}_dx"]) < TOL
    assert abs(O.loss_value("cce", g["loss_y"], g["loss_p"]) - g["loss_cce"]) < 1e-13
    assert nrel(O.loss_derivative("cce", g["loss_y"], g["loss_p"]), g["loss_cce_d"]) < TOL
    assert abs(O.loss_value("bce", g["loss_yb"], g["loss_pb"]) - g["loss_bce"]) < 1e-13
    assert nrel(O.loss_derivative("bce", g["loss_yb"], g["loss_pb"]), g["loss_bce_d"]) < TOL
    assert abs(O.loss_value("mse", g["loss_y"], g["loss_p"]) - g["loss_mse"]) < 1e-13
    assert nrel(O.loss_derivative("mse", g["loss_y"], g["loss_p"]), g["loss_mse_d"]) < TOL
    assert nrel(O.clip_l2(g["clip_in"]), g["clip_big"]) < TOL
    assert np.array_equal(O.clip_l2(g["clip_in"] * 1e-2), g["clip_small"])


@pytest.mark.parametrize("gi", range(6))
def test_conv2d(golden, gi):
    g = golden("ops")
    H, W, C, F, kh, kw, sh, sw, same = [int(v) for v in g[f"conv{gi}_cfg"]]
    pad = "same" if same else "valid"
    y = O.conv2d_fwd(g[f"conv{gi}_x"], g[f"conv{gi}_w"], g[f"conv{gi}_b"], (sh, sw), pad)
    assert nrel(y, g[f"conv{gi}_y"]) < TOL
    dx, dw, db = O.conv2d_bwd(g[f"conv{gi}_x"], g[f"conv{gi}_w"], g[f"conv{gi}_err"], (sh, sw), pad)
    assert nrel(dx, g[f"conv{gi}_dx"]) < TOL
    assert nrel(dw, g[f"conv{gi}_dw"]) < TOL
    assert db.shape == g[f"conv{gi}_db"].shape == (F,)  # Q9
    assert nrel(db, g[f"conv{gi}_db"]) < TOL


@pytest.mark.parametrize("gi", range(4))
def test_maxpool(golden, gi):
    g = golden("ops")
    H, W, C, ph, pw, sh, sw = [int(v) for v in g[f"pool{gi}_cfg"]]
    assert np.array_equal(O.maxpool_fwd(g[f"pool{gi}_x"], (ph, pw), (sh, sw)), g[f"pool{gi}_y"])
    assert nrel(O.maxpool_bwd(g[f"pool{gi}_x"], g[f"pool{gi}_err"], (ph, pw), (sh, sw)), g[f"pool{gi}_dx"]) < TOL


def test_maxpool_all_zero_ties(golden):
    g = golden("ops")
    dx = O.maxpool_bwd(np.zeros((1, 2, 2, 1)), np.ones((1, 1, 1, 1)), (2, 2), (2, 2))
    assert np.array_equal(dx, g["pool_allzero_dx"]) and dx.sum() == 4  # Q6


def test_batchnorm(golden):
    g = golden("ops")
    shp = g["bn_x1"].shape[1:]
    gamma, beta, rm, rv = np.ones(shp), np.zeros(shp), np.zeros(shp), np.ones(shp)
    y1, rm, rv, cache = O.bn_fwd(g["bn_x1"], gamma, beta, rm, rv, 0.9, 1e-5, True)
    assert nrel(y1, g["bn_y1"]) < TOL and nrel(rm, g["bn_rm1"]) < TOL and nrel(rv, g["bn_rv1"]) < TOL
    dx, dg, dbt = O.bn_bwd(g["bn_e1"], gamma, cache, 1e-5)
    assert nrel(dx, g["bn_dx1"]) < 1e-9 and nrel(dg, g["bn_dgamma1"]) < TOL and nrel(dbt, g["bn_dbeta1"]) < TOL
    y2, rm, rv, _ = O.bn_fwd(g["bn_x2"], gamma, beta, rm, rv, 0.9, 1e-5, True)
    assert nrel(y2, g["bn_y2"]) < TOL and nrel(rm, g["bn_rm2"]) < TOL and nrel(rv, g["bn_rv2"]) < TOL
    yi, _, _, _ = O.bn_fwd(g["bn_x1"], gamma, beta, rm, rv, 0.9, 1e-5, False)
    assert nrel(yi, g["bn_yinf"]) < TOL
    shp = g["bn2d_x"].shape[1:]
    y, _, _, cache = O.bn_fwd(g["bn2d_x"], np.ones(shp), np.zeros(shp), np.zeros(shp), np.ones(shp), 0.8, 1e-3, True)
    assert nrel(y, g["bn2d_y"]) < TOL
    assert nrel(O.bn_bwd(g["bn2d_e"], np.ones(shp), cache, 1e-3)[0], g["bn2d_dx"]) < 1e-9


@pytest.mark.parametrize("gi", range(6))
def test_conv2d_transpose(golden, gi):
    g = golden("ops")
    H, W, C, F, kh, kw, sh, sw, same = [int(v) for v in g[f"convT{gi}_cfg"]]
    pad = "same" if same else "valid"
    y = O.convT_fwd(g[f"convT{gi}_x"], g[f"convT{gi}_w"], g[f"convT{gi}_b"], (sh, sw), pad)
    assert nrel(y, g[f"convT{gi}_y"]) < TOL
    dx, dw, db = O.convT_bwd(g[f"convT{gi}_x"], g[f"convT{gi}_w"], g[f"convT{gi}_err"], (sh, sw))
    assert nrel(dx, g[f"convT{gi}_dx"]) < TOL
    assert nrel(dw, g[f"convT{gi}_dw"]) < TOL
    assert db.shape == (1, 1, 1, F) and nrel(db, g[f"convT{gi}_db"]) < TOL


def test_adam_sgd(golden):
    g = golden("ops")
    p = [g[f"adam_p{i}_init"].copy() for i in range(4)]
    gs = [g[f"adam_g{i}"] for i in range(6)]
    opt = O.OracleAdam()
    opt.update(5, p[0], gs[0], p[1], gs[1])
    opt.update(3, p[2], gs[2], p[3], gs[3])
    opt.update(5, p[0], gs[4], p[1], gs[5])
    assert opt.t == int(g["adam_t"]) == 3
    for i in range(4):
        assert nrel(p[i], g[f"adam_p{i}_final"]) < TOL
    assert nrel(opt.m_w[5], g["adam_m_w5"]) < TOL and nrel(opt.v_w[5], g["adam_v_w5"]) < TOL
    assert nrel(opt.m_b[5], g["adam_m_b5"]) < TOL and nrel(opt.v_b[5], g["adam_v_b5"]) < TOL
    w = g["sgd_w_init"].copy()
    O.OracleSGD(0.05).update(0, w, g["sgd_g"])
    assert nrel(w, g["sgd_w_final"]) < TOL


def test_reference_closed_form_adam():
    """Replays /root/reference/tests/test_optimizers.py:32-52 (Adam step 1) and :16-22 (SGD)."""
    w = np.array([[0.1, -0.2], [0.4, 0.5]]); b = np.array([0.1, -0.3])
    gw = np.array([[0.01, -0.02], [0.04, 0.05]]); gb = np.array([0.01, -0.03])
    w0, b0 = w.copy(), b.copy()
    O.OracleAdam(learning_rate=0.01).update(0, w, gw, b, gb)
    np.testing.assert_array_almost_equal(w, w0 - 0.01 * gw / (np.abs(gw) + 1e-8), decimal=6)
    np.testing.assert_array_almost_equal(b, b0 - 0.01 * gb / (np.abs(gb) + 1e-8), decimal=6)
    w, b = w0.copy(), b0.copy()
    O.OracleSGD(0.01).update(0, w, gw, b, gb)
    np.testing.assert_array_almost_equal(w, w0 - 0.01 * gw)


def _run_model(g, cfg, width=None):
    specs, _ = O.config_specs(cfg)
    if width:
        for s in specs:
            if s.get("units") == 4096:
                s["units"] = width
    m = O.OracleSequential(specs, "cce", O.OracleAdam(), seed=42)
    steps = int(g["steps"])
    for s in range(steps):
        trace = []
        loss = m.train_on_batch(g["x"], g["y"], trace=trace)
        assert abs(loss - float(g[f"s{s}_loss"])) < 1e-11 * max(1, abs(loss))
        check_packed(g, f"s{s}_pred", m.predictions, 1e-10)
        n = len(m.layers)
        for li, L in enumerate(m.layers):
            if "w" in L:
                # step tolerances: Adam amplifies fp64 noise at |g|~eps -> 1e-7 on weights
                check_packed(g, f"s{s}_L{li}_dw", L["dw"], 1e-9)
                check_packed(g, f"s{s}_L{li}_db", L["db"].reshape(g_shape(g, f"s{s}_L{li}_db")), 1e-9)
                check_packed(g, f"s{s}_L{li}_w", L["w"], 1e-7)
                check_packed(g, f"s{s}_L{li}_b", L["b"], 1e-7)
            if L["type"] == "bn":
                check_packed(g, f"s{s}_L{li}_rm", L["running_mean"], 1e-10)
                check_packed(g, f"s{s}_L{li}_rv", L["running_var"], 1e-10)
                check_packed(g, f"s{s}_L{li}_dgamma", L["dgamma"], 1e-9)
        if s == 0:
            for i, e in enumerate(trace):
                li = n - 1 - i
                key = f"s0_L{li}_inerr"
                if key in g or key + "__sub" in g:
                    check_packed(g, key, e, 1e-9)
    assert m.optimizer.t == int(g["adam_t"])
    keys = [int(k) for k in g["adam_keys"]]
    assert sorted(m.optimizer.m_w.keys()) == keys
    for k in keys:
        check_packed(g, f"adam_m_w{k}", m.optimizer.m_w[k], 1e-9)
        check_packed(g, f"adam_v_w{k}", m.optimizer.v_w[k], 1e-9)
    check_packed(g, "final_pred_eval", m.forward(g["x"], training=False), 1e-7)


def g_shape(g, name):
    return g[name].shape if name in g else tuple(g[name + "__shape"])


@pytest.mark.parametrize("name,cfg,width", [("cfg1_b32", 1, None), ("cfg2_b8", 2, None), ("cfg3_b4", 3, None),
                                            ("cfg4w256_b16", 4, 256)])
def test_model_traces(golden, name, cfg, width):
    _run_model(golden(name), cfg, width)


def test_gan_trace(golden):
    g = golden("gan_b4")
    Gs, Ds = O.gan_specs()
    G = O.OracleSequential(Gs, "bce", O.OracleAdam(), seed=7)
    D = O.OracleSequential(Ds, "bce", O.OracleAdam(), seed=7)
    B = int(g["batch"])
    for s in range(int(g["steps"])):
        dl, gl = O.gan_train_on_batch(G, D, g["real"], B, 32, 7)
        assert abs(dl - float(g[f"s{s}_d_loss"])) < 1e-10 and abs(gl - float(g[f"s{s}_g_loss"])) < 1e-10
        for nm, M in (("G", G), ("D", D)):
            for li, L in enumerate(M.layers):
                if "w" in L:
                    check_packed(g, f"s{s}_{nm}{li}_dw", L["dw"], 1e-9, what=f"step{s} ")
                    check_packed(g, f"s{s}_{nm}{li}_w", L["w"], 1e-7, what=f"step{s} ")
    assert G.optimizer.t == int(g["G_t"]) and D.optimizer.t == int(g["D_t"])
