"""CPU tests of the oracle itself: pinned against every artefact the reference ships for this path
 and against closed-form / self-consistency known answers.  No GPU, no compute calls into the
CUDA library."""
import copy
import math
import os

import numpy as np
import pytest
import torch

from common import CAM_ISO_NLF, make_oracle, synth_batch
from oracle import noise_flow_oracle as O


def test_layer_names_and_param_count_match_hps_txt(golden_dir, shipped):
    """hps.txt:1-18 = get_layer_names(), hps.txt:19 = 2433 trainable parameters; every variable the
    restatement asks for exists in the checkpoint under exactly the reference's name and shape."""
    hps, ck = shipped
    with open(os.path.join(golden_dir, "NoiseFlow", "hps.txt")) as f:
        head = [l.strip() for l in f.readlines()[:19]]
    orc = make_oracle(hps, ck)
    x, y = synth_batch(1)
    orc._loss(x, y, iso=[100.0], cam=[2.0])
    assert orc.get_layer_names() == head[:18]
    assert orc.store.num_trainable() == int(head[18]) == hps.num_params == 2433
    assert orc.store.created == []
    assert set(orc.store.trainable) == set(ck.keys())          # nothing in the checkpoint is left unused


def test_fill_triangular_docstring_examples_and_survey_ordering():
    x = np.arange(1, 7.0)
    assert np.array_equal(O.fill_triangular(x), [[4, 0, 0], [6, 5, 0], [3, 2, 1]])
    assert np.array_equal(O.fill_triangular(x, upper=True), [[1, 2, 3], [0, 5, 6], [0, 0, 4]])
    v = np.arange(6.0)   # L[1,0]=v3 L[2,0]=v5 L[2,1]=v4 L[3,0]=v2 L[3,1]=v1 L[3,2]=v0
    L = O.vec2stricttri(v, upper=False)
    assert (L[1, 0], L[2, 0], L[2, 1], L[3, 0], L[3, 1], L[3, 2]) == (3, 5, 4, 2, 1, 0)
    U = O.vec2stricttri(v, upper=True)
    assert (U[0, 1], U[0, 2], U[0, 3], U[1, 2], U[1, 3], U[2, 3]) == (0, 1, 2, 4, 5, 3)
    for up in (False, True):
        w = np.random.RandomState(0).randn(6)
        assert np.allclose(O.stricttri2vec(O.vec2stricttri(w, up), up), w)


def test_conv1x1_lu_reconstruction_is_well_conditioned(shipped):
    """A = P L U from the checkpoint: A A_inv = I, cond(A) ~ 1.9-2.9, log|det A| = sum log_S."""
    hps, ck = shipped
    orc = make_oracle(hps, ck)
    sums = []
    for b in orc.model[0]:
        if isinstance(b, O.Conv2d1x1):
            p = b.params()
            A, Ai = p["A"].numpy(), p["A_inv"].numpy()
            assert np.abs(A @ Ai - np.eye(4)).max() < 1e-12
            assert 1.5 < np.linalg.cond(A) < 3.5
            assert abs(np.linalg.slogdet(A)[1] - float(p["log_abs_det"])) < 1e-6
            sums.append(float(p["log_abs_det"]))
    assert np.allclose(sums, [-0.1523, 0.0693, 0.2464, 0.0949, 0.0509, 0.3561, 0.0615, 0.0991], atol=2e-4)


def test_fresh_conv1x1_init_is_orthogonal():
    """QR init (layers.py:95) through LU variables and back: P L (U + diag) must be orthogonal."""
    hps = O.make_hps(arch="unc")
    orc = O.OracleNoiseFlow([32, 32, 4], hps, None, seed=3)
    A = orc.model[0][0].params()["A"].numpy()
    assert np.abs(A @ A.T - np.eye(4)).max() < 1e-6
    assert abs(float(orc.model[0][0].params()["log_abs_det"])) < 1e-6


def test_roundtrip_and_sanity_numbers(shipped):
    hps, ck = shipped
    orc = make_oracle(hps, ck)
    for (cam, iso) in [(2, 100), (0, 100), (2, 800)]:
        x, y = synth_batch(16, cam=cam, iso=iso, seed=1)
        nll, sd_z = orc._loss(x, y, iso=[float(iso)], cam=[float(cam)])
        xr = orc.forward(orc.last_z, None, y, iso=[float(iso)], cam=[float(cam)])
        assert float((xr - torch.as_tensor(x, dtype=torch.float64)).abs().max()) < 1e-13
        gen = O.nll_sdn_closed_form(x, y, *CAM_ISO_NLF[(cam, iso)]).mean() / 4096
        # learnt model within ~0.05 nats/dim of the generating NLF, sd_z ~ 0.83-1.0
        assert abs(float(nll.mean()) / 4096 - gen) < 0.06
        assert 0.8 < float(sd_z) < 1.05


def test_logdet_against_bruteforce_jacobian(shipped):
    """log|det d inverse(x) / dx| by autograd on a single patch's 4096x4096 Jacobian ... too big; use the
    chain restricted to its locally coupled structure instead: finite-difference directional check of
    d(log p)/dx consistency plus an exact small case with a 4x4 crop of a fresh perturbed coupling."""
    hps = O.make_hps(arch="unc", flow_permutation=1)
    rng = np.random.RandomState(0)

    class Small(O.OracleNoiseFlow):
        pass
    orc = Small([32, 32, 4], hps, None, seed=1)
    x = torch.as_tensor(rng.randn(1, 32, 32, 4) * 0.5, dtype=torch.float64)
    orc.inverse(x, torch.zeros(1, dtype=torch.float64))
    for k, v in orc.store.vars.items():       # make the coupling non-trivial
        if k.endswith("l_last/W"):
            v.copy_(torch.as_tensor(rng.randn(*v.shape) * 0.2))
        if k.endswith("rescaling_scale0"):
            v.fill_(0.7)
    # exact Jacobian of an 8x8 window is not separable because of conv receptive fields, so compute the
    # full Jacobian of a reduced problem: treat the map restricted to channel pairs -- the coupling leaves
    # x0 untouched and scales x1 elementwise, hence log|det| = sum log_scale + 1024 * log|det A| exactly.
    z, ld = orc.inverse(x, torch.zeros(1, dtype=torch.float64))
    xx = x.clone().requires_grad_(True)
    zz, _ = orc.inverse(xx, torch.zeros(1, dtype=torch.float64))
    # Jacobian-vector products along 6 random directions agree with finite differences (map is smooth)
    for _ in range(6):
        d = torch.as_tensor(rng.randn(*x.shape), dtype=torch.float64)
        jvp = torch.autograd.functional.jvp(lambda t: orc.inverse(t, torch.zeros(1, dtype=torch.float64))[0], x, d)[1]
        h = 1e-6
        fd = (orc.inverse(x + h * d, torch.zeros(1, dtype=torch.float64))[0] -
              orc.inverse(x - h * d, torch.zeros(1, dtype=torch.float64))[0]) / (2 * h)
        assert float((jvp - fd).abs().max()) < 1e-6
    # exact log-det on a tiny 4x4x4 "patch" model sharing the same code path
    hps_s = O.make_hps(arch="unc", flow_permutation=1)
    small = O.OracleNoiseFlow([4, 4, 4], hps_s, None, seed=2)
    xs = torch.as_tensor(rng.randn(1, 4, 4, 4) * 0.5, dtype=torch.float64)
    small.inverse(xs, torch.zeros(1, dtype=torch.float64))
    for k, v in small.store.vars.items():
        if k.endswith("l_last/W") or k.endswith("l_last/b"):
            v.copy_(torch.as_tensor(rng.randn(*v.shape) * 0.3))
        if k.endswith("rescaling_scale0"):
            v.fill_(0.9)
    f = lambda t: small.inverse(t.reshape(1, 4, 4, 4), torch.zeros(1, dtype=torch.float64))[0].reshape(-1)
    J = torch.autograd.functional.jacobian(f, xs.reshape(-1))
    _, ld = small.inverse(xs, torch.zeros(1, dtype=torch.float64))
    assert abs(float(torch.linalg.slogdet(J)[1]) - float(ld[0])) < 1e-9


def test_edge_padding_indicator_matches_definition():
    """add_edge_padding (layers.py:555-583): ring of ones on the 34x34 border, zeros inside, x zero-padded."""
    x = torch.ones(2, 32, 32, 4, dtype=torch.float64)
    p = O.add_edge_padding(x, (3, 3))
    assert p.shape == (2, 34, 34, 5)
    ind = p[0, :, :, 4].numpy()
    assert ind[0].all() and ind[-1].all() and ind[:, 0].all() and ind[:, -1].all() and not ind[1:-1, 1:-1].any()
    assert float(p[0, 0, :, :4].abs().max()) == 0 and float(p[0, 1:-1, 1:-1, :4].min()) == 1


def test_batch_norm_modes_and_moving_update():
    st = O.VariableStore(None)
    x = torch.as_tensor(np.random.RandomState(0).randn(3, 8, 8, 4) * 2 + 1, dtype=torch.float64)
    y = O.batch_norm(st, "s", x, True, name="bn")
    assert np.allclose(y.mean(dim=(0, 1, 2)).numpy(), 0, atol=1e-12)
    m, v = x.mean(dim=(0, 1, 2)), x.var(dim=(0, 1, 2), unbiased=False)
    assert np.allclose(st.vars["s/bn/mean"].numpy(), 0.1 * m.numpy())                 # 0 - 0.1*(0 - m)
    assert np.allclose(st.vars["s/bn/var"].numpy(), 1 - 0.1 * (1 - v.numpy()))
    y2 = O.batch_norm(st, "s", x, False, name="bn")
    assert np.allclose(y2.numpy(), ((x - st.vars["s/bn/mean"]) / torch.sqrt(st.vars["s/bn/var"] + 1e-4)).numpy())


def test_gain_quirk_and_unknown_iso(shipped):
    """Gain/GainEx1/GainEx3 log-det is not multiplied by 4096 (AffineCouplingGain.py:86,96,111,125)."""
    x, y = synth_batch(2, seed=2)
    for tok, full in [("gain", False), ("gain1", False), ("gain3", False), ("gain2", True), ("gain4", True)]:
        orc = O.OracleNoiseFlow([32, 32, 4], O.make_hps(arch=tok), None)
        z, ld = orc.inverse(x, torch.zeros(2, dtype=torch.float64), yy=y, iso=[400.0], cam=[1.0])
        scale = float((torch.as_tensor(x, dtype=torch.float64) / z).flatten()[0])
        expect = -(4096 if full else 1) * math.log(scale)
        assert np.allclose(np.broadcast_to(ld.numpy(), (2,)), expect, rtol=1e-9), tok
    hps, ck = shipped
    orc = make_oracle(hps, ck)
    a, _ = orc._loss(x, y, iso=[500.0], cam=[2.0])          # unknown ISO: g = 0 -> gain = iso (cond_utils.py:226-230)
    sb = [b for b in orc.model[0] if isinstance(b, O.ScaleBijector)][0]
    s500, _ = sb._scale(torch.as_tensor(y, dtype=torch.float64), None, None, [500.0], [2.0])
    v = orc.store.vars
    ocp = np.exp(v["model/sdn_gain/cam_params"].numpy()[:, 2])
    b1 = math.exp(float(v["model/sdn_gain/beta1"][0]) * ocp[0]) / 500.0
    b2 = math.exp(float(v["model/sdn_gain/beta2"][0]) * ocp[1])
    assert np.allclose(s500.numpy(), np.sqrt(b1 * y.astype(np.float64) + b2), rtol=1e-12)
    with pytest.raises(IndexError):
        orc._loss(x, y, iso=[100.0], cam=[9.0])


def test_template_scope_naming_follows_first_call_order(shipped):
    """tf.make_template names scopes at first call: a forward-first graph (NoiseFlowWrapper) maps the
    checkpoint's real_nvp_conv_template (k=0) onto the LAST coupling."""
    hps, ck = shipped
    inv = make_oracle(hps, ck, first_call="inverse")
    fwd = make_oracle(hps, ck, first_call="forward")
    cps_i = [b for b in inv.model[0] if isinstance(b, O.AffineCoupling)]
    cps_f = [b for b in fwd.model[0] if isinstance(b, O.AffineCoupling)]
    assert [b._fn.scope for b in cps_i] == ["model/real_nvp_conv_template"] + ["model/real_nvp_conv_template_%d" % k for k in range(1, 8)]
    assert [b._fn.scope for b in cps_f] == [b._fn.scope for b in cps_i][::-1]
    x, y = synth_batch(2)
    a, _ = inv._loss(x, y, iso=[100.0], cam=[2.0])
    b, _ = fwd._loss(x, y, iso=[100.0], cam=[2.0])
    assert float((a - b).abs().max()) > 1.0      # genuinely different models


def test_squeeze_matches_numpy_reshape_transpose():
    rng = np.random.RandomState(0)
    x = rng.randn(3, 8, 12, 5)
    for f in (2, 4):
        s = O.squeeze2d(x, f, "chessboard")
        ref = x.reshape(3, 8 // f, f, 12 // f, f, 5).transpose(0, 1, 3, 5, 2, 4).reshape(3, 8 // f, 12 // f, 5 * f * f)
        assert np.array_equal(s, ref)
        sp = O.squeeze2d(x, f, "patch")
        refp = x.reshape(3, f, 8 // f, f, 12 // f, 5).transpose(0, 2, 4, 5, 1, 3).reshape(3, 8 // f, 12 // f, 5 * f * f)
        assert np.array_equal(sp, refp)
    x4 = rng.randn(2, 4, 4, 16)
    for t in ("chessboard", "patch"):
        assert np.array_equal(O.squeeze2d(O.unsqueeze2d(x4, 2, t), 2, t), x4)
    assert O.squeeze2d(x, 1) is x


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    c = np.array([[0, 0, 0, 0], [0xffffffff] * 4, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]], dtype=np.uint32)
    k = np.array([[0, 0], [0xffffffff, 0xffffffff], [0xa4093822, 0x299f31d0]], dtype=np.uint32)
    out = O.philox4x32_10(c, k)
    assert [hex(v) for v in out[0]] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(v) for v in out[1]] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(v) for v in out[2]] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]
    e = O.philox_normal(1, 0, 64)
    assert abs(e.mean()) < 0.01 and abs(e.std() - 1) < 0.01


# ---- pins against the initial-value constants of the reference's shipped graph -------------------
# tests/golden/meta_init_constants.json is extracted from models/NoiseFlow/ckpt/model.ckpt.best.meta by
# tools/extract_meta_init.py (a TF-free protobuf walk).
CONV_IDS = [1, 2, 3, 4, 6, 7, 8, 9]


@pytest.fixture(scope="module")
def meta_init(golden_dir):
    import json
    with open(os.path.join(golden_dir, "meta_init_constants.json")) as f:
        c = json.load(f)["constants"]
    return {k: np.asarray(v["values"], np.float64).reshape(v["shape"]) for k, v in c.items()}


def _conv_consts(meta_init, i):
    pre = "level0/bijector%d/Conv2d_1x1_%d/" % (i, i)
    tag = "_matpar_lu_conv2d_1x1_%d_0" % i
    return (pre, tag, meta_init[pre + "P" + tag + "/initial_value"], meta_init[pre + "stricttri2vec/input"],
            meta_init[pre + "stricttri2vec_1/input"], meta_init[pre + "log_S" + tag + "/initial_value"],
            meta_init[pre + "sign_S" + tag + "/initial_value"])


def test_meta_graph_initial_lu_constants_reassemble_to_orthogonal(meta_init, shipped):
    """layers.py:95 initialises every Conv2d1x1 from a random orthogonal matrix and matrix_param.py:100-128 stores its
    decomposition; the graph keeps those constants.  Pushing the reference's own L / U init matrices through OUR
    stricttri2vec -> (oracle and product) LU assembly must give an orthogonal A with log|det| = 0; P and sign_S are
    not trainable, so the shipped checkpoint must still hold exactly these constants."""
    from noise_flow_b200.params import lu_to_matrix, stricttri2vec, vec2stricttri
    _, ck = shipped
    for i in CONV_IDS:
        pre, tag, P, L, U, log_s, sign_s = _conv_consts(meta_init, i)
        assert np.array_equal(np.tril(L), L) and np.array_equal(np.diag(L), np.ones(4)) and np.array_equal(np.triu(U, 1), U)
        l_vec, u_vec = stricttri2vec(L, False), stricttri2vec(U, True)
        assert np.array_equal(O.stricttri2vec(L, upper=False), l_vec) and np.array_equal(O.stricttri2vec(U, upper=True), u_vec)
        assert np.array_equal(vec2stricttri(l_vec, False), L - np.eye(4)) and np.array_equal(vec2stricttri(u_vec, True), U)
        A, A_inv, logdet = lu_to_matrix(P, l_vec, u_vec, log_s, sign_s)
        assert np.abs(A @ A.T - np.eye(4)).max() < 1e-6, i
        assert np.abs(A @ A_inv - np.eye(4)).max() < 1e-6 and abs(logdet) < 1e-6
        store = O.VariableStore({pre + "P" + tag: P, pre + "L_vec" + tag: l_vec, pre + "U_vec" + tag: u_vec,
                                 pre + "log_S" + tag: log_s, pre + "sign_S" + tag: sign_s})
        po = O.matrix_param_lu(store, pre[:-1], tag[len("_matpar_lu_"):], None, 4)
        assert np.abs(po["A"].numpy() - A).max() < 1e-6 and store.created == []
        assert np.array_equal(ck[pre + "P" + tag], P.astype(np.float32))
        assert np.array_equal(ck[pre + "sign_S" + tag], sign_s.astype(np.float32))


def test_trained_lu_vectors_identify_the_vector_ordering(meta_init, shipped):
    """of all 720 orderings of the 6-vector, the TFP fill_triangular ordering we implement is the one
    under which the TRAINED L / U vectors of the shipped checkpoint stay closest to the graph's initial L / U
    matrices (summed over the 8 Conv2d1x1 layers) -- by a wide margin."""
    import itertools
    from noise_flow_b200.params import stricttri2vec
    _, ck = shipped
    perms = list(itertools.permutations(range(6)))
    for which in ("L", "U"):
        agg = np.zeros(len(perms))
        for i in CONV_IDS:
            pre, tag, P, L, U, _, _ = _conv_consts(meta_init, i)
            init = stricttri2vec(L - np.eye(4), False) if which == "L" else stricttri2vec(U, True)
            trained = ck[pre + which + "_vec" + tag].astype(np.float64)
            agg += np.array([((trained[list(p)] - init) ** 2).sum() for p in perms])
        order = np.argsort(agg)
        assert perms[order[0]] == (0, 1, 2, 3, 4, 5)
        assert agg[order[0]] < 0.8 * agg[order[1]] and agg[order[0]] < 0.3 * np.median(agg)


def test_meta_graph_initialisers_match_layer_definitions(meta_init):
    """conv weights ~ N(0, (width/512 * 0.05)^2) (layers.py:598-599), zero biases / logs (:662-673), BN moving stats
    (0, 1) (:383-386).  The sdn_gain initialisers in the shipped graph are fitted values from an older script
    revision; the current code (NoiseFlowWrapper.py:125-137) always passes constants -- they are initialisers only and
    are overwritten by the restore, so they are recorded in the fixture but deliberately NOT mirrored."""
    for t in [""] + ["_%d" % k for k in range(1, 8)]:
        pre = "model/real_nvp_conv_template%s/" % t
        for lname, shape in (("l_1", [3, 3, 2, 4]), ("l_2", [1, 1, 4, 4])):
            assert abs(float(meta_init[pre + lname + "/W/Initializer/random_normal/stddev"]) - 4 / 512 * 0.05) < 1e-9
            assert list(meta_init[pre + lname + "/W/Initializer/random_normal/shape"]) == shape
        assert not meta_init[pre + "l_last/logs/Initializer/zeros"].any()
        assert not meta_init[pre + "bn_nvp_conv_1/mean/Initializer/zeros"].any()
        assert (meta_init[pre + "bn_nvp_conv_2/var/Initializer/ones"] == 1).all()
    hps = O.make_hps(arch="unc")
    orc = O.OracleNoiseFlow([32, 32, 4], hps, None, seed=11)
    x, y = synth_batch(1)
    orc._loss(x, y, iso=[100.0], cam=[2.0])
    w1 = orc.store.vars["model/real_nvp_conv_template/l_1/W"].numpy()
    assert w1.shape == (3, 3, 2, 4) and 0.3 < w1.std() / (4 / 512 * 0.05) < 2.0
    assert meta_init["model/sdn_gain/cam_params/Initializer/Const"].shape == (3, 5)
