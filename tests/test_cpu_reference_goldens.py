"""The oracle (and the host logic) against golden vectors produced by the REFERENCE'S OWN Python.

`tests/golden/ref_*.npz` were generated in the build container by `oracle/make_reference_goldens.py`: the reference's
unmodified `borealisflows/*.py` (including its `NoiseFlowWrapper` class) executed over the TF-1.12 API stand-in of
`oracle/tf1_shim.py` in double precision, following the graph-construction order and `sess.run` calls of
`train_noise_flow.py` / `NoiseFlowWrapper.py`.  Here the independently written restatement `oracle/noise_flow_oracle.py`
(fp64) must reproduce them to round-off, which is what pins it; the GPU suite then checks the CUDA path against the same
files (tests/test_gpu_reference_goldens.py)."""
import os

import numpy as np
import pytest
import torch

from common import make_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name))


@pytest.fixture(scope="module")
def tg():
    return _load("ref_training_graph.npz")


@pytest.fixture(scope="module")
def shipped():
    from noise_flow_b200 import hps_loader, load_checkpoint
    hps = hps_loader(os.path.join(GOLD, "NoiseFlow", "hps.txt"))
    ck = load_checkpoint(os.path.join(GOLD, "NoiseFlow", "ckpt", "model.ckpt.best"))
    return hps, ck


def _oracle_grads(hps, variables, x, y, iso, cam, is_training):
    orc = make_oracle(hps, variables)
    orc._loss(x[:1], y[:1], iso=[iso], cam=[cam], is_training=False)
    params = {k: v for k, v in orc.store.vars.items() if orc.store.trainable.get(k, False)}
    for v in params.values():
        v.requires_grad_(True)
    loss, sd_z = orc.loss(x, y, iso=[iso], cam=[cam], is_training=is_training)
    loss.backward()
    return float(loss.detach()), float(sd_z.detach()), {k: (v.grad.numpy().copy() if v.grad is not None else np.zeros(tuple(v.shape))) for k, v in params.items()}, orc


def test_graph_inventory_matches_checkpoint_and_hps(tg, shipped):
    """The graph the reference's code builds: layer names == hps.txt:1-18, variables == the shipped checkpoint's 143
    tensors (nothing missing, nothing unused), 2433 trainable parameters (hps.txt:19) -- and our model says the same."""
    hps, ck = shipped
    with open(os.path.join(GOLD, "NoiseFlow", "hps.txt")) as f:
        lines = [l.strip() for l in f.readlines()]
    assert list(tg["layer_names"]) == lines[:18]
    assert int(tg["num_params"]) == 2433 == int(lines[18])
    assert set(tg["var_names"]) == set(ck) and len(tg["ckpt_unused"]) == 0
    shapes = dict(zip(tg["var_names"], tg["var_shapes"]))
    for k, v in ck.items():
        assert shapes[k] == ",".join(map(str, v.shape)), k
    orc = make_oracle(hps, ck)
    assert orc.get_layer_names() == list(tg["layer_names"])
    trainable = dict(zip(tg["var_names"], tg["var_trainable"]))
    orc._loss(tg["x"][:1], tg["y"][:1], iso=[100.0], cam=[2.0])
    assert {k for k, t in trainable.items() if t} == {k for k, t in orc.store.trainable.items() if t}
    # conv initialiser: N(0, (width/512*0.05)^2) (layers.py:598-599) -> std 3.9e-4 over the 8 l_1/W tensors
    assert abs(float(tg["init_std_l_1_W"]) - 4 / 512 * 0.05) < 1e-4


def test_oracle_reproduces_reference_loss_z_and_samples(tg, shipped):
    hps, ck = shipped
    orc = make_oracle(hps, ck)
    x, y, eps = tg["x"], tg["y"], tg["eps"]
    nll, sd_z = orc._loss(x, y, iso=[100.0], cam=[2.0])
    assert np.abs(nll.numpy() - tg["nll"]).max() < 1e-8             # |nll| ~ 1.2e4: round-off of fp64
    assert abs(float(sd_z) - float(tg["sd_z"])) < 1e-12
    assert abs(float(nll.mean()) - float(tg["loss"])) < 1e-8
    z, obj = orc.inverse(x, torch.zeros(len(x), dtype=torch.float64), yy=y, iso=[100.0], cam=[2.0])
    assert np.abs(z.numpy() - tg["z"]).max() < 1e-11 and np.abs(obj.numpy() - tg["logdet"]).max() < 1e-8
    assert float(tg["roundtrip_err"]) < 1e-14                       # forward(inverse(x)) == x in the reference
    for temp in (1.0, 0.6):
        xs = orc.sample(eps, temp, y, iso=[100.0], cam=[2.0]).numpy()
        assert np.abs(xs - tg["sample_T%g" % temp]).max() < 1e-13
    xs = orc.sample(eps, 1.0, y, iso=[100.0], cam=[2.0], is_training=True).numpy()
    assert np.abs(xs - tg["sample_batch_T1"]).max() < 1e-12
    # the moving statistics that run moved: m <- m - 0.1 (m - batch) (layers.py:394-395), all 32 of them
    moved = [k[len("after_sample_batch/"):] for k in tg.files if k.startswith("after_sample_batch/")]
    assert len(moved) == 32
    for k in moved:
        assert np.abs(orc.store.vars[k].numpy() - tg["after_sample_batch/" + k]).max() < 1e-12, k


@pytest.mark.parametrize("is_training", [False, True])
def test_oracle_gradients_match_reference_graph(tg, shipped, is_training):
    """d loss / d every trainable variable: torch autograd through the oracle vs tf.gradients of the reference graph."""
    hps, ck = shipped
    loss, sd_z, grads, _ = _oracle_grads(hps, ck, tg["x"], tg["y"], 100.0, 2.0, is_training)
    prefix = "grad_batch/" if is_training else "grad_moving/"
    names = [k[len(prefix):] for k in tg.files if k.startswith(prefix)]
    assert sum(tg[prefix + k].size for k in names) == 2433 and set(names) == set(grads)
    if is_training:
        assert abs(loss - float(tg["train_loss"])) < 1e-8 and abs(sd_z - float(tg["train_sd_z"])) < 1e-12
    for k in names:
        g = tg[prefix + k]
        assert np.abs(grads[k].reshape(g.shape) - g).max() <= 1e-9 * max(1.0, np.abs(g).max()), k


def test_adam_and_moving_averages_match_reference_train_step(tg, shipped):
    """sess.run([train_op, loss, sd_z], is_training=True) twice: our TF-rule Adam on the reference's gradients must
    land on the reference's variables, and the BatchNorm statistics must move as `layers.py:394-395` says."""
    from noise_flow_b200.train import AdamOptimizer
    hps, ck = shipped
    assert int(tg["train_bn_updates"]) == 32
    grads = {k[len("grad_batch/"):]: tg[k] for k in tg.files if k.startswith("grad_batch/")}
    vs = {k: np.asarray(v, dtype=np.float64) for k, v in ck.items()}
    opt = AdamOptimizer(learning_rate=1e-4)
    opt.apply_gradients(vs, grads)
    changed = {k[len("after_step/"):] for k in tg.files if k.startswith("after_step/")}
    moving = {k for k in ck if k.endswith("/mean") or k.endswith("/var")}
    unused = {k for k, g in grads.items() if not np.any(g)}        # rescaling_scale0 of the two scale layers
    assert unused == {"level0/bijector0/rescaling_scale0", "level0/bijector5/rescaling_scale0"}
    assert changed == (set(grads) - unused) | moving
    for k in set(grads) - unused:
        # the first Adam step moves every parameter by lr * g / (|g| + eps): fp32 storage on our side
        assert np.abs(vs[k].reshape(tg["after_step/" + k].shape) - tg["after_step/" + k]).max() < 2e-7 * max(1.0, np.abs(ck[k]).max()), k
    # moving statistics after the step == oracle's after one batch-statistics forward
    _, _, _, orc = _oracle_grads(hps, ck, tg["x"], tg["y"], 100.0, 2.0, True)
    for k in ck:
        if k.endswith("/mean") or k.endswith("/var"):
            assert np.abs(orc.store.vars[k].detach().numpy() - tg["after_step/" + k]).max() < 1e-12, k
    # second step: the oracle on the reference's post-step-1 variables must give the reference's second loss
    v1 = {k: tg["after_step/" + k] if ("after_step/" + k) in tg.files else ck[k] for k in ck}
    loss2, _, _, _ = _oracle_grads(hps, v1, tg["x"], tg["y"], 100.0, 2.0, True)
    assert abs(loss2 - float(tg["train_loss_step2"])) < 1e-7


def test_wrapper_graph_template_order_and_samples(shipped):
    """The reference's NoiseFlowWrapper builds only the sampling op, so `tf.make_template` hands the scopes out in
    latent->data order and Saver.restore loads net k into coupling 7-k.  The goldens come from the
    reference's class itself; the oracle in `first_call="forward"` order with batch statistics must reproduce them."""
    hps, ck = shipped
    wg = _load("ref_wrapper_graph.npz")
    scopes = dict(s.split(":") for s in wg["template_scopes"])
    couplings = sorted(int(i) for i in scopes)
    assert len(couplings) == 8
    for rank, i in enumerate(couplings):
        k = 7 - rank
        assert scopes[str(i)] == "model/real_nvp_conv_template" + ("_%d" % k if k else "")
    assert set(wg["var_names"]) == set(ck)
    orc = make_oracle(hps, ck, first_call="forward")
    xs = orc.sample(wg["eps"], float(wg["temp"]), wg["y"], nlf0=[float(wg["b1"])], nlf1=[float(wg["b2"])],
                    iso=[float(wg["iso"])], cam=[float(wg["cam"])], is_training=True).numpy()
    assert np.abs(xs - wg["sample"]).max() < 1e-12 and wg["sample"].dtype == np.float64
    moved = [k[len("after_call/"):] for k in wg.files if k.startswith("after_call/")]
    assert len(moved) == 32
    for k in moved:
        assert np.abs(orc.store.vars[k].numpy() - wg["after_call/" + k]).max() < 1e-12, k
    xs2 = orc.sample(wg["eps_call2"], float(wg["temp"]), wg["y_call2"], iso=[800.0], cam=[0.0], is_training=True).numpy()
    assert np.abs(xs2 - wg["sample_call2"]).max() < 1e-12
    # and the training-order oracle does NOT (the two orders really differ)
    other = make_oracle(hps, ck, first_call="inverse").sample(wg["eps"], float(wg["temp"]), wg["y"], iso=[100.0], cam=[2.0],
                                                               is_training=True).numpy()
    assert np.abs(other - wg["sample"]).max() > 1e-3


def _arch_cases(fname="ref_arch_cases.npz"):
    ac = _load(fname)
    tags = sorted({k.split("::")[0] for k in ac.files})
    return ac, tags


@pytest.mark.parametrize("fname,tag", [("ref_arch_cases.npz", t) for t in _arch_cases()[1]] +
                         [("ref_wide_cases.npz", t) for t in _arch_cases("ref_wide_cases.npz")[1]])
def test_oracle_reproduces_reference_arch_cases(fname, tag):
    """Every token `noise_flow_arch` parses (all sdn* / gain* layers incl. the log-det quirks and the unknown-ISO
    fall-backs, the three `flow_permutation` settings), perturbed variables, both BatchNorm modes; coupling-net widths
    4 ... 512 (ref_wide_cases.npz: 64 / 128 / 256 / 512, the reference's default `--width`)."""
    from noise_flow_b200 import make_hps
    ac, _ = _arch_cases(fname)
    g = {k.split("::", 1)[1]: ac[k] for k in ac.files if k.startswith(tag + "::")}
    hps = make_hps(arch=str(g["arch"]), flow_permutation=int(g["flow_permutation"]), width=int(g["width"]))
    variables = {k[len("var/"):]: v for k, v in g.items() if k.startswith("var/")}
    orc = make_oracle(hps, variables)
    assert orc.get_layer_names() == list(g["layer_names"])
    a = dict(nlf0=[float(g["nlf0"])], nlf1=[float(g["nlf1"])], iso=[float(g["iso"])], cam=[float(g["cam"])])
    nll, sd_z = orc._loss(g["x"], g["y"], **a)
    assert not orc.store.created, "variables the reference graph does not have: %s" % orc.store.created[:4]
    assert set(orc.store.vars) == set(variables)
    assert orc.store.num_trainable() == int(g["num_params"])
    scale = max(1.0, np.abs(g["nll"]).max())
    assert np.abs(nll.numpy() - g["nll"]).max() < 1e-9 * scale and abs(float(sd_z) - float(g["sd_z"])) < 1e-9
    z, obj = orc.inverse(g["x"], torch.zeros(len(g["x"]), dtype=torch.float64), yy=g["y"], **a)
    assert np.abs(z.numpy() - g["z"]).max() < 1e-6 * max(1.0, np.abs(g["z"]).max())       # stored as fp32
    assert np.abs(obj.numpy() - g["logdet"]).max() < 1e-9 * scale
    xs = orc.sample(g["eps"], 0.6, g["y"], **a).numpy()
    assert np.abs(xs - g["sample_T0.6"]).max() < 1e-6 * max(1.0, np.abs(g["sample_T0.6"]).max())
    nll_b, sd_b = orc._loss(g["x"], g["y"], is_training=True, **a)
    assert np.abs(nll_b.numpy() - g["nll_batch"]).max() < 1e-9 * scale and abs(float(sd_b) - float(g["sd_z_batch"])) < 1e-9


def test_squeeze_matches_reference_bit_exact():
    from oracle.noise_flow_oracle import squeeze2d, unsqueeze2d
    sq = _load("ref_squeeze.npz")
    x = torch.from_numpy(sq["x"])
    for factor in (1, 2):
        for kind in ("chessboard", "patch", "bogus"):
            s = squeeze2d(x, factor, kind)
            assert np.array_equal(s.numpy(), sq["squeeze_%d_%s" % (factor, kind)])
            key = "unsqueeze_%d_%s" % (factor, kind)
            if key in sq.files:
                assert np.array_equal(unsqueeze2d(s, factor, kind).numpy(), sq[key])


def test_metrics_match_reference_numpy_code():
    """The evaluation metrics next to the path: the oracle's restatements and the host-side functions of
    `noise_flow_b200.metrics` against the reference's own `sidd_utils` / `PatchStatsCalculator` code (ref_metrics.npz)."""
    from noise_flow_b200 import metrics
    from oracle.noise_flow_oracle import calc_baselines, get_histogram, kl_div_forward
    g = _load("ref_metrics.npz")
    x, y, xs, edges = g["x"], g["y"], g["x_sampled"], g["bin_edges"]
    assert len(edges) == 67                                      # 64 bins on [-0.1, 0.1] + two catch-all bins
    hp, hq = get_histogram(x, edges), get_histogram(xs, edges)
    assert np.array_equal(hp, g["hist_p"]) and np.array_equal(hq, g["hist_q"])          # counts / n: exact
    assert abs(kl_div_forward(hp, hq) - float(g["kl_forward"])) < 1e-15
    for fn, key in ((metrics.kl_div_forward, "kl_forward"), (metrics.kl_div_inverse, "kl_inverse"), (metrics.kl_div_sym, "kl_sym")):
        assert abs(fn(g["hist_p"], g["hist_q"]) - float(g[key])) < 1e-15, key
    assert np.allclose(g["kl_3_data_edges"], [g["kl_forward"], g["kl_inverse"], g["kl_sym"]], rtol=1e-12, atol=1e-15)
    nll_g, nll_s = calc_baselines(x, y, float(g["nlf0"]), float(g["nlf1"]), float(g["var_gauss"]))
    # the reference averages per minibatch (two minibatches of three patches): same as the mean over the six patches
    assert abs(nll_g.mean() - float(g["nll_gauss_mean"])) < 1e-9 * abs(float(g["nll_gauss_mean"]))
    assert abs(nll_s.mean() - float(g["nll_sdn_mean"])) < 1e-9 * abs(float(g["nll_sdn_mean"]))
    # bits per dimension as logged by the driver (sidd_utils.py:879-881)
    bpd = (float(g["nll_sdn_mean"]) / 4096 + np.log(256)) / np.log(2.0)
    assert abs(bpd - float(g["bpd_of_nll_sdn"])) < 1e-12
    # Bayer packing used around the sampler (sample_noise_flow.py:74-79): pack -> unpack is the identity
    assert np.array_equal(g["unpacked"], g["bayer"]) and g["packed"].shape == (4, 6, 4)


def test_host_model_spec_reproduces_reference_legacy_graphs():
    """The product's host-side assembly (noise_flow_b200.params.ModelSpec) of the legacy revnet2d models: bijector names,
    exactly the reference graph's variables (nothing created, nothing unused), trainable-parameter count."""
    from noise_flow_b200 import make_hps
    from noise_flow_b200.params import ModelSpec
    lc = _load("ref_legacy_cases.npz")
    for tag in sorted({k.split("::")[0] for k in lc.files}):
        g = {k.split("::", 1)[1]: lc[k] for k in lc.files if k.startswith(tag + "::")}
        flags = {str(f): True for f in g["flags"]}
        hps = make_hps(arch=None, depth=int(g["depth"]), sidd_cond=str(g["sidd_cond"]), flow_permutation=int(g["flow_permutation"]), **flags)
        variables = {k[len("var/"):]: v for k, v in g.items() if k.startswith("var/")}
        spec = ModelSpec(hps, variables)
        spec.assign_template_scopes("inverse")
        spec.create_scale_variables()
        assert spec.get_layer_names() == list(g["layer_names"]), tag
        assert not spec.store.created and set(spec.store.vars) == set(variables), tag
        assert spec.store.num_trainable() == int(g["num_params"]), tag


def _legacy_tags():
    lc = _load("ref_legacy_cases.npz")
    return sorted({k.split("::")[0] for k in lc.files})


@pytest.mark.parametrize("tag", _legacy_tags())
def test_oracle_reproduces_reference_legacy_revnet2d_cases(tag):
    """`hps.arch` unset -> `revnet2d` (noise_flow_model.py:237-392): the clean-image-conditioned couplings CondY / CondYG /
    CondXY / CondXYG (the G variants with ISO-conditioned convolutions), CamSdn, the ISO-polynomial SdnGain / FitSdnGain2
    layers and the append_* options; the goldens are the reference's own classes executed over the TF stand-in (the CUDA
    path runs the same cases in tests/test_gpu_reference_goldens.py)."""
    from types import SimpleNamespace
    lc = _load("ref_legacy_cases.npz")
    g = {k.split("::", 1)[1]: lc[k] for k in lc.files if k.startswith(tag + "::")}
    hps = SimpleNamespace(arch=None, depth=int(g["depth"]), sidd_cond=str(g["sidd_cond"]), flow_permutation=int(g["flow_permutation"]),
                          width=4, decomp="LU", squeeze_factor=1, squeeze_type="chessboard", n_levels=1, gain_init=-5.0,
                          x_shape=[None, 32, 32, 4], append_sdn2=False, append_sdn_first=False, append_cY=False, append_sdn=False)
    for f in g["flags"]:
        setattr(hps, str(f), True)
    variables = {k[len("var/"):]: v for k, v in g.items() if k.startswith("var/")}
    orc = make_oracle(hps, variables)
    assert orc.get_layer_names() == list(g["layer_names"])
    a = dict(nlf0=[float(g["nlf0"])], nlf1=[float(g["nlf1"])], iso=[float(g["iso"])], cam=[float(g["cam"])])
    nll, sd_z = orc._loss(g["x"], g["y"], **a)
    assert not orc.store.created, "variables the reference graph does not have: %s" % orc.store.created[:4]
    assert set(orc.store.vars) == set(variables) and orc.store.num_trainable() == int(g["num_params"])
    scale = max(1.0, np.abs(g["nll"]).max())
    assert np.abs(nll.numpy() - g["nll"]).max() < 1e-9 * scale and abs(float(sd_z) - float(g["sd_z"])) < 1e-9 * max(1.0, float(g["sd_z"]))
    if str(g["sidd_cond"]) == "uncond":
        xs = orc.sample(g["eps"], 0.6).numpy()
    else:
        xs = orc.sample(g["eps"], 0.6, g["y"], **a).numpy()
    assert np.abs(xs - g["sample_T0.6"]).max() < 1e-6 * max(1.0, np.abs(g["sample_T0.6"]).max())       # stored as fp32
    nll_b, sd_b = orc._loss(g["x"], g["y"], is_training=True, **a)
    assert np.abs(nll_b.numpy() - g["nll_batch"]).max() < 1e-9 * max(1.0, np.abs(g["nll_batch"]).max())
    assert abs(float(sd_b) - float(g["sd_z_batch"])) < 1e-9 * max(1.0, float(g["sd_z_batch"]))
