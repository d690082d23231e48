"""GPU parity of the train-step machinery (BASELINE config 5): d(mean NLL)/d(every trainable variable) against
torch autograd through the CPU oracle, in both BatchNorm modes, and one Adam step with TensorFlow's update rule."""
import copy
import math

import numpy as np
import pytest
import torch

from common import make_oracle, synth_batch

pytestmark = pytest.mark.gpu


def _oracle_loss_and_grads(hps, ck, x, y, iso, cam, is_training):
    orc = make_oracle(hps, ck)
    orc._loss(x[:1], y[:1], iso=[iso], cam=[cam], is_training=False)          # create every variable
    params = {k: v for k, v in orc.store.vars.items() if orc.store.trainable.get(k, False)}
    for v in params.values():
        v.requires_grad_(True)
    loss, sd_z = orc.loss(x, y, iso=[iso], cam=[cam], is_training=is_training)
    loss.backward()
    grads = {k: (v.grad.numpy().copy() if v.grad is not None else np.zeros(tuple(v.shape))) for k, v in params.items()}
    return float(loss.detach()), float(sd_z.detach()), grads, orc


def _check(grads, grads_o, rel):
    assert set(grads) == set(grads_o)
    worst, bad = 0.0, []
    for k, go in sorted(grads_o.items()):
        g = grads[k]
        assert g.shape == go.shape, k
        # biases in front of a batch-statistics BatchNorm have an exactly-zero gradient: absolute floor 1e-3
        # (typical gradient magnitudes here are 1..1000), relative tolerance otherwise
        scale = max(np.abs(go).max(), 1e-3 / rel)
        if (k.endswith("/l_1/b") or k.endswith("/l_2/b")) and np.abs(go).max() < 1e-6:
            scale = max(scale, 2e-2 / rel)      # exact zero in the oracle: ours is fp32 summation noise, bound 2e-2 absolute
        err = np.abs(g - go).max() / scale
        worst = max(worst, err)
        if not err < rel:
            bad.append((k, float(err), g.ravel()[:3], go.ravel()[:3]))
    assert not bad, "\n".join("%s err %.3g got %s want %s" % b for b in bad[:40])
    return worst


@pytest.mark.parametrize("is_training", [False, True])
def test_gradients_match_oracle_autograd(shipped, is_training):
    from noise_flow_b200 import NoiseFlow
    from noise_flow_b200.train import loss_and_grad
    hps, ck = shipped
    x, y = synth_batch(6, cam=2, iso=100, seed=91)
    nf = NoiseFlow([32, 32, 4], is_training, copy.copy(hps), variables=ck, device="cuda:0", first_call="inverse")
    loss, sd_z, grads = loss_and_grad(nf, x, y, iso=[100.0], cam=[2.0], is_training=is_training)
    loss_o, sd_o, grads_o, _ = _oracle_loss_and_grads(hps, ck, x, y, 100.0, 2.0, is_training)
    assert abs(loss - loss_o) / 4096 < 1e-4 and abs(sd_z - sd_o) < 1e-4
    assert sum(g.size for g in grads_o.values()) == 2433
    worst = _check(grads, grads_o, rel=5e-4)
    print("max relative gradient error (per tensor, vs max |g|): %.2e" % worst)


def test_gradients_other_arch_per_patch_rows():
    """Fresh perturbed model with permutation instead of 1x1 conv, a stand-alone scale mix, per-patch (cam, iso)."""
    from noise_flow_b200 import NoiseFlow, make_hps
    from noise_flow_b200.train import loss_and_grad
    hps = make_hps(arch="sdn4|unc|gain2|unc", flow_permutation=1)
    nf0 = NoiseFlow([32, 32, 4], False, copy.copy(hps), device="cuda:0", seed=3, first_call="inverse")
    rng = np.random.RandomState(7)
    vs = {k: v.copy() for k, v in nf0.variables.items()}
    for k in vs:
        if k.endswith("/l_1/W") or k.endswith("/l_2/W"):
            vs[k] = (rng.randn(*vs[k].shape) * 0.5).astype(np.float32)
        elif k.endswith("/l_last/W"):
            vs[k] = (rng.randn(*vs[k].shape) * 0.1).astype(np.float32)
        elif k.endswith("/b") or k.endswith("/logs"):
            vs[k] = (rng.randn(*vs[k].shape) * 0.2).astype(np.float32)
        elif "rescaling_scale" in k:
            vs[k] = np.float32(0.5)
    x, y = synth_batch(5, cam=2, iso=800, seed=93)
    x = (x * 3).astype(np.float32)
    nf = NoiseFlow([32, 32, 4], True, copy.copy(hps), variables=vs, device="cuda:0", first_call="inverse")
    loss, sd_z, grads = loss_and_grad(nf, x, y, iso=[800.0], cam=[1.0], is_training=True)
    loss_o, _, grads_o, _ = _oracle_loss_and_grads(hps, vs, x, y, 800.0, 1.0, True)
    assert abs(loss - loss_o) / 4096 < 1e-4
    _check(grads, grads_o, rel=5e-4)


def test_adam_train_step_matches_tf_update_rule(shipped):
    from noise_flow_b200 import NoiseFlow
    from noise_flow_b200.train import AdamOptimizer, train_step
    hps, ck = shipped
    x, y = synth_batch(4, seed=95)
    nf = NoiseFlow([32, 32, 4], True, copy.copy(hps), variables=ck, device="cuda:0", first_call="inverse")
    opt = AdamOptimizer(learning_rate=1e-4)
    loss0, _ = train_step(nf, opt, x, y, iso=[100.0], cam=[2.0])
    _, _, grads_o, _ = _oracle_loss_and_grads(hps, ck, x, y, 100.0, 2.0, True)
    # first Adam step: m = (1-b1) g, v = (1-b2) g^2, lr_t = lr*sqrt(1-b2)/(1-b1)  ->  step = lr * g / (|g| + eps')
    for k, go in grads_o.items():
        if np.abs(go).max() < 1e-6:      # exactly-zero gradients (biases in front of batch-stat BN): Adam's
            continue                     # normalised step amplifies fp32 noise there, in TF as much as here
        lr_t = 1e-4 * math.sqrt(1 - 0.999) / (1 - 0.9)
        expect = ck[k].astype(np.float64) - lr_t * (0.1 * go) / (np.sqrt(0.001 * go * go) + 1e-8)
        got = nf.variables[k].astype(np.float64)
        big = np.abs(go) > 1e-3 * max(np.abs(go).max(), 1e-12)          # where the sign of g is well determined
        assert np.abs(got - expect)[big].max() < 2e-6 if big.any() else True, k
    # the step goes downhill
    loss1, _ = train_step(nf, opt, x, y, iso=[100.0], cam=[2.0])
    assert np.isfinite(loss1) and loss1 < loss0 + 1.0
