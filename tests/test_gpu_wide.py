"""Coupling nets wider than the shipped width 4 (``--width``, sidd/ArgParser.py:43, job_noise_flow.sh:19): the CTA-per-patch
CUDA-core kernel (csrc/nf_wide.cu, widths 8 / 16 / 32) and the tensor-core kernels (csrc/nf_wide_tc.cu, widths 32 / 64 / 128,
and csrc/nf_wide_tcs.cu, widths 256 / 512 with streamed weights: tcgen05.mma with activations and accumulators in tensor
memory; the default where they exist) against the CPU oracle in
both directions and both BatchNorm modes."""
import copy

import numpy as np
import pytest
import torch

from common import make_oracle, synth_batch

pytestmark = pytest.mark.gpu


def _perturbed_model(width, arch="sdn5|unc|unc|gain4|unc", perm=1, seed=11):
    """Fresh model with non-trivial weights (the reference initialises l_last to zero = identity coupling)."""
    from noise_flow_b200 import NoiseFlow, make_hps
    hps = make_hps(arch=arch, width=width, flow_permutation=perm)
    nf0 = NoiseFlow([32, 32, 4], False, copy.copy(hps), device="cuda:0", seed=seed, first_call="inverse")
    rng = np.random.RandomState(seed)
    vs = {k: v.copy() for k, v in nf0.variables.items()}
    for k in vs:
        if k.endswith("/l_1/W"):
            vs[k] = (rng.randn(*vs[k].shape) * 0.5).astype(np.float32)
        elif k.endswith("/l_2/W"):
            vs[k] = (rng.randn(*vs[k].shape) * (1.0 / np.sqrt(width))).astype(np.float32)
        elif k.endswith("/l_last/W"):
            vs[k] = (rng.randn(*vs[k].shape) * (0.2 / np.sqrt(width))).astype(np.float32)
        elif k.endswith("/b") or k.endswith("/logs"):
            vs[k] = (rng.randn(*vs[k].shape) * 0.2).astype(np.float32)
        elif k.endswith("/mean"):
            vs[k] = (rng.randn(*vs[k].shape) * 0.1).astype(np.float32)
        elif k.endswith("/var"):
            vs[k] = (0.5 + rng.rand(*vs[k].shape)).astype(np.float32)
        elif "rescaling_scale" in k:
            vs[k] = np.float32(0.5)
    return hps, vs


@pytest.mark.parametrize("width,tensor_cores", [(8, None), (16, None), (32, False), (32, True), (64, True), (128, True),
                                                (256, True), (512, True)])
def test_wide_log_prob_and_sample_match_oracle(width, tensor_cores):
    from noise_flow_b200 import NoiseFlow
    hps, vs = _perturbed_model(width)
    nf = NoiseFlow([32, 32, 4], False, copy.copy(hps), variables=vs, device="cuda:0", first_call="inverse")
    if tensor_cores is not None:
        nf.set_tensor_cores(tensor_cores)
    orc = make_oracle(hps, vs)
    x, y = synth_batch(5, cam=2, iso=100, seed=31)
    x = (x * 20).astype(np.float32)                 # O(1) inputs so that the coupling nets are exercised
    nll, sd_z, z = nf._loss(x, y, iso=[100.0], cam=[2.0], return_z=True)
    nll_o, sd_o = orc._loss(x, y, iso=[100.0], cam=[2.0])
    assert np.abs(nll.cpu().numpy() - nll_o.numpy()).max() / 4096 < 1e-4          # tolerance: 1e-4 nats / dim
    assert abs(float(sd_z) - float(sd_o)) < 1e-4
    z_o, _ = orc.inverse(x, torch.zeros(5, dtype=torch.float64), y, iso=[100.0], cam=[2.0])
    assert np.abs(z.cpu().numpy() - z_o.numpy()).max() < 2e-4 * max(1.0, float(z_o.abs().max()))
    # sampling direction with injected eps, and the round trip
    eps = np.random.RandomState(5).randn(5, 32, 32, 4).astype(np.float32)
    xs = nf.sample(y, 0.6, y, iso=[100.0], cam=[2.0], eps=eps).cpu().numpy()
    xo = orc.sample(eps, 0.6, y, iso=[100.0], cam=[2.0]).numpy()
    assert np.abs(xs - xo).max() < 2e-4 * max(1.0, np.abs(xo).max())
    back = nf.forward(z, None, y, iso=[100.0], cam=[2.0]).cpu().numpy()
    assert np.abs(back - x).max() < 1e-4 * max(1.0, np.abs(x).max())


def test_wide_per_bijector_and_permutation():
    """run_layers over single bijectors (partial programs upload their own blob), channel permutation, per-patch rows."""
    from noise_flow_b200 import NoiseFlow
    hps, vs = _perturbed_model(16, arch="unc|sdn4|unc|gain4", perm=0, seed=13)
    nf = NoiseFlow([32, 32, 4], False, copy.copy(hps), variables=vs, device="cuda:0", first_call="inverse")
    orc = make_oracle(hps, vs)
    x, y = synth_batch(4, cam=2, iso=100, seed=33)
    x = (x * 20).astype(np.float32)
    cams, isos = [2.0, 0.0, 1.0, 4.0], [100.0, 800.0, 400.0, 3200.0]
    nll = nf._loss(x, y, iso=isos, cam=cams)[0].cpu().numpy()
    for k in range(4):
        nll_o, _ = orc._loss(x[k:k + 1], y[k:k + 1], iso=[isos[k]], cam=[cams[k]])
        assert abs(nll[k] - float(nll_o[0])) / 4096 < 1e-4
    n_layers = len(nf.get_layer_names())
    cur = torch.as_tensor(x).cuda()
    total = torch.zeros(4, device="cuda")
    for l in range(n_layers):
        cur, ld = nf.run_layers(l, l + 1, "inverse", cur, y, iso=[100.0], cam=[2.0])
        total += ld
    z, ld_all = nf.inverse(x, None, y, iso=[100.0], cam=[2.0])
    assert float((cur - z).abs().max()) < 1e-4 * max(1.0, float(z.abs().max()))
    assert float((total - ld_all).abs().max()) / 4096 < 1e-5


@pytest.mark.parametrize("width,tensor_cores", [(8, None), (32, False), (32, True), (64, True), (128, True), (256, True),
                                                (512, True)])
def test_wide_batch_statistics_mode_matches_oracle(width, tensor_cores):
    """is_training=True: BatchNorm on the statistics of the batch, moving statistics updated (layers.py:388-398)."""
    from noise_flow_b200 import NoiseFlow
    hps, vs = _perturbed_model(width, arch="sdn5|unc|gain4|unc")
    nf = NoiseFlow([32, 32, 4], True, copy.copy(hps), variables={k: v.copy() for k, v in vs.items()}, device="cuda:0",
                   first_call="inverse")
    if tensor_cores is not None:
        nf.set_tensor_cores(tensor_cores)
    orc = make_oracle(hps, vs)
    x, y = synth_batch(6, cam=2, iso=100, seed=35)
    x = (x * 20).astype(np.float32)
    nll, sd_z = nf._loss(x, y, iso=[100.0], cam=[2.0], is_training=True)
    nll_o, sd_o = orc._loss(x, y, iso=[100.0], cam=[2.0], is_training=True)
    assert np.abs(nll.cpu().numpy() - nll_o.numpy()).max() / 4096 < 1e-4
    for k, v in orc.store.vars.items():
        if k.endswith("/mean") or k.endswith("/var"):
            assert np.allclose(nf.variables[k], v.detach().numpy(), rtol=1e-4, atol=1e-5), k
    eps = np.random.RandomState(6).randn(6, 32, 32, 4).astype(np.float32)
    xs = nf.sample(y, 0.6, y, iso=[100.0], cam=[2.0], eps=eps, is_training=True).cpu().numpy()
    xo = orc.sample(eps, 0.6, y, iso=[100.0], cam=[2.0], is_training=True).numpy()
    assert np.abs(xs - xo).max() < 5e-4 * max(1.0, np.abs(xo).max())


def test_wide_philox_sampling_statistics_and_large_batch():
    """In-kernel Philox sampling on more patches than CTAs; mean NLL of the sampled noise is finite and stable."""
    from noise_flow_b200 import NoiseFlow
    hps, vs = _perturbed_model(8)
    nf = NoiseFlow([32, 32, 4], False, copy.copy(hps), variables=vs, device="cuda:0", first_call="inverse")
    n = 700
    y = torch.rand((n, 32, 32, 4), device="cuda")
    xs = nf.sample(y, 1.0, y, iso=[100.0], cam=[2.0], seed=3, offset=0)
    xs2 = nf.sample(y, 1.0, y, iso=[100.0], cam=[2.0], seed=3, offset=0)
    assert torch.equal(xs, xs2)                                  # counter-based RNG: same (seed, offset) -> same draw
    nll, sd_z, z = nf._loss(xs, y, iso=[100.0], cam=[2.0], return_z=True)
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1.0) < 5e-3 and abs(float(sd_z) - 1.0) < 2e-2


def test_tensor_core_kernel_equals_cuda_core_kernel_on_many_patches():
    """Width 32 has both kernels: more patches than one wave of the tensor-core kernel (148 CTAs x 4 groups), ragged last
    round, per-patch conditioning rows, explicit and in-kernel (Philox) noise."""
    from noise_flow_b200 import NoiseFlow
    hps, vs = _perturbed_model(32, arch="sdn5|unc|unc|gain4|unc")
    nf = NoiseFlow([32, 32, 4], False, copy.copy(hps), variables=vs, device="cuda:0", first_call="inverse")
    n = 1501
    g = torch.Generator(device="cuda").manual_seed(1)
    y = torch.rand((n, 32, 32, 4), device="cuda", generator=g)
    x = torch.randn((n, 32, 32, 4), device="cuda", generator=g) * 0.3
    cams = (np.arange(n) % 5).astype(np.float64)
    isos = np.asarray([100.0, 400.0, 800.0, 1600.0, 3200.0])[(np.arange(n) // 5) % 5]
    res = {}
    for tc in (True, False):
        nf.set_tensor_cores(tc)
        nll, sd_z, z = nf._loss(x, y, iso=isos, cam=cams, return_z=True)
        xs = nf.sample(y, 1.0, y, iso=isos, cam=cams, seed=3, offset=0)
        res[tc] = (nll, z, xs)
    assert float((res[True][0] - res[False][0]).abs().max()) / 4096 < 2e-5
    assert float((res[True][1] - res[False][1]).abs().max()) < 2e-4 * max(1.0, float(res[False][1].abs().max()))
    assert float((res[True][2] - res[False][2]).abs().max()) < 2e-4 * max(1.0, float(res[False][2].abs().max()))
    nf.set_tensor_cores(True)
    assert torch.equal(nf._loss(x, y, iso=isos, cam=cams)[0], res[True][0])       # fixed reduction order: bit-identical reruns


def test_wide_unsupported_widths_and_switches_raise():
    from noise_flow_b200 import NoiseFlow, make_hps
    with pytest.raises(RuntimeError):
        NoiseFlow([32, 32, 4], False, make_hps(arch="unc", width=24), device="cuda:0", first_call="inverse")
    nf = NoiseFlow([32, 32, 4], False, make_hps(arch="unc", width=64), device="cuda:0", first_call="inverse")
    with pytest.raises(RuntimeError):
        nf.set_tensor_cores(False)            # width 64 has no CUDA-core kernel


@pytest.mark.parametrize("width,is_training", [(8, True), (8, False), (16, True), (32, True), (32, False)])
def test_wide_gradients_match_oracle_autograd(width, is_training):
    """The train step at coupling-net widths 8 / 16 / 32 (the reference trains any `--width`, train_noise_flow.py:187-198):
    d(mean NLL) / d(every trainable variable) from the CTA-per-patch backward kernels (csrc/nf_train_wide.cu) against torch
    autograd through the fp64 oracle, both BatchNorm modes; `DeviceTrainer` hands out the wide trainer."""
    from noise_flow_b200 import NoiseFlow
    from noise_flow_b200.train import DeviceTrainer, WideTrainer
    from test_gpu_train import _check, _oracle_loss_and_grads
    hps, vs = _perturbed_model(width, arch="sdn5|unc|gain4|unc")
    nf = NoiseFlow([32, 32, 4], is_training, copy.copy(hps), variables={k: v.copy() for k, v in vs.items()}, device="cuda:0",
                   first_call="inverse")
    x, y = synth_batch(5, cam=2, iso=100, seed=95)
    x = (x * 20).astype(np.float32)
    tr = DeviceTrainer(nf, max_batch=8)
    assert isinstance(tr, WideTrainer)
    tr.loss_and_grad(x, y, iso=[100.0], cam=[2.0], is_training=is_training)
    loss, sd_z = tr.loss()
    loss_o, sd_o, grads_o, _ = _oracle_loss_and_grads(hps, vs, x, y, 100.0, 2.0, is_training)
    assert abs(loss - loss_o) / 4096 < 1e-4 and abs(sd_z - sd_o) < 1e-4
    worst = _check(tr.gradients(), grads_o, rel=5e-4)
    print("width %d, is_training %s: max relative gradient error %.2e" % (width, is_training, worst))


def test_wide_adam_steps_follow_the_oracle():
    """Two Adam steps at width 16 (batch-statistics BatchNorm, TF update rule, moving averages) against the same steps taken
    with the oracle's autograd gradients."""
    from noise_flow_b200 import NoiseFlow
    from noise_flow_b200.train import AdamOptimizer, DeviceTrainer
    from test_gpu_train import _oracle_loss_and_grads
    hps, vs = _perturbed_model(16, arch="sdn5|unc|gain4|unc")
    nf = NoiseFlow([32, 32, 4], True, copy.copy(hps), variables={k: v.copy() for k, v in vs.items()}, device="cuda:0",
                   first_call="inverse")
    tr = DeviceTrainer(nf, learning_rate=1e-3, max_batch=8)
    x, y = synth_batch(4, cam=2, iso=100, seed=97)
    x = (x * 20).astype(np.float32)
    ref_vars = {k: v.copy() for k, v in vs.items()}
    opt = AdamOptimizer(learning_rate=1e-3)
    for step in range(2):
        loss, _ = tr.step(x, y, iso=[100.0], cam=[2.0])
        loss_o, _, grads_o, orc = _oracle_loss_and_grads(hps, ref_vars, x, y, 100.0, 2.0, True)
        assert abs(loss - loss_o) / 4096 < 1e-4
        opt.apply_gradients(ref_vars, grads_o)
        for k, v in orc.store.vars.items():          # the oracle moved its BatchNorm statistics during the training-mode pass
            if k.endswith("/mean") or k.endswith("/var"):
                ref_vars[k] = v.detach().numpy().astype(np.float32)
        got = tr.variables()
        for k, want in ref_vars.items():
            d = np.abs(got[k].astype(np.float64) - want).max()
            if k.endswith("/l_1/b") or k.endswith("/l_2/b"):
                assert d <= (step + 1) * 1e-3 * 2.01 + 1e-7, (step, k, d)     # exact-zero gradients: +-lr random walk on both sides
            else:
                assert d < 2e-4 * max(1.0, np.abs(want).max()) + 2e-5, (step, k, d)


def test_unsupported_training_is_refused_loudly():
    from noise_flow_b200 import NoiseFlow
    from noise_flow_b200.train import DeviceTrainer, loss_and_grad
    hps, vs = _perturbed_model(64, arch="unc")
    nf = NoiseFlow([32, 32, 4], True, copy.copy(hps), variables=vs, device="cuda:0", first_call="inverse")
    x, y = synth_batch(2)
    with pytest.raises(NotImplementedError):
        DeviceTrainer(nf)                               # widths 4 / 8 / 16 / 32
    with pytest.raises(RuntimeError):
        loss_and_grad(nf, x, y, iso=[100.0], cam=[2.0])


def test_wide_model_through_saved_artefacts_and_wrapper(tmp_path):
    """Write hps.txt + a TF-V2 checkpoint of a width-16 model, reload it through ``NoiseFlowWrapper`` (the reference's
    public sampling API) and compare the fused (moving-statistics) sampler with the oracle on injected noise."""
    from noise_flow_b200 import NoiseFlow, hps_logger, save_checkpoint
    from noise_flow_b200.NoiseFlowWrapper import NoiseFlowWrapper
    hps, vs = _perturbed_model(16, arch="sdn5|unc|unc|gain4|unc|unc")
    d = tmp_path / "WideFlow"
    (d / "ckpt").mkdir(parents=True)
    nf = NoiseFlow([32, 32, 4], False, copy.copy(hps), variables=vs, device="cuda:0", first_call="inverse")
    hps_logger(str(d / "hps.txt"), hps, nf.get_layer_names(), nf.num_trainable_params())      # borealisflows/utils.py:110-119
    save_checkpoint(str(d / "ckpt" / "model.ckpt.best"), nf.variables)
    w = NoiseFlowWrapper(str(d), sampling_temperature=0.6, template_order="training", bn_mode="moving", seed=5)
    assert int(w.hps.width) == 16 and w.nf_model.get_layer_names() == nf.get_layer_names()
    y = np.random.RandomState(8).rand(7, 32, 32, 4).astype(np.float32)
    out = w.sample_noise_nf(y, 0.0, 0.0, 800, 2)
    assert out.shape == (7, 32, 32, 4) and out.dtype == np.float32 and np.isfinite(out).all()
    eps = np.random.RandomState(9).randn(7, 32, 32, 4).astype(np.float32)
    xs = w.nf_model.sample(y, 0.6, y, iso=[800.0], cam=[2.0], eps=eps).cpu().numpy()
    xo = make_oracle(hps, vs).sample(eps, 0.6, y, iso=[800.0], cam=[2.0]).numpy()
    assert np.abs(xs - xo).max() < 2e-4 * max(1.0, np.abs(xo).max())
    # the wrapper's default (batch statistics, is_training=True as the reference feeds) runs on wide nets too
    w2 = NoiseFlowWrapper(str(d), sampling_temperature=0.6)
    out2 = w2.sample_noise_nf(y, 0.0, 0.0, 800, 2)
    assert out2.shape == (7, 32, 32, 4) and np.isfinite(out2).all()
