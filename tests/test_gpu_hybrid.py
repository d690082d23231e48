"""GPU parity tests of the hybrid chain kernel (csrc/nf_hybrid.cu: conv-3 of every coupling net on tcgen05 with fp16
hi/lo-split operands, the rest fp32) against the CPU oracle and against the all-fp32 CUDA-core kernel.
Contract tolerance: |NLL - oracle| < 1e-4 nats/dim; the tests hold it two orders tighter."""
import copy

import numpy as np
import pytest
import torch

from common import make_oracle, synth_batch

pytestmark = pytest.mark.gpu


def _nf(hps, ck, mode):
    from noise_flow_b200 import NoiseFlow
    nf = NoiseFlow([32, 32, 4], False, copy.copy(hps), variables=ck, device="cuda:0", first_call="inverse")
    nf.set_tensor_cores(mode)
    return nf


@pytest.mark.parametrize("cam,iso,n", [(2, 100, 40), (0, 1600, 7), (2, 3200, 1), (1, 800, 13)])
def test_hybrid_log_prob_matches_oracle(shipped, cam, iso, n):
    hps, ck = shipped
    x, y = synth_batch(n, cam=cam, iso=iso, seed=170 + cam)
    nf = _nf(hps, ck, "hybrid")
    nll, sd_z, z = nf._loss(x, y, iso=[float(iso)], cam=[float(cam)], return_z=True)
    orc = make_oracle(hps, ck)
    nll_o, sd_o = orc._loss(x, y, iso=[float(iso)], cam=[float(cam)])
    err = np.abs(nll.cpu().numpy() - nll_o.numpy()).max() / 4096
    zo = orc.last_z.numpy()
    zerr = np.abs(z.cpu().numpy() - zo).max()
    # the all-fp32 kernel on the same inputs: the hybrid path must not be visibly further from the oracle
    nll32, _, z32 = _nf(hps, ck, False)._loss(x, y, iso=[float(iso)], cam=[float(cam)], return_z=True)
    zerr32 = np.abs(z32.cpu().numpy() - zo).max()
    print("hybrid: max |dNLL| = %.3e nats/dim, max |dz| = %.3e (all-fp32 kernel: %.3e)" % (err, zerr, zerr32))
    assert err < 1e-6, err
    assert zerr < 2e-5 * (1 + np.abs(zo).max()), (zerr, zerr32)
    assert abs(float(sd_z) - float(sd_o)) < 1e-5
    assert np.abs(nll.cpu().numpy() - nll32.cpu().numpy()).max() / 4096 < 1e-6


def test_hybrid_sample_and_roundtrip(shipped):
    hps, ck = shipped
    x, y = synth_batch(20, seed=177)
    eps = np.random.RandomState(178).randn(20, 32, 32, 4).astype(np.float32)
    nf = _nf(hps, ck, "hybrid")
    xs = nf.sample(y, 0.6, y, iso=[800.0], cam=[2.0], eps=eps).cpu().numpy()
    xo = make_oracle(hps, ck).sample(eps, 0.6, y, iso=[800.0], cam=[2.0]).numpy()
    assert np.abs(xs - xo).max() < 2e-5 * (1 + np.abs(xo).max())
    z, _ = nf.inverse(x, None, yy=y, iso=[100.0], cam=[2.0])
    xr = nf.forward(z, None, yy=y, iso=[100.0], cam=[2.0]).cpu().numpy()
    assert np.abs(xr - x).max() < 2e-5 * (1 + np.abs(x).max())


def test_hybrid_philox_sampler_matches_fp32_kernel(shipped):
    """In-kernel Philox noise: same counter layout as the all-fp32 kernel, so the two samplers draw the same eps."""
    hps, ck = shipped
    _, y = synth_batch(50, seed=179)
    a = _nf(hps, ck, "hybrid").sample(y, 1.0, y, iso=[400.0], cam=[1.0], seed=1234).cpu().numpy()
    b = _nf(hps, ck, False).sample(y, 1.0, y, iso=[400.0], cam=[1.0], seed=1234).cpu().numpy()
    assert np.abs(a - b).max() < 2e-5 * (1 + np.abs(b).max())


@pytest.mark.parametrize("n", [1, 11, 12, 13, 148 * 12 + 5, 5000])
def test_hybrid_batch_shapes_match_fp32_kernel(shipped, n):
    """Ragged tails (n not a multiple of the 12 resident patches), more patches than one round, per-patch rows."""
    hps, ck = shipped
    g = torch.Generator(device="cuda:0").manual_seed(9 + n)
    y = torch.rand((n, 32, 32, 4), device="cuda:0", generator=g)
    x = torch.randn((n, 32, 32, 4), device="cuda:0", generator=g) * torch.sqrt(0.000479 * y + 0.000002)
    a, sa = _nf(hps, ck, "hybrid")._loss(x, y, iso=[100.0], cam=[2.0])
    b, sb = _nf(hps, ck, False)._loss(x, y, iso=[100.0], cam=[2.0])
    assert a.shape == b.shape
    assert float((a - b).abs().max()) / 4096 < 1e-6
    assert abs(float(sa) - float(sb)) < 1e-5


def test_hybrid_is_deterministic(shipped):
    hps, ck = shipped
    x, y = synth_batch(100, seed=181)
    nf = _nf(hps, ck, "hybrid")
    a, _ = nf._loss(x, y, iso=[100.0], cam=[2.0])
    b, _ = nf._loss(x[:37], y[:37], iso=[100.0], cam=[2.0])
    c, _ = nf._loss(x, y, iso=[100.0], cam=[2.0])
    assert torch.equal(a, c)
    assert torch.equal(a[:37], b)   # a patch's result does not depend on the batch around it


def test_hybrid_stress_model_accuracy():
    """A model with O(1) random coupling-net weights (the stress case of test_gpu_parity.py): activations of a few hundred
    and log-scales up to 0.8 per coupling amplify conv-3's 22-bit operands (fp16 hi + lo) through the chain.  Stated bound:
    5e-5 relative to the largest value, for z and for a sampled patch; the all-fp32 kernel is printed next to it."""
    from noise_flow_b200 import NoiseFlow, make_hps
    hps = make_hps(arch="sdn|unc|gain|unc", flow_permutation=0)
    nf0 = NoiseFlow([32, 32, 4], False, copy.copy(hps), device="cuda:0", seed=4, first_call="inverse")
    rng = np.random.RandomState(5)
    vs = {k: v.copy() for k, v in nf0.variables.items()}
    for k in vs:
        if k.endswith("/l_1/W") or k.endswith("/l_2/W"):
            vs[k] = (rng.randn(*vs[k].shape) * 0.5).astype(np.float32)
        elif k.endswith("/l_last/W"):
            vs[k] = (rng.randn(*vs[k].shape) * 0.1).astype(np.float32)
        elif k.endswith("/b") or k.endswith("/logs") or k.endswith("/mean"):
            vs[k] = (rng.randn(*vs[k].shape) * 0.2).astype(np.float32)
        elif k.endswith("/var"):
            vs[k] = (rng.rand(*vs[k].shape) + 0.05).astype(np.float32)
        elif "rescaling_scale" in k:
            vs[k] = np.float32(0.3 + 0.5 * rng.rand())
    orc = make_oracle(hps, vs)
    x, y = synth_batch(6, cam=2, iso=800, seed=19)
    x = (x * 5).astype(np.float32)
    kw = dict(nlf0=[0.003], nlf1=[0.00002], iso=[800.0], cam=[3.0])
    eps = rng.randn(6, 32, 32, 4).astype(np.float32)
    nll_o, _ = orc._loss(x, y, **kw)
    zo = orc.last_z.numpy()
    xo = orc.sample(eps, 0.8, y, **kw).numpy()
    err = {}
    for mode in ("hybrid", False):
        nf = NoiseFlow([32, 32, 4], False, copy.copy(hps), variables=vs, device="cuda:0", first_call="inverse")
        nf.set_tensor_cores(mode)
        nll, _, z = nf._loss(x, y, return_z=True, **kw)
        xs = nf.sample(y, 0.8, y, eps=eps, **kw).cpu().numpy()
        err[mode] = (np.abs(nll.cpu().numpy() - nll_o.numpy()).max() / 4096, np.abs(z.cpu().numpy() - zo).max() / (1 + np.abs(zo).max()),
                     np.abs(xs - xo).max() / (1 + np.abs(xo).max()))
    print("stress model, (|dNLL| nats/dim, rel |dz|, rel |dx|): hybrid %.2e %.2e %.2e   all-fp32 %.2e %.2e %.2e" % (err["hybrid"] + err[False]))
    assert err["hybrid"][0] < 2e-5 and err["hybrid"][1] < 5e-5 and err["hybrid"][2] < 5e-5


def test_auto_mode_routes_sampling_to_the_hybrid_kernel(shipped):
    hps, ck = shipped
    x, y = synth_batch(30, seed=183)
    eps = np.random.RandomState(184).randn(30, 32, 32, 4).astype(np.float32)
    auto, hyb, f32 = _nf(hps, ck, "auto"), _nf(hps, ck, "hybrid"), _nf(hps, ck, "default")
    kw = dict(iso=[100.0], cam=[2.0])
    assert torch.equal(auto.sample(y, 0.6, y, eps=eps, **kw), hyb.sample(y, 0.6, y, eps=eps, **kw))
    assert torch.equal(auto._loss(x, y, **kw)[0], f32._loss(x, y, **kw)[0])
