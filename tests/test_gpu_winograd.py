"""GPU parity tests of the Winograd chain kernel (csrc/nf_wino.cu: all-fp32, the two 3x3 convolutions of every coupling net
in a vertical F(2,3) form, two rows per step) against the CPU oracle and the direct-form kernel, at the direct kernel's tolerances."""
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
def test_winograd_log_prob_matches_oracle(shipped, cam, iso, n):
    hps, ck = shipped
    x, y = synth_batch(n, cam=cam, iso=iso, seed=270 + cam)
    nll, sd_z, z = _nf(hps, ck, "winograd")._loss(x, y, iso=[float(iso)], cam=[float(cam)], return_z=True)
    orc = make_oracle(hps, ck)
    nll_o, sd_o = orc._loss(x, y, iso=[float(iso)], cam=[float(cam)])
    zo = orc.last_z.numpy()
    err, zerr = np.abs(nll.cpu().numpy() - nll_o.numpy()).max() / 4096, np.abs(z.cpu().numpy() - zo).max()
    nll32, _, z32 = _nf(hps, ck, False)._loss(x, y, iso=[float(iso)], cam=[float(cam)], return_z=True)
    print("winograd: max |dNLL| = %.3e nats/dim, max |dz| = %.3e (direct form: %.3e)" % (err, zerr, np.abs(z32.cpu().numpy() - zo).max()))
    assert err < 3e-6, err
    assert zerr < 2e-5 * (1 + np.abs(zo).max())
    assert abs(float(sd_z) - float(sd_o)) < 1e-5


def test_winograd_sample_roundtrip_and_philox(shipped):
    hps, ck = shipped
    x, y = synth_batch(20, seed=277)
    eps = np.random.RandomState(278).randn(20, 32, 32, 4).astype(np.float32)
    nf = _nf(hps, ck, "winograd")
    xs = nf.sample(y, 0.6, y, iso=[800.0], cam=[2.0], eps=eps).cpu().numpy()
    xo = make_oracle(hps, ck).sample(eps, 0.6, y, iso=[800.0], cam=[2.0]).numpy()
    assert np.abs(xs - xo).max() < 2e-5 * (1 + np.abs(xo).max())
    z, _ = nf.inverse(x, None, yy=y, iso=[100.0], cam=[2.0])
    xr = nf.forward(z, None, yy=y, iso=[100.0], cam=[2.0]).cpu().numpy()
    assert np.abs(xr - x).max() < 2e-5 * (1 + np.abs(x).max())
    a = nf.sample(y, 1.0, y, iso=[400.0], cam=[0.0], seed=1234, offset=0).cpu().numpy()    # same Philox counters in both kernels
    b = _nf(hps, ck, False).sample(y, 1.0, y, iso=[400.0], cam=[0.0], seed=1234, offset=0).cpu().numpy()
    assert np.abs(a - b).max() < 2e-5 * (1 + np.abs(b).max())


@pytest.mark.parametrize("n", [1, 15, 16, 17, 148 * 16 + 5, 5000])
def test_winograd_batch_shapes_match_direct_kernel(shipped, n):
    hps, ck = shipped
    g = torch.Generator(device="cuda:0").manual_seed(19 + n)
    y = torch.rand((n, 32, 32, 4), device="cuda:0", generator=g)
    x = torch.randn((n, 32, 32, 4), device="cuda:0", generator=g) * torch.sqrt(0.000479 * y + 0.000002)
    a, sa = _nf(hps, ck, "winograd")._loss(x, y, iso=[100.0], cam=[2.0])
    b, sb = _nf(hps, ck, False)._loss(x, y, iso=[100.0], cam=[2.0])
    assert float((a - b).abs().max()) / 4096 < 3e-6
    assert abs(float(sa) - float(sb)) < 1e-5


def test_default_is_the_winograd_kernel_in_both_directions(shipped):
    hps, ck = shipped
    x, y = synth_batch(30, seed=283)
    eps = np.random.RandomState(284).randn(30, 32, 32, 4).astype(np.float32)
    dflt, wino, direct = _nf(hps, ck, "default"), _nf(hps, ck, "winograd"), _nf(hps, ck, "direct")
    kw = dict(iso=[100.0], cam=[2.0])
    assert torch.equal(dflt._loss(x, y, **kw)[0], wino._loss(x, y, **kw)[0])
    assert torch.equal(dflt.sample(y, 0.6, y, eps=eps, **kw), wino.sample(y, 0.6, y, eps=eps, **kw))
    assert not torch.equal(dflt._loss(x, y, **kw)[0], direct._loss(x, y, **kw)[0])      # (different rounding: really another kernel)
    # batch-statistics probes stay on the direct-form kernel whatever the mode
    a, _ = wino._loss(x, y, is_training=True, **kw)
    b, _ = direct._loss(x, y, is_training=True, **kw)
    assert float((a - b).abs().max()) / 4096 < 3e-6
