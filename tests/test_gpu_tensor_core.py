"""GPU parity tests of the tensor-core (tcgen05) formulation of the chain against the CPU oracle and against the
fp32 CUDA-core kernel.  Contract tolerance: |NLL - oracle| < 1e-4 nats/dim; z / x within 1e-4 relative."""
import copy

import numpy as np
import pytest
import torch

from common import make_oracle, synth_batch

pytestmark = pytest.mark.gpu


def _nf(hps, ck, tc):
    from noise_flow_b200 import NoiseFlow
    nf = NoiseFlow([32, 32, 4], False, copy.copy(hps), variables=ck, device="cuda:0", first_call="inverse")
    nf.set_tensor_cores(tc)
    return nf


@pytest.mark.parametrize("cam,iso,n", [(2, 100, 40), (0, 1600, 7), (2, 3200, 1)])
def test_tc_log_prob_matches_oracle(shipped, cam, iso, n):
    hps, ck = shipped
    x, y = synth_batch(n, cam=cam, iso=iso, seed=70 + cam)
    nf = _nf(hps, ck, True)
    nll, sd_z, z = nf._loss(x, y, iso=[float(iso)], cam=[float(cam)], return_z=True)
    orc = make_oracle(hps, ck)
    nll_o, sd_o = orc._loss(x, y, iso=[float(iso)], cam=[float(cam)])
    err = np.abs(nll.cpu().numpy() - nll_o.numpy()).max() / 4096
    zerr = np.abs(z.cpu().numpy() - orc.last_z.numpy()).max()
    print("tensor-core path: max |dNLL| = %.3e nats/dim, max |dz| = %.3e" % (err, zerr))
    assert err < 1e-4, err
    assert zerr < 1e-4 * (1 + np.abs(orc.last_z.numpy()).max())
    assert abs(float(sd_z) - float(sd_o)) < 1e-4
    # and against the fp32 CUDA-core kernel on the same inputs
    nll32, _ = _nf(hps, ck, False)._loss(x, y, iso=[float(iso)], cam=[float(cam)])
    assert np.abs(nll.cpu().numpy() - nll32.cpu().numpy()).max() / 4096 < 1e-4


def test_tc_sample_and_roundtrip(shipped):
    hps, ck = shipped
    x, y = synth_batch(20, seed=77)
    eps = np.random.RandomState(78).randn(20, 32, 32, 4).astype(np.float32)
    nf = _nf(hps, ck, True)
    xs = nf.sample(y, 0.6, y, iso=[800.0], cam=[2.0], eps=eps).cpu().numpy()
    xo = make_oracle(hps, ck).sample(eps, 0.6, y, iso=[800.0], cam=[2.0]).numpy()
    assert np.abs(xs - xo).max() < 1e-4 * (1 + 100 * np.abs(xo).max())
    z, _ = nf.inverse(x, None, yy=y, iso=[100.0], cam=[2.0])
    xr = nf.forward(z, None, yy=y, iso=[100.0], cam=[2.0]).cpu().numpy()
    assert np.abs(xr - x).max() < 1e-4 * (1 + 100 * np.abs(x).max())


def test_tc_large_batch_matches_fp32_kernel(shipped):
    hps, ck = shipped
    n = 5000   # > 148 CTAs x 3 groups: persistent loop, ragged tail
    g = torch.Generator(device="cuda:0").manual_seed(9)
    y = torch.rand((n, 32, 32, 4), device="cuda:0", generator=g)
    x = torch.randn((n, 32, 32, 4), device="cuda:0", generator=g) * torch.sqrt(0.000479 * y + 0.000002)
    a, _ = _nf(hps, ck, True)._loss(x, y, iso=[100.0], cam=[2.0])
    b, _ = _nf(hps, ck, False)._loss(x, y, iso=[100.0], cam=[2.0])
    assert float((a - b).abs().max()) / 4096 < 1e-4
