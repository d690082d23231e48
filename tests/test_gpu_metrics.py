"""GPU tests of the evaluation metrics next to the path: baselines (float, tolerance) and histograms (integer work:
bit-exact against numpy)."""
import numpy as np
import pytest
import torch

from common import CAM_ISO_NLF, synth_batch

pytestmark = pytest.mark.gpu


def test_baselines_match_reference_formula():
    from noise_flow_b200.metrics import calc_baselines
    from oracle import noise_flow_oracle as O
    b1, b2 = CAM_ISO_NLF[(2, 800)]
    x, y = synth_batch(37, cam=2, iso=800, seed=3)
    g, s = calc_baselines(x, y, [b1], [b2], 1.7e-3)
    go, so = O.calc_baselines(x, y, np.float32(b1), np.float32(b2), np.float32(1.7e-3))
    assert np.abs(g.cpu().numpy() - go).max() / 4096 < 2e-6
    assert np.abs(s.cpu().numpy() - so).max() / 4096 < 2e-6
    # known answer: the sdn baseline equals the NLL of a camsdn-only flow (same closed form)
    assert np.abs(so - O.nll_sdn_closed_form(x, y, np.float32(b1), np.float32(b2))).max() < 1e-6


@pytest.mark.parametrize("n", [1, 4097, 300000])
def test_histogram_bit_exact_and_kl(n):
    from noise_flow_b200.metrics import default_bin_edges, get_histogram, kl_div_3_data, kl_div_forward
    from oracle import noise_flow_oracle as O
    rng = np.random.RandomState(n)
    edges = default_bin_edges()
    assert len(edges) == 67
    d = (rng.randn(n) * 0.04).astype(np.float32)
    d[: min(n, 8)] = np.array([-0.1, 0.1, -1000.0, 1000.0, 2000.0, -0.1 + 0.2 / 64, 0.0, np.nan], np.float32)[: min(n, 8)]
    h, centers = get_histogram(d, edges)
    ho = O.get_histogram(d, edges)
    assert np.array_equal(h, ho)                      # integer counts / n: bit-exact
    assert len(centers) == 66
    q = (rng.randn(n) * 0.05).astype(np.float32)
    hq, _ = get_histogram(torch.as_tensor(q, device="cuda:0"), edges)
    assert np.array_equal(hq, O.get_histogram(q, edges))
    if n > 1000:
        assert abs(kl_div_forward(h, hq) - O.kl_div_forward(ho, O.get_histogram(q, edges))) < 1e-12
        fwd, inv, sym = kl_div_3_data(d[8:], q, edges)
        assert fwd > 0 and inv > 0 and abs(sym - (fwd + inv) / 2) < 1e-15
