"""Evaluation metrics the reference computes next to the flow's NLL, on the device.

* :func:`calc_baselines` -- Gaussian and camera-NLF NLL baselines (reference ``sidd/PatchStatsCalculator.py:92-123``)
* :func:`get_histogram`, :func:`kl_div_forward`, :func:`kl_div_inverse`, :func:`kl_div_sym`, :func:`kl_div_3_data`
  -- marginal KL divergence between real and sampled noise (``sidd/sidd_utils.py:1202-1274``); the histogram
  (the only part that touches the data) runs on the GPU and is bit-identical to ``np.histogram``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _stream(dev) -> int:
    return int(torch.cuda.current_stream(dev).cuda_stream)


def _dev_f32(a, device=None):
    if not isinstance(a, torch.Tensor):
        a = torch.as_tensor(np.asarray(a))
    if not a.is_cuda:
        a = a.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    return a.to(torch.float32).contiguous()


def calc_baselines(x, y, nlf0, nlf1, var_gauss, device=None):
    """Per-patch ``(nll_gauss[N], nll_sdn[N])`` of ``calc_baselines``: ``0.5*(log 2pi + log v + x^2/v)`` summed over
    the patch with ``v = var_gauss`` (scalar, the reference's ``stats['sc_in_vr']``) and ``v = y*nlf0 + nlf1``."""
    lib = _lib.load()
    x, y = _dev_f32(x, device), _dev_f32(y, device)
    if x.shape != y.shape or x.dim() != 4 or tuple(x.shape[1:]) != (32, 32, 4):
        raise ValueError("x and y must both be [N, 32, 32, 4]")
    n = x.shape[0]
    g = torch.empty(n, device=x.device, dtype=torch.float32)
    s = torch.empty(n, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.nf_baseline_nll(x.data_ptr(), y.data_ptr(), float(np.asarray(nlf0).reshape(-1)[0]),
                                       float(np.asarray(nlf1).reshape(-1)[0]), float(var_gauss), n, g.data_ptr(),
                                       s.data_ptr(), _stream(x.device)), "nf_baseline_nll")
    return g, s


def default_bin_edges():
    """The 66-bin edges of ``kldiv_patch_set`` (sidd_utils.py:1044-1045)."""
    bw = 0.2 / 64
    return np.concatenate(([-1000.0], np.arange(-0.1, 0.1 + 1e-9, bw), [1000.0]), axis=0)


def get_histogram(data, bin_edges=None, left_edge=0.0, right_edge=1.0, n_bins=1000, device=None):
    """sidd_utils.py:1266-1274 -> ``(hist / n, bin_centers)``; ``data`` any shape, float32 on the device."""
    lib = _lib.load()
    data_range = right_edge - left_edge
    bin_width = data_range / n_bins
    if bin_edges is None:
        bin_edges = np.arange(left_edge, right_edge + bin_width, bin_width)
    bin_edges = np.asarray(bin_edges, dtype=np.float64)
    bin_centers = bin_edges[:-1] + (bin_width / 2.0)
    d = _dev_f32(data, device).reshape(-1)
    n = d.numel()
    edges = torch.as_tensor(bin_edges, device=d.device)
    counts = torch.zeros(len(bin_edges) - 1, device=d.device, dtype=torch.int64)
    with torch.cuda.device(d.device):
        _lib.check(lib.nf_histogram(d.data_ptr(), n, edges.data_ptr(), len(bin_edges) - 1, counts.data_ptr(),
                                    _stream(d.device)), "nf_histogram")
    return counts.cpu().numpy() / n, bin_centers


def kl_div_forward(p, q):
    """sidd_utils.py:1202-1209."""
    p, q = np.asarray(p, dtype=np.float64), np.asarray(q, dtype=np.float64)
    idx = ~(np.isnan(p) | np.isinf(p) | np.isnan(q) | np.isinf(q))
    p, q = p[idx], q[idx]
    idx = (p > 0) & (q > 0)
    p, q = p[idx], q[idx]
    return np.sum(p * np.log(p / q))


def kl_div_inverse(p, q):
    """sidd_utils.py:1212-1219."""
    p, q = np.asarray(p, dtype=np.float64), np.asarray(q, dtype=np.float64)
    idx = ~(np.isnan(p) | np.isinf(p) | np.isnan(q) | np.isinf(q))
    p, q = p[idx], q[idx]
    idx = (p > 0) & (q > 0)
    p, q = p[idx], q[idx]
    return np.sum(q * np.log(q / p))


def kl_div_sym(p, q):
    """sidd_utils.py:1222-1223."""
    return (kl_div_forward(p, q) + kl_div_inverse(p, q)) / 2.0


def kl_div_3_data(p_data, q_data, bin_edges=None, left_edge=0.0, right_edge=1.0, n_bins=1000, device=None):
    """sidd_utils.py:1247-1263 -> ``(kl_fwd, kl_inv, kl_sym)`` between two sets of data points."""
    if bin_edges is None:
        data_range = right_edge - left_edge
        bin_width = data_range / n_bins
        bin_edges = np.arange(left_edge, right_edge + bin_width, bin_width)
    p, _ = get_histogram(p_data, bin_edges, left_edge, right_edge, n_bins, device)
    q, _ = get_histogram(q_data, bin_edges, left_edge, right_edge, n_bins, device)
    idx = (p > 0) & (q > 0)
    p, q = p[idx], q[idx]
    logp, logq = np.log(p), np.log(q)
    kl_fwd = np.sum(p * (logp - logq))
    kl_inv = np.sum(q * (logq - logp))
    return kl_fwd, kl_inv, (kl_fwd + kl_inv) / 2.0
