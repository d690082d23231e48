"""Shared helpers for the test-suite: SIDD-shaped synthetic data and oracle construction."""
import copy
import os

import numpy as np
import torch

CAM_ISO_NLF = {  # reference cam_iso_nlf.txt (camera NLF beta1, beta2); cam index per sidd_utils.py:262
    (0, 100): (0.000702, 0.000003), (0, 400): (0.000735, 0.000003), (0, 800): (0.000723, 0.000003),
    (0, 1600): (0.000687, 0.000002), (1, 100): (0.000228, 0.000002), (1, 800): (0.001764, 0.000012),
    (2, 100): (0.000479, 0.000002), (2, 400): (0.001774, 0.000002), (2, 800): (0.003696, 0.000002),
    (2, 1600): (0.008211, 0.000002), (2, 3200): (0.019930, 0.000002), (3, 100): (0.000473, 0.000003),
    (3, 800): (0.003476, 0.000052), (4, 100): (0.000739, 0.000002), (4, 800): (0.003356, 0.000063),
}


def synth_batch(n, cam=2, iso=100, seed=0, dtype=np.float32):
    """y ~ U[0,1), x = N(0,1)*sqrt(beta1*y+beta2) with the camera NLF of (cam, iso) (sidd_utils.py:1017-1020)."""
    rng = np.random.RandomState(seed)
    y = rng.rand(n, 32, 32, 4)
    b1, b2 = CAM_ISO_NLF[(cam, iso)]
    x = rng.randn(n, 32, 32, 4) * np.sqrt(b1 * y + b2)
    return x.astype(dtype), y.astype(dtype)


def make_oracle(hps, variables, dtype=torch.float64, first_call="inverse", seed=0):
    """OracleNoiseFlow with template scopes named in ``first_call`` order (tf.make_template semantics)."""
    from oracle.noise_flow_oracle import AffineCoupling, OracleNoiseFlow
    nf = OracleNoiseFlow([32, 32, 4], copy.copy(hps), variables, dtype=dtype, seed=seed)
    cps = [b for b in nf.model[0] if isinstance(b, AffineCoupling)]
    for b in (cps if first_call == "inverse" else reversed(cps)):
        b._fn.ensure_scope("model")
    return nf
