"""``hps.txt`` reader / writer -- the reference's flat hyper-parameter file.

Mirrors ``NoiseFlowWrapper.hps_loader`` (reference ``borealisflows/NoiseFlowWrapper.py:96-138``) and
``hps_logger`` (``borealisflows/utils.py:110-119``): a CSV of ``key,value`` pairs; the first lines of
the shipped file are the layer names and the parameter count (single-field rows, skipped on load).
"""
from __future__ import annotations

import csv
from types import SimpleNamespace

import numpy as np


class Hps(SimpleNamespace):
    """Attribute bag, like the reference's ad-hoc ``class Hps: pass``."""

    def get(self, key, default=None):
        return getattr(self, key, default)


def default_param_inits(arch: str):
    """``NoiseFlowWrapper.py:125-137`` / ``train_noise_flow.py:205-215``: ``param_inits`` is never parsed
    from ``hps.txt``; it is always recomputed and only used as variable initialisers."""
    npcam = 1 if ("sdn6" in arch and "sdn5" not in arch) else 3
    c_i = 1.0
    beta1_i = -5.0 / c_i
    beta2_i = 0.0
    gain_params_i = np.full([5], -5.0 / c_i)
    cam_params_i = np.full([npcam, 5], 1.0)
    return (c_i, beta1_i, beta2_i, gain_params_i, cam_params_i)


def hps_loader(path: str) -> Hps:
    """Same coercion rules as the reference: int, else float, else 'True'/'False', else the raw string."""
    hps = Hps()
    with open(path, "r", newline="") as f:
        for pair in csv.reader(f):
            if len(pair) < 2:
                continue
            val = pair[1]
            try:
                val = int(val)
            except ValueError:
                try:
                    val = float(val)
                except ValueError:
                    if val == "True":
                        val = True
                    elif val == "False":
                        val = False
            setattr(hps, pair[0], val)
    if not hasattr(hps, "arch") or hps.arch in ("", None, "None"):
        hps.arch = None                  # legacy revnet2d model (noise_flow_model.py:63-68): depth / sidd_cond / append_* decide
    hps.param_inits = default_param_inits(hps.arch or "")
    return hps


def hps_logger(path: str, hps, layer_names, num_params) -> None:
    """``borealisflows/utils.py:110-119``: layer names, parameter count, then ``key,value`` rows."""
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        for n in layer_names:
            w.writerow([n])
        w.writerow([num_params])
        for k, v in vars(hps).items():
            w.writerow([k, v])


def make_hps(**kw) -> Hps:
    """Hot-path keys with the shipped model's values (reference ``models/NoiseFlow/hps.txt``)."""
    d = dict(arch="sdn5|unc|unc|unc|unc|gain4|unc|unc|unc|unc", width=4, flow_permutation=1, decomp="LU",
             squeeze_factor=1, squeeze_type="chessboard", n_levels=1, depth=-1, gain_init=-5.0,
             sidd_cond="mix", x_shape=[None, 32, 32, 4])
    d.update(kw)
    if not d.get("arch"):                # legacy revnet2d model: the reference's ArgParser defaults for the keys it reads
        for k, v in (("append_sdn2", False), ("append_sdn_first", False), ("append_cY", False), ("append_sdn", False)):
            d.setdefault(k, v)
        d["arch"] = None
    h = Hps(**d)
    if not hasattr(h, "param_inits"):
        h.param_inits = default_param_inits(h.arch or "")
    return h
