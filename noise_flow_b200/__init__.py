"""B200-native Noise Flow density / sampling engine (drop-in for BorealisAI/noise_flow's hot path).

Like the reference package, the wrapper class lives in its own module:
``from noise_flow_b200.NoiseFlowWrapper import NoiseFlowWrapper`` (reference: ``borealisflows.NoiseFlowWrapper``).
"""
from .hps import Hps, hps_loader, hps_logger, make_hps  # noqa: F401
from .tf_checkpoint import load_checkpoint, save_checkpoint  # noqa: F401

__all__ = ["NoiseFlow", "Hps", "hps_loader", "hps_logger", "make_hps", "load_checkpoint",
           "save_checkpoint", "squeeze2d", "unsqueeze2d"]


def __getattr__(name):   # lazy: importing the package must not require torch / the CUDA library
    if name in ("NoiseFlow", "squeeze2d", "unsqueeze2d"):
        from . import noise_flow_model
        return getattr(noise_flow_model, name)
    raise AttributeError(name)
