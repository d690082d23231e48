"""B200-native Noise Flow density / sampling engine (drop-in for BorealisAI/noise_flow's hot path)."""
from .hps import Hps, hps_loader, hps_logger, make_hps  # noqa: F401
from .tf_checkpoint import load_checkpoint, save_checkpoint  # noqa: F401

__all__ = ["NoiseFlow", "NoiseFlowWrapper", "Hps", "hps_loader", "hps_logger", "make_hps", "load_checkpoint",
           "save_checkpoint", "squeeze2d", "unsqueeze2d"]


def __getattr__(name):   # lazy: importing the package must not require torch / the CUDA library
    if name in ("NoiseFlow", "squeeze2d", "unsqueeze2d"):
        from . import noise_flow_model
        return getattr(noise_flow_model, name)
    if name == "NoiseFlowWrapper":
        from .NoiseFlowWrapper import NoiseFlowWrapper
        return NoiseFlowWrapper
    raise AttributeError(name)
