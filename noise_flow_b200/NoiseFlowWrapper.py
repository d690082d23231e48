"""``NoiseFlowWrapper`` -- drop-in for ``borealisflows/NoiseFlowWrapper.py:19-138``.

``NoiseFlowWrapper(path, sampling_temperature=0.6).sample_noise_nf(batch_x, b1, b2, iso, cam)`` returns a
float32 numpy array ``[N, 32, 32, 4]`` of synthetic noise for the clean patches ``batch_x``, exactly like
the reference (which is called this way from ``sample_noise_flow.py:71`` and from 32 sampler threads in
``train_dncnn_noiseflow.py:142-160``).  ``path`` holds ``hps.txt`` and ``ckpt/model.ckpt.best.*``.

Two reference behaviours are reproduced deliberately and can be switched off:

* ``template_order='reference'`` -- the wrapper builds only the sampling graph, so ``tf.make_template``
  names the coupling-net scopes in latent->data order and ``Saver.restore`` therefore loads the checkpoint's
  ``real_nvp_conv_template`` (trained as the net of ``unc_1``) into ``unc_9``, ``..._1`` into ``unc_8`` and so
  on.  ``template_order='training'`` assigns them the way the training graph did.
* ``bn_mode='batch'`` (default) is what ``sample_noise_nf`` feeds (``is_training: True``,
  NoiseFlowWrapper.py:85-86): BatchNorm uses the statistics of the current batch of patches -- so a sample
  depends on which other patches share its batch -- and every call also moves the stored statistics
  (layers.py:394-395).  ``bn_mode='moving'`` uses the checkpoint's moving statistics: patches are then
  independent and the whole chain runs as ONE fused kernel (about 3x faster, and what the benchmark times).
"""
from __future__ import annotations

import logging
import os

import numpy as np
import torch

from .hps import hps_loader
from .noise_flow_model import NoiseFlow
from .tf_checkpoint import load_checkpoint


class NoiseFlowWrapper:
    def __init__(self, path, sampling_temperature=0.6, template_order="reference", bn_mode="batch",
                 device=None, seed=0):
        self.logger = logging.getLogger(__name__)
        self.nf_path = path
        self.nf_model = None
        self.x_shape = None
        self.is_cond = True                                                        # NoiseFlowWrapper.py:37
        self.temp = sampling_temperature
        if template_order not in ("reference", "training"):
            raise ValueError("template_order must be 'reference' or 'training'")
        if bn_mode not in ("moving", "batch"):
            raise ValueError("bn_mode must be 'moving' or 'batch'")
        self.template_order = template_order
        self.bn_mode = bn_mode
        self._device = device
        self._seed = seed

        self.hps = self.hps_loader(os.path.join(self.nf_path, "hps.txt"))
        self.ckpt_dir = os.path.join(self.nf_path, "ckpt")
        self.model_checkpoint_path = os.path.join(self.ckpt_dir, "model.ckpt.best")
        self.load_noise_flow_model()

    def load_noise_flow_model(self):
        self.x_shape = [None, 32, 32, 4]                                           # NoiseFlowWrapper.py:47
        if not hasattr(self.hps, "x_shape"):
            setattr(self.hps, "x_shape", self.x_shape)
        self.logger.info("Building Noise Flow")
        variables = load_checkpoint(self.model_checkpoint_path)                    # Saver.restore (:77)
        first_call = "forward" if self.template_order == "reference" else "inverse"
        self.nf_model = NoiseFlow(self.x_shape[1:], self.bn_mode == "batch", self.hps, variables=variables,
                                  first_call=first_call, device=self._device, seed=self._seed)
        missing = self.nf_model.spec.store.created
        if missing:   # Saver.restore would raise NotFoundError
            raise RuntimeError("checkpoint %s lacks variables %s" % (self.model_checkpoint_path, missing[:4]))

    def sample_noise_nf(self, batch_x, b1, b2, iso, cam):
        """NoiseFlowWrapper.py:81-87.  ``b1``/``b2`` are fed as nlf0/nlf1 (ignored by the shipped arch)."""
        x = self.sample_sidd_tf(batch_x, b1, b2, iso, cam)
        return self.nf_model.to_numpy(x)

    def sample_sidd_tf(self, batch_x, b1=0.0, b2=0.0, iso=100, cam=0):
        """NoiseFlowWrapper.py:89-94: returns the device tensor (the reference returns the TF op)."""
        y = batch_x if isinstance(batch_x, torch.Tensor) else np.asarray(batch_x, dtype=np.float32)
        if self.is_cond:
            return self.nf_model.sample(y, self.temp, y, [b1], [b2], [iso], [cam])
        return self.nf_model.sample(y, self.temp)

    def hps_loader(self, path):
        return hps_loader(path)
