"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck / synccheck).
Run on the GPU box:  compute-sanitizer --tool racecheck python tools/gpu/sanitize_smoke.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import synth_batch  # noqa: E402
from noise_flow_b200 import NoiseFlow, hps_loader, load_checkpoint, make_hps  # noqa: E402

g = os.path.join(ROOT, "tests", "golden", "NoiseFlow")
hps = hps_loader(os.path.join(g, "hps.txt"))
ck = load_checkpoint(os.path.join(g, "ckpt", "model.ckpt.best"))
n = int(os.environ.get("NF_SANITIZE_N", "37"))          # not a multiple of the warps per CTA: exercises the inactive-warp tail
x, y = synth_batch(n, seed=1)
nf = NoiseFlow([32, 32, 4], False, hps, variables=ck, device="cuda:0", first_call="inverse")
nll, sdz, z = nf._loss(x, y, iso=[100.0], cam=[2.0], return_z=True)
xs = nf.sample(y, 0.6, y, iso=[100.0], cam=[2.0], seed=3, offset=0)
xr = nf.forward(z, None, yy=y, iso=[100.0], cam=[2.0])
nll_b, _ = nf._loss(x, y, iso=[100.0], cam=[2.0], is_training=True)            # batch-statistics probes
zz, ld = nf.run_layers(0, 3, "inverse", x, yy=y, iso=[100.0], cam=[2.0])
nf2 = NoiseFlow([32, 32, 4], False, make_hps(arch="sdn5|gain4"), device="cuda:0", first_call="inverse")
nll_s, _ = nf2._loss(x, y, iso=[100.0], cam=[2.0])                              # streaming kernel
if os.environ.get("NF_SANITIZE_TRAIN", "1") == "1":
    from noise_flow_b200.train import AdamOptimizer, train_step
    nf3 = NoiseFlow([32, 32, 4], True, hps, variables=dict(ck), device="cuda:0", first_call="inverse")
    loss, sd = train_step(nf3, AdamOptimizer(1e-4), x, y, iso=[100.0], cam=[2.0])
    print("train loss", float(loss))
if os.environ.get("NF_SANITIZE_TC", "0") == "1":
    nf.set_tensor_cores(True)
    nll_t, _ = nf._loss(x, y, iso=[100.0], cam=[2.0])
    print("tc max diff", float((nll_t - nll).abs().max()))
torch.cuda.synchronize()
print("ok", float(nll.mean()) / 4096, float(np.abs(xr.cpu().numpy() - x).max()), float(nll_s.mean()) / 4096,
      float(nll_b.mean()) / 4096, tuple(xs.shape), tuple(zz.shape))
