"""Diagnostic for the tensor-core wide-net kernel (csrc/nf_wide_tc.cu): errors against the CPU oracle (fp64) at widths
32 / 64 / 128, both directions, and against the CUDA-core kernel on more patches than one wave (width 32)."""
import copy, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from common import make_oracle, synth_batch
from test_gpu_wide import _perturbed_model
from noise_flow_b200 import NoiseFlow

widths = [int(w) for w in (sys.argv[1:] or ["32", "64", "128"])]
for width in widths:
    hps, vs = _perturbed_model(width)
    nf = NoiseFlow([32, 32, 4], False, copy.copy(hps), variables=vs, device="cuda:0", first_call="inverse")
    orc = make_oracle(hps, vs)
    n = 5
    x, y = synth_batch(n, cam=2, iso=100, seed=31)
    x = (x * 20).astype(np.float32)
    t0 = time.time()
    nll, sd_z, z = nf._loss(x, y, iso=[100.0], cam=[2.0], return_z=True)
    torch.cuda.synchronize()
    nll_o, sd_o = orc._loss(x, y, iso=[100.0], cam=[2.0])
    z_o, _ = orc.inverse(x, torch.zeros(n, dtype=torch.float64), y, iso=[100.0], cam=[2.0])
    print("width %3d  nll err/dim %.3e  (nll/dim %s)  sd_z err %.2e  z err %.3e (max |z| %.2f)" % (
        width, np.abs(nll.cpu().numpy() - nll_o.numpy()).max() / 4096, (nll_o.numpy() / 4096)[:2], abs(float(sd_z) - float(sd_o)),
        np.abs(z.cpu().numpy() - z_o.numpy()).max(), float(z_o.abs().max())), flush=True)
    eps = np.random.RandomState(5).randn(n, 32, 32, 4).astype(np.float32)
    xs = nf.sample(y, 0.6, y, iso=[100.0], cam=[2.0], eps=eps).cpu().numpy()
    xo = orc.sample(eps, 0.6, y, iso=[100.0], cam=[2.0]).numpy()
    back = nf.forward(z, None, y, iso=[100.0], cam=[2.0]).cpu().numpy()
    print("           sample err %.3e (max |x| %.3f)   round trip %.3e" % (np.abs(xs - xo).max(), np.abs(xo).max(), np.abs(back - x).max()), flush=True)
    if width == 32:
        m = 1500
        g = torch.Generator(device="cuda").manual_seed(1)
        yb = torch.rand((m, 32, 32, 4), device="cuda", generator=g)
        xb = torch.randn((m, 32, 32, 4), device="cuda", generator=g) * 0.3
        nll_tc, _, z_tc = nf._loss(xb, yb, iso=[100.0], cam=[2.0], return_z=True)
        nf.set_tensor_cores(False)
        nll_cc, _, z_cc = nf._loss(xb, yb, iso=[100.0], cam=[2.0], return_z=True)
        nf.set_tensor_cores(True)
        print("           %d patches, tensor-core vs CUDA-core kernel: nll diff/dim %.3e  z diff %.3e" % (
            m, float((nll_tc - nll_cc).abs().max()) / 4096, float((z_tc - z_cc).abs().max())), flush=True)
        xs1 = nf.sample(yb, 1.0, yb, iso=[100.0], cam=[2.0], seed=3, offset=0)
        nf.set_tensor_cores(False)
        xs2 = nf.sample(yb, 1.0, yb, iso=[100.0], cam=[2.0], seed=3, offset=0)
        nf.set_tensor_cores(True)
        print("           Philox sampling, tensor-core vs CUDA-core: diff %.3e (max |x| %.3f)" % (float((xs1 - xs2).abs().max()), float(xs2.abs().max())), flush=True)
    for nb in (4096,):
        g = torch.Generator(device="cuda").manual_seed(2)
        yb = torch.rand((nb, 32, 32, 4), device="cuda", generator=g)
        xb = torch.randn((nb, 32, 32, 4), device="cuda", generator=g) * 0.3
        for it in range(2):
            torch.cuda.synchronize(); t0 = time.time()
            nf._loss(xb, yb, iso=[100.0], cam=[2.0])
            torch.cuda.synchronize(); dt = time.time() - t0
        print("           log_prob %d patches: %.2f ms  -> %.3f M patches/s" % (nb, dt * 1e3, nb / dt / 1e6), flush=True)
