"""Latency of the reference's public sampling call, NoiseFlowWrapper.sample_noise_nf (numpy in, numpy out), per batch size
and BatchNorm mode.  sample_noise_flow.py calls it with ONE patch at a time, train_dncnn_noiseflow.py with minibatches."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import numpy as np, torch
from noise_flow_b200.NoiseFlowWrapper import NoiseFlowWrapper
path = os.path.join(ROOT, "tests", "golden", "NoiseFlow")
rng = np.random.RandomState(0)
for mode in ("batch", "moving"):
    w = NoiseFlowWrapper(path, sampling_temperature=0.6, bn_mode=mode)
    for n in (1, 16, 128, 1024, 8192):
        y = rng.rand(n, 32, 32, 4).astype(np.float32)
        for _ in range(5):
            w.sample_noise_nf(y, 0.0, 0.0, 100, 2)
        torch.cuda.synchronize()
        reps = 100 if n <= 1024 else 20
        t0 = time.perf_counter()
        for _ in range(reps):
            out = w.sample_noise_nf(y, 0.0, 0.0, 100, 2)
        dt = (time.perf_counter() - t0) / reps
        print("bn_mode=%-6s n=%5d  %8.3f ms/call  %10.0f patches/s" % (mode, n, dt * 1e3, n / dt), flush=True)
