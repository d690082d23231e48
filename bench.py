#!/usr/bin/env python
"""Headline benchmark: 32x32 patches/sec of the fused Noise Flow ``log_prob`` (NLL) path.

    python bench.py --gpus N --steps K --warmup W            # this engine (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port), host cores

A "step" is one ``NoiseFlow.loss`` over a batch of ``--batch`` synthetic SIDD-shaped patches per GPU with the
shipped S-Ax4-G-Ax4 weights: fused chain kernel (x, y -> nll, sd_z) + deterministic fp64 batch reduction
(+ one NCCL all-reduce of [sum nll, sum sd_z, n] when N > 1).  ``value`` keeps inputs resident in HBM;
``e2e`` times the host-buffer C-ABI call (pinned host x, y -> H2D -> kernel -> D2H nll) for the same batch.
Prints ONE JSON line on rank 0 (contract in the task statement / DESIGN.md "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

ALG_BYTES_LOG_PROB = 32 * 32 * 4 * 4 * 2 + 8        # read x, y; write nll + sd_z (no conditioning-row array here)
ALG_BYTES_SAMPLE = 32 * 32 * 4 * 4 * 2              # read y, write x (in-kernel Philox)
CONV_FLOP_PER_PATCH = 2 * 1024 * 8 * (72 + 16 + 144 + 16)   # 2 x MACs of the 8 couplings
FALLBACK_HBM_GBS = 6650.0                            # B200_PROFILING.md fallback


def load_model_files():
    from noise_flow_b200 import hps_loader, load_checkpoint
    g = os.path.join(ROOT, "tests", "golden", "NoiseFlow")
    return hps_loader(os.path.join(g, "hps.txt")), load_checkpoint(os.path.join(g, "ckpt", "model.ckpt.best"))


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_rate(hps, ck, seconds_target=12.0, threads=None):
    """patches/s of the oracle port (torch-CPU fp32, all host threads) on a bounded sample."""
    from common import make_oracle, synth_batch
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    orc = make_oracle(hps, ck, dtype=torch.float32)
    x, y = synth_batch(256, seed=3)
    orc._loss(x, y, iso=[100.0], cam=[2.0])                      # warm-up
    t0 = time.perf_counter()
    orc._loss(x, y, iso=[100.0], cam=[2.0])
    dt = time.perf_counter() - t0
    n = int(min(16384, max(256, 256 * seconds_target / max(dt, 1e-3))))          # 16384 patches: ~6 GB of activations on the host
    reps = int(max(1, min(8, round(seconds_target * 256 / max(dt, 1e-3) / n))))   # bounded sample: ~seconds_target of CPU work
    x, y = synth_batch(n, seed=4)
    t0 = time.perf_counter()
    for _ in range(reps):
        orc._loss(x, y, iso=[100.0], cam=[2.0])
    dt = time.perf_counter() - t0
    return reps * n / dt, threads, ("log_prob of %d x %d synthetic S6/ISO-100 patches, oracle port torch-CPU fp32, %.1f s"
                                    % (reps, n, dt))


def run_reference(args):
    """--impl reference: the reference's own CPU path on the box's host cores, all threads.  TF 1.12 cannot run and
    /root/reference does not exist on the GPU box, so the timed code is the oracle port (torch-CPU fp32) -- the same
    restatement that reproduces the reference's own Python to fp64 round-off (tests/test_cpu_reference_goldens.py).
    Same config / metric / unit as our arm; every step is a bounded sample (--ref-patches) of that workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from common import make_oracle, synth_batch
    hps, ck = load_model_files()
    if args.arch:
        hps.arch = args.arch
    threads = os.cpu_count()
    torch.set_num_threads(threads)
    orc = make_oracle(hps, ck, dtype=torch.float32)
    per_step = args.ref_patches
    x, y = synth_batch(per_step, seed=5)
    eps = np.random.RandomState(6).randn(per_step, 32, 32, 4).astype(np.float32)
    if args.mode == "sample":
        step = lambda: orc.sample(eps, 0.6, y, iso=[100.0], cam=[2.0])     # noqa: E731
    else:
        step = lambda: orc.loss(x, y, iso=[100.0], cam=[2.0])             # noqa: E731
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = per_step * args.steps / dt
    mode = "sample" if args.mode == "sample" else "log_prob"
    sample = "%d steps x %d patches of the workload on %d host threads" % (args.steps, per_step, threads)
    out = {"impl": "reference", "metric": {"log_prob": "patches_per_sec_nll", "sample": "patches_per_sec_sample"}[mode],
           "value": val, "unit": "patches/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "%s Noise Flow (shipped weights, arch %s), batch %d 32x32x4 patches per GPU, "
                                  "cam S6 / ISO 100" % (mode, hps.arch, args.batch),
                      "per_gpu_batch": args.batch, "global_batch": args.gpus * args.batch, "parallelism": "dp%d" % args.gpus,
                      "width": 4, "reference_sample": sample,
                      "note": "TF 1.12/TFP 0.5 reference cannot run here; timed arm = oracle port (torch-CPU fp32)"},
           "cpu_baseline": {"value": val, "unit": "patches/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="patches per GPU per step")
    ap.add_argument("--mode", default="log_prob", choices=["log_prob", "sample", "train"])
    ap.add_argument("--ref-patches", type=int, default=1024)
    ap.add_argument("--warps", type=int, default=0, help="resident patches per CTA (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--tc", type=int, default=0, help="1: tensor-core (tcgen05) coupling convolutions")
    ap.add_argument("--cta-warps", type=int, default=0, help="train mode: warps per patch-CTA (0 = automatic, 8, 16)")
    ap.add_argument("--fused", type=int, default=1, help="train mode: 1 = one cooperative kernel per loss+gradient, 0 = one launch per pass")
    ap.add_argument("--trainer", default="device", choices=["device", "host"],
                    help="train mode: device-resident step (nf_trainer_*) or the host-synchronous path (train_step)")
    ap.add_argument("--width", type=int, default=4,
                    help="coupling-net width (reference --width): 4 = shipped weights; 8 / 16 / 32 = CTA-per-patch kernel on a "
                         "randomly initialised net of the shipped arch")
    ap.add_argument("--clean", default="uniform", choices=["uniform", "dark"],
                    help="clean patches y ~ U[0, 1) (default, the quoted workload) or the dark variant y ~ Beta(2, 5)")
    ap.add_argument("--arch", default=None, help="override hps.arch, e.g. \"sdn5|gain4\" (HBM-bound streaming kernel)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout (ONE JSON line)
        dist.init_process_group("nccl", device_id=dev)
    from noise_flow_b200 import NoiseFlow, _lib
    from noise_flow_b200.distributed import allreduce_sums
    hps, ck = load_model_files()
    if args.arch:
        hps.arch = args.arch
    if args.width != 4:     # no shipped checkpoint at other widths: reference initialisers, then non-trivial net weights
        import numpy as _np
        hps.width = args.width
        nf0 = NoiseFlow([32, 32, 4], False, hps, variables={k: v for k, v in ck.items() if "real_nvp_conv_template" not in k},
                        first_call="inverse", device=dev, seed=0)
        rng = _np.random.RandomState(0)
        ck = {k: v.copy() for k, v in nf0.variables.items()}
        for k in ck:
            if k.endswith("/l_1/W"):
                ck[k] = (rng.randn(*ck[k].shape) * 0.5).astype(_np.float32)
            elif k.endswith("/l_2/W"):
                ck[k] = (rng.randn(*ck[k].shape) / _np.sqrt(args.width)).astype(_np.float32)
            elif k.endswith("/l_last/W"):
                ck[k] = (rng.randn(*ck[k].shape) * 0.05 / _np.sqrt(args.width)).astype(_np.float32)
        del nf0
    nf = NoiseFlow([32, 32, 4], False, hps, variables=ck, first_call="inverse", device=dev)
    if args.warps:
        nf.set_launch(args.warps, 0)
    if args.tc:
        nf.set_tensor_cores(True)
    lib, eng = _lib.load(), nf._engine
    B = args.batch
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    y = torch.rand((B, 32, 32, 4), device=dev, generator=g)
    if args.clean == "dark":      # SURVEY 8d's second clean-image distribution: y ~ Beta(2, 5) = G1 / (G1 + G2), mean 0.29
        torch.manual_seed(4321 + rank)
        conc = torch.tensor([2.0, 5.0], device=dev).view(2, 1, 1, 1, 1).expand(2, B, 32, 32, 4)
        ga = torch._standard_gamma(conc)
        y = (ga[0] / (ga[0] + ga[1])).contiguous()
        del ga, conc
    x = torch.randn((B, 32, 32, 4), device=dev, generator=g) * torch.sqrt(0.000479 * y + 0.000002)
    nll = torch.empty(B, device=dev)
    sdz = torch.empty(B, device=dev)
    xs = torch.empty_like(x) if args.mode == "sample" else None
    sums = torch.zeros(3, device=dev, dtype=torch.float64)
    stream = torch.cuda.current_stream(dev).cuda_stream
    row = 2 * 5 + 0   # S6, ISO 100

    def kernel_only(i):
        if args.mode == "train":
            return
        if args.mode == "log_prob":
            _lib.check(lib.nf_log_prob(eng.handle, x.data_ptr(), y.data_ptr(), None, row, B, nll.data_ptr(),
                                       sdz.data_ptr(), None, stream))
        else:
            _lib.check(lib.nf_sample(eng.handle, y.data_ptr(), None, row, B, 0.6, None, 7, i, rank * B,
                                     xs.data_ptr(), stream))

    train_opt = trainer = None
    last_loss = [None]
    if args.mode == "train":
        from noise_flow_b200 import train as nf_train
        from noise_flow_b200.train import AdamOptimizer, DeviceTrainer, train_step
        if args.trainer == "device":
            trainer = DeviceTrainer(nf, learning_rate=1e-4, max_batch=B, cta_warps=args.cta_warps, fused=bool(args.fused))
        else:
            train_opt = AdamOptimizer(learning_rate=1e-4)

    def step(i):
        if args.mode == "train":     # BASELINE config 5: one Adam step (batch-stat BN forward + backward + update)
            if trainer is not None:  # loss and sd_z come back to the host every step, as sess.run returns them
                last_loss[0] = trainer.step(x, y, iso=[100.0], cam=[2.0])
            else:
                last_loss[0] = train_step(nf, train_opt, x, y, iso=[100.0], cam=[2.0])
            return
        kernel_only(i)
        if args.mode == "log_prob":
            _lib.check(lib.nf_reduce_sums(nll.data_ptr(), sdz.data_ptr(), B, sums.data_ptr(), stream))
            if world > 1:
                allreduce_sums(sums)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.mode == "train":
        nf_train.TIMINGS = {}
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    # dominant kernel alone (CUDA events on the launching stream), for the roofline line
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for i in range(args.steps):
        kernel_only(1000 + i)
    k1.record()
    torch.cuda.synchronize(dev)
    kms = k0.elapsed_time(k1) / args.steps
    t = torch.tensor([ms, kms], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, kms = float(t[0]), float(t[1])
    mean_nll = float(sums[0] / sums[2]) / 4096 if args.mode == "log_prob" else None

    # ---- e2e: host buffers through the C-ABI host entry point (copies inside the timed region)
    e2e = None
    if not args.no_e2e and args.mode == "train" and trainer is not None:
        # the reference's train thread feeds numpy minibatches (sidd/MiniBatchSampler.py): pinned host batch -> device,
        # one Adam step, loss and sd_z back on the host, every step
        hx_t = torch.empty((B, 32, 32, 4), dtype=torch.float32, pin_memory=True)
        hy_t = torch.empty((B, 32, 32, 4), dtype=torch.float32, pin_memory=True)
        hx_t.copy_(x); hy_t.copy_(y)
        e_steps = args.steps

        def e2e_train_step():
            xd, yd = hx_t.to(dev, non_blocking=True), hy_t.to(dev, non_blocking=True)
            return trainer.step(xd, yd, iso=[100.0], cam=[2.0])
        e2e_train_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_train_step()
        barrier()
        edt = time.perf_counter() - t0
        tt = torch.tensor([edt], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        edt = float(tt[0])
        e2e = {"value": world * B * e_steps / edt, "unit": "patches/s", "h2d_bytes_per_step": 2 * B * 16384,
               "d2h_bytes_per_step": 24, "steps": e_steps,
               "path": "DeviceTrainer.step on a pinned host minibatch: H2D of x and y, Adam step, loss/sd_z read back"}
    if not args.no_e2e and args.mode != "train":
        import ctypes as C
        nb = B * 4096 * 4
        hx_t = torch.empty((B, 32, 32, 4), dtype=torch.float32, pin_memory=True)
        hy_t = torch.empty((B, 32, 32, 4), dtype=torch.float32, pin_memory=True)
        hn_t = torch.empty((B,), dtype=torch.float32, pin_memory=True)
        hx_t.copy_(x); hy_t.copy_(y)
        torch.cuda.synchronize(dev)
        hx, hy, hn = hx_t.data_ptr(), hy_t.data_ptr(), hn_t.data_ptr()
        hsums = (C.c_double * 3)()
        e_steps = max(2, min(args.steps, 5))

        def e2e_step():
            if args.mode == "log_prob":
                _lib.check(lib.nf_log_prob_host(eng.handle, hx, hy, None, row, B, hn, None, None, hsums))
            else:
                _lib.check(lib.nf_sample_host(eng.handle, hy, None, row, B, 0.6, None, 7, 0, hx))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        barrier()
        edt = time.perf_counter() - t0
        tt = torch.tensor([edt], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        edt = float(tt[0])
        h2d = 2 * nb if args.mode == "log_prob" else nb
        d2h = B * 4 if args.mode == "log_prob" else nb
        e2e = {"value": world * B * e_steps / edt, "unit": "patches/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e_steps,
               "path": "nf_%s_host: pinned host buffers, 4096-patch chunks in flight on 4 streams" % args.mode}
        del hx_t, hy_t, hn_t

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return 0
    value = world * B * args.steps / (ms * 1e-3)
    peak, peak_src = measured_peak()
    alg = ALG_BYTES_LOG_PROB if args.mode == "log_prob" else ALG_BYTES_SAMPLE
    achieved = B * alg / (kms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("%s_%d" % (args.mode, B))
        except Exception:
            traffic = None
    n_couplings = hps.arch.split("|").count("unc")
    conv_flop = CONV_FLOP_PER_PATCH * n_couplings / 8.0
    if args.width != 4:     # 18 W + W^2 + 36 W + 16 MAC per pixel and coupling
        wd = args.width
        conv_flop = 2.0 * 1024 * n_couplings * (18 * wd + wd * wd + 36 * wd + 16)
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    out = {"metric": {"log_prob": "patches_per_sec_nll", "sample": "patches_per_sec_sample",
                      "train": "patches_per_sec_adam_step"}[args.mode],
           "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": "%s Noise Flow (shipped weights, arch %s), batch %d 32x32x4 patches per GPU, "
                                  "cam S6 / ISO 100" % (args.mode, hps.arch, B),
                      "per_gpu_batch": B, "global_batch": world * B, "parallelism": "dp%d" % world,
                      "width": args.width, "clean": args.clean,
                      "l2": "inputs (%.1f GiB per GPU) exceed the 126 MB L2" % (2 * B * 16384 / 2 ** 30),
                      "mean_nll_per_dim": mean_nll},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "peak_source": peak_src, "kernel": "nf_scale_stream_kernel" if "unc" not in hps.arch else ("nf_chain_kernel" if args.width == 4 else "nf_wide_chain_kernel"),
                        "kernel_ms": kms, "alg_bytes_per_patch": alg,
                        "note": ("binding roof is the FP32 FMA pipe (see roofline_fp32), not HBM" if n_couplings else
                                 "scale-layer-only chain: streaming kernel, HBM-bound")},
           "roofline_fp32": {"bound": "fp32_fma", "achieved": B * conv_flop / (kms * 1e-3) / 1e12,
                             "peak": fp32_peak, "unit": "TFLOP/s",
                             "frac": B * conv_flop / (kms * 1e-3) / 1e12 / fp32_peak,
                             "peak_source": "148 SMs x 128 FMA/clk x 2 x median SM clock under load"},
           "e2e": e2e, "gpu_launches": world * args.steps * (2 if args.mode == "log_prob" else 1), "clocks": clocks}
    if args.mode == "train":   # a step is ~70 small launches + host chain rules: no single-kernel roofline applies
        n_cp = max(n_couplings, 1)
        out["roofline"] = None
        out["roofline_fp32"] = None
        out["config"]["trainer"] = args.trainer
        out["config"]["cta_warps"] = args.cta_warps
        out["config"]["fused"] = args.fused
        out["config"]["loss_per_dim_last_step"] = None if last_loss[0] is None else float(last_loss[0][0]) / 4096
        if trainer is not None:
            out["gpu_launches"] = world * args.steps * trainer.launches_per_step(True)
            out["config"]["note"] = ("one sess.run([train_op, loss, sd_z]) equivalent, device-resident: LU assembly + scale tables, "
                                     "batch-stat BN forward (3 passes per coupling), backward (3 passes per coupling), chain rules, "
                                     "Adam, BN moving averages -- one stream, no synchronisation except the 24-byte loss read-back"
                                     + ("; one NCCL all-reduce of the reduce buffer" if world > 1 else ""))
        else:
            out["gpu_launches"] = world * args.steps * (n_cp * 6 + (len(hps.arch.split("|")) - n_cp) * 2 + 2)
            out["config"]["host_ms_per_step"] = {k: round(1e3 * v / args.steps, 3) for k, v in (nf_train.TIMINGS or {}).items()}
            out["config"]["note"] = ("one sess.run([train_op, loss, sd_z]) equivalent: batch-stat BN forward (2 probes + apply per "
                                     "coupling), backward (3 passes per coupling), host LU/scale chain rules, Adam, re-fold")
    if world == 1 and not args.no_cpu_baseline and args.mode != "train":
        v, cores, sample = cpu_oracle_rate(hps, ck)
        out["cpu_baseline"] = {"value": v, "unit": "patches/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    try:
        rc = main()
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        rc = 1
    sys.stdout.flush()
    sys.exit(rc)
