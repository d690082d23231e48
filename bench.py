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


def cpu_oracle_rate(hps, ck, seconds_target=12.0, threads=None, per_step=1024):
    """patches/s of the oracle port (torch-CPU fp32, all host threads) on a bounded sample: steps of `per_step` patches -- the SAME
    step size as `--impl reference` (--ref-patches), so the two CPU numbers of a round agree (larger steps run slower on the
    host: 16 384-patch steps measured 5.1 k patches/s against 10 k at 1024, round 1)."""
    from common import make_oracle, synth_batch
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    orc = make_oracle(hps, ck, dtype=torch.float32)
    x, y = synth_batch(per_step, seed=4)
    orc._loss(x, y, iso=[100.0], cam=[2.0])                      # warm-up
    t0 = time.perf_counter()
    orc._loss(x, y, iso=[100.0], cam=[2.0])
    dt1 = time.perf_counter() - t0
    reps = int(max(2, min(200, round(seconds_target / max(dt1, 1e-3)))))
    t0 = time.perf_counter()
    for _ in range(reps):
        orc._loss(x, y, iso=[100.0], cam=[2.0])
    dt = time.perf_counter() - t0
    return reps * per_step / dt, threads, ("log_prob, %d steps x %d synthetic S6/ISO-100 patches, oracle port torch-CPU fp32, %.1f s"
                                           % (reps, per_step, dt))


def workload_config(mode, arch, B, world, width, clean):
    """`config` of the JSON line -- the SAME dict for this engine's arm and for `--impl reference` (the driver compares them)."""
    return {"workload": "%s Noise Flow (shipped weights, arch %s), batch %d 32x32x4 patches per GPU, cam S6 / ISO 100" % (mode, arch, B),
            "per_gpu_batch": B, "global_batch": world * B, "parallelism": "dp%d" % world, "width": width, "clean": clean,
            "l2": "inputs (%.1f GiB per GPU) exceed the 126 MB L2" % (2 * B * 16384 / 2 ** 30)}


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
           "config": workload_config(mode, hps.arch, args.batch, args.gpus, 4, args.clean),
           "reference_note": "TF 1.12/TFP 0.5 reference cannot run here; timed arm = oracle port (torch-CPU fp32); every step is a "
                             "bounded sample of the workload: " + sample,
           "cpu_baseline": {"value": val, "unit": "patches/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------------------
# BASELINE.json configs 2-5 next to the headline line: each entry is a short timed run of its own (CUDA events around
# K launches after W warm-up launches, barrier + synchronize on both sides, max over ranks), with its own roofline.
# ----------------------------------------------------------------------------------------------------------------
def _timed(fn, steps, warmup, dev, world):
    """ms per step of fn(i), device-timed on torch's current stream, max over ranks."""
    for i in range(warmup):
        fn(i)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t[0])


def _hbm_roofline(n, alg_bytes, ms, kernel, traffic=None, note=None):
    peak, src = measured_peak()
    ach = n * alg_bytes / (ms * 1e-3) / 1e9
    r = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
         "peak_source": src, "kernel": kernel, "kernel_ms": ms, "alg_bytes_per_patch": alg_bytes}
    if note:
        r["note"] = note
    return r


def _fp32_roofline(n, flop_per_patch, ms, sm_mhz):
    peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    ach = n * flop_per_patch / (ms * 1e-3) / 1e12
    return {"bound": "fp32_fma", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "peak_source": "148 SMs x 128 FMA/clk x 2 x median SM clock under load",
            "flop_count": "ALGORITHMIC: the direct-form convolutions of the reference (4.06 MFLOP per patch for 8 couplings); the default "
                          "(Winograd) kernel executes 29 % fewer multiplies for the same result, so this is an effective rate"}


def _traffic(key):
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(tp)).get(key)
    except Exception:
        return None


def h2d_ceiling(dev, world, mib=256, reps=4):
    """What bounds `e2e`: pinned host -> device copy bandwidth with EVERY rank copying at the same time (the ranks of one box
    share the host's memory system and PCIe root complexes).  Best of three shapes (1 / 2 / 4 streams), two timed passes each:
    one shape alone under-reads the link on some boxes (47.6 GB/s on two streams next to 54.7 GB/s through nf_log_prob_host).
    Returns (aggregate GB/s over all ranks, this rank's GB/s)."""
    nbytes = mib << 20
    hs = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(4)]
    ds = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(4)]
    ss = [torch.cuda.Stream(device=dev) for _ in range(4)]
    for h in hs:
        h.fill_(1)
    best_agg, best_mine = 0.0, 0.0
    for ns in (1, 2, 4):
        def run():
            for _ in range(reps * 2 // ns if ns <= 2 else max(1, reps // 2)):
                for h, d, st in zip(hs[:ns], ds[:ns], ss[:ns]):
                    with torch.cuda.stream(st):
                        d.copy_(h, non_blocking=True)
        copies = (reps * 2 // ns if ns <= 2 else max(1, reps // 2)) * ns
        run()
        torch.cuda.synchronize(dev)
        for _ in range(2):
            if world > 1:
                torch.distributed.barrier()
            t0 = time.perf_counter()
            run()
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            mine = copies * nbytes / dt / 1e9
            if world > 1:
                torch.distributed.barrier()
                tmax = torch.tensor([dt], device=dev, dtype=torch.float64)
                torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
                agg = world * copies * nbytes / float(tmax[0]) / 1e9
            else:
                agg = mine
            if agg > best_agg:
                best_agg, best_mine = agg, mine
    del hs, ds
    return best_agg, best_mine


def wide_model(hps, ck, width, dev):
    """A net of the shipped arch at another coupling width: no shipped checkpoint exists, so reference initialisers for the
    structure and O(1)-activation random weights for the coupling nets."""
    import copy as _copy
    from noise_flow_b200 import NoiseFlow
    h = _copy.copy(hps)
    h.width = width
    nf0 = NoiseFlow([32, 32, 4], False, h, variables={k: v for k, v in ck.items() if "real_nvp_conv_template" not in k},
                    first_call="inverse", device=dev, seed=0)
    rng = np.random.RandomState(0)
    vs = {k: v.copy() for k, v in nf0.variables.items()}
    for k in vs:
        if k.endswith("/l_1/W"):
            vs[k] = (rng.randn(*vs[k].shape) * 0.5).astype(np.float32)
        elif k.endswith("/l_2/W"):
            vs[k] = (rng.randn(*vs[k].shape) / np.sqrt(width)).astype(np.float32)
        elif k.endswith("/l_last/W"):
            vs[k] = (rng.randn(*vs[k].shape) * 0.05 / np.sqrt(width)).astype(np.float32)
    del nf0
    return h, vs


def tensor_roofline(n, width, n_couplings, ms):
    """Tensor-pipe roofline of the wide-net kernels.  `achieved` = ALGORITHMIC flops (2 x the fp32 MACs of the three
    convolutions: 18 W + W^2 + 36 W per pixel and coupling) / time; `issued` counts what the bf16 (hi, lo) emulation really
    sends through tcgen05.mma (conv-1 K = 64, conv-2 three products + bias chunk, conv-3 N = 96 + 48 of which 72 useful)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        peak, src = float(json.load(open(p))["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    except Exception:
        peak, src = 1590.0, "fallback (B200_PROFILING.md)"
    alg = 2.0 * 1024 * n_couplings * (18 * width + width * width + 36 * width)
    issued = 2.0 * 1024 * n_couplings * (64 * width + 3 * width * width + 16 * width + width * 144)
    ach = n * alg / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
            "peak_source": src, "kernel": "nf_wide_tc_kernel" if width <= 128 else "nf_wide_tcs_kernel", "kernel_ms": ms,
            "alg_flop_per_patch": alg, "issued_tflops": n * issued / (ms * 1e-3) / 1e12,
            "issued_frac": n * issued / (ms * 1e-3) / 1e12 / peak,
            "note": "fp32-grade results from bf16 (hi, lo) split operands: 3 tensor-core products per fp32 MAC"}


def shard_check(nf, lib, dev, rank, world, n_global=8192):
    """sharded mean NLL == single-GPU mean NLL.  ONE seeded global batch; rank r evaluates its contiguous
    slice, one all-reduce of [sum nll, sum sd_z, n] (fp64); rank 0 also evaluates the whole batch alone."""
    from noise_flow_b200.distributed import allreduce_sums, shard_range
    g = torch.Generator(device=dev).manual_seed(777)           # same seed on every rank: the same global batch
    y = torch.rand((n_global, 32, 32, 4), device=dev, generator=g)
    x = torch.randn((n_global, 32, 32, 4), device=dev, generator=g) * torch.sqrt(0.000479 * y + 0.000002)
    lo, hi = shard_range(n_global, rank, world)
    nf._loss(x[lo:hi], y[lo:hi], iso=[100.0], cam=[2.0])
    sums = nf._tls.last_sums.clone()
    allreduce_sums(sums)
    out = None
    if rank == 0:
        nf._loss(x, y, iso=[100.0], cam=[2.0])
        one = nf._tls.last_sums
        out = {"n_global": n_global, "world": world,
               "mean_nll_per_dim_sharded": float(sums[0] / sums[2]) / 4096, "mean_nll_per_dim_single": float(one[0] / one[2]) / 4096,
               "abs_diff_nats_per_dim": abs(float(sums[0] / sums[2]) - float(one[0] / one[2])) / 4096,
               "sd_z_abs_diff": abs(float(sums[1] / sums[2]) - float(one[1] / one[2])),
               "count_equal": float(sums[2]) == float(one[2]) == float(n_global), "tolerance": 1e-7}
        out["ok"] = bool(out["abs_diff_nats_per_dim"] < 1e-7 and out["count_equal"])
    del x, y
    return out


def run_also(args, nf, hps, ck, x, y, dev, rank, world, sm_mhz, row):
    """The other BASELINE.json configs, every one at `world` ranks (so the driver's scaling run records them at 8 GPUs)."""
    from noise_flow_b200 import NoiseFlow, _lib
    from noise_flow_b200.distributed import allreduce_sums
    import copy as _copy
    lib, eng = _lib.load(), nf._engine
    B = x.shape[0]
    stream = torch.cuda.current_stream(dev).cuda_stream
    nll = torch.empty(B, device=dev)
    sdz = torch.empty(B, device=dev)
    sums = torch.zeros(3, device=dev, dtype=torch.float64)
    out = {}
    conv_flop = CONV_FLOP_PER_PATCH * hps.arch.split("|").count("unc") / 8.0

    def log_prob_step(handle, xx, yy, n, reduce=True):
        _lib.check(lib.nf_log_prob(handle, xx.data_ptr(), yy.data_ptr(), None, row, n, nll.data_ptr(), sdz.data_ptr(), None, stream))
        if reduce:
            _lib.check(lib.nf_reduce_sums(nll.data_ptr(), sdz.data_ptr(), n, sums.data_ptr(), stream))
            if world > 1:
                allreduce_sums(sums)

    # ---- config 3: sample_noise_flow inverse pass, 65536 patches conditioned on clean + cam / ISO, in-kernel Philox
    xs = torch.empty_like(x)
    ms = _timed(lambda i: _lib.check(lib.nf_sample(eng.handle, y.data_ptr(), None, row, B, 0.6, None, 7, i, rank * B, xs.data_ptr(), stream)),
                20, 3, dev, world)
    out["sample_%d" % B] = {
        "metric": "patches_per_sec_sample", "value": world * B / (ms * 1e-3), "unit": "patches/s", "ms_per_step": ms, "steps": 20,
        "config": {"workload": "sample (temperature 0.6, in-kernel Philox) Noise Flow, batch %d per GPU, cam S6 / ISO 100" % B,
                   "baseline_config": 3, "per_gpu_batch": B},
        "roofline": _hbm_roofline(B, ALG_BYTES_SAMPLE, ms, "nf_chain_wino_kernel<false>" if args.tc in (-1, 0, 4) else "nf_chain_kernel<false>",
                                  _traffic("sample_%d" % B), "binding roof is the FP32 FMA pipe (roofline_fp32)"),
        "roofline_fp32": _fp32_roofline(B, conv_flop, ms, sm_mhz), "gpu_launches": world * 20}
    # the same sampler and log_prob on the hybrid kernel (conv-3 on tcgen05; opt-in: nf_model_set_tensor_cores 2)
    nf.set_tensor_cores(2)
    ms = _timed(lambda i: _lib.check(lib.nf_sample(eng.handle, y.data_ptr(), None, row, B, 0.6, None, 7, i, rank * B, xs.data_ptr(), stream)),
                20, 3, dev, world)
    out["sample_%d_hybrid_kernel" % B] = {
        "metric": "patches_per_sec_sample", "value": world * B / (ms * 1e-3), "unit": "patches/s", "ms_per_step": ms, "steps": 20,
        "config": {"workload": "sample, hybrid kernel (conv-3 on tcgen05; nf_model_set_tensor_cores 2), batch %d per GPU" % B, "per_gpu_batch": B},
        "roofline": _hbm_roofline(B, ALG_BYTES_SAMPLE, ms, "nf_chain_hyb_kernel<false>", None, "binding roof is the FP32 FMA pipe (roofline_fp32)"),
        "roofline_fp32": _fp32_roofline(B, conv_flop, ms, sm_mhz), "gpu_launches": world * 20}
    del xs
    nf.set_tensor_cores(2)
    ms = _timed(lambda i: log_prob_step(eng.handle, x, y, B), 20, 3, dev, world)
    out["log_prob_%d_hybrid_kernel" % B] = {
        "metric": "patches_per_sec_nll", "value": world * B / (ms * 1e-3), "unit": "patches/s", "ms_per_step": ms, "steps": 20,
        "config": {"workload": "log_prob, hybrid kernel (conv-3 on tcgen05; nf_model_set_tensor_cores 2), batch %d per GPU" % B, "per_gpu_batch": B},
        "roofline": _hbm_roofline(B, ALG_BYTES_LOG_PROB, ms, "nf_chain_hyb_kernel<true>", None, "binding roof is the FP32 FMA pipe (roofline_fp32)"),
        "roofline_fp32": _fp32_roofline(B, conv_flop, ms, sm_mhz), "gpu_launches": world * 20 * 2}
    nf.set_tensor_cores(5)   # and on the direct-form all-fp32 kernel (round 1-4's default; the library default is the Winograd form)
    xs = torch.empty_like(x)
    ms = _timed(lambda i: _lib.check(lib.nf_sample(eng.handle, y.data_ptr(), None, row, B, 0.6, None, 7, i, rank * B, xs.data_ptr(), stream)),
                20, 3, dev, world)
    out["sample_%d_direct_kernel" % B] = {
        "metric": "patches_per_sec_sample", "value": world * B / (ms * 1e-3), "unit": "patches/s", "ms_per_step": ms, "steps": 20,
        "config": {"workload": "sample, direct-form all-fp32 kernel (nf_model_set_tensor_cores 5), batch %d per GPU" % B, "per_gpu_batch": B},
        "roofline": _hbm_roofline(B, ALG_BYTES_SAMPLE, ms, "nf_chain_kernel<false>", None, "binding roof is the FP32 FMA pipe (roofline_fp32)"),
        "roofline_fp32": _fp32_roofline(B, conv_flop, ms, sm_mhz), "gpu_launches": world * 20}
    del xs
    ms = _timed(lambda i: log_prob_step(eng.handle, x, y, B), 20, 3, dev, world)
    out["log_prob_%d_direct_kernel" % B] = {
        "metric": "patches_per_sec_nll", "value": world * B / (ms * 1e-3), "unit": "patches/s", "ms_per_step": ms, "steps": 20,
        "config": {"workload": "log_prob, direct-form all-fp32 kernel (nf_model_set_tensor_cores 5), batch %d per GPU" % B, "per_gpu_batch": B},
        "roofline": _hbm_roofline(B, ALG_BYTES_LOG_PROB, ms, "nf_chain_kernel<true>", None, "binding roof is the FP32 FMA pipe (roofline_fp32)"),
        "roofline_fp32": _fp32_roofline(B, conv_flop, ms, sm_mhz), "gpu_launches": world * 20 * 2}
    nf.set_tensor_cores(0 if args.tc < 0 else args.tc)
    # ---- configs 2 and 4: log_prob at 4096 (config 2) and the batch sweep 1k / 16k / 64k / 256k per GPU (config 4); batches
    # smaller than the resident one walk through it slice by slice, so consecutive launches never re-read L2-resident inputs
    for n in (1024, 4096, 16384, 262144):
        if n <= B:
            nsl = B // n
            fn = lambda i, n=n, nsl=nsl: log_prob_step(eng.handle, x[(i % nsl) * n:(i % nsl + 1) * n], y[(i % nsl) * n:(i % nsl + 1) * n], n)   # noqa: E731
            steps = 50
            ms = _timed(fn, steps, 3, dev, world)
            l2 = "slices of the resident %d-patch batch in rotation (footprint %.1f GiB > L2)" % (B, 2 * B * 16384 / 2 ** 30)
        else:
            try:
                reps = n // B
                xb, yb = x.repeat(reps, 1, 1, 1), y.repeat(reps, 1, 1, 1)
                nll, sdz = torch.empty(n, device=dev), torch.empty(n, device=dev)
            except RuntimeError:
                continue
            steps = 5
            ms = _timed(lambda i: log_prob_step(eng.handle, xb, yb, n), steps, 3, dev, world)
            del xb, yb
            nll, sdz = torch.empty(B, device=dev), torch.empty(B, device=dev)
            l2 = "inputs (%.1f GiB per GPU) exceed the 126 MB L2" % (2 * n * 16384 / 2 ** 30)
        out["log_prob_%d" % n] = {
            "metric": "patches_per_sec_nll", "value": world * n / (ms * 1e-3), "unit": "patches/s", "ms_per_step": ms, "steps": steps,
            "config": {"workload": "log_prob Noise Flow (kernel + fp64 reduce%s), batch %d per GPU" % (" + all-reduce" if world > 1 else "", n),
                       "baseline_config": 2 if n == 4096 else 4, "per_gpu_batch": n, "l2": l2},
            "roofline": _hbm_roofline(n, ALG_BYTES_LOG_PROB, ms, "nf_chain_wino_kernel<true>" if args.tc in (-1, 0, 3, 4) else "nf_chain_kernel<true>", None,
                                      "whole step (kernel + reduce) timed; binding roof is the FP32 FMA pipe (roofline_fp32)"),
            "roofline_fp32": _fp32_roofline(n, conv_flop, ms, sm_mhz), "gpu_launches": world * steps * 2}
    # ---- the HBM-bound streaming kernel: the reference's sdn5|gain4 baseline model (job_noise_flow.sh:53); the only chain the
    # north star's ">= 70 % of HBM" target applies to
    h2 = _copy.copy(hps)
    h2.arch = "sdn5|gain4"
    nf2 = NoiseFlow([32, 32, 4], False, h2, variables=ck, first_call="inverse", device=dev)
    ms = _timed(lambda i: log_prob_step(nf2._engine.handle, x, y, B, reduce=False), 50, 3, dev, world)
    out["stream_sdn5_gain4_%d" % B] = {
        "metric": "patches_per_sec_nll", "value": world * B / (ms * 1e-3), "unit": "patches/s", "ms_per_step": ms, "steps": 50,
        "config": {"workload": "log_prob of the scale-layer-only model sdn5|gain4 (job_noise_flow.sh:53), batch %d per GPU" % B,
                   "per_gpu_batch": B, "l2": "inputs (%.1f GiB per GPU) exceed the 126 MB L2" % (2 * B * 16384 / 2 ** 30)},
        "roofline": _hbm_roofline(B, ALG_BYTES_LOG_PROB, ms, "nf_scale_stream_kernel", None, "HBM-bound: ~60 flop per 32 KiB patch"),
        "gpu_launches": world * 50}
    del nf2
    # ---- wide coupling nets on the tensor cores: width 32 ("for Noise Flow it is 32", job_noise_flow.sh:19) and the
    # reference's default width 512 (sidd/ArgParser.py:43)
    for width, nb in ((32, 16384), (512, 2048)):
        hw, vw = wide_model(hps, ck, width, dev)
        nfw = NoiseFlow([32, 32, 4], False, hw, variables=vw, first_call="inverse", device=dev)
        steps = 10 if width == 32 else 3
        ms = _timed(lambda i: log_prob_step(nfw._engine.handle, x[:nb], y[:nb], nb, reduce=False), steps, 3, dev, world)
        out["wide_w%d_log_prob_%d" % (width, nb)] = {
            "metric": "patches_per_sec_nll", "value": world * nb / (ms * 1e-3), "unit": "patches/s", "ms_per_step": ms, "steps": steps,
            "config": {"workload": "log_prob, shipped arch at coupling-net width %d (random weights), batch %d per GPU" % (width, nb),
                       "per_gpu_batch": nb, "width": width,
                       "l2": "inputs %.2f GiB per GPU > L2" % (2 * nb * 16384 / 2 ** 30)},
            "roofline": tensor_roofline(nb, width, hw.arch.split("|").count("unc"), ms), "gpu_launches": world * steps}
        del nfw
    # ---- config 5: one Adam step (batch-statistics BN forward + backward + update) on a synthetic SIDD-shaped minibatch of
    # 138 patches per GPU (job_noise_flow.sh:37), one NCCL all-reduce of the reduce buffer when world > 1
    from noise_flow_b200.train import DeviceTrainer
    nb = 138
    nft = NoiseFlow([32, 32, 4], True, _copy.copy(hps), variables={k: v.copy() for k, v in ck.items()}, first_call="inverse", device=dev)
    tr = DeviceTrainer(nft, learning_rate=1e-4, max_batch=nb)
    xt, yt = x[:nb].contiguous(), y[:nb].contiguous()
    ms = _timed(lambda i: tr.step(xt, yt, iso=[100.0], cam=[2.0]), 100, 5, dev, world)
    import ctypes as C
    nbar, us = C.c_int(0), C.c_float(0.0)
    _lib.check(lib.nf_trainer_barriers_per_step(tr.handle, 1, C.byref(nbar)))
    _lib.check(lib.nf_probe_grid_barrier(nb, 200, C.byref(us), stream))
    n_cp = hps.arch.split("|").count("unc")
    # algorithmic work: forward 248 MAC per pixel and coupling, backward 2x that (input gradients + parameter gradients)
    train_flop = 3 * 2.0 * 1024 * n_cp * 248
    fp = _fp32_roofline(nb, train_flop, ms, sm_mhz)
    floor_ms = nbar.value * us.value * 1e-3
    out["train_adam_step_%d" % nb] = {
        "metric": "patches_per_sec_adam_step", "value": world * nb / (ms * 1e-3), "unit": "patches/s", "ms_per_step": ms, "steps": 100,
        "config": {"workload": "one Adam step (sess.run([train_op, loss, sd_z]) equivalent, device-resident), %d patches per GPU" % nb,
                   "baseline_config": 5, "per_gpu_batch": nb, "global_batch": world * nb,
                   "collective": "one NCCL all-reduce of [grads | sums | batch statistics] (fp64)" if world > 1 else "none (1 GPU)"},
        "roofline": {"bound": "latency", "achieved": ms, "peak": floor_ms, "unit": "ms", "frac": floor_ms / ms if ms > 0 else None,
                     "traffic": None, "kernel": "td_step_kernel<8>", "grid_barriers_per_step": nbar.value,
                     "us_per_grid_barrier": us.value,
                     "note": "a %d-patch step is a latency chain: floor = grid barriers x measured barrier cost on a %d-CTA cooperative "
                             "grid; frac = floor / measured step" % (nb, nb)},
        "roofline_fp32": fp, "gpu_launches": world * 100 * tr.launches_per_step(True)}
    del tr, nft
    return out


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
    ap.add_argument("--tc", type=int, default=-1, help="width 4: -1 = library default (hybrid kernel for sampling, all-fp32 for log_prob), "
                    "0 = all-fp32 kernel, 2 = hybrid kernel (conv-3 on tcgen05) in both directions, 1 = older all-TC experiment")
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
    ap.add_argument("--no-also", action="store_true", help="skip the `also` block (BASELINE configs 2-5 next to the headline)")
    ap.add_argument("--check-shard", action="store_true",
                    help="add `shard_check`: one seeded global batch, sharded mean NLL (all-reduce) vs rank 0's single-GPU pass")
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
        # ONE JSON line on stdout: NCCL prints its version banner there at NCCL_DEBUG >= VERSION (WARN and INFO included);
        # whatever level the environment asks for goes to stderr instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    from noise_flow_b200 import NoiseFlow, _lib
    from noise_flow_b200.distributed import allreduce_sums
    hps, ck = load_model_files()
    if args.arch:
        hps.arch = args.arch
    if args.width != 4:     # no shipped checkpoint at other widths: reference initialisers, then non-trivial net weights
        hps, ck = wide_model(hps, ck, args.width, dev)
    nf = NoiseFlow([32, 32, 4], False, hps, variables=ck, first_call="inverse", device=dev)
    if args.warps:
        nf.set_launch(args.warps, 0)
    if args.tc >= 0:
        nf.set_tensor_cores(args.tc)
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
        del hx_t, hy_t, hn_t
        ceil_agg, ceil_mine = h2d_ceiling(dev, world)
        moved = world * (h2d + d2h) * e_steps / edt / 1e9          # both directions share nothing on PCIe, but the host side does
        e2e = {"value": world * B * e_steps / edt, "unit": "patches/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e_steps,
               "path": "nf_%s_host: pinned host buffers, 4096-patch chunks in flight on 4 streams" % args.mode,
               "achieved_gbs": moved, "h2d_gbs": world * h2d * e_steps / edt / 1e9, "ceiling_gbs": ceil_agg,
               "ceiling_gbs_rank0": ceil_mine, "frac_of_ceiling": world * h2d * e_steps / edt / 1e9 / ceil_agg,
               "ceiling_note": "ceiling = pinned host -> device copies of 256 MiB per rank (best of 1 / 2 / 4 streams), all %d ranks at once, "
                               "measured in this run; frac = this run's H2D rate / that" % world}

    # ---- the other BASELINE configs (every rank takes part), sharding check
    also = shard = None
    default_workload = args.mode == "log_prob" and args.width == 4 and not args.arch and args.tc < 0 and args.clean == "uniform"
    sm_mhz_now = 1965.0
    if rank == 0 and clocks and clocks.get("sm_mhz"):
        sm_mhz_now = float(clocks["sm_mhz"])
    if default_workload and not args.no_also and B >= 16384:
        also = run_also(args, nf, hps, ck, x, y, dev, rank, world, sm_mhz_now, row)
    if args.check_shard or (default_workload and not args.no_also and world > 1):
        shard = shard_check(nf, lib, dev, rank, world)
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
           "config": workload_config(args.mode, hps.arch, B, world, args.width, args.clean), "mean_nll_per_dim": mean_nll,
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "peak_source": peak_src, "kernel": "nf_scale_stream_kernel" if "unc" not in hps.arch else (("nf_chain_hyb_kernel" if (args.tc == 2 or (args.tc == 3 and args.mode == "sample")) else
                                                                                                                         ("nf_chain_wino_kernel" if (args.tc in (-1, 0, 4) or (args.tc == 3 and args.mode != "sample")) else "nf_chain_kernel"))
                                                                                                                        if args.width == 4 else "nf_wide_chain_kernel"),
                        "kernel_ms": kms, "alg_bytes_per_patch": alg,
                        "note": ("binding roof is the FP32 FMA pipe (see roofline_fp32), not HBM" if n_couplings else
                                 "scale-layer-only chain: streaming kernel, HBM-bound")},
           "roofline_fp32": {"bound": "fp32_fma", "achieved": B * conv_flop / (kms * 1e-3) / 1e12,
                             "peak": fp32_peak, "unit": "TFLOP/s",
                             "frac": B * conv_flop / (kms * 1e-3) / 1e12 / fp32_peak,
                             "peak_source": "148 SMs x 128 FMA/clk x 2 x median SM clock under load"},
           "e2e": e2e, "gpu_launches": world * args.steps * (2 if args.mode == "log_prob" else 1), "clocks": clocks}
    out["roofline"]["traffic_source"] = ("profiles/traffic.json: dram bytes of one `ncu --set full` capture of this kernel at this batch "
                                         "(a constant per build, not re-measured by this run)") if traffic is not None else None
    if args.width != 4 and args.mode != "train" and n_couplings:
        from noise_flow_b200.csrc_info import WIDE_TC_WIDTHS
        if args.width in WIDE_TC_WIDTHS:        # tensor-core kernels: the binding roof is the tensor pipe
            out["roofline_hbm"] = out["roofline"]
            out["roofline"] = tensor_roofline(B, args.width, n_couplings, kms)
    if also is not None:
        out["also"] = also
    if shard is not None:
        out["shard_check"] = shard
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
        v, cores, sample = cpu_oracle_rate(hps, ck, per_step=args.ref_patches)
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
