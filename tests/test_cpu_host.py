"""CPU tests of the host logic: artefact readers, arch parsing / variable naming, LU + scale-table folding
against the oracle, the C-ABI surface (load + symbols + argument validation, no compute), and the
world_size-2 gloo path of the multi-GPU plumbing."""
import copy
import ctypes as C
import os
import re
import struct
import subprocess
import sys

import numpy as np
import pytest
import torch

from common import make_oracle, synth_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------ artefacts
def test_checkpoint_reader_on_shipped_bundle(golden_dir):
    from noise_flow_b200.tf_checkpoint import _mask_crc, crc32c, load_checkpoint, read_index
    prefix = os.path.join(golden_dir, "NoiseFlow", "ckpt", "model.ckpt.best")
    ents = read_index(prefix + ".index")
    ck = load_checkpoint(prefix)
    assert len(ck) == 143 and sum(v.size for v in ck.values()) == 2721
    assert os.path.getsize(prefix + ".data-00000-of-00001") == 2721 * 4
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    for name, e in ents.items():                                                # TF's own masked crc32c pins the bytes
        raw = data[e["offset"]:e["offset"] + e["size"]]
        assert e["crc32c"] == _mask_crc(crc32c(raw)), name
    assert ck["model/real_nvp_conv_template/l_last/W"].shape == (3, 3, 5, 4)
    assert ck["level0/bijector1/rescaling_scale0"].shape == ()
    assert abs(float(ck["model/sdn_gain/gain_val"][0]) - 1.0694) < 1e-4


def test_checkpoint_writer_roundtrip(tmp_path, shipped):
    from noise_flow_b200.tf_checkpoint import load_checkpoint, save_checkpoint
    _, ck = shipped
    save_checkpoint(str(tmp_path / "ckpt" / "m.ckpt"), ck)
    back = load_checkpoint(str(tmp_path / "ckpt" / "m.ckpt"))
    assert list(back) == sorted(ck) and all(np.array_equal(back[k], ck[k]) and back[k].shape == ck[k].shape for k in ck)


def test_reference_checkpoint_identical_to_golden_copy(golden_dir):
    ref = "/root/reference/models/NoiseFlow"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present on this machine")
    for rel in ("hps.txt", "ckpt/model.ckpt.best.index", "ckpt/model.ckpt.best.data-00000-of-00001"):
        assert open(os.path.join(ref, rel), "rb").read() == open(os.path.join(golden_dir, "NoiseFlow", rel), "rb").read()


def test_hps_loader_and_logger(tmp_path, golden_dir):
    from noise_flow_b200.hps import hps_loader, hps_logger
    hps = hps_loader(os.path.join(golden_dir, "NoiseFlow", "hps.txt"))
    assert hps.arch == "sdn5|unc|unc|unc|unc|gain4|unc|unc|unc|unc" and hps.width == 4 and hps.flow_permutation == 1
    assert hps.decomp == "LU" and hps.squeeze_factor == 1 and hps.n_levels == 1 and hps.gain_init == -5.0
    assert hps.pre_init is True and hps.init_sdn is False and hps.mb_qsize == "" and hps.lr == 0.0001
    c_i, b1, b2, gp, cp = hps.param_inits                      # always recomputed (NoiseFlowWrapper.py:125-137)
    assert (c_i, b1, b2) == (1.0, -5.0, 0.0) and gp.shape == (5,) and cp.shape == (3, 5) and (cp == 1).all()
    hps_logger(str(tmp_path / "hps.txt"), hps, ["a", "b"], 7)
    h2 = hps_loader(str(tmp_path / "hps.txt"))
    assert h2.arch == hps.arch and h2.width == 4


# ------------------------------------------------------------------------------------------ model assembly
def test_modelspec_names_scopes_and_counts(shipped):
    from noise_flow_b200.params import ModelSpec
    hps, ck = shipped
    ms = ModelSpec(copy.copy(hps), ck)
    ms.assign_template_scopes("inverse")
    ms.create_scale_variables()
    assert ms.get_layer_names() == ["sdn_0"] + sum((["Conv2d_1x1_%d" % i, "unc_%d" % i] for i in (1, 2, 3, 4)), []) + \
        ["gain_5"] + sum((["Conv2d_1x1_%d" % i, "unc_%d" % i] for i in (6, 7, 8, 9)), [])
    assert ms.store.created == [] and ms.store.num_trainable() == 2433
    assert set(ms.store.trainable) == set(ck)
    fresh = ModelSpec(copy.copy(hps), None, seed=1)
    fresh.assign_template_scopes("inverse")
    fresh.create_scale_variables()
    assert set(fresh.store.vars) == set(ck) and fresh.store.num_trainable() == 2433
    for k, v in ck.items():
        assert fresh.store.vars[k].shape == v.shape, k


def test_lu_and_scale_tables_match_oracle(shipped):
    from noise_flow_b200.params import ISO_VALS, ModelSpec
    from oracle import noise_flow_oracle as O
    hps, ck = shipped
    ms = ModelSpec(copy.copy(hps), ck)
    ms.assign_template_scopes("inverse")
    orc = make_oracle(hps, ck)
    y = torch.full((1, 1, 1, 1), 0.37, dtype=torch.float64)
    for l, b in zip(ms.layers, orc.model[0]):
        if l.kind == "conv1x1":
            a, ai, lad = ms.conv1x1_matrices(l)
            p = b.params()
            assert np.abs(a - p["A"].numpy()).max() < 1e-12 and np.abs(ai - p["A_inv"].numpy()).max() < 1e-12
            assert abs(lad - float(p["log_abs_det"])) < 1e-12
        elif l.kind == "scale":
            tab = ms.scale_table(l)
            assert tab.shape == (25, 2)
            for cam in range(5):
                for k, iso in enumerate(ISO_VALS):
                    s, _ = b._scale(y, None, None, [iso], [float(cam)])
                    row = tab[cam * 5 + k].astype(np.float64)
                    if l.token.startswith("sdn"):
                        assert abs(float(s.flatten()[0]) ** 2 - (row[0] * 0.37 + row[1])) < 1e-6 * (row[0] + row[1])
                    else:
                        assert abs(float(s.flatten()[0]) - row[0]) < 1e-6
    # SURVEY appendix B anchor: (S6, ISO 100) -> a = 0.022698, b = 1.07e-4
    sdn = ms.scale_table(ms.layers[0])[2 * 5 + 0]
    assert abs(sdn[0] - 0.022698) < 2e-6 and abs(sdn[1] - 1.07e-4) < 1e-6


@pytest.mark.parametrize("arch,perm,names", [
    ("unc|sdn2|gain", 0, ["permute", "unc_0", "sdn_1", "gain_2"]),
    ("sdn|bogus|unc", 2, ["sdn_0", "unc_2"]),
])
def test_arch_parsing_variants(arch, perm, names):
    from noise_flow_b200 import make_hps
    from noise_flow_b200.params import ModelSpec
    ms = ModelSpec(make_hps(arch=arch, flow_permutation=perm), None)
    assert ms.get_layer_names() == names


def test_unsupported_configs_raise():
    from noise_flow_b200 import make_hps
    from noise_flow_b200.params import ModelSpec
    with pytest.raises(NotImplementedError):
        ModelSpec(make_hps(squeeze_factor=2), None)
    with pytest.raises(NotImplementedError):
        ModelSpec(make_hps(n_levels=2), None)
    with pytest.raises(NotImplementedError):
        ModelSpec(make_hps(decomp="NONE"), None)


# ------------------------------------------------------------------------------------------ C-ABI surface
def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "noiseflow_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nf_[a-z0-9_]+)\s*\(", txt)))


def test_library_loads_and_exports_every_declared_symbol():
    from noise_flow_b200 import _lib
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "library does not export %s" % s
    assert sorted(_lib.SIGNATURES) == syms                 # the ctypes table covers the header exactly
    assert lib.nf_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (nf_[a-z0-9_]+)", out))
    assert exported == set(syms)                           # nothing undeclared leaks out of the library
    code = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in code


def test_c_abi_argument_validation_without_gpu():
    from noise_flow_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.nf_model_create(16, 16, 16, 4, C.byref(h)) == -2 and b"32x32x4" in lib.nf_last_error()
    assert lib.nf_model_create(32, 32, 4, 48, C.byref(h)) == -2       # widths: 4, 8, 16, 32, 64, 128, 256, 512
    assert lib.nf_model_create(32, 32, 4, 12, C.byref(h)) == -2              # wide kernel: widths 8 / 16 / 32
    hw = C.c_void_p()
    assert lib.nf_model_create(32, 32, 4, 32, C.byref(hw)) == 0 and hw.value and lib.nf_model_destroy(hw) == 0
    assert lib.nf_model_create(32, 32, 4, 4, C.byref(h)) == 0 and h.value
    perm = (C.c_int32 * 4)(0, 0, 1, 2)
    assert lib.nf_model_add_permute(h, perm) == -1
    tab = (C.c_float * 4)(1.0, 0.0, 2.0, 0.0)
    assert lib.nf_model_add_scale(h, 7, 1, tab, 2) == -1
    assert lib.nf_model_add_scale(h, 2, 1, tab, 2) == 0
    assert lib.nf_model_add_scale(h, 2, 1, tab, 99) == -1
    assert lib.nf_model_num_layers(h) == 1
    # launching before finalize is a state error, not a crash
    assert lib.nf_log_prob(h, 1, 1, None, 0, 1, 1, None, None, None) == -4
    assert lib.nf_model_set_launch(h, 17, 0) == -1 and lib.nf_model_set_launch(h, 8, 0) == 0
    assert lib.nf_model_destroy(h) == 0


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[/.]noise_flow_oracle", re.M)
    for root, _, files in os.walk(os.path.join(ROOT, "noise_flow_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                assert not pat.search(open(os.path.join(root, f)).read()), f


# ------------------------------------------------------------------------------------------ multi-GPU plumbing on gloo
def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from noise_flow_b200.distributed import global_means, shard_range
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    n = 1001
    nll = torch.arange(n, dtype=torch.float64) * 0.5 - 3.0
    sdz = torch.cos(torch.arange(n, dtype=torch.float64))
    lo, hi = shard_range(n, rank, world)
    sums = torch.stack([nll[lo:hi].sum(), sdz[lo:hi].sum(), torch.tensor(float(hi - lo), dtype=torch.float64)])
    m, s = global_means(sums)
    q.put((rank, lo, hi, float(m), float(s), float(nll.mean()), float(sdz.mean())))
    dist.destroy_process_group()


def test_sharded_mean_nll_world_size_2_gloo():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert res[0][1:3] == (0, 501) and res[1][1:3] == (501, 1001)
    for r in res:
        assert abs(r[3] - r[5]) < 1e-12 and abs(r[4] - r[6]) < 1e-12      # every rank holds the global means


def test_shipped_checkpoint_shows_the_bias_random_walk_of_the_reference(shipped):
    """`l_1/b` and `l_2/b` sit in front of a batch-statistics BatchNorm (layers.py:469-489): their exact gradient is zero
    and they are initialised to zero (layers.py:602), yet the reference's own trained checkpoint holds values of 0.01 - 0.5.
    TensorFlow's fp32 gradient is summation noise there and Adam normalises noise to steps of about +-lr: a random walk of
    1e-4 * sqrt(2000 epochs x 112 iterations) ~ 0.05 per thread.  The device trainer does the same (it does NOT zero these
    gradients structurally); tests/test_gpu_reference_goldens.py::test_two_adam_steps_match_reference_train_op bounds the
    walk by lr per step."""
    _, ck = shipped
    b = np.concatenate([v.reshape(-1) for k, v in ck.items() if k.endswith("/l_1/b") or k.endswith("/l_2/b")])
    assert b.size == 64 and np.all(b != 0.0)
    assert 0.05 < float(np.sqrt(np.mean(b * b))) < 0.5


def test_lu_chain_rule_closed_form_equals_autograd(shipped):
    """train.lu_chain (closed form of d loss / d (L_vec, U_vec, log_S) for A = P L U, matrix_param.py:117-130)
    against torch autograd through the same parameterisation, on the shipped LU variables."""
    from noise_flow_b200.train import lu_chain, lu_chain_autograd
    _, ck = shipped
    rs = np.random.RandomState(5)
    for i in (1, 2, 3, 4, 6, 7, 8, 9):
        dA = rs.randn(4, 4)
        a = lu_chain(ck, "level0/bijector%d/Conv2d_1x1_%d" % (i, i), "conv2d_1x1_%d_0" % i, dA)
        b = lu_chain_autograd(ck, "level0/bijector%d/Conv2d_1x1_%d" % (i, i), "conv2d_1x1_%d_0" % i, dA)
        assert set(a) == set(b)
        for k in a:
            assert a[k].shape == ck[k].shape and np.abs(a[k] - b[k]).max() < 1e-12


def test_device_train_program_layout(shipped):
    """Host half of the device-resident train step: flat variable layout and the op list of ``nf_trainer_create``."""
    from noise_flow_b200 import make_hps
    from noise_flow_b200.params import ModelSpec
    from noise_flow_b200.train import build_train_program, tri_positions
    hps, ck = shipped
    spec = ModelSpec(hps, ck)
    spec.assign_template_scopes("inverse")
    spec.create_scale_variables()
    names, off, n_vars, mask, ops = build_train_program(spec)
    assert n_vars == 2721 and int(mask.sum()) == 2433                       # checkpoint floats / hps.txt num_params
    assert [op.kind for op in ops] == [2, 1, 1, 1, 1, 2, 1, 1, 1, 1]        # sdn5 | 4 x (1x1 + unc) | gain4 | 4 x ...
    assert [op.token for op in ops if op.kind == 2] == [5, 14]
    assert all(op.mix_kind == 1 for op in ops if op.kind == 1)
    cp = ops[1]
    assert off["model/real_nvp_conv_template/l_1/W"] == cp.off_w1 and off["level0/bijector1/rescaling_scale0"] == cp.off_scale
    assert not mask[cp.off_bn1_mean] and not mask[cp.off_P] and mask[cp.off_L] and mask[cp.off_logS]
    # every offset range is disjoint from every other (no variable is addressed twice by accident)
    seen = np.zeros(n_vars, dtype=np.int32)
    for op in ops:
        if op.kind == 1:
            for f, sz in (("off_w1", 72), ("off_b1", 4), ("off_w2", 16), ("off_b2", 4), ("off_w3", 180), ("off_b3", 4),
                          ("off_logs", 4), ("off_scale", 1), ("off_bn1_mean", 4), ("off_bn1_var", 4), ("off_bn2_mean", 4),
                          ("off_bn2_var", 4), ("off_P", 16), ("off_L", 6), ("off_U", 6), ("off_logS", 4), ("off_signS", 4)):
                o = getattr(op, f)
                assert o >= 0
                seen[o:o + sz] += 1
    assert seen.max() == 1
    # triangle positions = where stricttri2vec reads: SURVEY 8(a7) ordering
    assert tri_positions(False) == [14, 13, 12, 4, 9, 8] and tri_positions(True) == [1, 2, 3, 11, 6, 7]
    # unsupported scale tokens are refused, not silently mis-trained
    spec2 = ModelSpec(make_hps(arch="sdn2|unc|gain2"), None)
    spec2.assign_template_scopes("inverse")
    spec2.create_scale_variables()
    with pytest.raises(NotImplementedError):
        build_train_program(spec2)
    spec3 = ModelSpec(make_hps(arch="sdn4|unc|gain4", flow_permutation=0), None)
    spec3.assign_template_scopes("inverse")
    spec3.create_scale_variables()
    ops3 = build_train_program(spec3)[4]
    assert [op.kind for op in ops3] == [2, 1, 2] and ops3[1].mix_kind == 2 and list(ops3[1].perm) == [3, 2, 1, 0]


def test_result_logger_writes_the_reference_tsv(tmp_path):
    """borealisflows/utils.py:89-107 + the column sets of train_noise_flow.py:336-348: header without trailing newline, every
    row prefixed by one, str.format of the values, append mode skips the header, a missing column raises KeyError."""
    from noise_flow_b200.utils import SAMPLE_COLUMNS, TEST_COLUMNS, TRAIN_COLUMNS, ResultLogger
    assert TRAIN_COLUMNS == ["epoch", "NLL", "NLL_G", "NLL_SDN", "sdz", "train_time"]
    assert TEST_COLUMNS[-1] == "msg" and SAMPLE_COLUMNS[-4:] == ["KLD_G", "KLD_NLF", "KLD_NF", "KLD_R"]
    p = str(tmp_path / "train.txt")
    lg = ResultLogger(p, TRAIN_COLUMNS)
    lg.log({"epoch": 1, "NLL": -2.5, "NLL_G": np.float32(-2.25), "NLL_SDN": -2.4, "sdz": 1.0, "train_time": 12})
    del lg
    lg = ResultLogger(p, TRAIN_COLUMNS, append=True)
    lg.log({"epoch": 2, "NLL": -2.75, "NLL_G": -2.25, "NLL_SDN": -2.4, "sdz": 0.5, "train_time": 11, "extra": "ignored"})
    with pytest.raises(KeyError):
        lg.log({"epoch": 3})
    del lg
    assert open(p).read() == ("epoch\tNLL\tNLL_G\tNLL_SDN\tsdz\ttrain_time\n1\t-2.5\t-2.25\t-2.4\t1.0\t12\n"
                              "2\t-2.75\t-2.25\t-2.4\t0.5\t11")       # rows START with the newline
    ref = "/root/reference/borealisflows/utils.py"
    if os.path.exists(ref):      # the reference's own class (its module imports matplotlib + tensorflow: executed over the stand-ins)
        import importlib, sys
        from oracle import tf1_shim
        tf1_shim.install()
        sys.path.insert(0, "/root/reference")
        RefLogger = importlib.import_module("borealisflows.utils").ResultLogger
        q = str(tmp_path / "ref.txt")
        rl = RefLogger(q, TRAIN_COLUMNS)
        rl.log({"epoch": 1, "NLL": -2.5, "NLL_G": np.float32(-2.25), "NLL_SDN": -2.4, "sdz": 1.0, "train_time": 12})
        del rl
        rl = RefLogger(q, TRAIN_COLUMNS, append=True)
        rl.log({"epoch": 2, "NLL": -2.75, "NLL_G": -2.25, "NLL_SDN": -2.4, "sdz": 0.5, "train_time": 11})
        del rl
        assert open(q).read() == open(p).read()


@pytest.mark.skipif(not os.path.exists("/root/reference/borealisflows/utils.py"), reason="needs the reference checkout")
def test_hps_logger_bytes_equal_the_reference_function(tmp_path, golden_dir):
    """hps.txt written by ours and by the reference's `hps_logger` (borealisflows/utils.py:110-119, executed over the
    import stand-ins) from the same Hps object are byte-identical, and the reference's `hps_loader` reads ours back."""
    import importlib
    from types import SimpleNamespace
    from noise_flow_b200.hps import hps_loader, hps_logger
    from oracle import tf1_shim
    tf1_shim.install()
    sys.path.insert(0, "/root/reference")
    ref_utils = importlib.import_module("borealisflows.utils")
    hps = hps_loader(os.path.join(golden_dir, "NoiseFlow", "hps.txt"))
    plain = SimpleNamespace(**{k: v for k, v in vars(hps).items() if k != "param_inits"})
    names = ["sdn_0", "Conv2d_1x1_1", "unc_1"]
    a, b = str(tmp_path / "ours.txt"), str(tmp_path / "ref.txt")
    hps_logger(a, plain, names, 2433)
    ref_utils.hps_logger(b, plain, names, 2433)
    assert open(a, "rb").read() == open(b, "rb").read()
    back = ref_utils.hps_loader(a)
    assert back.arch == hps.arch and back.width == "4" and back.flow_permutation == "1"      # utils.hps_loader keeps strings
