"""ctypes binding of the C-ABI in ``include/noiseflow_b200.h``.

There is deliberately NO fallback: if the shared library is missing or a call fails, a RuntimeError is
raised.  ctypes releases the GIL around every call, which preserves the reference's threading contract
(many Python threads driving one model: ``train_noise_flow.py:38-47``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnoiseflow_b200.so")

NF_OK = 0
c_float_p = C.POINTER(C.c_float)


class NfCouplingWeights(C.Structure):
    _fields_ = [(n, c_float_p) for n in ("l1_w", "l1_b", "bn1_mean", "bn1_var", "l2_w", "l2_b", "bn2_mean",
                                          "bn2_var", "last_w", "last_b", "last_logs")] + \
               [("rescaling_scale", C.c_float), ("bn_eps", C.c_float)]


class NfTrainOp(C.Structure):
    """``nf_train_op`` (include/noiseflow_b200.h): one op of the device-resident train program; offsets index the
    flat variable array, -1 = absent."""
    _fields_ = [("kind", C.c_int32), ("mix_kind", C.c_int32)] + \
               [(n, C.c_int32) for n in ("off_P", "off_L", "off_U", "off_logS", "off_signS")] + \
               [("perm", C.c_int32 * 4)] + \
               [(n, C.c_int32) for n in ("off_w1", "off_b1", "off_w2", "off_b2", "off_w3", "off_b3", "off_logs", "off_scale",
                                         "off_bn1_mean", "off_bn1_var", "off_bn2_mean", "off_bn2_var", "token",
                                         "off_beta1", "off_beta2", "off_gain_params", "off_cam_params", "off_gain_val")] + \
               [("c_i", C.c_float)]


# name -> (restype, argtypes); every symbol the header declares is listed here (tests check both ways)
SIGNATURES = {
    "nf_abi_version": (C.c_int, []),
    "nf_last_error": (C.c_char_p, []),
    "nf_device_info": (C.c_int, [C.POINTER(C.c_int)] * 4),
    "nf_model_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "nf_model_destroy": (C.c_int, [C.c_void_p]),
    "nf_model_add_conv1x1": (C.c_int, [C.c_void_p, c_float_p, c_float_p, C.c_float]),
    "nf_model_add_permute": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "nf_model_add_affine_coupling": (C.c_int, [C.c_void_p, C.POINTER(NfCouplingWeights)]),
    "nf_model_add_cond_coupling": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(NfCouplingWeights)]),
    "nf_model_set_cond_coupling": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(NfCouplingWeights)]),
    "nf_model_add_scale": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p, C.c_int]),
    "nf_model_finalize": (C.c_int, [C.c_void_p]),
    "nf_model_num_layers": (C.c_int, [C.c_void_p]),
    "nf_model_set_conv1x1": (C.c_int, [C.c_void_p, C.c_int, c_float_p, c_float_p, C.c_float]),
    "nf_model_set_affine_coupling": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(NfCouplingWeights)]),
    "nf_model_set_scale": (C.c_int, [C.c_void_p, C.c_int, c_float_p, C.c_int]),
    "nf_model_begin_update": (C.c_int, [C.c_void_p]),
    "nf_model_end_update": (C.c_int, [C.c_void_p]),
    "nf_model_set_launch": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "nf_model_set_bs_small": (C.c_int, [C.c_void_p, C.c_int]),
    "nf_model_set_tensor_cores": (C.c_int, [C.c_void_p, C.c_int]),
    "nf_log_prob": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nf_inverse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                             C.c_void_p, C.c_void_p, C.c_void_p]),
    "nf_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                             C.c_void_p, C.c_void_p, C.c_void_p]),
    "nf_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_void_p,
                            C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "nf_run_layers": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nf_chain_batch_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                                       C.c_float, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nf_grad_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "nf_train_workspace_floats": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "nf_loss_and_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nf_trainer_create": (C.c_int, [C.POINTER(NfTrainOp), C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                    C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_float, C.POINTER(C.c_void_p)]),
    "nf_trainer_destroy": (C.c_int, [C.c_void_p]),
    "nf_trainer_reduce_len": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "nf_trainer_loss_and_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int,
                                           C.c_void_p, C.c_void_p]),
    "nf_trainer_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                   C.c_int, C.c_void_p]),
    "nf_trainer_get_vars": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "nf_trainer_set_vars": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "nf_trainer_launches_per_step": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "nf_trainer_barriers_per_step": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "nf_probe_grid_barrier": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_float), C.c_void_p]),
    "nf_trainer_set_graph": (C.c_int, [C.c_void_p, C.c_int]),
    "nf_trainer_set_cta_warps": (C.c_int, [C.c_void_p, C.c_int]),
    "nf_trainer_set_fused": (C.c_int, [C.c_void_p, C.c_int]),
    "nf_reduce_sums": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "nf_baseline_nll": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int64, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "nf_histogram": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "nf_squeeze2d": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nf_unsqueeze2d": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nf_log_prob_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nf_sample_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_void_p,
                                 C.c_uint64, C.c_uint64, C.c_void_p]),
    "nf_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "nf_host_free": (C.c_int, [C.c_void_p]),
}

_lib = None


def load():
    """Load ``libnoiseflow_b200.so`` (once) and declare every prototype.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "%s is missing: the CUDA extension is not built (run `python -m noise_flow_b200.build`); "
            "there is no CPU fallback by design" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError here = header / library drift
        fn.restype = res
        fn.argtypes = args
    if lib.nf_abi_version() != 1:
        raise RuntimeError("libnoiseflow_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != NF_OK:
        msg = load().nf_last_error()
        raise RuntimeError("noiseflow_b200 %s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))
