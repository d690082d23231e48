"""Train-step support for the reference's driver (``train_noise_flow.py:187-198,50-77``): gradient of the batch-mean
NLL with respect to every trainable TF variable, Adam with TensorFlow's update rule, and the data-parallel
reduction.

The device does the heavy part (``nf_loss_and_grad``: forward with batch-statistics BatchNorm, three backward passes
per coupling, fp64 gradient accumulation).  What is left for the host is the chain rule through the two tiny
parameterisations that live on the host anyway -- the 4x4 LU construction (``matrix_param.py:117-138``) and the
per-(camera, ISO) scale scalars (``cond_utils.py``) -- done here with torch autograd on a handful of scalars.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .params import GAIN_TOKENS, ISO_VALS, SDN_TOKENS, stricttri2vec, vec2stricttri

_T64 = torch.float64


# ------------------------------------------------------------------------------------------------ host chain rules
def _tri_index(n: int, upper: bool) -> np.ndarray:
    m = n * (n - 1) // 2
    return vec2stricttri(np.arange(1, m + 1, dtype=np.float64), upper).astype(np.int64)


def lu_chain(v: Dict[str, np.ndarray], vscope: str, pname: str, dA: np.ndarray) -> Dict[str, np.ndarray]:
    """d loss / d (L_vec, U_vec, log_S) from d loss / d A for ``A = P L U`` (matrix_param.py:117-130), closed form:
    with ``U' = U + diag(sign_S exp(log_S))``:  ``dL = strict_lower(P^T G U'^T)``, ``dU' = L^T P^T G``,
    ``dU = strict_upper(dU')``, ``dlog_S = diag(dU') * sign_S * exp(log_S)``; the 6-vectors take the entries of the
    matrix gradients at their own positions (``stricttri2vec`` is a permutation)."""
    names = {k: "%s/%s_matpar_lu_%s" % (vscope, k, pname) for k in ("P", "L_vec", "U_vec", "log_S", "sign_S")}
    p = np.asarray(v[names["P"]], np.float64)
    n = p.shape[0]
    s_diag = np.asarray(v[names["sign_S"]], np.float64) * np.exp(np.asarray(v[names["log_S"]], np.float64))
    l = vec2stricttri(np.asarray(v[names["L_vec"]], np.float64), upper=False) + np.eye(n)
    u = vec2stricttri(np.asarray(v[names["U_vec"]], np.float64), upper=True) + np.diag(s_diag)
    ptg = p.T @ np.asarray(dA, np.float64)
    d_l = ptg @ u.T
    d_u = l.T @ ptg
    return {names["L_vec"]: stricttri2vec(np.tril(d_l, -1), upper=False),
            names["U_vec"]: stricttri2vec(np.triu(d_u, 1), upper=True),
            names["log_S"]: np.diag(d_u) * s_diag}


def lu_chain_autograd(v: Dict[str, np.ndarray], vscope: str, pname: str, dA: np.ndarray) -> Dict[str, np.ndarray]:
    """Same as :func:`lu_chain` through torch autograd (cross-check used by the tests)."""
    names = {k: "%s/%s_matpar_lu_%s" % (vscope, k, pname) for k in ("P", "L_vec", "U_vec", "log_S", "sign_S")}
    p = torch.as_tensor(v[names["P"]], dtype=_T64)
    sign_s = torch.as_tensor(v[names["sign_S"]], dtype=_T64)
    l_vec = torch.as_tensor(v[names["L_vec"]], dtype=_T64).requires_grad_(True)
    u_vec = torch.as_tensor(v[names["U_vec"]], dtype=_T64).requires_grad_(True)
    log_s = torch.as_tensor(v[names["log_S"]], dtype=_T64).requires_grad_(True)
    n = p.shape[0]
    li, ui = torch.as_tensor(_tri_index(n, False)), torch.as_tensor(_tri_index(n, True))
    l = torch.cat([l_vec.new_zeros(1), l_vec])[li] + torch.eye(n, dtype=_T64)
    u = torch.cat([u_vec.new_zeros(1), u_vec])[ui] + torch.diag(sign_s * torch.exp(log_s))
    a = p @ (l @ u)
    (a * torch.as_tensor(dA, dtype=_T64)).sum().backward()
    return {names["L_vec"]: l_vec.grad.numpy(), names["U_vec"]: u_vec.grad.numpy(), names["log_S"]: log_s.grad.numpy()}


def _scale_row_torch(token, tv, hps, cam, iso):
    """Differentiable twin of ``params.scale_row`` for the trainable tokens: returns (a, b) or (g, None)."""
    gain_init = float(getattr(hps, "gain_init", 0.0))
    sig = torch.sigmoid

    def ladder(fmt):
        key = int(iso) if float(iso) in ISO_VALS else 800
        return tv["model/" + fmt % key][0]

    def onehot(gp):
        for k, val in enumerate(ISO_VALS):
            if float(iso) == val:
                return gp[k]
        return gp.new_zeros(())

    if token == "sdn":
        return sig(tv["model/b1"][0]), sig(tv["model/b2"][0])
    if token == "sdn1":
        r_gain = torch.exp(1e-2 * ladder("r_gain_param_%05d")) * iso
        return sig(tv["model/b1"][0]) / r_gain, sig(tv["model/b2"][0])
    if token in ("sdn2", "sdn3"):
        gain = torch.exp(1e-1 * ladder("gain_param_%05d")) * iso
        b1, b2 = sig(tv["model/b1"][0]), sig(tv["model/b2"][0])
        return (b1, gain * b2) if token == "sdn2" else (gain * b1, gain * gain * b2)
    if token == "sdn4":
        s = "model/sdn_gain"
        gain = torch.exp(onehot(tv[s + "/gain_params"])) * iso
        return torch.exp(tv[s + "/beta1"][0]) / gain, torch.exp(tv[s + "/beta2"][0])
    if token in ("sdn5", "sdn6"):
        c_i = hps.param_inits[0]
        s = "model/sdn_gain"
        ocp = torch.exp(c_i * tv[s + "/cam_params"][:, int(cam)])
        gsel = onehot(tv[s + "/gain_params"])
        beta1, beta2 = tv[s + "/beta1"][0], tv[s + "/beta2"][0]
        if token == "sdn5":
            gain = torch.exp(c_i * gsel * ocp[2]) * iso
            return torch.exp(c_i * beta1 * ocp[0]) / gain, torch.exp(c_i * beta2 * ocp[1])
        gain = torch.exp(c_i * gsel * ocp[0]) * iso
        return torch.exp(c_i * beta1) / gain, torch.exp(c_i * beta2)
    if token == "gain":
        return sig(tv["model/g1"][0]) * iso + sig(tv["model/g2"][0]), None
    if token == "gain1":
        return torch.exp(1e-5 * tv["model/g1"][0]) * iso + torch.exp(1e-5 * tv["model/g2"][0]), None
    if token == "gain2":
        return torch.exp(1e-1 * ladder("gain_param_%05d")) * iso, None
    if token == "gain3":
        return torch.exp(1e-5 * ladder("gain_param_%05d")), None
    if token == "gain4":
        return tv["model/sdn_gain/gain_val"][0], None
    return None, None      # camsdn: no trainable parameters


def scale_chain(spec, layer, dtable: np.ndarray, extra_rows) -> Dict[str, np.ndarray]:
    """d loss / d (scale-layer variables) from d loss / d (table rows)."""
    rows = [(float(cam), iso) for cam in range(5) for iso in ISO_VALS] + [(r[0], r[1]) for r in (extra_rows or [])]
    used = [k for k in range(min(len(rows), dtable.shape[0])) if dtable[k].any()]
    if not used or layer.token == "camsdn":
        return {}
    names = [k for k in spec.store.vars if k.startswith("model/") and "real_nvp_conv_template" not in k]
    tv = {k: torch.as_tensor(spec.store.vars[k], dtype=_T64).clone().requires_grad_(True) for k in names}
    total = None
    for k in used:
        a, b = _scale_row_torch(layer.token, tv, spec.hps, rows[k][0], rows[k][1])
        if a is None:
            continue
        term = a * float(dtable[k, 0])
        if b is not None:
            term = term + b * float(dtable[k, 1])
        total = term if total is None else total + term
    if total is None or not total.requires_grad:
        return {}
    total.backward()
    return {k: t.grad.numpy() for k, t in tv.items() if t.grad is not None}


# ------------------------------------------------------------------------------------------------ loss + gradients
TIMINGS = None   # set to a dict to accumulate host wall-clock seconds per phase of a step (bench.py --mode train)
_t_last = [0.0]


def _tick(name):
    if TIMINGS is not None:
        import time
        now = time.perf_counter()
        if name is not None:
            TIMINGS[name] = TIMINGS.get(name, 0.0) + now - _t_last[0]
        _t_last[0] = now


def loss_and_grad(nf, x, y, nlf0=None, nlf1=None, iso=None, cam=None, is_training=True, refold=True, bn_update=True):
    """``(loss, sd_z, grads)``: ``loss = mean_n nll_n`` (``NoiseFlow.loss``) and ``grads[tf_variable_name]`` (float64
    numpy, same shapes as the variables) for every trainable variable.  With ``is_training`` (the reference's train
    thread, ``train_noise_flow.py:64-71``) BatchNorm runs on batch statistics and the moving statistics are updated."""
    nf.build("inverse")
    nf._fresh()          # the engine's moving statistics feed the is_training=False gradient
    _tick(None)
    eng, spec = nf._engine, nf.spec
    lib = eng.lib
    x = nf._dev(x, "x")
    cond = getattr(nf.hps, "sidd_cond", "mix")
    yy = nf._dev(y, "y") if (cond is not None and cond != "uncond") else None
    n = x.shape[0]
    if n == 0:
        raise ValueError("empty batch")
    rows, drow = nf._rows(n, nlf0, nlf1, iso, cam)
    n_layers = len(spec.layers)
    offs = (C.c_int64 * (n_layers + 1))()
    _lib.check(lib.nf_grad_layout(eng.handle, offs), "nf_grad_layout")
    nws = C.c_int64()
    _lib.check(lib.nf_train_workspace_floats(eng.handle, n, C.byref(nws)), "nf_train_workspace_floats")
    ws = torch.empty(nws.value, device=nf.device, dtype=torch.float32)
    dscr = torch.zeros(512, device=nf.device, dtype=torch.float64)
    flat = np.zeros(max(int(offs[n_layers]), 1), dtype=np.float64)
    cps = [l for l in spec.layers if l.kind == "coupling"]
    wd = int(spec.width)
    bstats = np.zeros((max(len(cps), 1), 4 * wd), dtype=np.float32)       # per coupling: mean1, var1, mean2, var2 ([W] each)
    sums = (C.c_double * 3)()
    p = lambda t: t.data_ptr() if t is not None else None
    _tick("setup")
    with torch.cuda.device(nf.device):
        _lib.check(lib.nf_loss_and_grad(eng.handle, x.data_ptr(), p(yy), p(rows), drow, n, 1 if is_training else 0,
                                        ws.data_ptr(), dscr.data_ptr(), flat.ctypes.data_as(C.c_void_p),
                                        bstats.ctypes.data_as(C.c_void_p), sums,
                                        int(torch.cuda.current_stream(nf.device).cuda_stream)), "nf_loss_and_grad")
    _tick("nf_loss_and_grad")
    grads: Dict[str, np.ndarray] = {}
    v = spec.store.vars

    def add(name, g):
        g = np.asarray(g, dtype=np.float64).reshape(v[name].shape)
        grads[name] = grads[name] + g if name in grads else g

    for idx, l in enumerate(spec.layers):
        blk = flat[offs[idx]:offs[idx + 1]]
        if l.kind == "conv1x1":
            for k, g in lu_chain(v, l.data["vscope"], l.data["pname"], blk.reshape(4, 4)).items():
                add(k, g)
            # the layer's own log-det, H*W*sum(log_S) per patch (layers.py:129-130), enters the loss as -ldj
            add("%s/log_S_matpar_lu_%s" % (l.data["vscope"], l.data["pname"]), np.full(4, -1024.0))
        elif l.kind == "coupling":
            s = l.data["template"]
            o = 0
            for name, size in (("/l_1/W", 18 * wd), ("/l_1/b", wd), ("/l_2/W", wd * wd), ("/l_2/b", wd),
                               ("/l_last/W", 36 * (wd + 1)), ("/l_last/b", 4), ("/l_last/logs", 4)):
                add(s + name, blk[o:o + size])
                o += size
            add(l.scope + "/rescaling_scale0", blk[o])
        elif l.kind == "scale":
            for k, g in scale_chain(spec, l, blk.reshape(-1, 2), nf._extra_rows).items():
                add(k, g)
    for name, trainable in spec.store.trainable.items():      # variables the loss does not depend on
        if trainable and name not in grads:
            grads[name] = np.zeros(v[name].shape, dtype=np.float64)
    _tick("chain_rules")
    if is_training and bn_update:
        nf._apply_bn_moving_update(bstats, refold=refold)    # refold=False: the caller re-folds after its optimizer step
    elif is_training:
        nf.last_batch_stats = bstats                         # bn_update=False: the caller averages them over ranks first
    _tick("bn_update_refold")
    loss = sums[0] / n
    sd_z = sums[1] / n
    return loss, sd_z, grads


# ------------------------------------------------------------------------------------------------ Adam (TF semantics)
class AdamOptimizer:
    """``tf.train.AdamOptimizer(learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8)`` (train_noise_flow.py:191-194):
    ``lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)``; ``m = b1 m + (1-b1) g``; ``v = b2 v + (1-b2) g^2``;
    ``var -= lr_t * m / (sqrt(v) + eps)``."""

    def __init__(self, learning_rate=1e-4, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
        self.t = 0
        self.m: Dict[str, np.ndarray] = {}
        self.v: Dict[str, np.ndarray] = {}

    def apply_gradients(self, variables: Dict[str, np.ndarray], grads: Dict[str, np.ndarray]):
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for k, g in grads.items():
            g = np.asarray(g, dtype=np.float64)
            m = self.m.setdefault(k, np.zeros_like(g))
            vv = self.v.setdefault(k, np.zeros_like(g))
            m += (1.0 - self.b1) * (g - m)
            vv += (1.0 - self.b2) * (g * g - vv)
            variables[k] = (variables[k].astype(np.float64) - lr_t * m / (np.sqrt(vv) + self.eps)).astype(np.float32)


def train_step(nf, optimizer: AdamOptimizer, x, y, nlf0=None, nlf1=None, iso=None, cam=None, group=None):
    """One ``sess.run([train_op, loss, sd_z], is_training=True)`` (train_noise_flow.py:50-77) on this rank's shard.
    With ``torch.distributed`` initialised, gradients and ``[sum nll, sum sd_z, n]`` are summed over ranks with ONE
    all-reduce (gradients are then divided by the world size: every rank's loss is the mean over ITS shard, and
    BatchNorm NORMALISES with per-rank statistics, which is the reference's per-``sess.run`` semantics).  The batch
    statistics ride in the same buffer: the moving averages move towards their rank average, exactly as
    ``DeviceTrainer.step`` does, so replicas -- and checkpoints written by any rank -- stay identical."""
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    loss, sd_z, grads = loss_and_grad(nf, x, y, nlf0, nlf1, iso, cam, is_training=True, refold=False, bn_update=not multi)
    names = sorted(grads)
    if multi:
        world = dist.get_world_size(group)
        n = int(np.asarray(x).shape[0]) if not isinstance(x, torch.Tensor) else x.shape[0]
        bstats = np.asarray(nf.last_batch_stats, dtype=np.float64)
        flat = np.concatenate([grads[k].reshape(-1) for k in names] + [np.array([loss * n, sd_z * n, float(n)]), bstats.reshape(-1)])
        t = torch.as_tensor(flat, dtype=torch.float64, device=nf.device if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        flat = t.cpu().numpy()
        o = 0
        for k in names:
            sz = grads[k].size
            grads[k] = flat[o:o + sz].reshape(grads[k].shape) / world
            o += sz
        loss, sd_z = flat[o] / flat[o + 2], flat[o + 1] / flat[o + 2]
        nf._apply_bn_moving_update((flat[o + 3:o + 3 + bstats.size].reshape(bstats.shape) / world).astype(np.float32), refold=False)
    trainable = {k: g for k, g in grads.items() if nf.spec.store.trainable.get(k, False)}
    _tick("allreduce")
    optimizer.apply_gradients(nf.spec.store.vars, trainable)
    _tick("adam")
    nf.refresh_parameters()
    _tick("refold")
    return loss, sd_z


# ------------------------------------------------------------------------------------------------ device-resident step
_DEVICE_TOKENS = {"sdn4": 4, "sdn5": 5, "sdn6": 6, "gain4": 14}          # nf_train_token


def tri_positions(upper: bool):
    """Flat position ``r*4+c`` of ``L_vec[k]`` / ``U_vec[k]`` in the strict triangle (matrix_param.py:31-57)."""
    pos = vec2stricttri(np.arange(1, 7, dtype=np.float64), upper).astype(np.int64)
    out = [0] * 6
    for r in range(4):
        for c in range(4):
            if pos[r, c] > 0:
                out[pos[r, c] - 1] = r * 4 + c
    return out


def build_train_program(spec):
    """``(names, offsets, n_vars, trainable_mask, ops)`` for ``nf_trainer_create``: the flat variable layout (every
    variable of the store, in store order) and one ``nf_train_op`` per kernel op in data -> latent order.  Host-only
    (no GPU needed).  Raises ``NotImplementedError`` for what the device-side chain rules do not cover."""
    store = spec.store
    names = list(store.vars.keys())
    o, off = {}, 0
    for k in names:
        o[k] = off
        off += int(store.vars[k].size)
    n_vars = off
    mask = np.zeros(n_vars, dtype=np.uint8)
    for k in names:
        if store.trainable.get(k, False):
            mask[o[k]:o[k] + store.vars[k].size] = 1
    ops, layers, i = [], spec.layers, 0
    while i < len(layers):
        l = layers[i]
        op = _lib.NfTrainOp()
        for f, _ in _lib.NfTrainOp._fields_:
            if f.startswith("off_"):
                setattr(op, f, -1)
        if l.kind in ("conv1x1", "permute"):
            if i + 1 >= len(layers) or layers[i + 1].kind != "coupling":
                raise NotImplementedError("stand-alone 1x1 conv / permutation (not followed by a coupling)")
            if l.kind == "conv1x1":
                op.mix_kind = 1
                s, p = l.data["vscope"], l.data["pname"]
                for f, k in (("off_P", "P"), ("off_L", "L_vec"), ("off_U", "U_vec"), ("off_logS", "log_S"), ("off_signS", "sign_S")):
                    setattr(op, f, o["%s/%s_matpar_lu_%s" % (s, k, p)])
            else:
                op.mix_kind = 2
                # nf_model_add_permute: forward y[i] = x[perm[i]], so data -> latent sends channel i to perm[i]
                op.perm = (C.c_int32 * 4)(*l.data["perm"])
            i += 1
            l = layers[i]
        if l.kind == "coupling":
            if spec.width != 4:
                raise NotImplementedError("train kernels are built for width 4")
            op.kind = 1
            t = l.data["template"]
            for f, k in (("off_w1", "/l_1/W"), ("off_b1", "/l_1/b"), ("off_w2", "/l_2/W"), ("off_b2", "/l_2/b"),
                         ("off_w3", "/l_last/W"), ("off_b3", "/l_last/b"), ("off_logs", "/l_last/logs"),
                         ("off_bn1_mean", "/bn_nvp_conv_1/mean"), ("off_bn1_var", "/bn_nvp_conv_1/var"),
                         ("off_bn2_mean", "/bn_nvp_conv_2/mean"), ("off_bn2_var", "/bn_nvp_conv_2/var")):
                setattr(op, f, o[t + k])
            op.off_scale = o[l.scope + "/rescaling_scale0"]
        elif l.kind == "scale":
            if l.token not in _DEVICE_TOKENS:
                raise NotImplementedError("scale token %r has no device-side chain rule (use train_step)" % l.token)
            op.kind = 2
            op.token = _DEVICE_TOKENS[l.token]
            s = "model/sdn_gain"
            if l.token == "gain4":
                op.off_gain_val = o[s + "/gain_val"]
            else:
                op.off_beta1, op.off_beta2, op.off_gain_params = o[s + "/beta1"], o[s + "/beta2"], o[s + "/gain_params"]
                if l.token in ("sdn5", "sdn6"):
                    op.off_cam_params = o[s + "/cam_params"]
                    op.c_i = float(spec.hps.param_inits[0])
                if l.token == "sdn6" and tuple(spec.store.vars[s + "/cam_params"].shape) != (1, 5):
                    raise NotImplementedError("sdn6 next to sdn5 (shared cam_params of another shape)")
        else:
            raise NotImplementedError(l.kind)
        ops.append(op)
        i += 1
    return names, o, n_vars, mask, ops


class DeviceTrainer:
    """(Coupling-net widths 8 / 16 / 32: the constructor returns a :class:`WideTrainer` -- same interface on the
    host-synchronous kernels.)

    The same train step with NOTHING on the host: every TF variable, the Adam slots and the step counter live in
    device memory inside an ``nf_trainer`` (include/noiseflow_b200.h); LU assembly, scale tables, batch-statistics
    BatchNorm, backward, chain rules, Adam and the BatchNorm moving averages are kernels on one stream, with no
    stream synchronisation inside a step.  ``step`` is one ``sess.run([train_op, loss, sd_z])``
    (train_noise_flow.py:64-71); data parallel = ONE all-reduce of the trainer's reduce buffer (gradients,
    ``[sum nll, sum sd_z, n]`` and the batch statistics that drive the moving averages).

    Supported: arch strings made of ``unc`` (with ``flow_permutation`` 0 or 1) and the ``sdn4 / sdn5 / sdn6 / gain4``
    scale layers, standard (camera, ISO) rows -- i.e. every configuration in the reference's ``job_noise_flow.sh``.
    Anything else raises ``NotImplementedError``; :func:`train_step` (host-side chain rules) covers those.

    ``sync_to_model()`` copies the trained variables back into ``nf.variables`` and re-folds the inference engine.
    """

    def __new__(cls, nf, *args, **kw):
        if cls is DeviceTrainer and int(nf.spec.width) != 4:
            return object.__new__(WideTrainer)
        return object.__new__(cls)

    def __init__(self, nf, learning_rate=1e-4, beta1=0.9, beta2=0.999, epsilon=1e-8, max_batch=256, group=None,
                 cuda_graph=True, cta_warps=0, fused=True):
        from .params import BN_EPS
        nf.build("inverse")
        self.nf, self.group = nf, group
        self.lr, self.b1, self.b2, self.eps = float(learning_rate), float(beta1), float(beta2), float(epsilon)
        spec = nf.spec
        self.lib = _lib.load()
        self.names, self.offsets, self.n_vars, mask, self.ops = build_train_program(spec)
        arr = (_lib.NfTrainOp * len(self.ops))(*self.ops)
        tri_lo = (C.c_int32 * 6)(*tri_positions(False))
        tri_up = (C.c_int32 * 6)(*tri_positions(True))
        flat = self._flatten()
        self.handle = C.c_void_p()
        self.max_batch = int(max_batch)
        with torch.cuda.device(nf.device):
            _lib.check(self.lib.nf_trainer_create(arr, len(self.ops), flat.ctypes.data_as(C.c_void_p),
                                                  mask.ctypes.data_as(C.c_void_p), self.n_vars, self.max_batch, tri_lo, tri_up,
                                                  BN_EPS, C.byref(self.handle)), "nf_trainer_create")
        n = C.c_int64()
        _lib.check(self.lib.nf_trainer_reduce_len(self.handle, C.byref(n)), "nf_trainer_reduce_len")
        self.red = torch.zeros(n.value, device=nf.device, dtype=torch.float64)
        self.steps = 0
        _lib.check(self.lib.nf_trainer_set_graph(self.handle, 1 if cuda_graph else 0), "nf_trainer_set_graph")
        _lib.check(self.lib.nf_trainer_set_cta_warps(self.handle, int(cta_warps)), "nf_trainer_set_cta_warps")
        _lib.check(self.lib.nf_trainer_set_fused(self.handle, 1 if fused else 0), "nf_trainer_set_fused")

    def _flatten(self):
        v = self.nf.spec.store.vars
        return np.concatenate([np.asarray(v[k], dtype=np.float32).reshape(-1) for k in self.names]) if self.names else np.zeros(0, np.float32)

    # ---- one step
    def _stream(self):
        return int(torch.cuda.current_stream(self.nf.device).cuda_stream)

    def loss_and_grad(self, x, y, nlf0=None, nlf1=None, iso=None, cam=None, is_training=True):
        """Enqueue loss + gradients of this rank's batch into ``self.red`` (no synchronisation)."""
        nf = self.nf
        x = nf._dev(x, "x")
        cond = getattr(nf.hps, "sidd_cond", "mix")
        yy = nf._dev(y, "y") if (cond is not None and cond != "uncond") else None
        n = x.shape[0]
        if n < 1 or n > self.max_batch:
            raise ValueError("batch size %d outside 1..max_batch=%d" % (n, self.max_batch))
        rows, drow = nf._rows(n, nlf0, nlf1, iso, cam)
        if drow >= 25 or (rows is not None and int(rows.max()) >= 25):
            raise NotImplementedError("non-standard (camera, ISO) conditioning row (use train_step)")
        self._keep = (x, yy, rows)       # keep the inputs alive until the stream has consumed them
        with torch.cuda.device(nf.device):
            _lib.check(self.lib.nf_trainer_loss_and_grad(self.handle, x.data_ptr(), yy.data_ptr() if yy is not None else None,
                                                         rows.data_ptr() if rows is not None else None, drow, n,
                                                         1 if is_training else 0, self.red.data_ptr(), self._stream()),
                       "nf_trainer_loss_and_grad")
        return n

    def step(self, x, y, nlf0=None, nlf1=None, iso=None, cam=None, sync=True):
        """One Adam step.  Returns ``(loss, sd_z)`` as Python floats (``sync=True``: one 24-byte read-back, the
        only synchronisation) or as a device tensor ``[sum nll, sum sd_z, n]`` (``sync=False``)."""
        import torch.distributed as dist
        self.loss_and_grad(x, y, nlf0, nlf1, iso, cam, is_training=True)
        world = 1
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            world = dist.get_world_size(self.group)
            dist.all_reduce(self.red, op=dist.ReduceOp.SUM, group=self.group)
        with torch.cuda.device(self.nf.device):
            _lib.check(self.lib.nf_trainer_apply(self.handle, self.red.data_ptr(), self.lr, self.b1, self.b2, self.eps, world, 1,
                                                 self._stream()), "nf_trainer_apply")
        self.steps += 1
        sums = self.red[self.n_vars:self.n_vars + 3]
        if not sync:
            return sums.clone()
        s = sums.cpu().numpy()
        return float(s[0] / s[2]), float(s[1] / s[2])

    def gradients(self) -> Dict[str, np.ndarray]:
        """Gradients of the last ``loss_and_grad`` by TF variable name (synchronises; tests and debugging)."""
        flat = self.red[:self.n_vars].cpu().numpy()
        v = self.nf.spec.store.vars
        return {k: flat[self.offsets[k]:self.offsets[k] + v[k].size].reshape(v[k].shape).copy() for k in self.names
                if self.nf.spec.store.trainable.get(k, False)}

    def loss(self):
        s = self.red[self.n_vars:self.n_vars + 3].cpu().numpy()
        return float(s[0] / s[2]), float(s[1] / s[2])

    def batch_stats(self) -> np.ndarray:
        return self.red[self.n_vars + 3:].cpu().numpy().reshape(-1, 16)

    def launches_per_step(self, is_training=True) -> int:
        n = C.c_int()
        _lib.check(self.lib.nf_trainer_launches_per_step(self.handle, 1 if is_training else 0, C.byref(n)), "launches")
        return n.value

    # ---- variables
    def variables(self) -> Dict[str, np.ndarray]:
        flat = np.empty(self.n_vars, dtype=np.float32)
        with torch.cuda.device(self.nf.device):
            _lib.check(self.lib.nf_trainer_get_vars(self.handle, flat.ctypes.data_as(C.c_void_p), self._stream()), "nf_trainer_get_vars")
        v = self.nf.spec.store.vars
        return {k: flat[self.offsets[k]:self.offsets[k] + v[k].size].reshape(v[k].shape).copy() for k in self.names}

    def sync_to_model(self):
        """Device variables -> ``nf.variables`` (in place) and re-fold the inference engine."""
        new = self.variables()
        v = self.nf.spec.store.vars
        with self.nf._lock:
            for k in self.names:
                v[k][...] = new[k]
        self.nf.refresh_parameters()

    def sync_from_model(self):
        flat = self._flatten()
        with torch.cuda.device(self.nf.device):
            _lib.check(self.lib.nf_trainer_set_vars(self.handle, flat.ctypes.data_as(C.c_void_p), self._stream()), "nf_trainer_set_vars")

    def __del__(self):
        try:
            if self.handle:
                self.lib.nf_trainer_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


class WideTrainer(DeviceTrainer):
    """``DeviceTrainer`` for coupling nets wider than 4 (``--width`` 8 / 16 / 32; the reference trains any width,
    train_noise_flow.py:187-198).  Same methods, built on the host-synchronous step: forward with batch-statistics BatchNorm
    through the wide chain kernels, backward through ``csrc/nf_train_wide.cu`` (one CTA per patch, three passes per
    coupling), chain rules to the LU / scale variables and TF-rule Adam on the host.  Data parallel: one all-reduce of
    [gradients | sums | batch statistics], replicas stay identical (:func:`train_step`)."""

    def __init__(self, nf, learning_rate=1e-4, beta1=0.9, beta2=0.999, epsilon=1e-8, max_batch=256, group=None, **_unused):
        nf.build("inverse")
        if int(nf.spec.width) not in (8, 16, 32):
            raise NotImplementedError("the train-step kernels are built for coupling-net widths 4 / 8 / 16 / 32, got %d" % int(nf.spec.width))
        self.nf, self.group = nf, group
        self.opt = AdamOptimizer(learning_rate, beta1, beta2, epsilon)
        self.max_batch = int(max_batch)
        self.steps = 0
        self.handle = None
        self._last = None

    def loss_and_grad(self, x, y, nlf0=None, nlf1=None, iso=None, cam=None, is_training=True):
        loss, sd_z, grads = loss_and_grad(self.nf, x, y, nlf0, nlf1, iso, cam, is_training=is_training, refold=False, bn_update=False)
        self._last = (float(loss), float(sd_z), grads)
        return int(np.asarray(x).shape[0]) if not isinstance(x, torch.Tensor) else int(x.shape[0])

    def step(self, x, y, nlf0=None, nlf1=None, iso=None, cam=None, sync=True):
        loss, sd_z = train_step(self.nf, self.opt, x, y, nlf0, nlf1, iso, cam, group=self.group)
        self.steps += 1
        self._last = (float(loss), float(sd_z), None)
        return float(loss), float(sd_z)

    def gradients(self) -> Dict[str, np.ndarray]:
        tr = self.nf.spec.store.trainable
        return {k: np.asarray(g) for k, g in (self._last[2] or {}).items() if tr.get(k, False)}

    def loss(self):
        return self._last[0], self._last[1]

    def batch_stats(self) -> np.ndarray:
        return np.asarray(self.nf.last_batch_stats)

    def launches_per_step(self, is_training=True) -> int:
        n_cp = sum(l.kind == "coupling" for l in self.nf.spec.layers)
        return n_cp * ((3 if is_training else 1) + 3) + (len(self.nf.spec.layers) - n_cp) * 2 + 2

    def variables(self) -> Dict[str, np.ndarray]:
        return {k: v.copy() for k, v in self.nf.spec.store.vars.items()}

    def sync_to_model(self):
        self.nf.refresh_parameters()

    def sync_from_model(self):
        pass

    def __del__(self):
        pass
