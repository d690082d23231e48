"""``NoiseFlow`` -- the reference's model API (``borealisflows/noise_flow_model.py:44-513``) on top of
the fused sm_100a kernels.  Same method names, argument order and naming trap as the reference:
``inverse`` = data -> latent (likelihood), ``forward`` = latent -> data (sampling).

Tensors are ``torch`` CUDA tensors (used as device memory only; numpy inputs are copied over), shaped
``[N, 32, 32, 4]`` float32.  Every compute call goes through the C-ABI in ``include/noiseflow_b200.h``;
there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from .params import (BN_EPS, MAX_ROWS, N_STD_ROWS, NO_FULL_SUM, SCALE_GAIN, SCALE_SDN, SDN_TOKENS, ModelSpec,
                     std_row)

LOG_2PI = float(np.log(2.0 * np.pi))
N_DIMS = 32 * 32 * 4


def _fp(a: np.ndarray):
    return a.ctypes.data_as(_lib.c_float_p)


class _Engine:
    """Owns one ``nf_model`` handle built from a :class:`ModelSpec` (host-side, needs no GPU to build)."""

    def __init__(self, spec: ModelSpec):
        self.lib = _lib.load()
        self.spec = spec
        self._keep: List[np.ndarray] = []
        self.handle = C.c_void_p()
        _lib.check(self.lib.nf_model_create(32, 32, 4, int(spec.width), C.byref(self.handle)), "nf_model_create")
        self.scale_layers: List[int] = []
        for idx, l in enumerate(spec.layers):
            self._add(idx, l)
        self._finalized = False

    def _coupling_struct(self, l, iso=100.0):
        w = self.spec.coupling_weights(l, iso)
        st = _lib.NfCouplingWeights()
        keep = []
        for k in ("l1_w", "l1_b", "bn1_mean", "bn1_var", "l2_w", "l2_b", "bn2_mean", "bn2_var", "last_w",
                  "last_b", "last_logs"):
            arr = np.ascontiguousarray(w[k], dtype=np.float32)
            keep.append(arr)
            setattr(st, k, _fp(arr))
        st.rescaling_scale = w["rescaling_scale"]
        st.bn_eps = BN_EPS
        return st, keep

    def _add(self, idx, l):
        lib, h = self.lib, self.handle
        if l.kind == "conv1x1":
            a, a_inv, lad = self.spec.conv1x1_matrices(l)
            a32, i32 = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(a_inv, np.float32)
            _lib.check(lib.nf_model_add_conv1x1(h, _fp(a32), _fp(i32), lad), "nf_model_add_conv1x1")
        elif l.kind == "permute":
            perm = (C.c_int32 * 4)(*l.data["perm"])
            _lib.check(lib.nf_model_add_permute(h, perm), "nf_model_add_permute")
        elif l.kind == "coupling":
            st, keep = self._coupling_struct(l)
            mode = int(l.data.get("mode", 0))
            if mode == 0:
                _lib.check(lib.nf_model_add_affine_coupling(h, C.byref(st)), "nf_model_add_affine_coupling")
            else:       # clean-image-conditioned couplings of the legacy revnet2d models
                _lib.check(lib.nf_model_add_cond_coupling(h, mode, C.byref(st)), "nf_model_add_cond_coupling")
            del keep
        elif l.kind == "scale":
            tab = np.ascontiguousarray(self.spec.scale_table(l), np.float32)
            kind = SCALE_SDN if l.token in SDN_TOKENS else SCALE_GAIN
            full = 0 if l.token in NO_FULL_SUM else 1
            _lib.check(lib.nf_model_add_scale(h, kind, full, _fp(tab), tab.shape[0]), "nf_model_add_scale")
            self.scale_layers.append(idx)
        else:
            raise ValueError(l.kind)

    def finalize(self):
        if not self._finalized:
            _lib.check(self.lib.nf_model_finalize(self.handle), "nf_model_finalize")
            self._finalized = True

    def update_scale_tables(self, extra):
        for idx in self.scale_layers:
            tab = np.ascontiguousarray(self.spec.scale_table(self.spec.layers[idx], extra), np.float32)
            _lib.check(self.lib.nf_model_set_scale(self.handle, idx, _fp(tab), tab.shape[0]), "nf_model_set_scale")

    def refresh_parameters(self, extra=None, iso=100.0):
        """Re-upload every layer from the variable store (after an optimizer step / BN update); ``iso`` = the ISO the
        ISO-conditioned templates (legacy ``G`` couplings) are evaluated at."""
        lib, h = self.lib, self.handle
        _lib.check(lib.nf_model_begin_update(h), "nf_model_begin_update")     # set every layer, fold / upload once
        try:
            for idx, l in enumerate(self.spec.layers):
                if l.kind == "conv1x1":
                    a, a_inv, lad = self.spec.conv1x1_matrices(l)
                    a32, i32 = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(a_inv, np.float32)
                    _lib.check(lib.nf_model_set_conv1x1(h, idx, _fp(a32), _fp(i32), lad), "nf_model_set_conv1x1")
                elif l.kind == "coupling":
                    st, keep = self._coupling_struct(l, iso)
                    if int(l.data.get("mode", 0)) == 0:
                        _lib.check(lib.nf_model_set_affine_coupling(h, idx, C.byref(st)), "nf_model_set_affine_coupling")
                    else:
                        _lib.check(lib.nf_model_set_cond_coupling(h, idx, C.byref(st)), "nf_model_set_cond_coupling")
                    del keep
            self.update_scale_tables(extra)
        finally:
            _lib.check(lib.nf_model_end_update(h), "nf_model_end_update")

    def __del__(self):
        try:
            if self.handle:
                self.lib.nf_model_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


def _stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _iso_guarded(fn):
    """Run a NoiseFlow method under :meth:`NoiseFlow._iso_guard` (a no-op unless the model has ISO-conditioned templates)."""
    import functools
    import inspect
    sig = inspect.signature(fn)

    @functools.wraps(fn)
    def wrapper(self, *args, **kw):
        if not self.spec.has_iso_templates:
            return fn(self, *args, **kw)
        with self._iso_guard(sig.bind(self, *args, **kw).arguments.get("iso")):
            return fn(self, *args, **kw)
    return wrapper


class NoiseFlow(object):
    """Drop-in for ``borealisflows.noise_flow_model.NoiseFlow`` (arch-string models, ``n_levels == 1``).

    Parameters mirror the reference constructor ``NoiseFlow(x_shape, is_training, hps)``; the extra keyword
    arguments replace what TensorFlow's graph/session machinery supplied:

    ``variables``   ``{tf_variable_name: ndarray}`` (e.g. from :func:`tf_checkpoint.load_checkpoint`) = Saver.restore;
                    missing names are initialised exactly as the reference initialises them.
    ``first_call``  which graph function is traced first and therefore fixes the ``real_nvp_conv_template[_k]``
                    checkpoint names (``tf.make_template`` names scopes at first call): ``'inverse'`` for a
                    training script (loss first, train_noise_flow.py:302) -- the default, and the order every
                    checkpoint is written under; ``'forward'`` reproduces a sample-only graph
                    (``NoiseFlowWrapper``).  The order is fixed here, explicitly; it does NOT depend on which
                    method of this object happens to be called first.
    """

    def __init__(self, x_shape, is_training, hps=None, variables: Optional[Dict[str, np.ndarray]] = None,
                 first_call: Optional[str] = None, device: Union[str, torch.device, None] = None, seed: int = 0):
        if list(x_shape) != [32, 32, 4]:
            raise NotImplementedError("kernels are built for x_shape [32, 32, 4], got %s" % (list(x_shape),))
        self.x_shape = list(x_shape)
        self.hps = hps
        self.depth = getattr(hps, "depth", -1)
        self.n_levels = getattr(hps, "n_levels", 1)
        self._is_training = is_training
        self.spec = ModelSpec(hps, variables, seed)
        self.model = [self.spec.layers]
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self.device = torch.device(device)
        self._engine: Optional[_Engine] = None
        self._lock = threading.Lock()
        self._tls = threading.local()
        self._stale = False
        self._extra_rows: List[tuple] = []      # slot k <-> conditioning row 25 + k (stable: see _row_for)
        self._extra_used: Dict[int, int] = {}
        self._extra_tick = 0
        self._seed = seed
        self._sample_calls = 0
        self._folded_iso = 100.0         # ISO the ISO-conditioned templates are currently folded at (legacy G couplings)
        self._iso_lock = threading.RLock()
        if first_call is not None:
            self.build(first_call)

    # ------------------------------------------------------------------ construction
    def build(self, first_call: str = "inverse"):
        """Create template variables (first-call naming) + scale variables and the device engine."""
        with self._lock:
            if self._engine is None:
                self.spec.assign_template_scopes(first_call)
                self.spec.create_scale_variables()
                eng = _Engine(self.spec)
                eng.finalize()
                self._engine = eng
        return self

    @property
    def variables(self) -> Dict[str, np.ndarray]:
        return self.spec.store.vars

    def num_trainable_params(self) -> int:
        return self.spec.store.num_trainable()

    def get_layer_names(self):                                                      # noise_flow_model.py:508-513
        return self.spec.get_layer_names()

    def refresh_parameters(self):
        if self._engine is not None:
            with self._lock:
                self._engine.refresh_parameters(self._extra_rows, self._folded_iso)
                self._stale = False

    class _NoGuard(object):
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    def _iso_guard(self, iso):
        """Models with ISO-conditioned templates (``real_nvp_conv_template_iso``, layers.py:501-547: ``w = B1 * iso[0] + B2``)
        evaluate their coupling nets at the minibatch's ISO: re-fold when it changes, and keep concurrent callers with
        different ISOs apart for the duration of the call.  Every other model: no-op."""
        if not self.spec.has_iso_templates:
            return NoiseFlow._NoGuard()
        arr = np.asarray(100.0 if iso is None else iso, dtype=np.float64).reshape(-1)
        if len(arr) > 1 and not np.all(arr == arr[0]):
            raise ValueError("ISO-conditioned coupling nets take ONE iso per call (the reference reads iso[0], layers.py:633)")
        self._iso_lock.acquire()
        try:
            if float(arr[0]) != self._folded_iso:
                self.build()
                self._folded_iso = float(arr[0])
                self.refresh_parameters()
        except BaseException:
            self._iso_lock.release()
            raise
        return self._iso_lock

    def _fresh(self):
        """Re-fold the engine if the BatchNorm moving statistics were moved since the last fold.  A batch-statistics
        call (is_training=True) only WRITES the moving statistics; folding them into the fused moving-statistics
        program is deferred to the first call that reads them."""
        if self._stale:
            self.refresh_parameters()

    def set_launch(self, warps_per_cta: int = 12, num_ctas: int = 0):
        self.build()
        _lib.check(self._engine.lib.nf_model_set_launch(self._engine.handle, warps_per_cta, num_ctas), "nf_model_set_launch")

    def set_tensor_cores(self, enable=True):
        """Width 4 -- which chain kernel runs (C ABI ``nf_model_set_tensor_cores``):
        ``"default"`` (0) = ``"winograd"`` (4): the all-fp32 vertical-Winograd kernel in both directions;
        ``False`` / ``"direct"`` (5): the direct-form all-fp32 kernel everywhere;
        ``"hybrid"`` (2): conv-3 of every coupling net on the tensor cores (tcgen05, fp16 hi/lo split operands) in both
        directions;  ``"auto"`` (3): hybrid for sampling / forward (where it is faster), default otherwise;
        ``True`` (1): the older experimental kernel with both 3x3 convolutions as bf16 hi/lo implicit GEMMs.
        Widths 32 / 64 / 128: the tensor-core kernel is the default; ``False`` selects the CUDA-core kernel (width 32 only)."""
        self.build()
        if isinstance(enable, bool):
            mode = (1 if enable else 5) if self.hps.width == 4 else int(enable)
        else:
            mode = {"default": 0, "hybrid": 2, "auto": 3, "winograd": 4, "direct": 5}.get(enable, enable)
        _lib.check(self._engine.lib.nf_model_set_tensor_cores(self._engine.handle, int(mode)), "nf_model_set_tensor_cores")
        return self

    def set_batch_stats_fused(self, enable: bool = True):
        """``is_training=True`` calls on small batches (one co-resident CTA per patch, <= 296 on a B200) run the whole chain
        incl. the BatchNorm probes as ONE cooperative kernel (default); ``False`` forces the layer-by-layer path."""
        self.build()
        _lib.check(self._engine.lib.nf_model_set_bs_small(self._engine.handle, 1 if enable else 0), "nf_model_set_bs_small")
        return self

    # ------------------------------------------------------------------ helpers
    def _check_training(self, is_training) -> bool:
        """Resolve the reference's ``is_training`` placeholder (constructor value unless overridden per call)."""
        t = self._is_training if is_training is None else is_training
        if callable(t):
            t = t()
        return bool(t)

    def _batch_stats_chain(self, direction, inp, yy, rows, drow, n, temp=1.0, seed=0, offset=0, patch_base=0,
                           want_nll=False, want_logdet=False):
        """``is_training == True``: BatchNorm normalises with the statistics of THIS batch (layers.py:388-398) and,
        as a side effect, moves the stored statistics towards them: ``train_m -= 0.1 * (train_m - m)`` (:394-395).
        The chain runs layer by layer (two probe launches + one apply launch per coupling)."""
        e = self._engine
        out = torch.empty((n, 32, 32, 4), device=self.device, dtype=torch.float32)
        w = int(self.spec.width)
        ws = torch.zeros(2 * w, device=self.device, dtype=torch.float64)
        nll = torch.empty(n, device=self.device, dtype=torch.float32) if want_nll else None
        sdz = torch.empty(n, device=self.device, dtype=torch.float32) if want_nll else None
        ld = torch.empty(n, device=self.device, dtype=torch.float32) if want_logdet else None
        cps = [l for l in self.spec.layers if l.kind == "coupling"]
        bstats = np.zeros((max(len(cps), 1), 4 * w), dtype=np.float32)      # per coupling: mean1, var1, mean2, var2
        p = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(self.device):
            _lib.check(e.lib.nf_chain_batch_stats(
                e.handle, direction, p(inp), p(yy), p(rows), drow, n, float(temp), int(seed), int(offset),
                int(patch_base), out.data_ptr(), p(ld), p(nll), p(sdz), ws.data_ptr(),
                bstats.ctypes.data_as(C.c_void_p), _stream_ptr(self.device)), "nf_chain_batch_stats")
        if n > 0:
            self._apply_bn_moving_update(bstats)
        return out, ld, nll, sdz

    def _apply_bn_moving_update(self, bstats, refold=True):
        """``train_m -= 0.1 * (train_m - m)`` (layers.py:394-395) for every coupling, then re-fold the engine."""
        cps = [l for l in self.spec.layers if l.kind == "coupling"]
        w = int(self.spec.width)
        with self._lock:
            v = self.spec.store.vars
            for k, l in enumerate(cps):
                s = l.data["template"]
                for j, name in enumerate(("bn_nvp_conv_1/mean", "bn_nvp_conv_1/var", "bn_nvp_conv_2/mean",
                                          "bn_nvp_conv_2/var")):
                    cur = v["%s/%s" % (s, name)]
                    cur -= np.float32(0.1) * (cur - bstats[k, w * j:w * j + w])
            if refold:
                self._stale = True            # lazily re-folded by the next moving-statistics call (_fresh)
        self.last_batch_stats = bstats

    # ---- host <-> device staging for the numpy-in / numpy-out calls of the reference API --------------------------
    # A pageable cudaMemcpy runs at a few GB/s; large numpy batches therefore go through per-thread PINNED staging
    # buffers (multi-threaded host copy into / out of them, DMA at PCIe speed).  One set per calling thread: the
    # reference drives one model from 16-32 Python threads (train_noise_flow.py:38-47).
    _PIN_MIN_BYTES = 1 << 20

    def _pinned(self, role, shape):
        tl = self._tls
        pool = getattr(tl, "pool", None)
        if pool is None:
            pool = tl.pool = {}
        n = int(np.prod(shape))
        buf = pool.get(role)
        if buf is None or buf.numel() < n:
            buf = pool[role] = torch.empty(max(n, 2 * (buf.numel() if buf is not None else 0)), dtype=torch.float32, pin_memory=True)
        return buf[:n].view(shape)

    def _dev(self, a, name):
        if a is None:
            return None
        if not isinstance(a, torch.Tensor):
            arr = np.asarray(a)
            if arr.ndim != 4 or tuple(arr.shape[1:]) != (32, 32, 4):
                raise ValueError("%s must have shape [N, 32, 32, 4], got %s" % (name, tuple(arr.shape)))
            if arr.size * 4 >= self._PIN_MIN_BYTES and self.device.type == "cuda":
                ev = getattr(self._tls, "h2d_done_" + name, None)
                if ev is not None:
                    ev.synchronize()                       # the previous upload from this staging buffer has finished
                pin = self._pinned("in_" + name, arr.shape)
                pin.copy_(torch.from_numpy(np.ascontiguousarray(arr)))        # dtype conversion + parallel host copy
                out = pin.to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                setattr(self._tls, "h2d_done_" + name, ev)
                return out
            a = torch.as_tensor(arr)
        a = a.to(device=self.device, dtype=torch.float32).contiguous()
        if a.dim() != 4 or tuple(a.shape[1:]) != (32, 32, 4):
            raise ValueError("%s must have shape [N, 32, 32, 4], got %s" % (name, tuple(a.shape)))
        return a

    def to_numpy(self, t: torch.Tensor) -> np.ndarray:
        """Device tensor -> fresh float32 numpy array (what ``sess.run`` returned); large results travel through the
        calling thread's pinned staging buffer."""
        t = t.detach()
        if not t.is_cuda or t.numel() * 4 < self._PIN_MIN_BYTES:
            return t.cpu().numpy()
        pin = self._pinned("out", tuple(t.shape))
        pin.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        out = torch.empty(tuple(t.shape), dtype=torch.float32)
        out.copy_(pin)                                     # parallel host copy out of the reusable pinned buffer
        return out.numpy()

    def _rows(self, n, nlf0, nlf1, iso, cam):
        """(rows tensor or None, default_row).  The reference feeds length-1 ``iso``/``cam`` lists per
        minibatch (sidd/MiniBatchSampler.py:60-64); per-patch arrays of length N are an extension."""
        def scal(v, default):
            if v is None:
                return default
            arr = np.asarray(v, dtype=np.float64).reshape(-1)
            return arr
        iso_a, cam_a = scal(iso, np.array([100.0])), scal(cam, np.array([0.0]))
        n0_a, n1_a = scal(nlf0, np.array([0.0])), scal(nlf1, np.array([1.0]))
        needs_nlf = any(l.kind == "scale" and l.token == "camsdn" for l in self.spec.layers)
        per_patch = max(len(iso_a), len(cam_a)) > 1
        if per_patch:
            if len(iso_a) not in (1, n) or len(cam_a) not in (1, n):
                raise ValueError("iso / cam must have length 1 or N")
            iso_f = np.broadcast_to(iso_a, (n,))
            cam_f = np.broadcast_to(cam_a, (n,))
            rows = np.empty(n, dtype=np.int32)
            keys = sorted(set(zip(cam_f.tolist(), iso_f.tolist())))
            pinned = set()                # non-standard rows this batch uses: never recycled while the batch is assembled
            for key in keys:
                r = self._row_for(key[0], key[1], float(n0_a[0]), float(n1_a[0]), needs_nlf, pinned)
                rows[(cam_f == key[0]) & (iso_f == key[1])] = r
            return torch.as_tensor(rows, device=self.device), 0
        return None, self._row_for(float(cam_a[0]), float(iso_a[0]), float(n0_a[0]), float(n1_a[0]), needs_nlf, set())

    def _row_for(self, cam, iso, n0, n1, needs_nlf, pinned) -> int:
        """Conditioning-table row of (cam, iso[, nlf0, nlf1]).  Rows 0..24 are the standard (camera, ISO) grid; anything else
        (unknown ISO, explicit camera NLF of `camsdn`) lives in one of MAX_ROWS - 25 = 7 extra slots.  A slot keeps its row id
        for as long as its key stays: a new key overwrites the least recently used slot IN PLACE, never one the current batch
        uses (more than 7 distinct non-standard keys in one batch raise)."""
        r = None if needs_nlf else std_row(cam, iso)
        if r is not None:
            return r
        key = (cam, iso, n0, n1) if needs_nlf else (cam, iso, 0.0, 1.0)
        n_slots = MAX_ROWS - N_STD_ROWS
        with self._lock:
            slots = self._extra_rows
            if key in slots:
                idx = slots.index(key)
            else:
                if len(slots) < n_slots:
                    slots.append(key)
                    idx = len(slots) - 1
                else:
                    free = [k for k in sorted(range(n_slots), key=lambda k: self._extra_used.get(k, 0)) if k not in pinned]
                    if not free:
                        raise ValueError("more than %d distinct non-standard (cam, iso[, nlf]) keys in one batch" % n_slots)
                    idx = free[0]
                    slots[idx] = key                 # in place: the other slots keep their row ids
                self._engine.update_scale_tables(slots)
            self._extra_tick += 1
            self._extra_used[idx] = self._extra_tick
            pinned.add(idx)
            return N_STD_ROWS + idx

    # ------------------------------------------------------------------ reference API
    @_iso_guarded
    def inverse(self, x, objective, yy=None, nlf0=None, nlf1=None, iso=None, cam=None, is_training=None):
        """noise_flow_model.py:394-428 -> ``(z, objective + sum of log-dets)``."""
        training = self._check_training(is_training)
        self.build("inverse")
        x, yy = self._dev(x, "x"), self._dev(yy, "yy")
        n = x.shape[0]
        rows, drow = self._rows(n, nlf0, nlf1, iso, cam)
        if training:
            z, ld, _, _ = self._batch_stats_chain(0, x, yy, rows, drow, n, want_logdet=True)
            if objective is None:
                return z, ld
            obj = objective if isinstance(objective, torch.Tensor) else torch.as_tensor(np.asarray(objective))
            return z, obj.to(self.device, torch.float32) + ld
        self._fresh()
        z = torch.empty_like(x)
        ld = torch.empty(n, device=self.device, dtype=torch.float32)
        e = self._engine
        with torch.cuda.device(self.device):
            _lib.check(e.lib.nf_inverse(e.handle, x.data_ptr(), yy.data_ptr() if yy is not None else None,
                                        rows.data_ptr() if rows is not None else None, drow, n,
                                        z.data_ptr(), ld.data_ptr(), _stream_ptr(self.device)), "nf_inverse")
        if objective is None:
            return z, ld
        obj = objective if isinstance(objective, torch.Tensor) else torch.as_tensor(np.asarray(objective))
        return z, obj.to(self.device, torch.float32) + ld

    @_iso_guarded
    def forward(self, z, eps_std=None, yy=None, nlf0=None, nlf1=None, iso=None, cam=None, is_training=None):
        """noise_flow_model.py:430-447 (``eps_std`` only matters for multi-level models, as in the reference)."""
        training = self._check_training(is_training)
        self.build()
        z, yy = self._dev(z, "z"), self._dev(yy, "yy")
        n = z.shape[0]
        rows, drow = self._rows(n, nlf0, nlf1, iso, cam)
        if training:
            return self._batch_stats_chain(1, z, yy, rows, drow, n)[0]
        self._fresh()
        x = torch.empty_like(z)
        e = self._engine
        with torch.cuda.device(self.device):
            _lib.check(e.lib.nf_forward(e.handle, z.data_ptr(), yy.data_ptr() if yy is not None else None,
                                        rows.data_ptr() if rows is not None else None, drow, n,
                                        x.data_ptr(), None, _stream_ptr(self.device)), "nf_forward")
        return x

    @_iso_guarded
    def sample(self, y, eps_std=None, yy=None, nlf0=None, nlf1=None, iso=None, cam=None, is_training=None,
               eps=None, seed=None, offset=None, patch_base: int = 0):
        """noise_flow_model.py:449-456: ``z = eps * eps_std``, ``x = forward(z)``.

        ``eps`` injects the standard-normal draw (parity tests); otherwise it is drawn in-kernel with
        Philox4x32-10 keyed by ``seed`` (default: the constructor seed) and ``offset`` (default: a per-model
        call counter, so successive calls give fresh noise like ``tf.random_normal``)."""
        training = self._check_training(is_training)
        self.build()
        same = yy is y
        y = self._dev(y, "y")
        yy = y if same else (self._dev(yy, "yy") if yy is not None else None)   # sidd_utils.py:1165-1170 passes y twice
        n = y.shape[0]
        rows, drow = self._rows(n, nlf0, nlf1, iso, cam)
        temp = 1.0 if eps_std is None else float(np.asarray(eps_std).reshape(-1)[0])
        eps = self._dev(eps, "eps")
        if offset is None:
            with self._lock:
                offset = self._sample_calls
                self._sample_calls += 1
        if training:
            return self._batch_stats_chain(1, eps, yy, rows, drow, n, temp, self._seed if seed is None else seed,
                                           offset, patch_base)[0]
        self._fresh()
        x = torch.empty_like(y)
        e = self._engine
        with torch.cuda.device(self.device):
            _lib.check(e.lib.nf_sample(e.handle, yy.data_ptr() if yy is not None else None,
                                       rows.data_ptr() if rows is not None else None, drow, n, temp,
                                       eps.data_ptr() if eps is not None else None,
                                       int(self._seed if seed is None else seed), int(offset), int(patch_base),
                                       x.data_ptr(), _stream_ptr(self.device)), "nf_sample")
        return x

    @_iso_guarded
    def _loss(self, x, y, nlf0=None, nlf1=None, iso=None, cam=None, reuse=False, is_training=None, return_z=False):
        """noise_flow_model.py:458-480 -> ``(nll[N], sd_z)``."""
        training = self._check_training(is_training)
        self.build("inverse")
        x = self._dev(x, "x")
        cond = getattr(self.hps, "sidd_cond", "mix")
        yy = self._dev(y, "y") if (cond is not None and cond != "uncond") else None     # :464-467
        n = x.shape[0]
        rows, drow = self._rows(n, nlf0, nlf1, iso, cam)
        sums = torch.empty(3, device=self.device, dtype=torch.float64)
        e = self._engine
        if training:
            z, _, nll, sdz = self._batch_stats_chain(0, x, yy, rows, drow, n, want_nll=True)
            with torch.cuda.device(self.device):
                _lib.check(e.lib.nf_reduce_sums(nll.data_ptr(), sdz.data_ptr(), n, sums.data_ptr(),
                                                _stream_ptr(self.device)), "nf_reduce_sums")
            self.last_sums = self._tls.last_sums = sums
            sd_z = (sums[1] / max(n, 1)).to(torch.float32)
            return (nll, sd_z, z) if return_z else (nll, sd_z)
        self._fresh()
        nll = torch.empty(n, device=self.device, dtype=torch.float32)
        sdz = torch.empty(n, device=self.device, dtype=torch.float32)
        z = torch.empty_like(x) if return_z else None
        with torch.cuda.device(self.device):
            st = _stream_ptr(self.device)
            _lib.check(e.lib.nf_log_prob(e.handle, x.data_ptr(), yy.data_ptr() if yy is not None else None,
                                         rows.data_ptr() if rows is not None else None, drow, n, nll.data_ptr(),
                                         sdz.data_ptr(), z.data_ptr() if z is not None else None, st), "nf_log_prob")
            _lib.check(e.lib.nf_reduce_sums(nll.data_ptr(), sdz.data_ptr(), n, sums.data_ptr(), st), "nf_reduce_sums")
        # [sum nll, sum sd_z, n] (fp64, deterministic).  The calling thread's copy is the one `loss` / `sharded_loss` read:
        # the reference drives one model from 16-32 threads; `last_sums` (last call of ANY thread) is a debugging aid only
        self.last_sums = self._tls.last_sums = sums
        sd_z = (sums[1] / max(n, 1)).to(torch.float32)          # tf.reduce_mean(tf.sqrt(var_z))  :478
        if return_z:
            return nll, sd_z, z
        return nll, sd_z

    def loss(self, x, y, nlf0=None, nlf1=None, iso=None, cam=None, reuse=False, is_training=None):
        """noise_flow_model.py:482-484 -> ``(mean NLL, sd_z)``; the mean is the fp64 deterministic reduction."""
        nll, sd_z = self._loss(x, y, nlf0, nlf1, iso, cam, reuse, is_training)
        return (self._tls.last_sums[0] / max(nll.shape[0], 1)).to(torch.float32), sd_z

    def log_prob(self, x, y, nlf0=None, nlf1=None, iso=None, cam=None, is_training=None):
        """``log p(x | y, cam, iso)`` per patch (= ``-nll``; named by BASELINE.json's north star)."""
        return -self._loss(x, y, nlf0, nlf1, iso, cam, is_training=is_training)[0]

    def prior(self, name, x):
        """noise_flow_model.py:486-506: standard-normal base measure ``(logp, sample)``."""
        n = x.shape[0]

        def logp(z1):
            z1 = self._dev(z1, "z")
            return (-0.5 * (LOG_2PI + z1 * z1)).sum(dim=(1, 2, 3))

        def sample(eps_std=None):
            eps = torch.randn((n, 32, 32, 4), device=self.device, dtype=torch.float32)
            return eps if eps_std is None else eps * float(np.asarray(eps_std).reshape(-1)[0])

        return logp, sample

    # ------------------------------------------------------------------ per-bijector access (tests, config 1)
    @_iso_guarded
    def run_layers(self, first: int, last: int, direction: str, x, yy=None, nlf0=None, nlf1=None, iso=None, cam=None):
        """``_inverse_and_log_det_jacobian`` / ``_forward_and_log_det_jacobian`` of bijectors first..last-1."""
        self.build("inverse")
        self._fresh()
        x, yy = self._dev(x, "x"), self._dev(yy, "yy")
        n = x.shape[0]
        rows, drow = self._rows(n, nlf0, nlf1, iso, cam)
        out = torch.empty_like(x)
        ld = torch.empty(n, device=self.device, dtype=torch.float32)
        e = self._engine
        with torch.cuda.device(self.device):
            _lib.check(e.lib.nf_run_layers(e.handle, first, last, 0 if direction == "inverse" else 1, x.data_ptr(),
                                           yy.data_ptr() if yy is not None else None,
                                           rows.data_ptr() if rows is not None else None, drow, n, out.data_ptr(),
                                           ld.data_ptr(), _stream_ptr(self.device)), "nf_run_layers")
        return out, ld


# ---------------------------------------------------------------------------------------------------
# squeeze2d / unsqueeze2d (borealisflows/utils.py:30-86) on the device, bit-exact
# ---------------------------------------------------------------------------------------------------
def _squeeze_common(x: torch.Tensor, factor: int, squeeze_type: str, inverse: bool) -> torch.Tensor:
    lib = _lib.load()
    if not x.is_cuda:
        raise RuntimeError("squeeze2d/unsqueeze2d run on the GPU; got a CPU tensor")
    assert factor >= 1
    if factor == 1:
        return x                                                                    # utils.py:32-33,65-66
    x = x.to(torch.float32).contiguous()
    n, h, w, c = x.shape
    st = 1 if squeeze_type == "patch" else 0                                        # unknown type -> chessboard
    if not inverse:
        assert h % factor == 0 and w % factor == 0
        out = torch.empty((n, h // factor, w // factor, c * factor * factor), device=x.device, dtype=x.dtype)
        H, W, Cc, fn = h, w, c, lib.nf_squeeze2d
    else:
        assert c >= 4 and c % 4 == 0
        H, W, Cc = h * factor, w * factor, c // (factor * factor)
        out = torch.empty((n, H, W, Cc), device=x.device, dtype=x.dtype)
        fn = lib.nf_unsqueeze2d
    with torch.cuda.device(x.device):
        _lib.check(fn(x.data_ptr(), n, H, W, Cc, factor, st, out.data_ptr(), _stream_ptr(x.device)), "squeeze")
    return out


def squeeze2d(x, factor=2, squeeze_type="chessboard"):
    return _squeeze_common(x, factor, squeeze_type, False)


def unsqueeze2d(x, factor=2, squeeze_type="chessboard"):
    return _squeeze_common(x, factor, squeeze_type, True)
