"""CPU ORACLE for the Noise Flow hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  The product (``noise_flow_b200``) never does.

PARITY PINNED against the reference's own Python: TensorFlow 1.12 / TFP 0.5 cannot be installed here (Python
3.12, no network) and the reference ships no tests, but its unmodified source files (``borealisflows/*.py`` including
the ``NoiseFlowWrapper`` class) are executed in the build container over a stand-in for the TF-1.12 API surface they
touch (``oracle/tf1_shim.py``: the TF primitives restated over torch fp64, a deferred graph with placeholders /
Session / Saver) by ``oracle/make_reference_goldens.py``, and the outputs are committed as ``tests/golden/ref_*.npz``.
This restatement reproduces them to fp64 round-off (tests/test_cpu_reference_goldens.py): NLL, z, log-det, samples,
both BatchNorm modes and their moving-average side effects, tf.gradients of all 2433 parameters, two Adam steps,
the wrapper's sampling-only graph, every ``sdn*`` / ``gain*`` token, squeeze / unsqueeze.  What remains restated on
BOTH sides -- and is therefore pinned only by TensorFlow's documented semantics -- is the handful of TF primitives
themselves (conv2d SAME/VALID, moments, make_template scope naming, fill_triangular, where / one_hot).
Additional pins: (a) the shipped checkpoint artefacts (tensor names / shapes / ``num_params`` / layer names) and the
initial-value constants of the shipped ``.meta`` graph (LU assembly and 6-vector ordering, initialisers),
(b) closed-form known answers derived from the reference's own formulas and (c) self-consistency (round trips,
brute-force Jacobians).  See DESIGN.md "Oracle".

It is a line-by-line restatement in torch-CPU (float64 by default, float32 on request) of

* ``borealisflows/noise_flow_model.py:71-235,394-541``  (arch parsing, inverse/forward/sample/loss)
* ``borealisflows/layers.py:74-145,251-401,452-498,555-613,651-674`` (Conv2d1x1, AffineCoupling,
  batch_norm, real_nvp_conv_template, add_edge_padding, conv2d, conv2d_zeros)
* ``borealisflows/matrix_param.py:23-140`` (LU parameterisation, fill_triangular ordering)
* ``borealisflows/utils.py:30-86`` (squeeze2d / unsqueeze2d)
* ``borealisflows/noise_flow_layers/*.py`` + ``cond_utils.py`` (all sdn*/gain* scale layers)

Variables live in a flat ``{tf_variable_name: np.ndarray}`` store with TF-1 scoping rules
(``get_variable`` creates with the reference initialiser when the name is absent), so loading the
shipped checkpoint exercises exactly the names the reference's ``Saver`` uses -- including the
first-call-order naming of ``tf.make_template`` scopes (see ``_Template``).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

ISO_VALS = (100.0, 400.0, 800.0, 1600.0, 3200.0)   # cond_utils.py:211,224
LOG_2PI = math.log(2.0 * math.pi)


# =============================================================================================
# variable store with TF-1 variable_scope semantics (only what the hot path uses)
# =============================================================================================
class VariableStore:
    """name -> torch tensor.  ``get(name, shape, init)`` mirrors ``tf.get_variable``."""

    def __init__(self, values: Optional[Dict[str, np.ndarray]] = None, dtype=torch.float64, seed=0):
        self.dtype = dtype
        self.vars: Dict[str, torch.Tensor] = {}
        self.trainable: Dict[str, bool] = {}
        self.created: List[str] = []
        self._scope_counts: Dict[str, int] = {}
        self.rng = np.random.RandomState(seed)
        if values:
            for k, v in values.items():
                self.vars[k] = torch.as_tensor(np.asarray(v), dtype=dtype).clone()

    def get(self, name, shape, init, trainable=True):
        if name not in self.vars:
            val = init() if callable(init) else init
            arr = np.broadcast_to(np.asarray(val, dtype=np.float64), tuple(shape)).copy()
            self.vars[name] = torch.as_tensor(arr, dtype=self.dtype)
            self.created.append(name)
        v = self.vars[name]
        assert tuple(v.shape) == tuple(shape), (name, tuple(v.shape), tuple(shape))
        self.trainable.setdefault(name, trainable)
        return v

    def unique_scope(self, prefix, default_name):
        """``tf.variable_scope(None, default_name=...)`` uniquification: name, name_1, name_2, ..."""
        key = prefix + "/" + default_name
        n = self._scope_counts.get(key, 0)
        self._scope_counts[key] = n + 1
        return key if n == 0 else "%s_%d" % (key, n)

    def num_trainable(self):
        return int(sum(self.vars[k].numel() for k, t in self.trainable.items() if t))


# =============================================================================================
# borealisflows/utils.py:30-86
# =============================================================================================
def squeeze2d(x, factor=2, squeeze_type="chessboard"):
    """utils.py:30-60.  ``x``: [N,H,W,C] (torch or numpy)."""
    assert factor >= 1
    if factor == 1:
        return x
    n, height, width, n_channels = x.shape
    assert height % factor == 0 and width % factor == 0
    perm = (0, 2, 4, 5, 1, 3) if squeeze_type == "patch" else (0, 1, 3, 5, 2, 4)
    if squeeze_type == "patch":
        x = x.reshape(n, factor, height // factor, factor, width // factor, n_channels)
    else:  # 'chessboard' and the unknown-type fallback (utils.py:52-57)
        x = x.reshape(n, height // factor, factor, width // factor, factor, n_channels)
    x = x.permute(*perm) if isinstance(x, torch.Tensor) else x.transpose(perm)
    return x.reshape(n, height // factor, width // factor, n_channels * factor * factor)


def unsqueeze2d(x, factor=2, squeeze_type="chessboard"):
    """utils.py:63-86."""
    assert factor >= 1
    if factor == 1:
        return x
    n, height, width, n_channels = x.shape
    assert n_channels >= 4 and n_channels % 4 == 0
    x = x.reshape(n, height, width, int(n_channels / factor ** 2), factor, factor)
    perm = (0, 4, 1, 5, 2, 3) if squeeze_type == "patch" else (0, 1, 4, 2, 5, 3)
    x = x.permute(*perm) if isinstance(x, torch.Tensor) else x.transpose(perm)
    return x.reshape(n, int(height * factor), int(width * factor), int(n_channels / factor ** 2))


# =============================================================================================
# borealisflows/matrix_param.py
# =============================================================================================
def fill_triangular(v, upper=False):
    """TFP ``fill_triangular`` (the op matrix_param.py:44 calls), vector [m] -> [n,n], m = n(n+1)/2.

    Published algorithm (tensorflow_probability/python/internal/distribution_util.py, v0.5; TFP is an un-vendored
    dependency of the reference): with ``tail = x[n:]``,
    lower: ``reshape(concat([tail, reverse(x)]), [n, n])`` then keep the lower triangle;
    upper: ``reshape(concat([x, reverse(tail)]), [n, n])`` then keep the upper triangle.
    Docstring example of the library: [1..6] -> [[4,0,0],[6,5,0],[3,2,1]] (lower), [[1,2,3],[0,5,6],[0,0,4]] (upper).
    """
    v = np.asarray(v)
    m = v.shape[-1]
    n = int(round((math.sqrt(8 * m + 1) - 1) / 2))
    assert n * (n + 1) // 2 == m
    tail = v[n:]
    if upper:
        full = np.concatenate([v, tail[::-1]]).reshape(n, n)
        return np.triu(full)
    full = np.concatenate([tail, v[::-1]]).reshape(n, n)
    return np.tril(full)


def fill_triangular_inverse(mat, upper=False):
    """TFP ``fill_triangular_inverse`` (matrix_param.py:87): the vector ``v`` such that
    ``fill_triangular(v, upper) == mat`` on the kept triangle.  Implemented by probing the forward
    map (n <= 16 here), which makes it the exact inverse by construction."""
    mat = np.asarray(mat)
    n = mat.shape[-1]
    m = n * (n + 1) // 2
    pos = fill_triangular(np.arange(1, m + 1, dtype=np.float64), upper=upper).astype(np.int64)
    out = np.zeros(m, dtype=mat.dtype)
    for i in range(n):
        for j in range(n):
            if pos[i, j]:
                out[pos[i, j] - 1] = mat[i, j]
    return out


def vec2stricttri(vec, upper):
    """matrix_param.py:31-57: fill_triangular of the (n-1)x(n-1) block, padded to strict nxn."""
    base = fill_triangular(vec, upper=upper)
    k = base.shape[0]
    out = np.zeros((k + 1, k + 1), dtype=base.dtype)
    if upper:
        out[:k, 1:] = base      # pad [[0,1],[1,0]]: one row below, one column left
    else:
        out[1:, :k] = base      # pad [[1,0],[0,1]]
    return out


def stricttri2vec(mat, upper):
    """matrix_param.py:60-97."""
    mat = np.asarray(mat)
    trim = mat[:-1, 1:] if upper else mat[1:, :-1]
    trim = np.triu(trim) if upper else np.tril(trim)
    return fill_triangular_inverse(trim, upper=upper)


def _t_vec2stricttri(vec: torch.Tensor, upper: bool) -> torch.Tensor:
    """Differentiable torch version of :func:`vec2stricttri` (index map derived from it)."""
    m = vec.shape[-1]
    idx = vec2stricttri(np.arange(1, m + 1, dtype=np.float64), upper).astype(np.int64)
    flat = torch.cat([vec.new_zeros(1), vec])
    return flat[torch.as_tensor(idx)]


def matrix_param_lu(store: VariableStore, scope: str, name: str, init_A_fn, n: int):
    """matrix_param.py:100-140.  Returns dict(A, A_inv, log_abs_det).
    ``init_A_fn`` is only evaluated when the variables do not exist yet (fresh initialisation)."""
    import scipy.linalg

    def _lu():
        if not hasattr(_lu, "c"):
            p_, l_, u_ = scipy.linalg.lu(init_A_fn())       # :102
            _lu.c = (p_, l_, u_)
        return _lu.c

    p = store.get("%s/P_matpar_lu_%s" % (scope, name), (n, n), lambda: _lu()[0], trainable=False)
    sign_s = store.get("%s/sign_S_matpar_lu_%s" % (scope, name), (n,),
                       lambda: np.sign(np.diag(_lu()[2])), trainable=False)          # :105,111
    log_s = store.get("%s/log_S_matpar_lu_%s" % (scope, name), (n,),
                      lambda: np.log(np.abs(np.diag(_lu()[2]))))                    # :106,114
    nv = n * (n - 1) // 2
    l_vec = store.get("%s/L_vec_matpar_lu_%s" % (scope, name), (nv,),
                      lambda: stricttri2vec(_lu()[1], upper=False))                  # :117-119
    u_vec = store.get("%s/U_vec_matpar_lu_%s" % (scope, name), (nv,),
                      lambda: stricttri2vec(np.triu(_lu()[2], k=1), upper=True))    # :124-126
    eye = torch.eye(n, dtype=store.dtype)
    l = _t_vec2stricttri(l_vec, upper=False) + eye                                   # :120-121
    u = _t_vec2stricttri(u_vec, upper=True) + torch.diag(sign_s * torch.exp(log_s))  # :127-128
    A = p @ (l @ u)                                                                   # :130
    p_inv = p.t()                                                                     # :133
    y = torch.linalg.solve_triangular(l, p_inv, upper=False)                          # :135-136
    A_inv = torch.linalg.solve_triangular(u, y, upper=True)
    return {"A": A, "A_inv": A_inv, "log_abs_det": log_s.sum()}                      # :138-140


def matrix_param_none(store: VariableStore, scope: str, name: str, init_A_fn, n: int):
    """matrix_param.py:23-29."""
    A = store.get("%s/A_matpar_none_%s" % (scope, name), (n, n), init_A_fn)
    return {"A": A, "A_inv": torch.linalg.inv(A), "log_abs_det": torch.linalg.slogdet(A)[1]}


# =============================================================================================
# borealisflows/layers.py
# =============================================================================================
def _conv_nhwc(x, w, padding):
    """``tf.nn.conv2d(x, w, [1,1,1,1], padding, 'NHWC')``: cross-correlation, filter [kh,kw,in,out]."""
    xt = x.permute(0, 3, 1, 2)
    wt = w.permute(3, 2, 0, 1)
    if padding == "SAME":
        pad = ((w.shape[0] - 1) // 2, (w.shape[1] - 1) // 2)
    else:
        pad = (0, 0)
    return F.conv2d(xt, wt, padding=pad).permute(0, 2, 3, 1)


class Conv2d1x1:
    """layers.py:74-145 (bias=False everywhere on the hot path, noise_flow_model.py:88)."""

    def __init__(self, store, scope, x_shape, decomp="LU", layer_id=0, order=0, name="conv2d_1x1"):
        self.name = name
        self.i0, self.i1, self.ic = x_shape
        self.store, self.scope = store, scope + "/" + name                           # :96
        self._decomp = decomp
        self._pname = "conv2d_1x1_%d_%d" % (layer_id, order)                        # :98-99
        # :95  random orthogonal init (only drawn when the variables do not exist yet)
        self._init = None

    def _init_A(self):
        if self._init is None:
            import scipy.linalg
            self._init = scipy.linalg.qr(self.store.rng.randn(self.ic, self.ic))[0].astype("float32")
        return self._init

    def params(self):
        if self._decomp == "NONE" or self.ic <= 1:                                   # matrix_param.py:196-204
            return matrix_param_none(self.store, self.scope, self._pname, self._init_A, self.ic)
        if self._decomp != "LU":
            raise NotImplementedError("decomp %s" % self._decomp)
        return matrix_param_lu(self.store, self.scope, self._pname, self._init_A, self.ic)

    def _forward(self, x):                                                            # :108-115
        return x @ self.params()["A_inv"]

    def _inverse(self, y):                                                            # :117-124
        return y @ self.params()["A"]

    def _inverse_log_det_jacobian(self, y):                                           # :129-130
        return self.params()["log_abs_det"] * (self.i0 * self.i1)

    def _forward_log_det_jacobian(self, x):                                           # :126-127
        return -self._inverse_log_det_jacobian(None)

    def _inverse_and_log_det_jacobian(self, y):                                       # :137-140
        return self._inverse(y), self._inverse_log_det_jacobian(y)

    def _forward_and_log_det_jacobian(self, x):                                       # :132-135
        return self._forward(x), self._forward_log_det_jacobian(x)


class Permute:
    """``tfb.Permute(permutation=range(C)[::-1])`` (noise_flow_model.py:80-84): y[..., i] = x[..., perm[i]]."""

    def __init__(self, n_channels, name="permute"):
        self.name = name
        self.perm = list(range(n_channels))[::-1]
        self.inv = list(np.argsort(self.perm))

    def _forward(self, x):
        return x[..., self.perm]

    def _inverse(self, y):
        return y[..., self.inv]

    def _inverse_and_log_det_jacobian(self, y):
        return self._inverse(y), torch.zeros((), dtype=y.dtype)

    def _forward_and_log_det_jacobian(self, x):
        return self._forward(x), torch.zeros((), dtype=x.dtype)


def batch_norm(store, scope, x, training, eps=1e-4, decay=0.1, name="batch_norm"):
    """layers.py:378-401.  Side effect in training mode: moving statistics are updated (:394-395)."""
    c = x.shape[-1]
    train_m = store.get("%s/%s/mean" % (scope, name), (c,), 0.0, trainable=False)
    train_v = store.get("%s/%s/var" % (scope, name), (c,), 1.0, trainable=False)
    if training:
        m = x.mean(dim=(0, 1, 2))                                                    # :393 tf.nn.moments
        v = ((x - m) ** 2).mean(dim=(0, 1, 2))                                       # population variance
        with torch.no_grad():
            train_m -= decay * (train_m - m)                                         # :394
            train_v -= decay * (train_v - v)                                         # :395
        return (x - m) / torch.sqrt(v + eps)                                         # :398
    return (x - train_m) / torch.sqrt(train_v + eps)                                 # :400


def add_edge_padding(x, filter_size):
    """layers.py:555-583: zero-pad 1 px and append an indicator channel that is 1 on the pad ring."""
    if filter_size[0] == 1 and filter_size[1] == 1:
        return x
    a = (filter_size[0] - 1) // 2
    b = (filter_size[1] - 1) // 2
    x = F.pad(x, (0, 0, b, b, a, a))                                                 # :563 (N,H,W,C): pad W by b, H by a
    pad = torch.zeros((1,) + tuple(x.shape[1:3]) + (1,), dtype=x.dtype)             # :567
    pad[:, :a, :, 0] = 1.0
    pad[:, -a:, :, 0] = 1.0
    pad[:, :, :b, 0] = 1.0
    pad[:, :, -b:, 0] = 1.0
    pad = pad.expand(x.shape[0], -1, -1, -1)                                         # :576
    return torch.cat([x, pad], dim=3)                                                # :577


def conv2d(store, scope, name, x, width, filter_size=(3, 3), edge_bias=False):
    """layers.py:586-613 (pad SAME, stride 1, skip 1, no weight-norm, edge_bias False by default)."""
    pad = "SAME"
    if edge_bias and pad == "SAME":
        x = add_edge_padding(x, filter_size)
        pad = "VALID"
    n_in = x.shape[3]
    std = width / 512 * 0.05                                                          # :599
    w = store.get("%s/%s/W" % (scope, name), tuple(filter_size) + (n_in, width),
                  lambda: store.rng.randn(*(tuple(filter_size) + (n_in, width))) * std)
    x = _conv_nhwc(x, w, pad)                                                         # :604
    return x + store.get("%s/%s/b" % (scope, name), (1, 1, 1, width), 0.0)           # :608


def conv2d_zeros(store, scope, name, x, width, filter_size=(3, 3), logscale_factor=3, edge_bias=True):
    """layers.py:651-674."""
    pad = "SAME"
    if edge_bias and pad == "SAME":
        x = add_edge_padding(x, filter_size)
        pad = "VALID"
    n_in = x.shape[3]
    w = store.get("%s/%s/W" % (scope, name), tuple(filter_size) + (n_in, width), 0.0)
    x = _conv_nhwc(x, w, pad)
    x = x + store.get("%s/%s/b" % (scope, name), (1, 1, 1, width), 0.0)
    x = x * torch.exp(store.get("%s/%s/logs" % (scope, name), (1, width), 0.0) * logscale_factor)
    return x


class _Template:
    """``tf.make_template`` (layers.py:498): the variable scope is created -- and uniquified as
    ``real_nvp_conv_template``, ``..._1``, ... under the *current* scope -- at the FIRST CALL,
    so checkpoint names depend on which of inverse()/forward() is traced first."""

    def __init__(self, store, default_name):
        self.store, self.default_name, self.scope = store, default_name, None

    def ensure_scope(self, current_scope):
        if self.scope is None:
            self.scope = self.store.unique_scope(current_scope, self.default_name)
        return self.scope


class RealNVPConvTemplate(_Template):
    """layers.py:452-498."""

    def __init__(self, store, x_shape, width):
        super().__init__(store, "real_nvp_conv_template")
        self.x_shape, self.width = x_shape, width

    def __call__(self, x, is_training, current_scope="model"):
        s = self.ensure_scope(current_scope)
        st = self.store
        ic = int(self.x_shape[2] / 2)
        num_output = 2 * ic                                                            # :467
        x = conv2d(st, s, "l_1", x, self.width)                                       # :469
        x = batch_norm(st, s, x, is_training, name="bn_nvp_conv_1")                   # :472-477
        x = torch.relu(x)                                                              # :478
        x = conv2d(st, s, "l_2", x, self.width, filter_size=(1, 1))                   # :480
        x = batch_norm(st, s, x, is_training, name="bn_nvp_conv_2")                   # :483-488
        x = torch.relu(x)                                                              # :489
        x = conv2d_zeros(st, s, "l_last", x, num_output)                              # :491
        shift, log_scale = x[..., :num_output // 2], x[..., num_output // 2:]         # :494 tf.split
        return shift, log_scale


def conv2d_iso(store, scope, name, x, width, iso0, filter_size=(3, 3)):
    """layers.py:616-648: filter and bias are affine in the ISO, ``W = B1 * iso[0] + B2``, ``b = C1 * iso[0] + C2``
    (all four N(0, 0.05^2)-initialised, :628); pad SAME, no edge bias."""
    n_in = x.shape[3]
    shp = tuple(filter_size) + (n_in, width)
    rnd = lambda sh: (lambda: store.rng.randn(*sh) * 0.05)
    b1 = store.get("%s/%s/B1" % (scope, name), shp, rnd(shp))
    b2 = store.get("%s/%s/B2" % (scope, name), shp, rnd(shp))
    x = _conv_nhwc(x, b1 * iso0 + b2, "SAME")                                         # :633,:639
    c1 = store.get("%s/%s/C1" % (scope, name), (1, 1, 1, width), rnd((1, 1, 1, width)))
    c2 = store.get("%s/%s/C2" % (scope, name), (1, 1, 1, width), rnd((1, 1, 1, width)))
    return x + (c1 * iso0 + c2)                                                       # :647


class RealNVPConvTemplateIso(_Template):
    """layers.py:501-547: ``real_nvp_conv_template`` with ISO-conditioned first two convolutions."""

    def __init__(self, store, x_shape, width):
        super().__init__(store, "real_nvp_conv_template_iso")
        self.x_shape, self.width = x_shape, width

    def __call__(self, x, iso, is_training, current_scope="model"):
        s = self.ensure_scope(current_scope)
        st = self.store
        iso0 = torch.as_tensor(np.asarray(iso, dtype=np.float64), dtype=st.dtype).reshape(-1)[0]
        num_output = 2 * int(self.x_shape[2] / 2)                                     # :516
        x = conv2d_iso(st, s, "l_1", x, self.width, iso0)                             # :518
        x = torch.relu(batch_norm(st, s, x, is_training, name="bn_nvp_conv_1"))       # :519-527
        x = conv2d_iso(st, s, "l_2", x, self.width, iso0, filter_size=(1, 1))         # :529
        x = torch.relu(batch_norm(st, s, x, is_training, name="bn_nvp_conv_2"))       # :530-538
        x = conv2d_zeros(st, s, "l_last", x, num_output)                              # :540
        return x[..., :num_output // 2], x[..., num_output // 2:]                     # :543


class CondCoupling:
    """The clean-image-conditioned couplings reachable through ``revnet2d`` (noise_flow_model.py:305-348):
    ``AffineCouplingCondY`` / ``CondYG`` -- the net sees only the clean patch and shifts / scales ALL channels
    (AffineCouplingCondY.py:44-72) -- and ``AffineCouplingCondXY`` / ``CondXYG`` -- the net sees
    ``concat(x0, yy)`` and transforms ``x1`` (AffineCouplingCondXY.py:45-79).  The ``G`` variants pass the ISO to a
    ``real_nvp_conv_template_iso``."""

    def __init__(self, mode, store, scope, x_shape, fn, layer_id=0, name="real_nvp"):
        assert mode in ("Y", "YG", "XY", "XYG")
        self.mode, self.name, self._fn, self.store = mode, name, fn, store
        self.i0, self.i1, self.ic = x_shape
        self.scale = store.get("%s/rescaling_scale%d" % (scope, layer_id), (), 1e-4)

    def _net(self, x, yy, iso, is_training):
        inp = yy if self.mode in ("Y", "YG") else torch.cat([x[..., :self.ic // 2], yy], dim=-1)
        if self.mode.endswith("G"):
            shift, log_scale = self._fn(inp, iso, is_training)
        else:
            shift, log_scale = self._fn(inp, is_training)
        return shift, self.scale * torch.tanh(log_scale)

    def _inverse_and_log_det_jacobian(self, y, yy, nlf0=None, nlf1=None, iso=None, cam=None, is_training=False):
        shift, ls = self._net(y, yy, iso, is_training)
        if self.mode in ("Y", "YG"):
            return y * torch.exp(ls) + shift, ls.sum(dim=(1, 2, 3))
        y0, y1 = y[..., :self.ic // 2], y[..., self.ic // 2:]
        return torch.cat([y0, y1 * torch.exp(ls) + shift], dim=-1), ls.sum(dim=(1, 2, 3))

    def _forward(self, x, yy, nlf0=None, nlf1=None, iso=None, cam=None, is_training=False):
        shift, ls = self._net(x, yy, iso, is_training)
        if self.mode in ("Y", "YG"):
            return (x - shift) * torch.exp(-ls)
        x0, x1 = x[..., :self.ic // 2], x[..., self.ic // 2:]
        return torch.cat([x0, (x1 - shift) * torch.exp(-ls)], dim=-1)


class AffineCoupling:
    """layers.py:251-375."""

    def __init__(self, store, scope, x_shape, fn, layer_id=0, name="real_nvp"):
        self.name = name
        self.i0, self.i1, self.ic = x_shape
        self._fn = fn
        self.store = store
        self.scale = store.get("%s/rescaling_scale%d" % (scope, layer_id), (), 1e-4)  # :271-273

    def _forward_and_log_det_jacobian(self, x, is_training=False):                    # :333-353
        x0 = x[..., :self.ic // 2]
        x1 = x[..., self.ic // 2:]
        shift, log_scale = self._fn(x0, is_training)
        log_scale = self.scale * torch.tanh(log_scale)                                # :342
        y1 = (x1 - shift) * torch.exp(-log_scale)                                     # :343-347
        y = torch.cat([x0, y1], dim=-1)
        return y, -log_scale.sum(dim=(1, 2, 3))                                       # :352

    def _inverse_and_log_det_jacobian(self, y, is_training=False):                    # :355-375
        y0 = y[..., :self.ic // 2]
        y1 = y[..., self.ic // 2:]
        shift, log_scale = self._fn(y0, is_training)
        log_scale = self.scale * torch.tanh(log_scale)                                # :362
        x1 = y1 * torch.exp(log_scale) + shift                                        # :363-367
        x = torch.cat([y0, x1], dim=-1)
        return x, log_scale.sum(dim=(1, 2, 3))                                        # :372

    def _forward(self, x, is_training=False):
        return self._forward_and_log_det_jacobian(x, is_training)[0]

    def _inverse(self, y, is_training=False):
        return self._inverse_and_log_det_jacobian(y, is_training)[0]


# =============================================================================================
# borealisflows/noise_flow_layers/cond_utils.py -- scale functions
# =============================================================================================
def _sigmoid(t):
    return torch.sigmoid(t)


def _iso_select(store, scope, fmt, iso, init, legacy_default=True):
    """Nested ``tf.cond`` ladders over iso[0] (e.g. cond_utils.py:71-89): unknown ISO -> the 800 branch."""
    gs = {int(v): store.get("%s/%s" % (scope, fmt % int(v)), (1,), init) for v in ISO_VALS}
    key = int(float(iso[0])) if float(iso[0]) in ISO_VALS else 800
    return gs[key]


def _iso_onehot_param(gain_params, iso):
    """``tf.where(tf.equal(iso_vals, iso))`` -> one_hot -> reduce_sum (cond_utils.py:170-172,226-228):
    an ISO that is not in the table yields an empty ``where`` and therefore g = 0."""
    g = gain_params.new_zeros(())
    for k, v in enumerate(ISO_VALS):
        if float(iso[0]) == v:
            g = g + gain_params[k]
    return g


def _cam_params(cam_params, cam, c_i):
    """cond_utils.py:216-220.  Unknown cam -> ``cam_idx[0]`` on an empty tensor -> error."""
    idx = [k for k in range(5) if float(cam[0]) == float(k)]
    if not idx:
        raise IndexError("camera id %r not in [0..4] (reference: tf.where -> empty -> cam_idx[0] fails)" % (cam,))
    return torch.exp(c_i * cam_params[:, idx[0]])


def scale_fn(kind, store, model_scope, yy, nlf0, nlf1, iso, cam, gain_init, param_inits):
    """Returns (scale tensor broadcastable to x, full_sum_logdet: bool).

    ``full_sum_logdet`` False reproduces the reference quirk that ``Gain``, ``GainEx1`` and
    ``GainEx3`` return ``-+log(scale)`` without summing over the 4096 dimensions
    (AffineCouplingGain.py:86,96,111,125 and the same lines of Ex1 / Ex3)."""
    g = store.get
    sc = model_scope
    iso_t = torch.as_tensor(np.asarray(iso, dtype=np.float64), dtype=store.dtype) if iso is not None else None
    if kind in ("sdngain", "fitsdngain2"):                                            # cond_utils.py:11-38 (ISO polynomials)
        i0 = iso_t.reshape(-1)[0]
        e = lambda n: torch.exp(g(sc + "/" + n, (1,), -6.0))
        if kind == "sdngain":                                                         # sdn_iso_model_params_3 (:11-24)
            beta1 = e("p1") * i0 ** 2 + e("p2") * i0 + e("p3")
            beta2 = e("q1") * i0 ** 3 + e("q2") * i0 ** 2 + e("q3") * i0 + e("q4")
        else:                                                                         # sdn_iso_model_params_2 (:27-38)
            beta1 = e("p2") * i0 + e("p3")
            beta2 = e("q2") * i0 ** 2 + e("q3") * i0 + e("q4")
        return torch.sqrt(beta1 * yy + beta2), True                                   # AffineCouplingSdnGain.py:46-47
    if kind == "sdn":                                                                 # cond_utils.py:41-52
        b1 = _sigmoid(g(sc + "/b1", (1,), -3.0))
        b2 = _sigmoid(g(sc + "/b2", (1,), 3.0))
        return torch.sqrt(b1 * yy + b2), True
    if kind == "sdn1":                                                                # :55-97
        c = 1e-2
        rg = _iso_select(store, sc, "r_gain_param_%05d", iso, 0.0 / c)
        r_gain = torch.exp(c * rg) * iso_t
        b1 = _sigmoid(g(sc + "/b1", (1,), -3.0))
        b2 = _sigmoid(g(sc + "/b2", (1,), 3.0))
        return torch.sqrt(b1 * yy / r_gain + b2), True
    if kind in ("sdn2", "sdn3"):                                                      # :100-162
        c = 1e-1
        gp = _iso_select(store, sc, "gain_param_%05d", iso, gain_init / c)
        gain = torch.exp(c * gp) * iso_t
        b1 = _sigmoid(g(sc + "/b1", (1,), -3.0))
        b2 = _sigmoid(g(sc + "/b2", (1,), 3.0))
        if kind == "sdn2":
            return torch.sqrt(gain * (b1 * yy / gain + b2)), True                     # :131
        return gain * torch.sqrt(b1 * yy / gain + b2), True                           # :161
    if kind == "sdn4":                                                                # :165-187
        c = 1
        s2 = sc + "/sdn_gain"
        g(s2 + "/gain_val", (1,), 1.0)
        gain_params = g(s2 + "/gain_params", (5,), gain_init / c)
        gain = torch.exp(c * _iso_onehot_param(gain_params, iso)) * iso_t
        beta1 = torch.exp(c * g(s2 + "/beta1", (1,), gain_init / c))
        beta2 = torch.exp(c * g(s2 + "/beta2", (1,), 0.0))
        return torch.sqrt(beta1 * yy / gain + beta2), True
    if kind in ("sdn5", "sdn6"):                                                      # :205-276
        (c_i, beta1_i, beta2_i, gain_params_i, cam_params_i) = param_inits
        npc = 3 if kind == "sdn5" else 1
        s2 = sc + "/sdn_gain"
        cam_params = g(s2 + "/cam_params", (npc, 5), lambda: np.asarray(cam_params_i))
        ocp = _cam_params(cam_params, cam, c_i)
        g(s2 + "/gain_val", (1,), 1.0)                                                # :223 (created, unused here)
        gain_params = g(s2 + "/gain_params", (5,), lambda: np.asarray(gain_params_i))
        gsel = _iso_onehot_param(gain_params, iso)
        beta1 = g(s2 + "/beta1", (1,), beta1_i)
        beta2 = g(s2 + "/beta2", (1,), beta2_i)
        if kind == "sdn5":
            gain = torch.exp(c_i * gsel * ocp[2]) * iso_t                             # :230
            beta1 = torch.exp(c_i * beta1 * ocp[0])                                   # :236
            beta2 = torch.exp(c_i * beta2 * ocp[1])                                   # :237
        else:
            gain = torch.exp(c_i * gsel * ocp[0]) * iso_t                             # :267
            beta1 = torch.exp(c_i * beta1)
            beta2 = torch.exp(c_i * beta2)
        return torch.sqrt(beta1 * yy / gain + beta2), True                            # :238 / :275
    if kind == "camsdn":                                                              # AffineCouplingCamSdn.py:47
        n0 = torch.as_tensor(np.asarray(nlf0, dtype=np.float64), dtype=store.dtype).reshape(-1, 1, 1, 1)
        n1 = torch.as_tensor(np.asarray(nlf1, dtype=np.float64), dtype=store.dtype).reshape(-1, 1, 1, 1)
        return torch.sqrt(yy * n0 + n1), True
    if kind == "gain":                                                                # cond_utils.py:319-330
        g1 = _sigmoid(g(sc + "/g1", (1,), -3.0))
        g2 = _sigmoid(g(sc + "/g2", (1,), 3.0))
        return g1 * iso_t + g2, False
    if kind == "gain1":                                                               # :333-351
        c = 1e-5
        g1 = g(sc + "/g1", (1,), -5.0 / c)
        g2 = g(sc + "/g2", (1,), 0.0 / c)
        return torch.exp(c * g1) * iso_t + torch.exp(c * g2), False
    if kind == "gain2":                                                               # :354-395
        c = 1e-1
        gp = _iso_select(store, sc, "gain_param_%05d", iso, gain_init / c)
        return torch.exp(c * gp) * iso_t, True
    if kind == "gain3":                                                               # :398-429
        c = 1e-5
        gp = _iso_select(store, sc, "gain_param_%05d", iso, -5.0 / c)
        return torch.exp(c * gp), False
    if kind == "gain4":                                                               # :432-440
        return g(sc + "/sdn_gain/gain_val", (1,), 1.0), True
    raise ValueError("unknown scale layer %r" % kind)


class ScaleBijector:
    """The 19 near-identical ``AffineCoupling{Sdn*,Gain*,CamSdn}`` templates
    (e.g. AffineCouplingSdnEx5.py:22-132, AffineCouplingGainEx4.py:23-127):
    forward ``y = x * scale``; inverse ``x = y / scale``; log-det ``+-sum(log scale)``."""

    def __init__(self, kind, store, scope, x_shape, layer_id, name, gain_init, param_inits):
        self.kind, self.name, self.store = kind, name, store
        self.i0, self.i1, self.ic = x_shape
        self.gain_init, self.param_inits = gain_init, param_inits
        store.get("%s/rescaling_scale%d" % (scope, layer_id), (), 1e-4)               # created, never used

    def _scale(self, yy, nlf0, nlf1, iso, cam):
        return scale_fn(self.kind, self.store, "model", yy, nlf0, nlf1, iso, cam,
                        self.gain_init, self.param_inits)

    def _logdet(self, scale, like, full):
        if full:
            return torch.log(scale + like * 0.0).sum(dim=(1, 2, 3))                   # GainEx4.py:86-91 "+ x*0.0"
        return torch.log(scale)                                                       # quirk: shape [1], no sum

    def _forward(self, x, yy, nlf0=None, nlf1=None, iso=None, cam=None):
        scale, _ = self._scale(yy, nlf0, nlf1, iso, cam)
        return x * scale

    def _inverse_and_log_det_jacobian(self, y, yy, nlf0=None, nlf1=None, iso=None, cam=None):
        scale, full = self._scale(yy, nlf0, nlf1, iso, cam)
        return y / scale, -self._logdet(scale, y, full)

    def _forward_and_log_det_jacobian(self, x, yy, nlf0=None, nlf1=None, iso=None, cam=None):
        scale, full = self._scale(yy, nlf0, nlf1, iso, cam)
        return x * scale, self._logdet(scale, x, full)


_SCALE_TOKENS = {"sdn": "sdn", "sdn1": "sdn1", "sdn2": "sdn2", "sdn3": "sdn3", "sdn4": "sdn4",
                 "sdn5": "sdn5", "sdn6": "sdn6", "gain": "gain", "gain1": "gain1", "gain2": "gain2",
                 "gain3": "gain3", "gain4": "gain4", "camsdn": "camsdn"}


# =============================================================================================
# borealisflows/noise_flow_model.py
# =============================================================================================
def default_param_inits(arch):
    """NoiseFlowWrapper.py:125-137 / train_noise_flow.py:205-215: always recomputed, never parsed."""
    npcam = 1 if ("sdn6" in arch and "sdn5" not in arch) else 3
    c_i = 1.0
    return (c_i, -5.0 / c_i, 0.0, np.full([5], -5.0 / c_i), np.full([npcam, 5], 1.0))


class OracleNoiseFlow:
    """``NoiseFlow`` (noise_flow_model.py:44-513), n_levels == 1: ``hps.arch`` models and, with ``hps.arch = None``, the
    legacy ``revnet2d`` assembly (oracle only: the CUDA engine implements the ``hps.arch`` path)."""

    def __init__(self, x_shape, hps, variables=None, dtype=torch.float64, seed=0):
        self.x_shape = list(x_shape)
        self.hps = hps
        self.dtype = dtype
        self.store = VariableStore(variables, dtype=dtype, seed=seed)
        if getattr(hps, "n_levels", 1) != 1:
            raise NotImplementedError("n_levels > 1 (split2d) is not on the hot path")
        if (not hasattr(hps, "param_inits") or isinstance(hps.param_inits, str)) and getattr(hps, "arch", None) is not None:
            hps.param_inits = default_param_inits(hps.arch)
        shape = list(self.x_shape)
        if hps.squeeze_factor != 1:                                                   # :58-60
            shape = [shape[0] // 2, shape[1] // 2, shape[2] * 4]
        if getattr(hps, "arch", None) is not None:                                   # :63-68
            self.model = [self.noise_flow_arch("level0", shape, hps.flow_permutation, hps.arch)]
        else:
            self.model = [self.revnet2d("level0", shape, hps.flow_permutation)]

    # ---- noise_flow_model.py:237-392 (legacy: only reachable with hps.arch unset)
    def revnet2d(self, name, x_shape, flow_permutation):
        h, st, depth = self.hps, self.store, self.hps.depth
        wide = list(x_shape[:-1]) + [x_shape[-1] * 2]                                # "double outputs" (:275,:313)
        tmpl = lambda shp: RealNVPConvTemplate(st, shp, h.width)
        tmpl_iso = lambda shp: RealNVPConvTemplateIso(st, shp, h.width)
        scale = lambda kind, scope, nm: ScaleBijector(kind, st, scope, x_shape, 0, nm, getattr(h, "gain_init", 0.0), None)
        b = []
        if getattr(h, "append_sdn2", False):                                          # :243-253
            b.append(scale("fitsdngain2", name + "/bijector_sdn2", "ac_fitSdnGain2_%d" % depth))
        if getattr(h, "append_sdn_first", False):                                     # :255-265
            b.append(scale("sdngain", name + "/bijector_sdn", "ac_fitSdnGain_%d" % depth))
        if getattr(h, "append_cY", False):                                            # :267-279
            b.append(CondCoupling("Y", st, name + "/bijector_cy", x_shape, tmpl(wide), name="ac_cY_first"))
        for i in range(depth):                                                        # :280-378
            scope = "%s/bijector%d" % (name, i)
            if flow_permutation == 0:
                b.append(Permute(x_shape[-1], name="permute"))
            elif flow_permutation == 1:
                b.append(Conv2d1x1(st, scope, x_shape, decomp=h.decomp, layer_id=i, name="Conv2d_1x1_%d" % i))
            cond = h.sidd_cond
            if cond == "condY":
                b.append(CondCoupling("Y", st, scope, x_shape, tmpl(wide), name="ac_cY_%d" % i))
            elif cond == "condYG":
                b.append(CondCoupling("YG", st, scope, x_shape, tmpl_iso(wide), name="ac_cYG_%d" % i))
            elif cond == "condXY":
                b.append(CondCoupling("XY", st, scope, x_shape, tmpl(x_shape), name="ac_cXY_%d" % i))
            elif cond == "condXYG":
                b.append(CondCoupling("XYG", st, scope, x_shape, tmpl_iso(x_shape), name="ac_cXYG_%d" % i))
            elif cond == "condSDN":
                b.append(scale("camsdn", scope, "ac_cSDN_%d" % i))
            elif cond == "fitSDN":
                b.append(scale("sdngain", scope, "ac_fitSDN_%d" % i))
            else:                                                                     # uncond | unc_sdn
                b.append(AffineCoupling(st, scope, x_shape, tmpl(x_shape), name="ac_unc_%d" % i))
        if getattr(h, "append_sdn", False):                                           # :379-390
            b.append(scale("sdngain", "%s/bijector%d" % (name, depth), "ac_fitSDN_%d" % depth))
        return b

    # ---- noise_flow_model.py:71-235
    def noise_flow_arch(self, name, x_shape, flow_permutation, arch):
        bijectors = []
        for i, lyr in enumerate(arch.split("|")):
            scope = "%s/bijector%d" % (name, i)
            if lyr == "unc":
                if flow_permutation == 0:
                    bijectors.append(Permute(x_shape[-1], name="permute"))            # :80-84
                elif flow_permutation == 1:
                    bijectors.append(Conv2d1x1(self.store, scope, x_shape, decomp=self.hps.decomp,
                                               layer_id=i, name="Conv2d_1x1_%d" % i))  # :85-90
                bijectors.append(AffineCoupling(
                    self.store, scope, x_shape,
                    RealNVPConvTemplate(self.store, x_shape, self.hps.width),
                    name="unc_%d" % i))                                                # :95-104
            elif lyr in _SCALE_TOKENS:
                pre = "gain" if lyr.startswith("gain") else "sdn"
                bijectors.append(ScaleBijector(_SCALE_TOKENS[lyr], self.store, scope, x_shape, 0,
                                               "%s_%d" % (pre, i), getattr(self.hps, "gain_init", 0.0),
                                               self.hps.param_inits))                  # :106-234
            # unknown tokens are silently skipped by the reference's if/elif chain
        return bijectors

    def get_layer_names(self):                                                         # :508-513
        return [b.name for b in self.model[0]]

    def _t(self, a):
        return torch.as_tensor(np.asarray(a), dtype=self.dtype) if not isinstance(a, torch.Tensor) else a.to(self.dtype)

    # ---- :394-428
    def inverse(self, x, objective, yy=None, nlf0=None, nlf1=None, iso=None, cam=None, is_training=False):
        z = self._t(x)
        yy = self._t(yy) if yy is not None else None
        sf = self.hps.squeeze_factor
        z = squeeze2d(z, sf, self.hps.squeeze_type)
        if yy is not None:
            yy = squeeze2d(yy, sf, self.hps.squeeze_type)
        for b in self.model[0]:
            if isinstance(b, ScaleBijector):
                z, ldj = b._inverse_and_log_det_jacobian(z, yy, nlf0, nlf1, iso, cam)
            elif isinstance(b, CondCoupling):
                z, ldj = b._inverse_and_log_det_jacobian(z, yy, nlf0, nlf1, iso, cam, is_training)
            elif isinstance(b, AffineCoupling):
                z, ldj = b._inverse_and_log_det_jacobian(z, is_training)
            else:
                z, ldj = b._inverse_and_log_det_jacobian(z)
            objective = objective + ldj                                                # :425
        return z, objective

    # ---- :430-447
    def forward(self, z, eps_std=None, yy=None, nlf0=None, nlf1=None, iso=None, cam=None, is_training=False):
        x = self._t(z)
        yy = self._t(yy) if yy is not None else None
        for b in reversed(self.model[0]):
            if isinstance(b, ScaleBijector):
                x = b._forward(x, yy, nlf0, nlf1, iso, cam)
            elif isinstance(b, CondCoupling):
                x = b._forward(x, yy, nlf0, nlf1, iso, cam, is_training)
            elif isinstance(b, AffineCoupling):
                x = b._forward(x, is_training)
            else:
                x = b._forward(x)
        return unsqueeze2d(x, self.hps.squeeze_factor, self.hps.squeeze_type)

    # ---- :449-456, :499-504, :525-541.  ``eps`` replaces tf.random_normal so parity tests can inject it.
    def sample(self, eps, eps_std=None, yy=None, nlf0=None, nlf1=None, iso=None, cam=None, is_training=False):
        eps = self._t(eps)
        z = eps * eps_std if eps_std is not None else eps                             # :501 (mean 0, logsd 0)
        return self.forward(z, eps_std, yy, nlf0, nlf1, iso, cam, is_training)

    # ---- :458-480
    def _loss(self, x, y, nlf0=None, nlf1=None, iso=None, cam=None, is_training=False):
        x = self._t(x)
        objective = torch.zeros(x.shape[0], dtype=self.dtype)
        cond = getattr(self.hps, "sidd_cond", "mix")
        if cond is not None and cond != "uncond":
            z, objective = self.inverse(x, objective, yy=y, nlf0=nlf0, nlf1=nlf1, iso=iso, cam=cam,
                                        is_training=is_training)
        else:
            z, objective = self.inverse(x, objective, is_training=is_training)
        logp = (-0.5 * (LOG_2PI + z ** 2)).sum(dim=(1, 2, 3))                        # :537-539 (mean 0, logsd 0)
        objective = objective + logp                                                  # :474
        nobj = -objective                                                             # :475
        var_z = ((z - z.mean(dim=(1, 2, 3), keepdim=True)) ** 2).mean(dim=(1, 2, 3))  # :477
        sd_z = torch.sqrt(var_z).mean()                                               # :478
        self.last_z = z
        return nobj, sd_z

    def loss(self, x, y, nlf0=None, nlf1=None, iso=None, cam=None, is_training=False):  # :482-484
        nll, sd_z = self._loss(x, y, nlf0, nlf1, iso, cam, is_training)
        return nll.mean(), sd_z


def make_hps(**kw):
    """Minimal hps namespace with the hot-path keys (SURVEY section 5 'Config / flags')."""
    d = dict(arch="sdn5|unc|unc|unc|unc|gain4|unc|unc|unc|unc", width=4, flow_permutation=1, decomp="LU",
             squeeze_factor=1, squeeze_type="chessboard", n_levels=1, depth=-1, gain_init=-5.0,
             sidd_cond="mix", x_shape=[None, 32, 32, 4])
    d.update(kw)
    return SimpleNamespace(**d)


# =============================================================================================
# closed-form baselines the reference itself uses as sanity checks
# =============================================================================================
def nll_sdn_closed_form(x, y, nlf0, nlf1):
    """sidd/PatchStatsCalculator.py:104-115: per-patch NLL of N(0, nlf0*y + nlf1)."""
    x = np.asarray(x, dtype=np.float64)
    var = np.asarray(y, dtype=np.float64) * nlf0 + nlf1
    return 0.5 * ((LOG_2PI + np.log(var)) + x ** 2 / var).reshape(x.shape[0], -1).sum(1)


# =============================================================================================
# Philox4x32-10 + Box-Muller: oracle of the in-kernel sampler RNG (include/noiseflow_b200.h, nf_sample)
# =============================================================================================
def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """counter: [..., 4] uint32, key: [..., 2] uint32 -> [..., 4] uint32 (Salmon et al. 2011)."""
    c = counter.astype(np.uint64).copy()
    k = key.astype(np.uint64).copy()
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c[..., 0]
        p1 = M1 * c[..., 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = np.stack([hi1 ^ c[..., 1] ^ k[..., 0], lo1, hi0 ^ c[..., 3] ^ k[..., 1], lo0], axis=-1)
        k = np.stack([(k[..., 0] + W0) & mask, (k[..., 1] + W1) & mask], axis=-1)
    return c.astype(np.uint32)


def philox_normal(seed: int, offset: int, n_patches: int, first_patch: int = 0) -> np.ndarray:
    """eps[n, h, w, c] exactly as the CUDA sampler draws it: one Philox call per pixel,
    counter = (pixel index h*32+w, patch index low, patch index high, offset), key = seed (lo, hi);
    the four 32-bit outputs feed two Box-Muller pairs -> channels (0,1) and (2,3)."""
    pix = np.arange(1024, dtype=np.uint64)
    pat = np.arange(first_patch, first_patch + n_patches, dtype=np.uint64)
    ctr = np.zeros((n_patches, 1024, 4), dtype=np.uint32)
    ctr[..., 0] = pix[None, :].astype(np.uint32)
    ctr[..., 1] = (pat[:, None] & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[..., 2] = (pat[:, None] >> np.uint64(32)).astype(np.uint32)
    ctr[..., 3] = np.uint32(offset & 0xFFFFFFFF)
    key = np.zeros((n_patches, 1024, 2), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    r = philox4x32_10(ctr, key).astype(np.float64)
    u = (r + 0.5) * (1.0 / 4294967296.0)          # (0,1) open interval
    out = np.empty((n_patches, 1024, 4), dtype=np.float64)
    for a in (0, 2):
        rad = np.sqrt(-2.0 * np.log(u[..., a]))
        ang = 2.0 * np.pi * u[..., a + 1]
        out[..., a] = rad * np.cos(ang)
        out[..., a + 1] = rad * np.sin(ang)
    return out.reshape(n_patches, 32, 32, 4)


# =============================================================================================
# evaluation metrics next to the path (sidd/PatchStatsCalculator.py:92-123, sidd/sidd_utils.py:1202-1274)
# =============================================================================================
def calc_baselines(x, y, nlf0, nlf1, var_gauss):
    """PatchStatsCalculator.py:100-115: per-patch Gaussian and camera-NLF NLL (before the batch mean)."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    vr = y * nlf0 + nlf1                                                              # :104
    nll_g = 0.5 * (np.log(2 * np.pi) + np.log(var_gauss) + x ** 2 / var_gauss)        # :107-108
    nll_s = 0.5 * (np.log(2 * np.pi) + np.log(vr) + x ** 2 / vr)                      # :112-113
    return nll_g.sum(axis=(1, 2, 3)), nll_s.sum(axis=(1, 2, 3))


def get_histogram(data, bin_edges):
    """sidd_utils.py:1266-1274."""
    n = np.prod(data.shape)
    hist, _ = np.histogram(data, bin_edges)
    return hist / n


def kl_div_forward(p, q):
    """sidd_utils.py:1202-1209."""
    idx = ~(np.isnan(p) | np.isinf(p) | np.isnan(q) | np.isinf(q))
    p, q = p[idx], q[idx]
    idx = (p > 0) & (q > 0)
    p, q = p[idx], q[idx]
    return np.sum(p * np.log(p / q))
