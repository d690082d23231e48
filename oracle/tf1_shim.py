"""TEST INFRASTRUCTURE -- a minimal *eager* stand-in for the TensorFlow 1.12 / TensorFlow-Probability 0.5 API
surface that the reference's hot-path Python touches, so that the reference's OWN source files
(`/root/reference/borealisflows/*.py`, unmodified, imported from where they lie) can be executed in this container and
golden input/output vectors can be generated from them (`oracle/make_reference_goldens.py` -> `tests/golden/ref_*.npz`).

TensorFlow itself cannot be installed here (Python 3.12, no network; the reference pins TF 1.12 / TFP 0.5,
README.md:15-19).  What this file restates is therefore ONLY the framework primitives -- `tf.nn.conv2d`,
`tf.nn.moments`, `tf.get_variable` / `tf.variable_scope` / `tf.make_template` naming rules, `tf.where`, `tf.one_hot`,
`tfd.fill_triangular`, ... -- each following TensorFlow's documented semantics (noted per function).  Everything the
reference authors wrote (architecture parsing, every bijector formula, BatchNorm, edge padding, LU parameterisation,
loss, sampling, all the quirks) runs from the reference's files, not from a restatement.

Execution model: a small deferred graph, like TF 1.x.  Every `Tensor` is a node (function + input nodes + control
dependencies) wrapping a `torch.Tensor` value (float64 by default: the goldens are the reference's formulas in double
precision; `set_float_dtype(torch.float32)` mimics TF's fp32 arithmetic).  Nodes are evaluated once when they are
created -- on the placeholders' one-element shadow values, only so that static shapes exist -- and re-evaluated, memoised
per run, by `Session.run(fetches, feed_dict)`; stateful ops (`assign_sub`, the optimiser) execute only inside a run;
`tf.cond` builds both branches and evaluates the taken one.  So `tf.placeholder`, `tf.Session`, `tf.train.Saver.restore`
and therefore the reference's `NoiseFlowWrapper` class itself work as written.  Variables are torch leaves, so
`tf.gradients` is torch autograd.

Nothing outside `tests/`, `oracle/` and `tools/` may import this module; the product never does.
"""
import contextlib
import math
import sys
import types

import numpy as np
import torch

_FLOAT = torch.float64


def set_float_dtype(dt):
    global _FLOAT
    _FLOAT = dt


# ----------------------------------------------------------------------------------------------------- shapes
class Dimension(int):
    """tf.Dimension: an int that prints like one (unknown dimensions do not occur in eager mode)."""
    @property
    def value(self):
        return int(self)


class TensorShape:
    def __init__(self, dims):
        self._d = [Dimension(d) for d in dims]

    def as_list(self):
        return [int(d) for d in self._d]

    def is_fully_defined(self):
        return True

    def concatenate(self, other):
        return TensorShape(self._d + list(TensorShape(other)._d if not isinstance(other, TensorShape) else other._d))

    @property
    def ndims(self):
        return len(self._d)

    def __len__(self):
        return len(self._d)

    def __iter__(self):
        return iter(self._d)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return TensorShape(self._d[k])
        return self._d[k]

    def __eq__(self, other):
        try:
            return self.as_list() == [int(d) for d in other]
        except TypeError:
            return False

    def __ne__(self, other):
        return not self.__eq__(other)

    def __repr__(self):
        return "TensorShape(%r)" % self.as_list()


# ----------------------------------------------------------------------------------------------------- dtypes
class DType:
    def __init__(self, name, is_float, torch_dtype):
        self.name, self.is_floating, self._torch = name, is_float, torch_dtype

    def __repr__(self):
        return "tf." + self.name

    def __eq__(self, other):
        return other is not None and _as_dtype(other).name == self.name

    def __hash__(self):
        return hash(self.name)


float32 = DType("float32", True, None)
float64 = DType("float64", True, None)
int32 = DType("int32", False, torch.int64)
int64 = DType("int64", False, torch.int64)
bool_ = DType("bool", False, torch.bool)
_DTYPES = {"float32": float32, "float64": float64, "int32": int32, "int64": int64, "bool": bool_}


def _as_dtype(d):
    if isinstance(d, DType):
        return d
    if isinstance(d, str):
        return _DTYPES[d]
    if d is None:
        return float32
    return _DTYPES[np.dtype(d).name]


def _torch_dtype(d):
    d = _as_dtype(d)
    return _FLOAT if d.is_floating else d._torch


# ------------------------------------------------------------------------------------------------------ graph
class _Scope:
    def __init__(self, name, reuse):
        self.name, self.reuse = name, reuse


AUTO_REUSE = "AUTO_REUSE"


class _Graph:
    """What a default tf.Graph carries: variables by full name (creation order kept), collections, the per-scope
    counters `variable_scope(None, default_name=...)` uses to make names unique, the active control dependencies, and
    the run counter that memoises node values."""

    def __init__(self):
        self.vars = {}
        self.collections = {}
        self.scope_stack = [_Scope("", False)]
        self.default_name_counts = {}
        self.ctrl_stack = []
        self.run_id = 0
        self.running = False               # True inside Session.run: stateful ops take effect
        self.random_normal_hook = None     # callable(shape) -> array: lets a test inject tf.random_normal draws
        self.update_log = []               # (variable name, new value) of every executed tf.assign_sub

    def as_default(self):
        return contextlib.nullcontext(self)


_G = _Graph()


def reset_default_graph():
    global _G
    _G = _Graph()


def get_default_graph():
    return _G


def Graph():
    return _Graph()


# ----------------------------------------------------------------------------------------------------- tensors
def _const(v):
    """python / numpy value -> torch.Tensor (floating data in the working precision)."""
    if isinstance(v, torch.Tensor):
        return v
    if isinstance(v, (bool, np.bool_)):
        return torch.tensor(bool(v))
    a = np.asarray(v)
    if a.dtype.kind == "f":
        return torch.tensor(a, dtype=_FLOAT)
    if a.dtype.kind in "iu":
        return torch.tensor(a.astype(np.int64))
    if a.dtype.kind == "b":
        return torch.tensor(a)
    raise TypeError("cannot convert %r to a tensor" % (v,))


def _resolve(x):
    """node inputs -> values: Tensors are evaluated, lists / tuples / slices are walked, the rest passes through."""
    if isinstance(x, Tensor):
        return x.t
    if isinstance(x, (list, tuple)):
        return type(x)(_resolve(e) for e in x)
    if isinstance(x, slice):
        return slice(_resolve(x.start), _resolve(x.stop), _resolve(x.step))
    return x


def _pair(a, b):
    """TF converts a python / numpy operand to the dtype of the tensor operand; integer tensors meeting floats are
    promoted (only index arithmetic does that here)."""
    a, b = _const(a), _const(b)
    if a.dtype != b.dtype and torch.bool not in (a.dtype, b.dtype):
        if a.dtype.is_floating_point != b.dtype.is_floating_point:
            a, b = (a, b.to(a.dtype)) if a.dtype.is_floating_point else (a.to(b.dtype), b)
    return a, b


class Tensor:
    __array_priority__ = 1000
    __array_ufunc__ = None          # numpy defers to the reflected operators below

    def __init__(self, fn, inputs=(), value=None, eager=True, name=None):
        self._fn, self._inputs = fn, tuple(inputs)
        self._ctrl = tuple(d for ds in _G.ctrl_stack for d in ds)
        self._graph = _G
        self._val, self._run = value, _G.run_id
        self.name = name or "Tensor:0"
        if fn is not None and value is None and eager:
            self._val = self._compute()
        elif fn is not None and not eager:
            self._run = -1

    def _compute(self):
        for d in self._ctrl:
            _resolve(d)
        return self._fn(*[_resolve(i) for i in self._inputs])

    @property
    def t(self):
        """value in the current run (memoised per Session.run; constants / variables / placeholders hold theirs)."""
        if self._fn is not None and self._run != self._graph.run_id:
            self._run = self._graph.run_id
            self._val = self._compute()
        return self._val

    # shape / dtype (static shape == shape of the current value; placeholders' unknown dimensions are 1 at build time)
    def get_shape(self):
        return TensorShape(self.t.shape)

    @property
    def shape(self):
        return TensorShape(self.t.shape)

    @property
    def dtype(self):
        if self.t.dtype == torch.bool:
            return bool_
        return float32 if self.t.dtype.is_floating_point else int64

    @property
    def op(self):
        return self

    def numpy(self):
        return self.t.detach().cpu().numpy()

    def __bool__(self):
        raise TypeError("Using a tf.Tensor as a Python bool is not allowed (TF 1.12 graph mode)")

    def __len__(self):
        raise TypeError("len() of a tf.Tensor is not defined (TF 1.12 graph mode)")

    def __iter__(self):
        raise TypeError("Tensor objects are not iterable in graph mode (TF 1.12)")

    def __getitem__(self, k):
        return Tensor(lambda t, kk: t[kk], (self, k))

    def _b(self, o, fn, swap=False):
        if swap:
            return Tensor(lambda a, b: fn(*_pair(b, a)), (self, o))
        return Tensor(lambda a, b: fn(*_pair(a, b)), (self, o))

    def __add__(self, o): return self._b(o, torch.add)
    def __radd__(self, o): return self._b(o, torch.add, True)
    def __sub__(self, o): return self._b(o, torch.sub)
    def __rsub__(self, o): return self._b(o, torch.sub, True)
    def __mul__(self, o): return self._b(o, torch.mul)
    def __rmul__(self, o): return self._b(o, torch.mul, True)
    def __truediv__(self, o): return self._b(o, torch.div)
    def __rtruediv__(self, o): return self._b(o, torch.div, True)
    def __pow__(self, o): return self._b(o, torch.pow)
    def __neg__(self): return Tensor(torch.neg, (self,))
    def __abs__(self): return Tensor(torch.abs, (self,))
    def __matmul__(self, o): return self._b(o, torch.matmul)
    def __lt__(self, o): return self._b(o, torch.lt)
    def __le__(self, o): return self._b(o, torch.le)
    def __gt__(self, o): return self._b(o, torch.gt)
    def __ge__(self, o): return self._b(o, torch.ge)
    __hash__ = object.__hash__      # TF 1.x tensors hash by identity; == is NOT overloaded in TF 1.12

    def __repr__(self):
        return "<shim Tensor shape=%s>" % (list(self.t.shape),)


def _T(value):
    """constant node"""
    return Tensor(None, (), value=_const(value))


def _op(fn, *inputs, eager=True):
    return Tensor(fn, inputs, eager=eager)


class Variable(Tensor):
    def __init__(self, name, value, trainable):
        super().__init__(None, (), value=value.detach().clone().requires_grad_(bool(trainable)), name=name + ":0")
        self.var_name, self.trainable = name, trainable

    @property
    def op(self):
        return types.SimpleNamespace(name=self.var_name)

    def load(self, value):
        """Assigns a new value.  The leaf is REPLACED (not written in place) so that autograd graphs of the current
        run stay valid; nodes read the variable through `.t` and see the new leaf from then on."""
        v = torch.as_tensor(np.asarray(value)) if not isinstance(value, torch.Tensor) else value
        v = v.detach().to(self._val.dtype).clone()
        assert tuple(v.shape) == tuple(self._val.shape), (self.var_name, tuple(v.shape), tuple(self._val.shape))
        self._val = v.requires_grad_(bool(self.trainable))


def placeholder(dtype, shape=None, name=None):
    """Fed by Session.run(feed_dict=...).  Until then it holds a shadow value (unknown dimensions = 1, zeros / False)
    whose only purpose is to give the graph-construction code static shapes."""
    dims = [1 if d is None else int(d) for d in (shape if shape is not None else [])]
    t = Tensor(None, (), value=torch.zeros(dims, dtype=_torch_dtype(dtype)), name=(name or "Placeholder") + ":0")
    t._placeholder_dtype = _torch_dtype(dtype)
    return t


class Session:
    """tf.Session: `run` starts a new evaluation (every node is recomputed at most once from the fed placeholders and
    the variables' current values) with stateful ops enabled."""

    def __init__(self, graph=None, config=None, target=""):
        self.graph = graph or _G

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def close(self):
        pass

    def run(self, fetches, feed_dict=None):
        g = _G
        g.run_id += 1
        for ph, val in (feed_dict or {}).items():
            assert isinstance(ph, Tensor) and hasattr(ph, "_placeholder_dtype"), "can only feed placeholders"
            ph._val = _const(val).to(ph._placeholder_dtype)
        g.running = True
        try:
            def ev(f):
                if isinstance(f, (list, tuple)):
                    return type(f)(ev(e) for e in f)
                if isinstance(f, dict):
                    return {k: ev(v) for k, v in f.items()}
                if f is None:
                    return None
                v = f.t
                return v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v
            return ev(fetches)
        finally:
            g.running = False


def global_variables_initializer():
    return _op(lambda: None, eager=False)       # variables take their initial value when they are created


checkpoint_reader = None        # callable(path) -> {variable name: ndarray}; set by the script that drives the shim


class Saver:
    """tf.train.Saver(): restore assigns, to every variable of the graph, the checkpoint tensor of the same name and
    fails if one is missing (TF raises NotFoundError)."""

    def __init__(self, var_list=None, **_):
        self._vars = var_list

    def restore(self, sess, save_path):
        assert checkpoint_reader is not None, "tf1_shim.checkpoint_reader is not set"
        values = checkpoint_reader(save_path)
        for v in (self._vars or list(_G.vars.values())):
            if v.var_name not in values:
                raise KeyError("Key %s not found in checkpoint" % v.var_name)
            v.load(np.asarray(values[v.var_name]).reshape(tuple(v._val.shape)))
        self.last_unused = sorted(set(values) - {v.var_name for v in _G.vars.values()})


# ----------------------------------------------------------------------------------------- variables and scopes
class variable_scope:
    """tf.variable_scope(name_or_scope, default_name=None, reuse=None).  Rules followed (TF 1.12
    `variable_scope.py`): a string opens `<current>/<name>`; a captured scope object re-enters its absolute name;
    `name_or_scope=None` opens `default_name` made unique among the default names already handed out in the current
    scope (`foo`, `foo_1`, ...); `reuse=None/False` inherits the enclosing flag, True / AUTO_REUSE override and are
    inherited by nested scopes (`self._reuse or self._old.reuse`)."""

    def __init__(self, name_or_scope, default_name=None, reuse=None, custom_getter=None, **_):
        self._arg, self._default, self._reuse = name_or_scope, default_name, reuse

    def __enter__(self):
        cur = _G.scope_stack[-1]
        if isinstance(self._arg, _Scope):
            name = self._arg.name
        else:
            if self._arg is None:
                key = (cur.name, self._default)
                n = _G.default_name_counts.get(key, 0)
                _G.default_name_counts[key] = n + 1
                leaf = self._default if n == 0 else "%s_%d" % (self._default, n)
            else:
                leaf = self._arg
            name = leaf if not cur.name else cur.name + "/" + leaf
        sc = _Scope(name, self._reuse or cur.reuse)
        _G.scope_stack.append(sc)
        return sc

    def __exit__(self, *a):
        sc = _G.scope_stack.pop()
        if not isinstance(self._arg, _Scope):
            # TF 1.12 `_pure_variable_scope.__exit__` -> `close_variable_subscopes`: leaving a scope entered by name
            # forgets which default names were handed out below it
            for key in _G.default_name_counts:
                if key[0] == sc.name or key[0].startswith(sc.name + "/"):
                    _G.default_name_counts[key] = 0
        return False


def get_variable_scope():
    return _G.scope_stack[-1]


@contextlib.contextmanager
def name_scope(name=None, default_name=None, values=None):
    yield name or default_name      # op names only; variable names are not affected (TF 1.x)


@contextlib.contextmanager
def control_dependencies(deps):
    _G.ctrl_stack.append(list(deps or []))
    try:
        yield
    finally:
        _G.ctrl_stack.pop()


def constant_initializer(value=0.0, dtype=None):
    def init(shape):
        v = np.asarray(value, dtype=np.float64)
        if v.size == 1:
            return torch.full(tuple(shape), float(v.reshape(-1)[0]), dtype=_FLOAT)
        assert v.size == int(np.prod(shape)), "constant_initializer: %d values for shape %s" % (v.size, shape)
        return torch.tensor(v.reshape(tuple(shape)), dtype=_FLOAT)      # row-major fill
    return init


class zeros_initializer:        # used both as `tf.zeros_initializer` (the class) and `tf.zeros_initializer()`
    def __call__(self, shape):
        return torch.zeros(tuple(shape), dtype=_FLOAT)


class ones_initializer:
    def __call__(self, shape):
        return torch.ones(tuple(shape), dtype=_FLOAT)


_INIT_RNG = np.random.RandomState(1234)


def random_normal_initializer(mean=0.0, stddev=1.0, seed=None, dtype=None):
    def init(shape):
        return torch.tensor(mean + stddev * _INIT_RNG.randn(*shape), dtype=_FLOAT)
    return init


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **_):
    sc = _G.scope_stack[-1]
    full = name if not sc.name else sc.name + "/" + name
    if full in _G.vars:
        if sc.reuse not in (True, AUTO_REUSE):
            raise ValueError("Variable %s already exists, disallowed. Did you mean to set reuse=True or "
                             "reuse=tf.AUTO_REUSE in VarScope?" % full)
        return _G.vars[full]
    if sc.reuse is True:
        raise ValueError("Variable %s does not exist, or was not created with tf.get_variable(). Did you mean to set "
                         "reuse=tf.AUTO_REUSE in VarScope?" % full)
    if initializer is None:
        raise NotImplementedError("get_variable(%s) without an initializer (glorot default) is not on the path" % full)
    if isinstance(initializer, type):
        initializer = initializer()
    if isinstance(initializer, Tensor):
        value = initializer.t
    elif callable(initializer):
        value = initializer([int(d) for d in shape])
    else:
        value = _const(initializer)
    if value.dtype.is_floating_point:
        value = value.to(_FLOAT)
    v = Variable(full, value, bool(trainable))
    _G.vars[full] = v
    return v


def trainable_variables():
    return [v for v in _G.vars.values() if v.trainable]


def global_variables():
    return list(_G.vars.values())


class GraphKeys:
    GLOBAL_VARIABLES = "variables"
    TRAINABLE_VARIABLES = "trainable_variables"
    UPDATE_OPS = "update_ops"


def get_collection(name, scope=None):
    if name == GraphKeys.GLOBAL_VARIABLES:
        return global_variables()
    if name == GraphKeys.TRAINABLE_VARIABLES:
        return trainable_variables()
    return list(_G.collections.get(name, []))


def add_to_collection(name, value):
    _G.collections.setdefault(name, []).append(value)


class _Template:
    """tf.make_template(name, fn) with create_scope_now_=False (TF 1.12 `template.py`): the variable scope is opened
    -- and its name made unique in the scope that is current THEN -- at the first call; later calls re-enter that
    captured scope with reuse=True."""

    def __init__(self, name, fn):
        self._name, self._fn, self._scope = name, fn, None

    def __call__(self, *a, **k):
        if self._scope is None:
            with variable_scope(None, default_name=self._name) as vs:
                self._scope = vs
                return self._fn(*a, **k)
        with variable_scope(self._scope, reuse=True):
            return self._fn(*a, **k)

    @property
    def variable_scope(self):
        return self._scope


def make_template(name, func, create_scope_now_=False, **kw):
    assert not create_scope_now_ and not kw
    return _Template(name, func)


# --------------------------------------------------------------------------------------------------------- ops
def convert_to_tensor(value, dtype=None, name=None):
    return value if isinstance(value, Tensor) else _T(value)


def constant(value, dtype=None, shape=None, name=None):
    t = _const(value)
    if dtype is not None:
        t = t.to(_torch_dtype(dtype))
    if shape is not None:
        t = t.expand(tuple(shape)).clone() if t.numel() == 1 else t.reshape(tuple(shape))
    return _T(t)


def _ints(shape):
    """run-time shape argument (already resolved: torch tensors / ints / TensorShape) -> list of ints"""
    if isinstance(shape, torch.Tensor):
        return [int(v) for v in shape.reshape(-1)]
    if isinstance(shape, TensorShape):
        return shape.as_list()
    return [int(s) for s in shape]


def reshape(x, shape, name=None):
    return _op(lambda t, s: _const(t).reshape(_ints(s)), x, shape)


def transpose(x, perm=None, name=None):
    return _op(lambda t: t.permute(*perm) if perm is not None else t.permute(*reversed(range(t.dim()))), x)


def concat(values, axis, name=None):
    def fn(ts):
        ts = [_const(t) for t in ts]
        if any(t.dtype.is_floating_point for t in ts):
            ts = [t.to(_FLOAT) for t in ts]
        return torch.cat(ts, dim=axis)
    return _op(fn, list(values))


def split(value, num_or_size_splits, axis=0, name=None):
    n = num_or_size_splits
    if isinstance(n, int):
        return [_op(lambda t, i=i: torch.chunk(t, n, dim=axis)[i], value) for i in range(n)]
    return [_op(lambda t, i=i: torch.split(t, list(n), dim=axis)[i], value) for i in range(len(n))]


def shape(x, name=None):
    return _op(lambda t: torch.tensor(list(_const(t).shape), dtype=torch.int64), x)


def zeros(shape, dtype=float32, name=None):
    return _op(lambda s: torch.zeros(_ints(s), dtype=_torch_dtype(dtype)), shape)


def ones(shape, dtype=float32, name=None):
    return _op(lambda s: torch.ones(_ints(s), dtype=_torch_dtype(dtype)), shape)


def zeros_like(x, dtype=None, name=None):
    return _op(lambda t: torch.zeros_like(_const(t), dtype=_torch_dtype(dtype) if dtype is not None else None), x)


def ones_like(x, dtype=None, name=None):
    return _op(lambda t: torch.ones_like(_const(t), dtype=_torch_dtype(dtype) if dtype is not None else None), x)


def eye(n, m=None, dtype=float32, **_):
    return _T(torch.eye(n, m if m is not None else n, dtype=_torch_dtype(dtype)))


def diag(v, name=None):
    return _op(lambda t: torch.diag(_const(t)), v)


def cast(x, dtype, name=None):
    return _op(lambda t: _const(t).to(_torch_dtype(dtype)), x)


def tile(x, multiples, name=None):
    return _op(lambda t, m: _const(t).repeat(*_ints(m)), x, multiples)


def pad(x, paddings, mode="CONSTANT", name=None, constant_values=0):
    flat = []
    for lo, hi in reversed([list(p) for p in paddings]):      # torch pads from the last dimension backwards
        flat += [int(lo), int(hi)]
    return _op(lambda t: torch.nn.functional.pad(_const(t), flat, value=constant_values), x)


def _ew(fn):
    def op(x, name=None):
        return _op(lambda t: fn(_const(t)), x)
    return op


exp = _ew(torch.exp)
log = _ew(torch.log)
sqrt = _ew(torch.sqrt)
tanh = _ew(torch.tanh)
abs = _ew(torch.abs)            # noqa: A001  (tf.abs)
sigmoid = _ew(torch.sigmoid)
relu = _ew(torch.relu)
identity = _ew(lambda t: t)
stop_gradient = _ew(lambda t: t.detach())


def _bin(fn):
    def op(a, b, name=None):
        return _op(lambda x, y: fn(*_pair(x, y)), a, b)
    return op


add = _bin(torch.add)
subtract = _bin(torch.sub)
multiply = _bin(torch.mul)
divide = _bin(torch.div)
equal = _bin(torch.eq)
greater_equal = _bin(torch.ge)
matmul = _bin(torch.matmul)


def Print(x, data=None, message=None, **_):
    return x


def _axes(axis):
    if axis is None:
        return None
    return tuple(axis) if isinstance(axis, (list, tuple)) else (int(axis),)


def reduce_sum(x, axis=None, keepdims=False, name=None, reduction_indices=None, keep_dims=None):
    ax = _axes(axis if axis is not None else reduction_indices)
    kd = bool(keepdims or keep_dims)
    return _op(lambda t: _const(t).sum() if ax is None else _const(t).sum(dim=ax, keepdim=kd), x)


def reduce_mean(x, axis=None, keepdims=False, name=None, reduction_indices=None, keep_dims=None):
    ax = _axes(axis if axis is not None else reduction_indices)
    kd = bool(keepdims or keep_dims)
    return _op(lambda t: _const(t).mean() if ax is None else _const(t).mean(dim=ax, keepdim=kd), x)


def where(condition, x=None, y=None, name=None):
    if x is None and y is None:
        return _op(lambda c: torch.nonzero(c), condition)     # [k, rank] int64 coordinates of the true elements
    return _op(lambda c, a, b: torch.where(c, *_pair(a, b)), condition, x, y)


def one_hot(indices, depth, on_value=1.0, off_value=0.0, axis=-1, dtype=None, name=None):
    def fn(idx):
        idx = _const(idx).to(torch.int64)
        ok = (idx >= 0) & (idx < int(depth))                  # out-of-range indices give all-off rows (TF semantics)
        out = torch.nn.functional.one_hot(torch.where(ok, idx, torch.zeros_like(idx)), int(depth)).to(_FLOAT)
        out = out * ok.unsqueeze(-1)
        return out * on_value + (1 - out) * off_value
    return _op(fn, indices)


class _Cond(Tensor):
    """tf.cond: both branches are built (as TF does), only the taken one is evaluated in a run."""

    def __init__(self, pred, a, b):
        self._pred, self._a, self._b2 = pred, a, b
        super().__init__(self._select, ())

    def _select(self):
        return (self._a if bool(_const(_resolve(self._pred))) else self._b2).t


def cond(pred, true_fn=None, false_fn=None, name=None, fn1=None, fn2=None, strict=False):
    a, b = (true_fn or fn1)(), (false_fn or fn2)()
    assert isinstance(a, Tensor) and isinstance(b, Tensor), "only single-tensor branches occur on the path"
    return _Cond(pred, a, b)


def gather(params, indices, axis=0, name=None):
    def fn(p, i):
        return torch.index_select(p, axis if axis >= 0 else p.dim() + axis, _const(i).to(torch.int64).reshape(-1))
    return _op(fn, params, indices)


def random_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):
    g = _G

    def fn(s):
        shp = _ints(s)
        if g.random_normal_hook is not None and g.running:
            e = torch.as_tensor(np.asarray(g.random_normal_hook(shp))).to(_FLOAT).reshape(shp)
        else:
            e = torch.randn(shp, dtype=_FLOAT)
        return e * stddev + mean
    return _op(fn, shape)


def assign_sub(ref, value, name=None):
    """ref <- ref - value when executed inside a run (graph construction leaves the variable alone); recorded in
    graph.update_log so tests can see the moving-average side effects of a run."""
    assert isinstance(ref, Variable)
    g = _G

    def fn(v):
        if g.running:
            ref.load(ref._val.detach() - v.detach())
            g.update_log.append((ref.var_name, ref._val.detach().clone().numpy()))
        return ref._val
    return _op(fn, value)


def matrix_set_diag(x, diagonal, name=None):
    def fn(t, d):
        mask = torch.eye(t.shape[-1], dtype=torch.bool)
        return torch.where(mask, torch.diag_embed(_const(d)).expand_as(t), t)
    return _op(fn, x, diagonal)


def _band(t, num_lower, num_upper):
    n, m = t.shape[-2], t.shape[-1]
    i = torch.arange(n).unsqueeze(1)
    j = torch.arange(m).unsqueeze(0)
    keep = torch.ones(n, m, dtype=torch.bool)
    if num_lower >= 0:
        keep &= (i - j) <= num_lower
    if num_upper >= 0:
        keep &= (j - i) <= num_upper
    return torch.where(keep, t, torch.zeros_like(t))


def matrix_band_part(x, num_lower, num_upper, name=None):
    return _op(lambda t: _band(_const(t), num_lower, num_upper), x)


def matrix_inverse(x, name=None):
    return _op(lambda t: torch.linalg.inv(t), x)


def triangular_solve(matrix, rhs, lower=True, adjoint=False, name=None):
    def fn(a, b):
        lo = lower
        if adjoint:
            a, lo = a.transpose(-1, -2), not lo
        return torch.linalg.solve_triangular(a, b, upper=not lo)
    return _op(fn, matrix, rhs)


def slogdet(x, name=None):
    return _op(lambda t: torch.linalg.slogdet(t)[0], x), _op(lambda t: torch.linalg.slogdet(t)[1], x)


def l2_normalize(x, axis=None, epsilon=1e-12, name=None, dim=None):
    ax = _axes(axis if axis is not None else dim)
    return _op(lambda t: t * torch.rsqrt(torch.clamp((t * t).sum(dim=ax, keepdim=True), min=epsilon)), x)


def moments(x, axes, shift=None, name=None, keep_dims=False):
    """tf.nn.moments (TF 1.12 `nn_impl.py`): mean, and the POPULATION variance as
    mean(squared_difference(x, stop_gradient(mean)))."""
    ax = tuple(axes)

    def mean_fn(t):
        m = t.mean(dim=ax, keepdim=True)
        return m if keep_dims else m.squeeze(ax)

    def var_fn(t):
        m = t.mean(dim=ax, keepdim=True)
        v = ((t - m.detach()) ** 2).mean(dim=ax, keepdim=True)
        return v if keep_dims else v.squeeze(ax)
    return _op(mean_fn, x), _op(var_fn, x)


def conv2d(input, filter, strides, padding, use_cudnn_on_gpu=True, data_format="NHWC", dilations=None, name=None):  # noqa: A002
    """tf.nn.conv2d: NHWC cross-correlation with an HWIO filter; SAME with stride 1 and an odd kernel pads
    (k-1)/2 zeros on every side, VALID pads nothing."""
    assert data_format == "NHWC" and list(strides) == [1, 1, 1, 1] and padding in ("SAME", "VALID")

    def fn(x, w):
        kh, kw = w.shape[0], w.shape[1]
        assert kh % 2 == 1 and kw % 2 == 1
        p = (kh // 2, kw // 2) if padding == "SAME" else (0, 0)
        return torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=p).permute(0, 2, 3, 1)
    return _op(fn, input, filter)


def atrous_conv2d(value, filters, rate, padding, name=None):
    raise NotImplementedError("atrous_conv2d (skip != 1) is not reachable from noise_flow_arch")


def flatten(x, name=None):
    return _op(lambda t: t.reshape(t.shape[0], -1), x)


def dense(*a, **k):
    raise NotImplementedError("tf.layers.dense (real_nvp_default_template) is not reachable from noise_flow_arch")


def gradients(ys, xs, grad_ys=None, name=None, **_):
    """tf.gradients: one node per x; the backward pass runs once per run (memoised on the first node evaluated)."""
    xs = list(xs)
    state = {"run": None, "grads": None}
    g = _G

    def all_grads(y):
        if state["run"] != g.run_id:
            state["grads"] = torch.autograd.grad(y, [v.t for v in xs], allow_unused=True, retain_graph=True)
            state["run"] = g.run_id
        return state["grads"]
    return [_op(lambda y, i=i: all_grads(y)[i], ys, eager=False) for i in range(len(xs))]


# ------------------------------------------------------------------------- TFP 0.5 / tf.contrib.distributions
def fill_triangular(x, upper=False, name=None):
    """tfp.distributions.fill_triangular (TFP 0.5 `distribution_util.py`): for a vector of m = n(n+1)/2 elements,
    upper: concat([x, reverse(x[n:])]) / lower: concat([x[n:], reverse(x)]), reshaped [n, n], band part kept."""
    def fn(t):
        t = _const(t)
        m = t.shape[-1]
        n = int(math.sqrt(0.25 + 2 * m) - 0.5)
        assert n * (n + 1) // 2 == m
        if upper:
            parts = [t, torch.flip(t[..., n:], dims=[-1])]
        else:
            parts = [t[..., n:], torch.flip(t, dims=[-1])]
        full = torch.cat(parts, dim=-1).reshape(tuple(t.shape[:-1]) + (n, n))
        return _band(full, 0 if upper else -1, -1 if upper else 0)
    return _op(fn, x)


def fill_triangular_inverse(x, upper=False, name=None):
    """tfp.distributions.fill_triangular_inverse (TFP 0.5): the inverse packing of `fill_triangular`."""
    def fn(t):
        t = _const(t)
        n = t.shape[-1]
        m = n * (n + 1) // 2
        if upper:
            initial, tri = t[..., 0, :], t[..., 1:, :]
        else:
            initial, tri = torch.flip(t[..., -1, :], dims=[-1]), t[..., :-1, :]
        consolidated = tri + torch.flip(tri, dims=[-1, -2])
        end = consolidated.reshape(tuple(t.shape[:-2]) + (n * (n - 1),))
        return torch.cat([initial, end[..., :m - n]], dim=-1)
    return _op(fn, x)


class Bijector:
    """tf.contrib.distributions.bijectors.Bijector: only what the reference's subclasses use -- constructor
    bookkeeping and `.name`.  (The reference calls the private `_forward` / `_inverse...` methods directly.)"""

    def __init__(self, graph_parents=None, is_constant_jacobian=False, validate_args=False, dtype=None,
                 forward_min_event_ndims=None, inverse_min_event_ndims=None, name=None):
        if not name:
            raise ValueError("the reference always names its bijectors")
        self._name = name
        self._is_constant_jacobian = is_constant_jacobian
        self._validate_args = validate_args

    @property
    def name(self):
        return self._name


class Permute(Bijector):
    """tfp.bijectors.Permute (TFP 0.5 `permute.py`): forward = gather(x, permutation, axis=-1), inverse = gather with
    the inverted permutation, both log-dets are the constant 0.  It has NO fused `_inverse_and_log_det_jacobian`, so the
    reference's try/except falls back to `_inverse` + `_inverse_log_det_jacobian` (noise_flow_model.py:419-425)."""

    def __init__(self, permutation, validate_args=False, name=None):
        super().__init__(forward_min_event_ndims=1, is_constant_jacobian=True, validate_args=validate_args,
                         name=name or "permute")
        self._perm = [int(p) for p in permutation]

    @property
    def permutation(self):
        return self._perm

    def _forward(self, x):
        return gather(x, self._perm, axis=-1)

    def _inverse(self, y):
        inv = [0] * len(self._perm)
        for i, p in enumerate(self._perm):
            inv[p] = i
        return gather(y, inv, axis=-1)

    def _inverse_log_det_jacobian(self, y):
        return constant(0.0)

    def _forward_log_det_jacobian(self, x):
        return constant(0.0)


# ------------------------------------------------------------------------------------- optimiser (TF 1.12 rule)
class AdamOptimizer:
    """tf.train.AdamOptimizer (TF 1.12 `adam.py` docstring): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m <- b1 m + (1-b1) g;
    v <- b2 v + (1-b2) g^2; var <- var - lr_t * m / (sqrt(v) + eps).  `minimize` returns the train op; evaluating it in
    a run applies one step (the forward pass it differentiates is the run's own, evaluated once)."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **_):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
        self.t, self.m, self.v = 0, {}, {}
        self.last_grads = {}

    def minimize(self, loss, var_list=None):
        vs = list(var_list or trainable_variables())
        gs = gradients(loss, vs)
        g = _G

        def step(lr, *grads):
            if not g.running:
                return None
            self.t += 1
            lr_t = float(_const(lr)) * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
            self.last_grads = {}
            for v, gr in zip(vs, grads):
                if gr is None:
                    continue
                self.last_grads[v.var_name] = gr.detach().clone().numpy()
                m = self.m.get(v.var_name, torch.zeros_like(gr))
                s = self.v.get(v.var_name, torch.zeros_like(gr))
                m = self.b1 * m + (1 - self.b1) * gr
                s = self.b2 * s + (1 - self.b2) * gr * gr
                self.m[v.var_name], self.v[v.var_name] = m, s
                v.load(v._val.detach() - lr_t * m / (s.sqrt() + self.eps))
            return None
        return _op(step, self.lr, *gs, eager=False)


# ------------------------------------------------------------------------------------------ module assembly
def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Registers `tensorflow`, `tensorflow_probability` and the few private TF modules `matrix_param.py` imports,
    plus inert stand-ins for the plotting / IO packages the reference imports at module level but never touches on
    this path (matplotlib, h5py, imageio).  Idempotent."""
    if "tensorflow" in sys.modules and getattr(sys.modules["tensorflow"], "__nf_shim__", False):
        return sys.modules["tensorflow"]
    g = globals()
    names = ["reshape", "transpose", "concat", "split", "shape", "zeros", "ones", "zeros_like", "ones_like", "eye", "diag",
             "cast", "tile", "pad", "exp", "log", "sqrt", "tanh", "abs", "add", "subtract", "multiply", "divide", "equal",
             "greater_equal", "matmul", "Print", "reduce_sum", "reduce_mean", "where", "one_hot", "cond", "gather",
             "random_normal", "assign_sub", "matrix_set_diag", "matrix_band_part", "matrix_inverse",
             "get_variable", "variable_scope", "get_variable_scope", "name_scope", "control_dependencies", "constant",
             "constant_initializer", "zeros_initializer", "ones_initializer", "random_normal_initializer",
             "convert_to_tensor", "placeholder", "make_template", "AUTO_REUSE", "float32", "float64", "int32", "int64",
             "get_collection", "add_to_collection", "trainable_variables", "global_variables", "GraphKeys", "gradients",
             "identity", "stop_gradient", "sigmoid", "Graph", "get_default_graph", "reset_default_graph", "Tensor",
             "Variable", "TensorShape", "Dimension", "Session", "global_variables_initializer"]
    tf = _module("tensorflow", **{n: g[n] for n in names})
    tf.__nf_shim__ = True
    tf.__version__ = "1.12.0-shim"
    tf.bool = bool_
    tf.nn = _module("tensorflow.nn", conv2d=conv2d, atrous_conv2d=atrous_conv2d, moments=moments, relu=relu, sigmoid=sigmoid,
                    l2_normalize=l2_normalize, tanh=tanh)
    tf.linalg = _module("tensorflow.linalg", triangular_solve=triangular_solve, slogdet=slogdet, inv=matrix_inverse)
    tf.layers = _module("tensorflow.layers", flatten=flatten, dense=dense)
    tf.summary = _module("tensorflow.summary", histogram=lambda *a, **k: None, scalar=lambda *a, **k: None)
    tf.train = _module("tensorflow.train", AdamOptimizer=AdamOptimizer, Saver=Saver,
                       get_checkpoint_state=lambda *a, **k: None)
    bij = _module("tensorflow.contrib.distributions.bijectors", Bijector=Bijector, Permute=Permute)
    dist = _module("tensorflow.contrib.distributions", bijectors=bij, fill_triangular=fill_triangular,
                   fill_triangular_inverse=fill_triangular_inverse)
    tf.contrib = _module("tensorflow.contrib", distributions=dist, layers=_module("tensorflow.contrib.layers"))
    ops = _module("tensorflow.python.framework.ops", name_scope=name_scope, convert_to_tensor=convert_to_tensor)
    array_ops = _module("tensorflow.python.ops.array_ops", reshape=reshape, concat=concat, shape=shape, pad=pad)
    gen_array_ops = _module("tensorflow.python.ops.gen_array_ops", matrix_band_part=matrix_band_part)
    fw = _module("tensorflow.python.framework", ops=ops)
    pops = _module("tensorflow.python.ops", array_ops=array_ops, gen_array_ops=gen_array_ops)
    tf.python = _module("tensorflow.python", framework=fw, ops=pops)
    tfp = _module("tensorflow_probability", distributions=dist, bijectors=bij)
    tfp.__version__ = "0.5.0-shim"

    class _Inert(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Inert(self.__name__ + "." + k)

        def __call__(self, *a, **k):
            return None

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "h5py", "imageio"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Inert(name)
    return tf
