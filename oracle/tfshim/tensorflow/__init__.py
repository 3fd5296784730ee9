"""Eager stand-in for the TensorFlow 1.13-1.15 symbols used by the reference's hot path -- TEST INFRASTRUCTURE ONLY.

Purpose: let the reference's OWN source files (``/root/reference/CMPC_model.py``, ``util/cell.py``, ``util/loss.py``,
``util/processing_tools.py``) execute UNMODIFIED in this image, where TensorFlow 1.x cannot be installed, so that the
wiring of the head (which tensor feeds which op, variable names / scopes / shapes, masks, reshapes, loss weights, the
optimizer recipe) is the reference's and not a restatement.  ``oracle/ref_runner.py`` drives it; ``tests/`` and
``tests/golden/make_ref_golden.py`` are the only users.  The product path never imports this package.

What stays a restatement: the arithmetic of each ``tf.*`` op below (TensorFlow itself is third-party, not vendored by the
reference; README pins "TensorFlow 1.5", the code needs 1.13-1.15).  Each op follows TF-1's published definition:

* ``tf.nn.conv2d`` / ``tf.nn.convolution``: NHWC x HWIO cross-correlation, stride 1, SAME zero padding;
* ``tf.nn.l2_normalize(x, axis=None, epsilon=1e-12)``: ``x * rsqrt(max(sum(x*x, axis, keepdims), eps))``; ``axis=None``
  reduces over EVERY axis (nn_impl.l2_normalize);
* ``tf.contrib.layers.layer_norm``: ``nn.moments`` over axes ``[1, rank)`` (begin_norm_axis=1), biased variance,
  ``variance_epsilon=1e-12``, ``gamma`` / ``beta`` over the last axis (begin_params_axis=-1), applied as
  ``x * (rsqrt(var + eps) * gamma) + (beta - mean * rsqrt(var + eps) * gamma)`` (nn.batch_normalization);
  variables live in ``variable_scope(scope, default_name='LayerNorm')``;
* ``tf.image.resize_bilinear`` (legacy kernel, align_corners=False, no half-pixel centres): ``src = dst * (in / out)``
  in float32, ``lo = floor(src)``, ``hi = min(lo + 1, in - 1)``, ``top + (bottom - top) * y_lerp`` with
  ``top = tl + (tr - tl) * x_lerp``;
* ``tf.nn.softmax``: max-subtracted exp / sum along ``axis``;
* ``tf.nn.sigmoid_cross_entropy_with_logits``: ``max(x, 0) - x * z + log1p(exp(-|x|))``;
* ``tf.nn.rnn_cell.LSTMCell(state_is_tuple=False)``: gates ``i, j, f, o = split([x, m] kernel + bias, 4)``,
  ``c = sigmoid(f + 1) c + sigmoid(i) tanh(j)``, ``m = sigmoid(o) tanh(c)``, state = concat(c, m), variables
  ``kernel`` (glorot uniform) / ``bias`` (zeros) under ``rnn/lstm_cell``;
* ``tf.nn.dynamic_rnn``: batch-major, scope ``rnn``, the cell is entered under the snake-cased class name, zero outputs
  and carried state for ``t >= sequence_length``; variables are created on the first step and shared by the others
  (the graph-mode while_loop traces the cell once, so default-named scopes -- ``LayerNorm``, ``LayerNorm_1`` ... --
  restart at every step);
* ``tf.train.AdamOptimizer`` (beta1 .9, beta2 .999, eps 1e-8, ``lr_t = lr sqrt(1 - b2^t) / (1 - b1^t)``,
  ``var -= lr_t m / (sqrt(v) + eps)``), ``compute_gradients`` = reverse-mode autodiff of the executed ops
  (``torch.autograd``), ``tf.train.polynomial_decay`` (cycle=False);
* ``tf.get_variable`` default initializer = glorot uniform; ``xavier_initializer_conv2d`` = glorot uniform with
  fan = receptive field x channels.

Graph mode is emulated eagerly: ``tf.placeholder`` returns the tensor that ``ref_runner`` fed for it (by creation order),
so constructing ``LSTM_model`` IS the ``sess.run``.  Tensors are a thin wrapper (``Tensor``) over ``torch`` CPU tensors
so that ``x.shape[-1].value``, ``var.op.name``, ``var.name`` behave as in TF.  ``tf.float32`` maps to the torch dtype
selected by ``_shim.float_dtype`` (float32 = what TF computes in; float64 = wiring check free of rounding noise).
"""
from __future__ import annotations

import math
import re
import types
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as _F

__version__ = "1.15.0-shim"


# ------------------------------------------------------------------------------------------------------------------
# shim state (set by oracle/ref_runner.py)
# ------------------------------------------------------------------------------------------------------------------
class _ShimState:
    def __init__(self):
        self.reset()

    def reset(self):
        self.float_dtype = torch.float32
        self.placeholder_feeds: List = []        # consumed in creation order by tf.placeholder
        self.placeholder_count = 0
        self.layers: Dict[str, "Tensor"] = {}    # backbone taps handed out by the deeplab_resnet stub
        self.params: Optional[Dict[str, torch.Tensor]] = None   # initial values by full variable name
        self.rnn_outputs_feed = None             # when set: outputs of the word LSTM (dynamic_rnn over an LSTMCell)
        self.variables: Dict[str, "Variable"] = {}
        self.var_order: List["Variable"] = []
        self.scope: List[str] = []
        self.default_counts: Dict[str, int] = {}
        self.optimizers: List["_AdamOptimizer"] = []
        self.init_seed = 0
        self.requires_grad = False


_shim = _ShimState()


# ------------------------------------------------------------------------------------------------------------------
# dtypes
# ------------------------------------------------------------------------------------------------------------------
class DType:
    def __init__(self, name, torch_dtype, is_float=False):
        self.name, self._torch, self.is_floating = name, torch_dtype, is_float

    @property
    def torch(self):
        if self.name == "float32":
            return _shim.float_dtype
        return self._torch

    @property
    def min(self):
        if self.name == "float32":
            return float(np.finfo(np.float32).min)
        if self.is_floating:
            return float(torch.finfo(self._torch).min)
        return int(torch.iinfo(self._torch).min)

    @property
    def max(self):
        if self.name == "float32":
            return float(np.finfo(np.float32).max)
        if self.is_floating:
            return float(torch.finfo(self._torch).max)
        return int(torch.iinfo(self._torch).max)

    def __repr__(self):
        return "tf." + self.name


float32 = DType("float32", torch.float32, True)
float64 = DType("float64", torch.float64, True)
int32 = DType("int32", torch.int32)
int64 = DType("int64", torch.int64)
bool = DType("bool", torch.bool)  # noqa: A001  (tf.bool)


# ------------------------------------------------------------------------------------------------------------------
# shapes and tensors
# ------------------------------------------------------------------------------------------------------------------
class Dimension:
    def __init__(self, v):
        self.value = None if v is None else int(v)

    def __int__(self):
        return self.value

    __index__ = __int__

    def __eq__(self, o):
        return self.value == (o.value if isinstance(o, Dimension) else o)

    def __hash__(self):
        return hash(self.value)

    def __mul__(self, o):
        return self.value * int(o)

    __rmul__ = __mul__

    def __repr__(self):
        return "Dimension(%s)" % self.value


class TensorShape:
    def __init__(self, dims):
        if isinstance(dims, TensorShape):
            dims = dims.as_list()
        self._dims = [d.value if isinstance(d, Dimension) else (None if d is None else int(d)) for d in dims]

    @property
    def ndims(self):
        return len(self._dims)

    rank = ndims

    def as_list(self):
        return list(self._dims)

    def __len__(self):
        return len(self._dims)

    def __iter__(self):
        return iter([Dimension(d) for d in self._dims])

    def __getitem__(self, k):
        if isinstance(k, slice):
            return TensorShape(self._dims[k])
        return Dimension(self._dims[k])

    def __eq__(self, o):
        return self.as_list() == TensorShape(o).as_list()

    def __repr__(self):
        return "TensorShape(%s)" % self._dims


def _shape_list(shape) -> List[int]:
    if isinstance(shape, TensorShape):
        return shape.as_list()
    if isinstance(shape, (int, np.integer, Dimension)):
        return [int(shape)]
    return [int(s) for s in shape]


def _t(x, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """unwrap to torch"""
    if isinstance(x, Tensor):
        v = x._v
    elif isinstance(x, torch.Tensor):
        v = x
    elif isinstance(x, np.ndarray):
        v = torch.from_numpy(x)
        if v.is_floating_point():
            v = v.to(_shim.float_dtype)
    else:
        v = torch.as_tensor(x)
        if v.is_floating_point():
            v = v.to(_shim.float_dtype)
    if dtype is not None and v.dtype != dtype:
        v = v.to(dtype)
    return v


def _pair(a, b):
    """binary-op operands: python scalars adopt the tensor's dtype like TF's convert_to_tensor(preferred dtype)"""
    ta, tb = isinstance(a, (Tensor, torch.Tensor)), isinstance(b, (Tensor, torch.Tensor))
    if ta and not tb and not isinstance(b, np.ndarray):
        va = _t(a)
        return va, torch.as_tensor(b, dtype=va.dtype)
    if tb and not ta and not isinstance(a, np.ndarray):
        vb = _t(b)
        return torch.as_tensor(a, dtype=vb.dtype), vb
    return _t(a), _t(b)


class _Op:
    def __init__(self, name):
        self.name = name


class Tensor:
    """tf.Tensor look-alike over a torch CPU tensor"""
    __array_priority__ = 100

    def __init__(self, v: torch.Tensor, name: Optional[str] = None):
        self._v = v
        self._name = name

    # TF surface ---------------------------------------------------------------------------------------------
    @property
    def shape(self):
        return TensorShape(list(self._v.shape))

    def get_shape(self):
        return self.shape

    @property
    def dtype(self):
        for d in (float32, float64, int32, int64, bool):
            if d.torch == self._v.dtype:
                return d
        return self._v.dtype

    @property
    def name(self):
        return (self._name or "Tensor") + ":0"

    @property
    def op(self):
        return _Op(self._name or "Tensor")

    def numpy(self):
        return self._v.detach().cpu().numpy()

    def __repr__(self):
        return "<shim tf.Tensor %s shape=%s dtype=%s>" % (self._name or "", tuple(self._v.shape), self._v.dtype)

    # operators ------------------------------------------------------------------------------------------------
    def __add__(self, o):
        a, b = _pair(self, o); return Tensor(a + b)

    def __radd__(self, o):
        a, b = _pair(o, self); return Tensor(a + b)

    def __sub__(self, o):
        a, b = _pair(self, o); return Tensor(a - b)

    def __rsub__(self, o):
        a, b = _pair(o, self); return Tensor(a - b)

    def __mul__(self, o):
        a, b = _pair(self, o); return Tensor(a * b)

    def __rmul__(self, o):
        a, b = _pair(o, self); return Tensor(a * b)

    def __truediv__(self, o):
        return divide(self, o)

    def __rtruediv__(self, o):
        return divide(o, self)

    __div__, __rdiv__ = __truediv__, __rtruediv__

    def __neg__(self):
        return Tensor(-self._v)

    def __gt__(self, o):
        a, b = _pair(self, o); return Tensor(a > b)

    def __ge__(self, o):
        a, b = _pair(self, o); return Tensor(a >= b)

    def __lt__(self, o):
        a, b = _pair(self, o); return Tensor(a < b)

    def __le__(self, o):
        a, b = _pair(self, o); return Tensor(a <= b)

    def __getitem__(self, k):
        return Tensor(self._v[k])

    def __hash__(self):
        return id(self)

    def __eq__(self, o):          # TF-1 tensors compare by identity (they are dict keys in train_op)
        return self is o


class Variable(Tensor):
    """tf.Variable(initial_value, trainable=True) / the object tf.get_variable returns"""

    def __init__(self, initial_value=None, trainable=True, name=None, _full_name=None):
        if _full_name is None:
            _full_name = _unique_name(name or "Variable")
        v = _t(initial_value).clone()
        key = _strip_root(_full_name)
        if _shim.params is not None and key in _shim.params:
            p = _shim.params[key]
            assert tuple(p.shape) == tuple(v.shape), (_full_name, tuple(p.shape), tuple(v.shape))
            v = p.detach().clone().to(v.dtype)
        if trainable and v.is_floating_point() and _shim.requires_grad:
            v.requires_grad_(True)
        super().__init__(v, _full_name)
        self.trainable = trainable
        _shim.variables[_full_name] = self
        _shim.var_order.append(self)


def _strip_root(full_name: str) -> str:
    return full_name[len("text_objseg/"):] if full_name.startswith("text_objseg/") else full_name


# ------------------------------------------------------------------------------------------------------------------
# variable scopes
# ------------------------------------------------------------------------------------------------------------------
def _scope_path() -> str:
    return "/".join(_shim.scope)


def _unique_name(default_name: str) -> str:
    """TF uniquifies default names per enclosing scope: LayerNorm, LayerNorm_1, ..."""
    base = (_scope_path() + "/" if _shim.scope else "") + default_name
    n = _shim.default_counts.get(base, 0)
    _shim.default_counts[base] = n + 1
    return base if n == 0 else "%s_%d" % (base, n)


class variable_scope:  # noqa: N801
    def __init__(self, name_or_scope=None, default_name=None, values=None, reuse=None, **_):
        self._name, self._default = name_or_scope, default_name

    def __enter__(self):
        if self._name is None:
            full = _unique_name(self._default)
            name = full.rsplit("/", 1)[-1]
        else:
            name = self._name
        _shim.scope.append(name)
        return self

    def __exit__(self, *exc):
        _shim.scope.pop()
        return False


def _glorot_uniform(shape, gen):
    shape = list(shape)
    if len(shape) < 1:
        fan_in = fan_out = 1
    elif len(shape) == 1:
        fan_in = fan_out = shape[0]
    elif len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * lim


class _Init:
    def __init__(self, fn):
        self.fn = fn

    def __call__(self, shape, gen):
        return self.fn(list(shape), gen)


def constant_initializer(value=0.0, **_):
    return _Init(lambda s, g: torch.full(s, float(value), dtype=torch.float64))


def zeros_initializer(**_):
    return _Init(lambda s, g: torch.zeros(s, dtype=torch.float64))


def ones_initializer(**_):
    return _Init(lambda s, g: torch.ones(s, dtype=torch.float64))


def random_normal_initializer(mean=0.0, stddev=1.0, **_):
    return _Init(lambda s, g: mean + stddev * torch.randn(s, generator=g, dtype=torch.float64))


def glorot_uniform_initializer(**_):
    return _Init(_glorot_uniform)


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **_):
    full = (_scope_path() + "/" if _shim.scope else "") + name
    if full in _shim.variables:           # re-entered scope (later RNN steps): the same variable
        return _shim.variables[full]
    shape = _shape_list(shape)
    init = initializer if initializer is not None else glorot_uniform_initializer()
    gen = torch.Generator().manual_seed(_shim.init_seed + len(_shim.var_order))
    val = init(shape, gen).to((dtype or float32).torch)
    return Variable(val, trainable=trainable, _full_name=full)


def trainable_variables():
    return [v for v in _shim.var_order if v.trainable]


def global_variables():
    return list(_shim.var_order)


# ------------------------------------------------------------------------------------------------------------------
# placeholders / conversion
# ------------------------------------------------------------------------------------------------------------------
def placeholder(dtype, shape=None, name=None):
    i = _shim.placeholder_count
    _shim.placeholder_count += 1
    if i >= len(_shim.placeholder_feeds) or _shim.placeholder_feeds[i] is None:
        val = torch.zeros(_shape_list(shape), dtype=dtype.torch)        # an unfed placeholder nobody reads
    else:
        val = _t(_shim.placeholder_feeds[i], dtype.torch)
        assert list(val.shape) == _shape_list(shape), ("feed shape", i, tuple(val.shape), shape)
    return Tensor(val, name or "Placeholder_%d" % i)


def convert_to_tensor(value, dtype=None, **_):
    if isinstance(value, Tensor) and dtype is None:
        return value
    return Tensor(_t(value, None if dtype is None else dtype.torch))


def cast(x, dtype, **_):
    return Tensor(_t(x).to(dtype.torch))


def to_float(x, **_):
    return cast(x, float32)


def stop_gradient(x, **_):
    return Tensor(_t(x).detach())


def zeros(shape, dtype=float32, **_):
    return Tensor(torch.zeros(_shape_list(shape), dtype=dtype.torch))


def shape(x, **_):
    return list(_t(x).shape)


# ------------------------------------------------------------------------------------------------------------------
# math
# ------------------------------------------------------------------------------------------------------------------
def _axes(axis, nd):
    if axis is None:
        return tuple(range(nd))
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) % nd for a in axis)
    return (int(axis) % nd,)


def reduce_sum(x, axis=None, keepdims=False, keep_dims=None, **_):
    v = _t(x)
    kd = keepdims if keep_dims is None else keep_dims
    if v.dtype == torch.bool:
        v = v.to(torch.int64)
    return Tensor(v.sum(dim=_axes(axis, v.dim()), keepdim=kd))


def reduce_mean(x, axis=None, keepdims=False, keep_dims=None, **_):
    v = _t(x)
    kd = keepdims if keep_dims is None else keep_dims
    return Tensor(v.mean(dim=_axes(axis, v.dim()), keepdim=kd))


def add(a, b, **_):
    a, b = _pair(a, b); return Tensor(a + b)


def add_n(xs, **_):
    out = _t(xs[0])
    for x in xs[1:]:
        out = out + _t(x)
    return Tensor(out)


def subtract(a, b, **_):
    a, b = _pair(a, b); return Tensor(a - b)


sub = subtract


def multiply(a, b, **_):
    if a is None or b is None:
        return None
    a, b = _pair(a, b); return Tensor(a * b)


def scalar_mul(s, x):
    return multiply(s, x)


def divide(a, b, **_):
    """tf.divide == python-3 true division: integer operands give float64"""
    a, b = _pair(a, b)
    if not a.is_floating_point() and not b.is_floating_point():
        a, b = a.to(torch.float64), b.to(torch.float64)
    return Tensor(a / b)


div = divide


def pow(a, b, **_):  # noqa: A001
    a, b = _pair(a, b); return Tensor(torch.pow(a, b))


def abs(x, **_):  # noqa: A001
    return Tensor(_t(x).abs())


def sigmoid(x, **_):
    return Tensor(torch.sigmoid(_t(x)))


def tanh(x, **_):
    return Tensor(torch.tanh(_t(x)))


def equal(a, b, **_):
    a, b = _pair(a, b); return Tensor(a == b)


def less(a, b, **_):
    a, b = _pair(a, b); return Tensor(a < b)


def logical_not(x, **_):
    return Tensor(~_t(x))


def logical_and(a, b, **_):
    return Tensor(_t(a) & _t(b))


def logical_or(a, b, **_):
    return Tensor(_t(a) | _t(b))


def matmul(a, b, transpose_a=False, transpose_b=False, **_):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return Tensor(torch.matmul(a, b))


# ------------------------------------------------------------------------------------------------------------------
# array ops
# ------------------------------------------------------------------------------------------------------------------
def reshape(x, shape, **_):
    return Tensor(_t(x).reshape(_shape_list(shape)))


def transpose(x, perm=None, **_):
    v = _t(x)
    if perm is None:
        perm = list(range(v.dim()))[::-1]
    return Tensor(v.permute(*perm))


def expand_dims(x, axis, **_):
    return Tensor(_t(x).unsqueeze(axis))


def concat(values, axis, **_):
    return Tensor(torch.cat([_t(v) for v in values], dim=int(axis)))


def stack(values, axis=0, **_):
    return Tensor(torch.stack([_t(v) for v in values], dim=int(axis)))


def tile(x, multiples, **_):
    return Tensor(_t(x).repeat(*[int(m) for m in multiples]))


def split(value, num_or_size_splits, axis=0, **_):
    v = _t(value)
    axis = int(axis)
    if isinstance(num_or_size_splits, int):
        assert v.shape[axis] % num_or_size_splits == 0
        parts = torch.split(v, v.shape[axis] // num_or_size_splits, dim=axis)
    else:
        parts = torch.split(v, list(num_or_size_splits), dim=axis)
    return [Tensor(p) for p in parts]


# ------------------------------------------------------------------------------------------------------------------
# tf.nn
# ------------------------------------------------------------------------------------------------------------------
def _conv_same(x, w, strides=None):
    x, w = _t(x), _t(w)
    if strides is not None:
        assert list(strides) == [1, 1, 1, 1], strides
    kh, kw, cin, cout = w.shape
    assert x.shape[-1] == cin, (tuple(x.shape), tuple(w.shape))
    assert kh % 2 == 1 and kw % 2 == 1
    if kh == 1 and kw == 1:
        return Tensor(torch.matmul(x, w.reshape(cin, cout)))
    y = _F.conv2d(x.permute(0, 3, 1, 2).contiguous(), w.permute(3, 2, 0, 1).contiguous(), padding=(kh // 2, kw // 2))
    return Tensor(y.permute(0, 2, 3, 1))


def _conv2d(input, filter=None, strides=None, padding="SAME", filters=None, **_):  # noqa: A002
    assert padding == "SAME"
    return _conv_same(input, filter if filter is not None else filters, strides)


def _convolution(input, filter, padding, strides=None, dilation_rate=None, data_format=None, **_):  # noqa: A002
    assert padding == "SAME" and data_format in (None, "NHWC")
    return _conv_same(input, filter)


def _l2_normalize(x, axis=None, epsilon=1e-12, name=None, dim=None):
    v = _t(x)
    if dim is not None:
        axis = dim
    sq = (v * v).sum(dim=_axes(axis, v.dim()), keepdim=True)
    return Tensor(v * torch.rsqrt(torch.clamp(sq, min=epsilon)))


def _softmax(logits, axis=None, name=None, dim=None):
    if dim is not None:
        axis = dim
    return Tensor(torch.softmax(_t(logits), dim=-1 if axis is None else int(axis)))


def _sigmoid_ce(_sentinel=None, labels=None, logits=None, name=None):
    x, z = _t(logits), _t(labels)
    return Tensor(torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-x.abs())))


def _embedding_lookup(params, ids, **_):
    return Tensor(_t(params)[_t(ids).long()])


def _snake(name: str) -> str:
    s = re.sub(r"(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    return re.sub(r"([a-z])([A-Z])", r"\1_\2", s).lower()


class LSTMStateTuple(tuple):
    def __new__(cls, c, h):
        return super().__new__(cls, (c, h))

    c = property(lambda self: self[0])
    h = property(lambda self: self[1])


class RNNCell:
    """tf.nn.rnn_cell.RNNCell (a Layer): calling it enters variable_scope(<snake-cased class name>)"""

    def __init__(self, trainable=True, name=None, dtype=None, _reuse=None, **_):
        self._scope_name = name or _snake(type(self).__name__)

    def __call__(self, inputs, state):
        # the graph-mode cell body is traced once: default-named sub-scopes restart at every call
        prefix = (_scope_path() + "/" if _shim.scope else "") + self._scope_name + "/"
        for k in [k for k in _shim.default_counts if k.startswith(prefix)]:
            del _shim.default_counts[k]
        with variable_scope(self._scope_name):
            return self.call(inputs, state)

    def zero_state(self, batch_size, dtype):
        def z(size):
            if isinstance(size, (TensorShape, list, tuple)) and not isinstance(size, LSTMStateTuple):
                return Tensor(torch.zeros([batch_size] + _shape_list(size), dtype=dtype.torch))
            return Tensor(torch.zeros([batch_size, int(size)], dtype=dtype.torch))
        ss = self.state_size
        if isinstance(ss, LSTMStateTuple):
            return LSTMStateTuple(z(ss[0]), z(ss[1]))
        return z(ss)


class LSTMCell(RNNCell):
    """tf.nn.rnn_cell.LSTMCell without peepholes / projection (rnn_cell_impl.LSTMCell.call)"""

    def __init__(self, num_units, forget_bias=1.0, state_is_tuple=True, **kw):
        super().__init__(**kw)
        self._scope_name = kw.get("name") or "lstm_cell"
        self._num_units, self._forget_bias, self._state_is_tuple = num_units, forget_bias, state_is_tuple

    @property
    def state_size(self):
        return LSTMStateTuple(self._num_units, self._num_units) if self._state_is_tuple else 2 * self._num_units

    @property
    def output_size(self):
        return self._num_units

    def call(self, inputs, state):
        n = self._num_units
        x = _t(inputs)
        if self._state_is_tuple:
            c_prev, m_prev = _t(state[0]), _t(state[1])
        else:
            s = _t(state)
            c_prev, m_prev = s[:, :n], s[:, n:]
        kernel = get_variable("kernel", [x.shape[-1] + n, 4 * n])
        bias = get_variable("bias", [4 * n], initializer=zeros_initializer())
        z = torch.matmul(torch.cat([x, m_prev], 1), kernel._v) + bias._v
        i, j, f, o = torch.split(z, n, dim=1)
        c = torch.sigmoid(f + self._forget_bias) * c_prev + torch.sigmoid(i) * torch.tanh(j)
        m = torch.sigmoid(o) * torch.tanh(c)
        new_state = LSTMStateTuple(Tensor(c), Tensor(m)) if self._state_is_tuple else Tensor(torch.cat([c, m], 1))
        return Tensor(m), new_state


def _dynamic_rnn(cell, inputs, sequence_length=None, initial_state=None, dtype=None, time_major=False, scope=None, **_):
    x = _t(inputs)
    assert not time_major
    B, T = x.shape[0], x.shape[1]
    if isinstance(cell, LSTMCell) and _shim.rnn_outputs_feed is not None:
        fed = _t(_shim.rnn_outputs_feed, (dtype or float32).torch)
        assert fed.shape[0] == B and fed.shape[1] == T
        return Tensor(fed), None
    with variable_scope(scope or "rnn"):
        state = initial_state if initial_state is not None else cell.zero_state(B, dtype or float32)
        seq = None if sequence_length is None else _t(sequence_length).long()
        outs = []
        for t in range(T):
            out, new_state = cell(Tensor(x[:, t]), state)
            if seq is not None:
                live = (t < seq)

                def sel(new, old, live=live):
                    nv, ov = _t(new), _t(old)
                    m = live.view([-1] + [1] * (nv.dim() - 1))
                    return Tensor(torch.where(m, nv, ov))
                out = sel(out, Tensor(torch.zeros_like(_t(out))))
                if isinstance(new_state, LSTMStateTuple):
                    new_state = LSTMStateTuple(sel(new_state[0], state[0]), sel(new_state[1], state[1]))
                else:
                    new_state = sel(new_state, state)
            state = new_state
            outs.append(_t(out))
        return Tensor(torch.stack(outs, 1)), state


rnn_cell = types.SimpleNamespace(RNNCell=RNNCell, LSTMCell=LSTMCell, LSTMStateTuple=LSTMStateTuple,
                                 BasicLSTMCell=LSTMCell)

nn = types.SimpleNamespace(
    conv2d=_conv2d, convolution=_convolution, l2_normalize=_l2_normalize, softmax=_softmax,
    relu=lambda x, **_: Tensor(torch.relu(_t(x))), tanh=tanh, sigmoid=sigmoid,
    sigmoid_cross_entropy_with_logits=_sigmoid_ce, embedding_lookup=_embedding_lookup,
    l2_loss=lambda x, **_: Tensor((_t(x) ** 2).sum() / 2), dynamic_rnn=_dynamic_rnn, rnn_cell=rnn_cell,
)


# ------------------------------------------------------------------------------------------------------------------
# tf.image
# ------------------------------------------------------------------------------------------------------------------
def _resize_bilinear(images, size, align_corners=False, **_):
    assert not align_corners
    x = _t(images)
    B, h, w, C = x.shape
    oh, ow = int(size[0]), int(size[1])

    def interp(n_in, n_out):
        scale = np.float32(n_in) / np.float32(n_out)
        src = np.arange(n_out, dtype=np.float32) * scale            # float32 like the TF kernel
        lo = np.floor(src).astype(np.int64)
        hi = np.minimum(lo + 1, n_in - 1)
        lerp = (src - lo.astype(np.float32)).astype(np.float32)
        return torch.from_numpy(lo), torch.from_numpy(hi), torch.from_numpy(lerp).to(x.dtype)

    ylo, yhi, yl = interp(h, oh)
    xlo, xhi, xl = interp(w, ow)
    xl = xl.view(1, 1, ow, 1)
    yl = yl.view(1, oh, 1, 1)
    top_rows, bot_rows = x[:, ylo], x[:, yhi]
    top = top_rows[:, :, xlo] + (top_rows[:, :, xhi] - top_rows[:, :, xlo]) * xl
    bot = bot_rows[:, :, xlo] + (bot_rows[:, :, xhi] - bot_rows[:, :, xlo]) * xl
    return Tensor(top + (bot - top) * yl)


image = types.SimpleNamespace(resize_bilinear=_resize_bilinear)


# ------------------------------------------------------------------------------------------------------------------
# tf.contrib.layers
# ------------------------------------------------------------------------------------------------------------------
def _layer_norm(inputs, center=True, scale=True, activation_fn=None, reuse=None, variables_collections=None,
                outputs_collections=None, trainable=True, begin_norm_axis=1, begin_params_axis=-1, scope=None):
    x = _t(inputs)
    with variable_scope(scope, default_name="LayerNorm"):
        pshape = list(x.shape[begin_params_axis:])
        beta = get_variable("beta", pshape, initializer=zeros_initializer()) if center else None
        gamma = get_variable("gamma", pshape, initializer=ones_initializer()) if scale else None
    axes = tuple(range(begin_norm_axis, x.dim()))
    mean = x.mean(dim=axes, keepdim=True)
    var = ((x - mean.detach()) ** 2).mean(dim=axes, keepdim=True)     # nn.moments: squared_difference(x, stop_gradient(mean))
    inv = torch.rsqrt(var + 1e-12)
    if gamma is not None:
        inv = inv * gamma._v
    out = x * inv + ((beta._v if beta is not None else 0.0) - mean * inv)
    if activation_fn is not None:
        return activation_fn(Tensor(out))
    return Tensor(out)


contrib = types.SimpleNamespace(layers=types.SimpleNamespace(
    layer_norm=_layer_norm,
    xavier_initializer_conv2d=lambda **_: glorot_uniform_initializer(),
    xavier_initializer=lambda **_: glorot_uniform_initializer()))


# ------------------------------------------------------------------------------------------------------------------
# tf.train
# ------------------------------------------------------------------------------------------------------------------
def _polynomial_decay(learning_rate, global_step, decay_steps, end_learning_rate=0.0001, power=1.0, cycle=False, **_):
    assert not cycle
    step = min(float(_t(global_step)), float(decay_steps))
    return Tensor(torch.tensor((learning_rate - end_learning_rate) * (1 - step / decay_steps) ** power + end_learning_rate,
                               dtype=torch.float64))


class _AdamOptimizer:
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **_):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
        self.m: Dict[str, torch.Tensor] = {}
        self.v: Dict[str, torch.Tensor] = {}
        self.t = 0
        self.raw_grads = None            # [(grad or None, var)] as compute_gradients returned them
        self.applied = None              # [(grad or None, var)] as handed to apply_gradients
        _shim.optimizers.append(self)

    def compute_gradients(self, loss, var_list=None, **_):
        var_list = list(var_list if var_list is not None else trainable_variables())
        live = [v for v in var_list if v._v.requires_grad]
        gs = torch.autograd.grad(_t(loss), [v._v for v in live], allow_unused=True)
        by_var = {id(v): g for v, g in zip(live, gs)}
        out = [(None if by_var.get(id(v)) is None else Tensor(by_var[id(v)]), v) for v in var_list]
        self.raw_grads = out
        return out

    def apply_gradients(self, grads_and_vars, global_step=None, **_):
        grads_and_vars = list(grads_and_vars)
        self.applied = grads_and_vars
        self.t += 1
        lr = float(_t(self.lr))
        lr_t = lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        with torch.no_grad():
            for g, var in grads_and_vars:
                if g is None:
                    continue
                gv = _t(g)
                k = var._name
                m = self.m.get(k, torch.zeros_like(gv))
                v = self.v.get(k, torch.zeros_like(gv))
                m = self.b1 * m + (1 - self.b1) * gv
                v = self.b2 * v + (1 - self.b2) * gv * gv
                self.m[k], self.v[k] = m, v
                var._v -= lr_t * m / (torch.sqrt(v) + self.eps)
            if global_step is not None:
                global_step._v += 1
        return Tensor(torch.zeros(()), "train")


train = types.SimpleNamespace(polynomial_decay=_polynomial_decay, AdamOptimizer=_AdamOptimizer)


# ------------------------------------------------------------------------------------------------------------------
# tf.summary (no-ops), tf.compat
# ------------------------------------------------------------------------------------------------------------------
summary = types.SimpleNamespace(scalar=lambda *a, **k: None, histogram=lambda *a, **k: None,
                                merge_all=lambda *a, **k: None, image=lambda *a, **k: None)

compat = types.SimpleNamespace(v1=types.SimpleNamespace(nn=nn, train=train, variable_scope=variable_scope,
                                                        get_variable=get_variable, placeholder=placeholder))
