"""Minimal stand-in for the `chainer` package (v7 semantics) over NumPy.

TEST INFRASTRUCTURE ONLY.  The reference (nogu-atsu/RGBD-GAN) is pure Python on
top of Chainer >= 7.0.0 + CuPy >= 7.0.0 (reference README.md:19-27; unpinned, the
last release of that line is Chainer 7.8.1).  Neither package is installed nor
installable in this image, so the reference cannot be imported as is.  This
module restates exactly the slice of Chainer's published behaviour that the hot
path touches, so that `tests/golden/make_golden.py` can import the reference's
own, UNMODIFIED source files from /root/reference and run them to produce the
golden vectors committed under tests/golden/.

What is restated (Chainer v7, file names relative to the chainer package):
  * variable.py: `Variable` (array/data/grad, reshape/transpose/__getitem__,
    `__array_priority__ = 200` so `ndarray <op> Variable` defers to Variable).
  * functions/math/basic_math.py: Add/AddConstant/Sub/SubFromConstant/Mul/
    MulConstant/Div/Neg incl. `_preprocess_rhs` (an ndarray operand is CAST TO
    THE VARIABLE'S DTYPE before the op -- int32 and bool operands become fp32)
    and `_preprocess_const` (python scalars become `x.dtype.type(value)`);
    broadcasting arithmetic with `sum_to` in backward;
    DivGrad: gx0 = gy / x1, gx1 = -gx0 * x0 / x1.
  * functions/math/clip.py: forward numpy.clip, backward gy * (min <= x <= max).
  * functions/math/matmul.py: numpy.matmul forward; gb = matmul(a^T, gy).
  * functions/array/get_item.py: forward x[slices]; backward zeros + numpy.add.at.
  * functions/array/{reshape,transpose,concat,scatter_add,broadcast}.py.
  * functions/loss/mean_absolute_error.py: abs(diff).sum()/size; gy*sign(diff)/size.
  * functions/loss/mean_squared_error.py: diff.dot(diff)/size; gy*diff*(2/size).
  * function_node.py / _backprop_utils: reverse walk by descending rank, gradients
    of a multiply-used variable are summed in arrival order.

Anything else on the import path of the reference modules (chainer.Chain,
chainer.links, ...) is a permissive stub: those are network layers, out of scope.
"""
import heapq
import sys
import types

import numpy


# --------------------------------------------------------------------------- xp
class _XP(types.ModuleType):
    """numpy, except `meshgrid` returns a list as it did in the NumPy 1.x the
    reference was written for (loss_functions.py:59-61 does `meshgrid(..) + [..]`)."""

    def __init__(self):
        super().__init__("numpy_compat")

    def __getattr__(self, name):
        return getattr(numpy, name)

    @staticmethod
    def meshgrid(*a, **k):
        return list(numpy.meshgrid(*a, **k))


xp_compat = _XP()


def get_array_module(*args):
    return xp_compat


def _force_array(x, dtype=None):
    if numpy.isscalar(x) or not isinstance(x, numpy.ndarray):
        return numpy.array(x, dtype)
    return x if dtype is None else x.astype(dtype, copy=False)


def _sum_to(x, shape):
    if x.shape == tuple(shape):
        return x
    ndim = len(shape)
    lead = x.ndim - ndim
    lead_axis = tuple(range(lead))
    axis = tuple(i + lead for i, sx in enumerate(shape) if sx == 1)
    y = x.sum(lead_axis + axis, keepdims=True)
    if lead > 0:
        y = y.squeeze(lead_axis)
    return y


# --------------------------------------------------------------------- Variable
class Variable:
    __array_priority__ = 200

    def __init__(self, data=None, name=None, grad=None, requires_grad=True):
        self.array = data
        self.name = name
        self.grad = grad
        self.requires_grad = requires_grad
        self.creator_node = None
        self.rank = 0

    # chainer aliases
    @property
    def data(self):
        return self.array

    @data.setter
    def data(self, v):
        self.array = v

    @property
    def creator(self):
        return self.creator_node

    shape = property(lambda s: s.array.shape)
    dtype = property(lambda s: s.array.dtype)
    ndim = property(lambda s: s.array.ndim)
    size = property(lambda s: s.array.size)

    def __len__(self):
        return len(self.array)

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = shape[0]
        return functions.reshape(self, shape)

    def transpose(self, *axes):
        if len(axes) == 0:
            axes = None
        elif len(axes) == 1 and (isinstance(axes[0], (tuple, list)) or axes[0] is None):
            axes = axes[0]
        return functions.transpose(self, axes)

    def __getitem__(self, slices):
        return functions.get_item(self, slices)

    def cleargrad(self):
        self.grad = None

    def backward(self):
        _backward(self)

    # arithmetic -- functions/math/basic_math.py
    def __neg__(self):
        return Neg().apply((self,))[0]

    def __add__(self, rhs):
        if numpy.isscalar(rhs):
            return AddConstant(rhs).apply((self,))[0]
        return Add().apply((self, _preprocess_rhs(self, rhs)))[0]

    __radd__ = __add__
    __iadd__ = __add__

    def __sub__(self, rhs):
        if numpy.isscalar(rhs):
            return AddConstant(-rhs).apply((self,))[0]
        return Sub().apply((self, _preprocess_rhs(self, rhs)))[0]

    def __rsub__(self, rhs):
        if numpy.isscalar(rhs):
            return SubFromConstant(rhs).apply((self,))[0]
        return Sub().apply((_preprocess_rhs(self, rhs), self))[0]

    def __mul__(self, rhs):
        if numpy.isscalar(rhs):
            return MulConstant(rhs).apply((self,))[0]
        return Mul().apply((self, _preprocess_rhs(self, rhs)))[0]

    __rmul__ = __mul__
    __imul__ = __mul__

    def __truediv__(self, rhs):
        if numpy.isscalar(rhs):
            return MulConstant(1. / rhs).apply((self,))[0]
        return Div().apply((self, _preprocess_rhs(self, rhs)))[0]

    def __rtruediv__(self, rhs):
        if numpy.isscalar(rhs):
            return DivFromConstant(rhs).apply((self,))[0]
        return Div().apply((_preprocess_rhs(self, rhs), self))[0]

    def __pow__(self, rhs):
        if not numpy.isscalar(rhs):
            raise TypeError("only Variable ** scalar is restated")
        return PowVarConst(rhs).apply((self,))[0]


def _preprocess_rhs(x, value):
    if isinstance(value, Variable):
        return value
    if not isinstance(value, numpy.ndarray):
        raise TypeError("Value must be a Variable, scalar or ndarray: %r" % type(value))
    return value.astype(x.dtype, copy=False)


def _preprocess_const(x, value):
    return x.dtype.type(value)


def as_variable(x):
    if isinstance(x, Variable):
        return x
    return Variable(x, requires_grad=False)


# ----------------------------------------------------------------- FunctionNode
class FunctionNode:
    """chainer.FunctionNode: forward(inputs)->tuple of arrays,
    backward(target_input_indexes, grad_outputs)->tuple of Variable|None."""

    inputs = None
    outputs = None
    rank = 0
    _retain_in = None
    _retain_out = None

    def check_type_forward(self, in_types):
        pass

    def retain_inputs(self, indexes):
        self._retain_in = tuple(indexes)

    def retain_outputs(self, indexes):
        self._retain_out = tuple(indexes)

    def get_retained_inputs(self):
        return tuple(self.inputs[i] for i in self._retain_in)

    def get_retained_outputs(self):
        return tuple(self.outputs[i] for i in self._retain_out)

    def apply(self, inputs):
        ins = tuple(as_variable(x) for x in inputs)
        outs = self.forward(tuple(v.array for v in ins))
        if not isinstance(outs, tuple):
            raise TypeError("forward must return a tuple")
        req = any(v.requires_grad for v in ins)
        self.inputs = ins
        self.rank = max([v.rank for v in ins] + [0])
        ret = []
        for o in outs:
            v = Variable(o, requires_grad=req)
            if req:
                v.creator_node = self
                v.rank = self.rank + 1
            ret.append(v)
        self.outputs = tuple(ret)
        return self.outputs

    def forward(self, inputs):
        raise NotImplementedError

    def backward(self, target_input_indexes, grad_outputs):
        raise NotImplementedError


def _arr(g):
    return g.array if isinstance(g, Variable) else g


def _reduce(gs):
    acc = _arr(gs[0])
    for g in gs[1:]:
        acc = acc + _arr(g)
    return acc


def _backward(loss):
    if loss.creator_node is None:
        return
    if loss.grad is None:
        loss.grad = numpy.ones_like(loss.array)
    grads = {id(loss): [loss.grad]}
    keep = {id(loss): loss}
    heap, seen, cnt = [], set(), 0

    def push(node):
        nonlocal cnt
        if id(node) not in seen:
            seen.add(id(node))
            heapq.heappush(heap, (-node.rank, cnt, node))
            cnt += 1

    push(loss.creator_node)
    while heap:
        _, _, node = heapq.heappop(heap)
        gys = []
        for o in node.outputs:
            g = grads.pop(id(o), None)
            gys.append(None if g is None else Variable(_reduce(g), requires_grad=False))
        idx = tuple(i for i, x in enumerate(node.inputs) if x.requires_grad)
        gxs = node.backward(idx, tuple(gys))
        if len(gxs) == len(node.inputs):
            gxs = tuple(gxs[i] for i in idx)
        for i, gx in zip(idx, gxs):
            if gx is None:
                continue
            x = node.inputs[i]
            grads.setdefault(id(x), []).append(_arr(gx))
            keep[id(x)] = x
            if x.creator_node is not None:
                push(x.creator_node)
    for k, g in grads.items():
        x = keep[k]
        if x.creator_node is None and x.requires_grad:
            tot = _reduce(g)
            x.grad = tot if x.grad is None or x is loss else x.grad + tot


# ------------------------------------------------------------------- arithmetic
class Neg(FunctionNode):
    def forward(self, x):
        return _force_array(-x[0]),

    def backward(self, idx, gy):
        return -_arr(gy[0]),


class Add(FunctionNode):
    def forward(self, x):
        self.shapes = (x[0].shape, x[1].shape)
        return _force_array(x[0] + x[1]),

    def backward(self, idx, gy):
        g = _arr(gy[0])
        return tuple(_sum_to(g, self.shapes[i]) for i in idx)


class AddConstant(FunctionNode):
    def __init__(self, value):
        self.value = value

    def forward(self, x):
        return _force_array(x[0] + _preprocess_const(x[0], self.value)),

    def backward(self, idx, gy):
        return _arr(gy[0]),


class Sub(FunctionNode):
    def forward(self, x):
        self.shapes = (x[0].shape, x[1].shape)
        return _force_array(x[0] - x[1]),

    def backward(self, idx, gy):
        g = _arr(gy[0])
        out = (_sum_to(g, self.shapes[0]), _sum_to(-g, self.shapes[1]))
        return tuple(out[i] for i in idx)


class SubFromConstant(FunctionNode):
    def __init__(self, value):
        self.value = value

    def forward(self, x):
        return _force_array(_preprocess_const(x[0], self.value) - x[0]),

    def backward(self, idx, gy):
        return -_arr(gy[0]),


class Mul(FunctionNode):
    def forward(self, x):
        self.x = x
        return _force_array(x[0] * x[1]),

    def backward(self, idx, gy):
        g = _arr(gy[0])
        x0, x1 = self.x
        out = []
        for i in idx:
            out.append(_sum_to(g * (x1 if i == 0 else x0), (x0 if i == 0 else x1).shape))
        return tuple(out)


class MulConstant(FunctionNode):
    def __init__(self, value):
        self.value = value

    def forward(self, x):
        return _force_array(_preprocess_const(x[0], self.value) * x[0]),

    def backward(self, idx, gy):
        g = _arr(gy[0])
        return _force_array(_preprocess_const(g, self.value) * g),


class Div(FunctionNode):
    def forward(self, x):
        self.x = x
        return _force_array(x[0] / x[1]),

    def backward(self, idx, gy):  # DivGrad.forward_cpu
        g = _arr(gy[0])
        x0, x1 = self.x
        gx0 = _force_array(g / x1)
        gx1 = _force_array(-gx0 * x0 / x1)
        out = (_sum_to(gx0, x0.shape), _sum_to(gx1, x1.shape))
        return tuple(out[i] for i in idx)


class PowVarConst(FunctionNode):
    """functions/math/basic_math.py: x ** c; backward c * x ** (c - 1) * gy"""

    def __init__(self, value):
        self.value = value

    def forward(self, x):
        self.x = x[0]
        return _force_array(x[0] ** _preprocess_const(x[0], self.value)),

    def backward(self, idx, gy):
        v = _preprocess_const(self.x, self.value)
        return _force_array(v * (self.x ** _preprocess_const(self.x, self.value - 1)) * _arr(gy[0])),


class ReLU(FunctionNode):
    """functions/activation/relu.py: maximum(x, 0); backward gy * (y > 0)"""

    def forward(self, x):
        self.y = _force_array(numpy.maximum(x[0], 0, dtype=x[0].dtype))
        return self.y,

    def backward(self, idx, gy):
        return _force_array(_arr(gy[0]) * (self.y > 0)),


class Mean(FunctionNode):
    """functions/math/average.py (axis=None): x.mean(); backward gy / size broadcast"""

    def forward(self, x):
        self.shape, self.dtype = x[0].shape, x[0].dtype
        return _force_array(x[0].mean(), x[0].dtype),

    def backward(self, idx, gy):
        n = int(numpy.prod(self.shape))
        return numpy.broadcast_to(_arr(gy[0]) * self.dtype.type(1.0 / n), self.shape).astype(self.dtype),


class DivFromConstant(FunctionNode):
    def __init__(self, value):
        self.value = value

    def forward(self, x):
        self.x = x[0]
        return _force_array(_preprocess_const(x[0], self.value) / x[0]),

    def backward(self, idx, gy):
        g = _arr(gy[0])
        v = _preprocess_const(self.x, self.value)
        return _force_array(-v * g / (self.x ** 2)),


# ------------------------------------------------------------------- functions
class Clip(FunctionNode):
    def __init__(self, x_min, x_max):
        self.x_min, self.x_max = x_min, x_max

    def forward(self, x):
        self.x = x[0]
        return _force_array(numpy.clip(x[0], self.x_min, self.x_max), x[0].dtype),

    def backward(self, idx, gy):
        cond = (self.x_min <= self.x) & (self.x <= self.x_max)
        return _force_array(_arr(gy[0]) * cond, self.x.dtype),


class MatMul(FunctionNode):
    def forward(self, x):
        self.x = x
        return _force_array(numpy.matmul(x[0], x[1])),

    def backward(self, idx, gy):
        g = _arr(gy[0])
        a, b = self.x
        out = []
        for i in idx:
            if i == 0:
                out.append(_sum_to(numpy.matmul(g, numpy.swapaxes(b, -1, -2)), a.shape))
            else:
                out.append(_sum_to(numpy.matmul(numpy.swapaxes(a, -1, -2), g), b.shape))
        return tuple(out)


class GetItem(FunctionNode):
    def __init__(self, slices):
        if isinstance(slices, list):
            slices = tuple(slices)
        elif not isinstance(slices, tuple):
            slices = slices,
        self.slices = tuple(_arr(s) for s in slices)

    def forward(self, x):
        self.in_shape, self.in_dtype = x[0].shape, x[0].dtype
        return _force_array(x[0][self.slices]),

    def backward(self, idx, gy):  # GetItemGrad.forward (CPU branch)
        gx = numpy.zeros(self.in_shape, self.in_dtype)
        numpy.add.at(gx, self.slices, _arr(gy[0]))
        return gx,


class Reshape(FunctionNode):
    def __init__(self, shape):
        self.shape = tuple(shape)

    def forward(self, x):
        self.in_shape = x[0].shape
        return x[0].reshape(self.shape),

    def backward(self, idx, gy):
        return _arr(gy[0]).reshape(self.in_shape),


class Transpose(FunctionNode):
    def __init__(self, axes):
        self.axes = axes

    def forward(self, x):
        return x[0].transpose(self.axes),

    def backward(self, idx, gy):
        inv = None
        if self.axes is not None:
            inv = tuple(numpy.argsort([a % len(self.axes) for a in self.axes]))
        return _arr(gy[0]).transpose(inv),


class Concat(FunctionNode):
    def __init__(self, axis):
        self.axis = axis

    def forward(self, xs):
        self.sizes = [x.shape[self.axis] for x in xs]
        return numpy.concatenate(xs, axis=self.axis),

    def backward(self, idx, gy):
        cuts = numpy.cumsum(self.sizes)[:-1]
        parts = numpy.split(_arr(gy[0]), cuts, axis=self.axis)
        return tuple(parts[i] for i in idx)


class ScatterAdd(FunctionNode):
    def __init__(self, slices):
        if isinstance(slices, list):
            slices = tuple(slices)
        elif not isinstance(slices, tuple):
            slices = slices,
        self.slices = slices

    def forward(self, x):
        a = x[0].copy()
        numpy.add.at(a, self.slices, x[1])
        self.b_shape = x[1].shape
        return a,

    def backward(self, idx, gy):
        g = _arr(gy[0])
        out = (g, _sum_to(g[self.slices], self.b_shape))
        return tuple(out[i] for i in idx)


class MeanAbsoluteError(FunctionNode):
    def forward(self, x):
        self.diff = x[0] - x[1]
        diff = self.diff.ravel()
        return numpy.array(abs(diff).sum() / diff.size, dtype=diff.dtype),

    def backward(self, idx, gy):
        g = _arr(gy[0])
        coeff = g * g.dtype.type(1. / self.diff.size)
        gx0 = numpy.broadcast_to(coeff, self.diff.shape) * numpy.sign(self.diff)
        out = (gx0, -gx0)
        return tuple(out[i] for i in idx)


class MeanSquaredError(FunctionNode):
    def forward(self, x):
        self.x = x
        diff = (x[0] - x[1]).ravel()
        return numpy.array(diff.dot(diff) / diff.size, dtype=diff.dtype),

    def backward(self, idx, gy):
        x0, x1 = self.x
        diff = x0 - x1
        gy0 = numpy.broadcast_to(_arr(gy[0]), diff.shape)
        gx0 = gy0 * diff * diff.dtype.type(2. / diff.size)
        out = (gx0, -gx0)
        return tuple(out[i] for i in idx)


class Cumsum(FunctionNode):
    """chainer.functions.cumsum (math/cumsum.py): forward xp.cumsum, backward flip(cumsum(flip(gy)))"""
    def __init__(self, axis=None):
        self.axis = axis

    def forward(self, x):
        return _force_array(numpy.cumsum(x[0], axis=self.axis), x[0].dtype),

    def backward(self, idx, gy):
        g = _arr(gy[0])
        return _force_array(numpy.flip(numpy.cumsum(numpy.flip(g, self.axis), axis=self.axis), self.axis), g.dtype),


class Sum(FunctionNode):
    """chainer.functions.sum (math/sum.py): forward x.sum(axis), backward broadcast of gy"""
    def __init__(self, axis=None, keepdims=False):
        self.axis = axis if (axis is None or isinstance(axis, tuple)) else (axis,)
        self.keepdims = keepdims

    def forward(self, x):
        self.in_shape = x[0].shape
        return _force_array(x[0].sum(axis=self.axis, keepdims=self.keepdims), x[0].dtype),

    def backward(self, idx, gy):
        g = _arr(gy[0])
        if not self.keepdims and self.axis is not None:
            for a in sorted(ax % len(self.in_shape) for ax in self.axis):
                g = numpy.expand_dims(g, a)
        return numpy.ascontiguousarray(numpy.broadcast_to(g, self.in_shape)),


class Sigmoid(FunctionNode):
    """chainer.functions.sigmoid (activation/sigmoid.py): forward_cpu tanh(x*0.5)*0.5+0.5, backward gy*y*(1-y)"""
    def forward(self, x):
        half = x[0].dtype.type(0.5)
        self.y = _force_array(numpy.tanh(x[0] * half) * half + half, x[0].dtype)
        return self.y,

    def backward(self, idx, gy):
        one = self.y.dtype.type(1)
        return _force_array(_arr(gy[0]) * self.y * (one - self.y), self.y.dtype),


class LeakyReLU(FunctionNode):
    """chainer.functions.leaky_relu (activation/leaky_relu.py): y = x; y[x < 0] *= slope; backward on y < 0"""
    def __init__(self, slope=0.2):
        self.slope = slope

    def forward(self, x):
        y = x[0].copy()
        y[x[0] < 0] *= self.slope
        self.y = y
        return y,

    def backward(self, idx, gy):
        g = _arr(gy[0]).copy()
        g[self.y < 0] *= self.slope
        return g,


class Softplus(FunctionNode):
    """chainer.functions.softplus (activation/softplus.py): forward_cpu (fmax(bx, 0) + log1p(exp(-fabs(bx)))) / beta,
    backward (SoftplusGrad.forward_cpu) gy * (1 - 1 / (1 + exp(beta * x)))"""
    def __init__(self, beta=1.0):
        self.beta = float(beta)
        self.beta_inv = float(1.0 / beta)

    def forward(self, x):
        self.x = x[0]
        bx = self.beta * x[0]
        y = (numpy.fmax(bx, 0) + numpy.log1p(numpy.exp(-numpy.fabs(bx)))) * self.beta_inv
        return _force_array(y.astype(x[0].dtype)),

    def backward(self, idx, gy):
        x = self.x
        gx = (1 - 1 / (1 + numpy.exp(self.beta * x))) * _arr(gy[0])
        return _force_array(gx.astype(x.dtype)),


functions = types.ModuleType("chainer.functions")
functions.softplus = lambda x, beta=1.0: Softplus(beta).apply((x,))[0]
functions.cumsum = lambda x, axis=None: Cumsum(axis).apply((x,))[0]
functions.sum = lambda x, axis=None, keepdims=False: Sum(axis, keepdims).apply((x,))[0]
functions.sigmoid = lambda x: Sigmoid().apply((x,))[0]
functions.leaky_relu = lambda x, slope=0.2: LeakyReLU(slope).apply((x,))[0]
functions.get_item = lambda x, slices: GetItem(slices).apply((x,))[0]
functions.reshape = lambda x, shape: Reshape(shape).apply((x,))[0]
functions.transpose = lambda x, axes=None: Transpose(axes).apply((x,))[0]
functions.clip = lambda x, x_min, x_max: Clip(x_min, x_max).apply((x,))[0]
functions.matmul = lambda a, b: MatMul().apply((a, b))[0]
functions.concat = lambda xs, axis=1: Concat(axis).apply(tuple(xs))[0]
functions.scatter_add = lambda a, slices, b: ScatterAdd(slices).apply((a, b))[0]
functions.mean_absolute_error = lambda x0, x1: MeanAbsoluteError().apply((x0, x1))[0]
functions.mean_squared_error = lambda x0, x1: MeanSquaredError().apply((x0, x1))[0]
functions.relu = lambda x: ReLU().apply((x,))[0]
functions.mean = lambda x: Mean().apply((x,))[0]


# --------------------------------------------------------- package registration
class _Stub(types.ModuleType):
    """Permissive module: any attribute is a dummy class (network layers etc.)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, cls)
        return cls


def install():
    """Register the shim as `chainer` in sys.modules (idempotent)."""
    if "chainer" in sys.modules and getattr(sys.modules["chainer"], "_is_shim", False):
        return sys.modules["chainer"]
    ch = _Stub("chainer")
    ch._is_shim = True
    ch.__version__ = "7.8.1-shim"
    ch.Variable = Variable
    ch.FunctionNode = FunctionNode
    ch.functions = functions
    ch.as_variable = as_variable
    cuda = types.ModuleType("chainer.backends.cuda")
    cuda.get_array_module = get_array_module
    backends = types.ModuleType("chainer.backends")
    backends.cuda = cuda
    backend = types.ModuleType("chainer.backend")
    backend.get_array_module = get_array_module
    ch.cuda, ch.backends, ch.backend = cuda, backends, backend
    ch.links = _Stub("chainer.links")
    sys.modules.update({
        "chainer": ch, "chainer.functions": functions, "chainer.cuda": cuda,
        "chainer.backends": backends, "chainer.backends.cuda": cuda,
        "chainer.backend": backend, "chainer.links": ch.links,
    })
    return ch
