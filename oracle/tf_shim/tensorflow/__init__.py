"""Minimal `tensorflow` stand-in, backed by torch-CPU tensors.  TEST INFRASTRUCTURE ONLY.

Purpose: TensorFlow is not installable in the build container, yet the reference's hot path
(/root/reference/wflib/IDEAL_model.py, tf2gan/loss.py) is pure `tf.*` op chains.  This module
implements exactly the op surface those two files touch, with TensorFlow's documented semantics, on
torch-CPU complex64/float32 tensors, so that the UNMODIFIED reference source can be imported and
executed by `oracle/gen_golden.py` to produce the golden vectors under `tests/golden/`.  Gradients of
the reference code are then available through torch autograd (`tf.stop_gradient` -> `.detach()`).

It is not a product dependency: nothing under `ideal-gan_b200/` imports it.  Elementary ops are IEEE
fp32 in both libraries; only reduction order inside matmul / qr / solve may differ from TF's Eigen
kernels (<= a few 1e-7 relative), which is far inside the 1e-5 parity tolerance.
"""
import builtins
import types

import numpy as _np
import torch as _t

__version__ = "2.8.2-shim"

float32 = _t.float32
float64 = _t.float64
complex64 = _t.complex64
complex128 = _t.complex128
int32 = _t.int32
int64 = _t.int64
bool = _t.bool  # noqa: A001  (mirrors tf.bool)

class Tensor(_t.Tensor):
    """torch tensor with TensorFlow's value semantics for augmented assignment: tf.Tensors are
    immutable, so `x *= y` in the reference rebinds `x` to a new tensor.  The default
    `__torch_function__` keeps this subclass through every op, so wrapping the inputs is enough."""

    def __iadd__(self, other):
        return _t.add(self, other)

    def __isub__(self, other):
        return _t.sub(self, other)

    def __imul__(self, other):
        return _t.mul(self, other)

    def __itruediv__(self, other):
        return _t.div(self, other)


def _is_t(x):
    return isinstance(x, _t.Tensor)


def convert_to_tensor(value, dtype=None, name=None):
    if _is_t(value):
        if not isinstance(value, Tensor):
            value = value.as_subclass(Tensor)
        return value if dtype is None or value.dtype == dtype else value.to(dtype)
    arr = _np.asarray(value)
    if dtype is None:
        if arr.dtype == _np.float64 and not isinstance(value, _np.ndarray):
            dtype = float32  # python floats default to float32 in TF
        elif arr.dtype == _np.int64 and not isinstance(value, _np.ndarray):
            dtype = int32
    out = _t.from_numpy(_np.ascontiguousarray(arr)).as_subclass(Tensor)
    return out if dtype is None else out.to(dtype)


def constant(value, dtype=None, shape=None, name=None):
    out = convert_to_tensor(value, dtype=dtype)
    if shape is not None:
        out = out.expand(*shape).clone()
    return out


def cast(x, dtype):
    if not _is_t(x):
        return _t.tensor(x, dtype=dtype)
    return x.to(dtype)


def _ints(shape):
    return [int(s) for s in shape]


def reshape(x, shape):
    return _t.reshape(x, _ints(shape))


def expand_dims(x, axis):
    return _t.unsqueeze(x, axis)


def squeeze(x, axis=None):
    return _t.squeeze(x) if axis is None else _t.squeeze(x, axis)


def tile(x, multiples):
    return x.repeat(*_ints(multiples))


def transpose(x, perm=None, conjugate=False):
    if perm is None:
        perm = list(builtins.range(x.dim()))[::-1]
    out = x.permute(*perm)
    return _t.conj_physical(out) if conjugate and out.is_complex() else out


def concat(values, axis):
    return _t.cat(list(values), dim=axis)


def stack(values, axis=0):
    return _t.stack(list(values), dim=axis)


def eye(n, dtype=float32):
    return _t.eye(int(n), dtype=dtype)


def ones(shape, dtype=float32):
    return _t.ones(_ints(shape), dtype=dtype)


def zeros(shape, dtype=float32):
    return _t.zeros(_ints(shape), dtype=dtype)


def ones_like(x):
    return _t.ones_like(x)


def zeros_like(x):
    return _t.zeros_like(x)


def range(start, limit=None, delta=1, dtype=None):  # noqa: A001
    if limit is None:
        start, limit = 0, start
    return _t.arange(start, limit, delta, dtype=dtype)


def _as_like(v, ref):
    return v if _is_t(v) else _t.tensor(v, dtype=ref.dtype)


def complex(real, imag):  # noqa: A001
    if not _is_t(real) and not _is_t(imag):
        real = _t.tensor(real, dtype=float32)
        imag = _t.tensor(imag, dtype=float32)
    elif not _is_t(real):
        real = _t.tensor(real, dtype=imag.dtype)
    elif not _is_t(imag):
        imag = _t.tensor(imag, dtype=real.dtype)
    real, imag = _t.broadcast_tensors(real, imag)
    return _t.complex(real.contiguous(), imag.contiguous())


def where(condition, x=None, y=None):
    if x is None and y is None:
        return _t.nonzero(condition)
    ref = x if _is_t(x) else y
    return _t.where(condition, _as_like(x, ref), _as_like(y, ref))


def maximum(x, y):
    ref = x if _is_t(x) else y
    return _t.maximum(_as_like(x, ref), _as_like(y, ref))


def minimum(x, y):
    ref = x if _is_t(x) else y
    return _t.minimum(_as_like(x, ref), _as_like(y, ref))


def abs(x):  # noqa: A001
    return _t.abs(x)


def square(x):
    return x * x


def sqrt(x):
    return _t.sqrt(x)


def pow(x, y):  # noqa: A001
    return _t.pow(x, y)


def clip_by_value(t, clip_value_min, clip_value_max, name=None):
    """tf.clip_by_value = minimum(maximum(t, lo), hi); NaN stays NaN in both libraries."""
    return _t.clamp(t, clip_value_min, clip_value_max)


def reduce_sum(x, axis=None, keepdims=False):
    if axis is None:
        return _t.sum(x)
    return _t.sum(x, dim=axis, keepdim=keepdims)


def reduce_mean(x, axis=None, keepdims=False):
    if axis is None:
        return _t.mean(x)
    return _t.mean(x, dim=axis, keepdim=keepdims)


def matmul(a, b, transpose_a=False, transpose_b=False, adjoint_a=False, adjoint_b=False):
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    if adjoint_a:
        a = _t.conj_physical(a.transpose(-1, -2))
    if adjoint_b:
        b = _t.conj_physical(b.transpose(-1, -2))
    return _t.matmul(a, b)


def stop_gradient(x):
    return x.detach()


class GradientTape:
    """`tf.GradientTape` on torch autograd: `watch` marks a leaf, `gradient` differentiates the recorded graph
    (oracle/tf_ref.py replays the golden fixtures through the same tape code real TensorFlow runs)."""

    def __init__(self, persistent=False, watch_accessed_variables=True):
        self._persistent = persistent

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def watch(self, x):
        x.requires_grad_(True)

    def gradient(self, target, sources):
        single = not isinstance(sources, (list, tuple))
        grads = _t.autograd.grad(target, [sources] if single else list(sources), allow_unused=True, retain_graph=self._persistent)
        return grads[0] if single else list(grads)


def function(func=None, **kwargs):
    """`@tf.function` (bare or with arguments): eager pass-through."""
    if func is None:
        return lambda f: f
    return func


def _divide_no_nan(x, y):
    ref = x if _is_t(x) else y
    x, y = _as_like(x, ref), _as_like(y, ref)
    safe = _t.where(y == 0, _t.ones_like(y), y)
    return _t.where(y == 0, _t.zeros_like(x / safe), x / safe)


def _assert_all_finite(x, message, name=None):
    if not builtins.bool(_t.isfinite(x).all()):
        raise ValueError("InvalidArgumentError: " + message)
    return x


math = types.SimpleNamespace(
    exp=_t.exp, log=_t.log, real=_t.real, imag=_t.imag, abs=_t.abs, angle=_t.angle,
    conj=_t.conj_physical, sqrt=_t.sqrt, square=square, cos=_t.cos, sin=_t.sin,
    reduce_prod=lambda x, axis=None: _t.prod(x) if axis is None else _t.prod(x, dim=axis),
    reduce_sum=reduce_sum, reduce_mean=reduce_mean, divide_no_nan=_divide_no_nan,
    bessel_i0e=_t.special.i0e, maximum=maximum, minimum=minimum, pow=pow,
)


def _qr(x, full_matrices=False):
    q, r = _t.linalg.qr(x, mode="complete" if full_matrices else "reduced")
    return q, r


def _diag(x):
    return _t.diag_embed(x)


linalg = types.SimpleNamespace(
    matmul=matmul, qr=_qr, solve=_t.linalg.solve, inv=_t.linalg.inv, diag=_diag,
    diag_part=lambda x: _t.diagonal(x, dim1=-2, dim2=-1),
    matvec=lambda a, b: _t.matmul(a, b.unsqueeze(-1)).squeeze(-1),
)

def reduce_prod(x, axis=None, keepdims=False):
    if axis is None:
        return _t.prod(x)
    return _t.prod(x, dim=axis, keepdim=keepdims)


def _total_variation(images, name=None):
    """tf.image.total_variation: images (n, H, W, C) -> (n,), or (H, W, C) -> scalar; sum of absolute neighbour differences."""
    if images.dim() == 3:
        return (images[1:] - images[:-1]).abs().sum() + (images[:, 1:] - images[:, :-1]).abs().sum()
    return (images[:, 1:] - images[:, :-1]).abs().sum(dim=(1, 2, 3)) + (images[:, :, 1:] - images[:, :, :-1]).abs().sum(dim=(1, 2, 3))


image = types.SimpleNamespace(total_variation=_total_variation)


nn = types.SimpleNamespace(relu=_t.relu)
debugging = types.SimpleNamespace(assert_all_finite=_assert_all_finite)


class _Layer:
    """`tf.keras.layers.Layer`: `__call__` forwards to `call` (the reference layers hold no weights)."""

    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)


class _Loss:
    """`tf.keras.losses.Loss`: `__call__(y_true, y_pred)` forwards to `call` (default reduction is a
    mean of an already-reduced scalar in the reference's subclasses)."""

    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, y_true, y_pred, sample_weight=None):
        return self.call(y_true, y_pred)


class _MeanSquaredError(_Loss):
    def call(self, y_true, y_pred):
        return _t.mean((y_pred - y_true) ** 2)


class _MeanAbsoluteError(_Loss):
    def call(self, y_true, y_pred):
        return _t.mean(_t.abs(y_pred - y_true))


losses = types.SimpleNamespace(Loss=_Loss, MeanSquaredError=_MeanSquaredError,
                               MeanAbsoluteError=_MeanAbsoluteError)
keras = types.ModuleType("tensorflow.keras")
keras.layers = types.ModuleType("tensorflow.keras.layers")
keras.layers.Layer = _Layer
keras.losses = types.ModuleType("tensorflow.keras.losses")
keras.losses.Loss = _Loss
keras.losses.MeanSquaredError = _MeanSquaredError
keras.losses.MeanAbsoluteError = _MeanAbsoluteError

import sys as _sys  # noqa: E402

_sys.modules.setdefault("tensorflow.keras", keras)
_sys.modules.setdefault("tensorflow.keras.layers", keras.layers)
_sys.modules.setdefault("tensorflow.keras.losses", keras.losses)
