"""NumPy-backed stand-in for ``jax.numpy`` -- TEST INFRASTRUCTURE ONLY.

jax/jaxlib are not installable in this image (no wheel, no network), so the
reference (``/root/reference/vivsim``) cannot be imported as-is.  This module
provides just enough of the ``jax.numpy`` surface for the *unmodified*
reference source to execute on NumPy, with JAX's x64-disabled dtype rules
imitated: every float64 / int64 result is demoted to float32 / int32
immediately after each primitive call, Python scalars are weak-typed (NumPy 2
NEP-50 already does that) and ``x.at[idx].set/add`` are functional updates.

What this pins: the reference's algorithm, indexing, operation order and
constants.  What it does not pin: XLA's own fp32 rounding / reduction order.

Used only by ``tests/golden/make_golden.py`` (run in the build container where
``/root/reference`` exists).  Nothing in the product imports it.
"""

import numpy as _np

pi = _np.pi
float32 = _np.float32
int32 = _np.int32
bool_ = _np.bool_
newaxis = None


def _demote_dtype(dt):
    dt = _np.dtype(dt)
    if dt == _np.float64:
        return _np.dtype(_np.float32)
    if dt == _np.int64:
        return _np.dtype(_np.int32)
    if dt == _np.complex128:
        return _np.dtype(_np.complex64)
    return dt


def _raw(x):
    if isinstance(x, Array):
        return x.view(_np.ndarray)
    if isinstance(x, (list, tuple)):
        return type(x)(_raw(v) for v in x)
    if isinstance(x, dict):
        return {k: _raw(v) for k, v in x.items()}
    return x


def _wrap(x):
    if isinstance(x, _np.ndarray):
        dt = _demote_dtype(x.dtype)
        if dt != x.dtype:
            x = x.astype(dt)
        return x.view(Array)
    if isinstance(x, _np.generic):
        return _wrap(_np.asarray(x))
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


class _AtIndexer:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtUpdater(self._arr, idx)


class _AtUpdater:
    def __init__(self, arr, idx):
        self._arr = arr
        self._idx = _raw(idx)

    def set(self, values):
        out = _np.array(_raw(self._arr), copy=True)
        out[self._idx] = _np.asarray(_raw(values)).astype(out.dtype, copy=False)
        return _wrap(out)

    def add(self, values):
        out = _np.array(_raw(self._arr), copy=True)
        vals = _np.asarray(_raw(values)).astype(out.dtype, copy=False)
        _np.add.at(out, self._idx, vals)
        return _wrap(out)


class Array(_np.ndarray):
    """ndarray subclass imitating an immutable, x64-disabled jax.Array."""

    __array_priority__ = 1000

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        ins = tuple(_raw(x) for x in inputs)
        if out is not None:
            kwargs["out"] = tuple(_raw(o) for o in out)
        return _wrap(getattr(ufunc, method)(*ins, **kwargs))

    def __array_function__(self, func, types, args, kwargs):
        return _wrap(func(*_raw(args), **_raw(kwargs)))

    def __getitem__(self, idx):
        return _wrap(_np.ndarray.__getitem__(self.view(_np.ndarray), _raw(idx)))

    def __setitem__(self, idx, value):
        raise TypeError("jax arrays are immutable; use .at[idx].set()")

    @property
    def at(self):
        return _AtIndexer(self)

    def astype(self, dtype, **kw):
        return _wrap(self.view(_np.ndarray).astype(_demote_dtype(dtype)))

    def block_until_ready(self):
        return self

    # in-place operators rebind in JAX; never mutate the operand
    def __iadd__(self, o): return self + o
    def __isub__(self, o): return self - o
    def __imul__(self, o): return self * o
    def __itruediv__(self, o): return self / o

    # python-scalar conversions on 0-d results
    def __iter__(self):
        if self.ndim == 0:
            raise TypeError("iteration over a 0-d array")
        return (self[i] for i in range(self.shape[0]))


ndarray = Array


def asarray(x, dtype=None):
    a = _np.asarray(_raw(x), dtype=None if dtype is None else _demote_dtype(dtype))
    return _wrap(a)


def array(x, dtype=None):
    a = _np.array(_raw(x), dtype=None if dtype is None else _demote_dtype(dtype))
    return _wrap(a)


def zeros(shape, dtype=float32):
    return _wrap(_np.zeros(shape, dtype=_demote_dtype(dtype)))


def ones(shape, dtype=float32):
    return _wrap(_np.ones(shape, dtype=_demote_dtype(dtype)))


def full(shape, fill_value, dtype=None):
    if dtype is None:
        dtype = _demote_dtype(_np.asarray(_raw(fill_value)).dtype)
    return _wrap(_np.full(shape, _raw(fill_value), dtype=_demote_dtype(dtype)))


def zeros_like(x, dtype=None):
    return _wrap(_np.zeros_like(_raw(x), dtype=dtype))


def ones_like(x, dtype=None):
    return _wrap(_np.ones_like(_raw(x), dtype=dtype))


def arange(*args, dtype=None, **kw):
    return _wrap(_np.arange(*args, dtype=None if dtype is None else _demote_dtype(dtype), **kw))


def linspace(*args, dtype=None, **kw):
    return _wrap(_np.linspace(*args, **kw).astype(_demote_dtype(dtype or _np.float32)))


def eye(n, dtype=float32):
    return _wrap(_np.eye(n, dtype=_demote_dtype(dtype)))


def isscalar(x):
    return _np.isscalar(x) or (hasattr(x, "ndim") and x.ndim == 0)


def ndim(x):
    return _np.ndim(_raw(x))


def einsum(subscripts, *operands, precision=None, **kw):
    return _wrap(_np.einsum(subscripts, *[_np.asarray(_raw(o)) for o in operands], **kw))


def tensordot(a, b, axes=2, precision=None):
    return _wrap(_np.tensordot(_np.asarray(_raw(a)), _np.asarray(_raw(b)), axes=axes))


def dot(a, b, precision=None):
    return _wrap(_np.dot(_np.asarray(_raw(a)), _np.asarray(_raw(b))))


def meshgrid(*xs, indexing="xy"):
    return [_wrap(m) for m in _np.meshgrid(*_raw(xs), indexing=indexing)]


class _Linalg:
    @staticmethod
    def norm(x, *a, **kw):
        return _wrap(_np.linalg.norm(_np.asarray(_raw(x)), *a, **kw))

    @staticmethod
    def solve(a, b):
        return _wrap(_np.linalg.solve(_np.asarray(_raw(a)), _np.asarray(_raw(b))))

    @staticmethod
    def inv(a):
        return _wrap(_np.linalg.inv(_np.asarray(_raw(a))))


linalg = _Linalg()


def _delegate(name):
    fn = getattr(_np, name)

    def wrapper(*args, **kwargs):
        return _wrap(fn(*_raw(args), **_raw(kwargs)))

    wrapper.__name__ = name
    return wrapper


for _name in (
    "concatenate stack sum abs sqrt cos sin where floor roll pad repeat cross take "
    "moveaxis swapaxes clip diag maximum minimum mean max min exp log reshape "
    "transpose expand_dims squeeze any all isnan isfinite cumsum prod sign square "
    "tile broadcast_to gradient argmax argmin"
).split():
    globals()[_name] = _delegate(_name)
