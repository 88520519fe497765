"""``jax.lax`` subset on NumPy -- TEST INFRASTRUCTURE ONLY (see numpy.py)."""

import numpy as _np

from . import numpy as jnp


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(int(lower), int(upper)):
        val = body_fun(i, val)
    return val


def scan(f, init, xs=None, length=None):
    carry = init
    ys = []
    n = length if xs is None else len(xs)
    for i in range(n):
        carry, y = f(carry, None if xs is None else xs[i])
        ys.append(y)
    if ys and ys[0] is not None:
        if isinstance(ys[0], tuple):
            ys = tuple(jnp.stack([y[k] for y in ys]) for k in range(len(ys[0])))
        else:
            ys = jnp.stack(ys)
    else:
        ys = None
    return carry, ys


def _clamped_starts(shape, starts, sizes):
    # XLA clamps start indices so that the slice stays inside the operand
    return [int(min(max(int(s), 0), dim - size)) for s, dim, size in zip(starts, shape, sizes)]


def dynamic_slice(operand, start_indices, slice_sizes):
    starts = _clamped_starts(operand.shape, start_indices, slice_sizes)
    idx = tuple(slice(s, s + n) for s, n in zip(starts, slice_sizes))
    return jnp.asarray(_np.asarray(operand)[idx].copy())


def dynamic_update_slice(operand, update, start_indices):
    starts = _clamped_starts(operand.shape, start_indices, update.shape)
    out = _np.array(_np.asarray(operand), copy=True)
    idx = tuple(slice(s, s + n) for s, n in zip(starts, update.shape))
    out[idx] = _np.asarray(update)
    return jnp.asarray(out)
