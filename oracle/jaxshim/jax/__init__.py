"""NumPy-backed stand-in for the ``jax`` package -- TEST INFRASTRUCTURE ONLY.

See ``jax/numpy.py`` in this directory for what this is and is not.  It exists
so that ``tests/golden/make_golden.py`` can execute the unmodified reference
source in a container where jax/jaxlib cannot be installed.
"""

from . import numpy  # noqa: F401
from . import lax  # noqa: F401
from .numpy import Array  # noqa: F401

__version__ = "0.0-numpy-shim"


def jit(fn=None, **_kw):
    if fn is None:
        return lambda f: f
    return fn


def block_until_ready(x):
    return x


def devices():
    return ["numpy-shim-cpu"]
