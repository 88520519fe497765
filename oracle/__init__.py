"""CPU oracle for the IB-LBM time step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy fp32 restatement of the algorithm implemented by the reference
(haimingz/vivsim v2.0.0, pure Python over jax.numpy).  Every function cites
the reference file:line it follows.  Only ``tests/``, ``__graft_entry__.smoke``
and ``bench.py``'s CPU-baseline / reference arm may import this package; the
product (``vivsim_b200``) never does and fails loudly without its CUDA library.

Parity status: the reference ships no tests or golden vectors and jax/jaxlib
cannot be installed in this image, so this oracle is pinned against the
reference's *own source executed unmodified on a NumPy-backed jax stand-in*
(``oracle/jaxshim``; fixtures in ``tests/golden/*.npz`` made by
``tests/golden/make_golden.py``).  That pins algorithm, indexing, operation
order and constants; XLA's fp32 rounding itself remains unpinned.
"""

from . import lattice, lbm, lbm3d, ib, ib3d, dyn, recipes  # noqa: F401
