"""D2Q9 operators with the names and signatures of the reference's ``vivsim.lbm``
(vivsim/lbm/__init__.py:1-42), running as sm_100a CUDA kernels behind the C ABI.

Arrays are fp32 torch CUDA tensors: f (9, NX, NY), rho (NX, NY), u / g (2, NX, NY)."""

from .. import _api

_api.bind(2, globals())
