"""D3Q19 operators with the names and signatures of the reference's ``vivsim.lbm3d``
(vivsim/lbm3d/__init__.py:1-44), running as sm_100a CUDA kernels behind the C ABI.

Arrays are fp32 torch CUDA tensors: f (19, NX, NY, NZ), rho (NX, NY, NZ), u / g (3, NX, NY, NZ)."""

from .. import _api

_api.bind(3, globals())
