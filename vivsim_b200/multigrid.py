"""Grid-refinement transfers with the names and signatures of the reference's ``vivsim.multigrid``
(vivsim/multigrid.py:1-170): D2Q9 blocks whose spacing differs by a factor 2 exchange the populations that cross
their common edge through a buffer layer.  The two transfers are line kernels behind the C ABI
(``vsb_mg_fine_to_coarse`` / ``vsb_mg_coarse_to_fine``); the rest is host arithmetic.

Arrays are fp32 torch CUDA tensors of shape (9, nx, ny)."""

import ctypes as C

import torch

from . import _lib as L


def init_grid(width, height, level=0, buffer_x=0, buffer_y=0, device="cuda"):
    """Zero populations, unit density and zero velocity for a block at refinement ``level``
    (multigrid.py:21-55).  Returns (f, rho, u)."""
    nx = int(width * 2 ** level) + buffer_x
    ny = int(height * 2 ** level) + buffer_y
    f = torch.zeros((9, nx, ny), dtype=torch.float32, device=device)
    rho = torch.ones((nx, ny), dtype=torch.float32, device=device)
    u = torch.zeros((2, nx, ny), dtype=torch.float32, device=device)
    return f, rho, u


def _blocks(f_fine, f_coarse):
    f_fine = L.dev(f_fine, name="f_fine")
    f_coarse = L.dev(f_coarse, name="f_coarse")
    for name, f in (("f_fine", f_fine), ("f_coarse", f_coarse)):
        if f.ndim != 3 or f.shape[0] != 9:
            raise ValueError(f"{name}: expected shape (9, nx, ny), got {tuple(f.shape)}")
    return f_fine, f_coarse


def _check_shapes(f_fine, f_coarse, dir):
    # the reference fails with a broadcasting ValueError when the edge lines do not match 2 : 1
    if dir in ("left", "right") and f_fine.shape[2] != 2 * f_coarse.shape[2]:
        raise ValueError(f"fine ny ({f_fine.shape[2]}) must be twice the coarse ny ({f_coarse.shape[2]})")
    if dir in ("up", "down") and f_fine.shape[1] != 2 * f_coarse.shape[1]:
        raise ValueError(f"fine nx ({f_fine.shape[1]}) must be twice the coarse nx ({f_coarse.shape[1]})")


def fine_to_coarse(f_fine, f_coarse, dir):
    """Coarse receiving edge line <- mean of the 2 x 2 fine cells it covers, for the three populations travelling
    towards ``dir`` ('left', 'right', 'up', 'down'); any other ``dir`` leaves f_coarse unchanged like the reference
    (multigrid.py:58-101).  Returns the updated copy of f_coarse."""
    f_fine, f_coarse = _blocks(f_fine, f_coarse)
    out = f_coarse.clone()
    if dir not in L.MG_DIR:
        return out
    _check_shapes(f_fine, f_coarse, dir)
    L.check(L.lib().vsb_mg_fine_to_coarse(f_fine.shape[1], f_fine.shape[2], L.ptr(f_fine), out.shape[1], out.shape[2],
                                          L.ptr(out), L.MG_DIR[dir], L.stream()))
    return out


def coarse_to_fine(f_coarse, f_fine, dir):
    """Fine receiving edge line <- piecewise-constant copy of the coarse edge line (multigrid.py:103-131).
    Returns the updated copy of f_fine."""
    f_fine, f_coarse = _blocks(f_fine, f_coarse)
    out = f_fine.clone()
    if dir not in L.MG_DIR:
        return out
    _check_shapes(f_fine, f_coarse, dir)
    L.check(L.lib().vsb_mg_coarse_to_fine(f_coarse.shape[1], f_coarse.shape[2], L.ptr(f_coarse), out.shape[1],
                                          out.shape[2], L.ptr(out), L.MG_DIR[dir], L.stream()))
    return out


def get_omega(nu, level=0):
    """Relaxation parameter at a refinement level (multigrid.py:134-149)."""
    omega_l0 = 1 / (3 * nu + 0.5)
    return 2 * omega_l0 / (2 ** (level + 1) + (1 - 2 ** level) * omega_l0)


def coord_to_indices(x, y, grid_start_x, grid_start_y, level=0):
    """Global coordinates -> local indices of a block at ``level`` (multigrid.py:152-170)."""
    return int((x - grid_start_x) * 2 ** level), int((y - grid_start_y) * 2 ** level)
