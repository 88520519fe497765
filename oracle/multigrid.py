"""NumPy fp32 restatement of the reference's ``vivsim/multigrid.py`` (TEST INFRASTRUCTURE ONLY).  Pinned against
tests/golden/multigrid.npz (the unmodified reference source executed on oracle/jaxshim)."""

import numpy as np

F32 = np.float32
# lbm/lattice.py:58-61
DIRS = {"left": [3, 7, 6], "right": [1, 5, 8], "up": [2, 5, 6], "down": [4, 7, 8]}


def fine_to_coarse(f_fine, f_coarse, dir):          # multigrid.py:58-101
    ff = np.asarray(f_fine, dtype=F32)
    fc = np.array(f_coarse, dtype=F32, copy=True)
    if dir not in DIRS:
        return fc
    q = DIRS[dir]
    if dir == "left":
        fc[q, -1] = F32(0.25) * (ff[q, 0, 0::2] + ff[q, 0, 1::2] + ff[q, 1, 0::2] + ff[q, 1, 1::2])
    elif dir == "right":
        fc[q, 0] = F32(0.25) * (ff[q, -1, 0::2] + ff[q, -1, 1::2] + ff[q, -2, 0::2] + ff[q, -2, 1::2])
    elif dir == "up":
        fc[q, :, 0] = F32(0.25) * (ff[q, 0::2, -1] + ff[q, 1::2, -1] + ff[q, 0::2, -2] + ff[q, 1::2, -2])
    else:
        fc[q, :, -1] = F32(0.25) * (ff[q, 0::2, 0] + ff[q, 1::2, 0] + ff[q, 0::2, 1] + ff[q, 1::2, 1])
    return fc


def coarse_to_fine(f_coarse, f_fine, dir):          # multigrid.py:103-131
    fc = np.asarray(f_coarse, dtype=F32)
    ff = np.array(f_fine, dtype=F32, copy=True)
    if dir not in DIRS:
        return ff
    q = DIRS[dir]
    if dir == "left":
        ff[q, -1, :] = np.repeat(fc[q, 0], 2, axis=-1)
    elif dir == "right":
        ff[q, 0, :] = np.repeat(fc[q, -1], 2, axis=-1)
    elif dir == "up":
        ff[q, :, 0] = np.repeat(fc[q, :, -1], 2, axis=-1)
    else:
        ff[q, :, -1] = np.repeat(fc[q, :, 0], 2, axis=-1)
    return ff


def get_omega(nu, level=0):                          # multigrid.py:134-149
    omega_l0 = 1 / (3 * nu + 0.5)
    return 2 * omega_l0 / (2 ** (level + 1) + (1 - 2 ** level) * omega_l0)


def coord_to_indices(x, y, grid_start_x, grid_start_y, level=0):   # multigrid.py:152-170
    return int((x - grid_start_x) * 2 ** level), int((y - grid_start_y) * 2 ** level)
