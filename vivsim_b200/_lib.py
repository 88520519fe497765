"""ctypes binding of libvivsim_b200.so (the C ABI declared in include/vivsim_b200.h).

There is no fallback: if the library is missing or the inputs are not CUDA tensors the call
raises.  torch is used only for device memory and streams."""

import ctypes as C
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# VIVSIM_B200_LIB selects a tuning variant of the same library (scripts/tune_variants.py); there is no other backend
LIB_PATH = os.environ.get("VIVSIM_B200_LIB") or os.path.join(HERE, "libvivsim_b200.so")

OK = 0
COLL = {"bgk": 0, "mrt": 1, "kbc": 2, "reg": 3}
FORCE = {None: 0, "none": 0, "edm": 1, "guo": 2}
BC = {"nee": 0, "nebb": 1, "equilibrium": 2, "bounce_back": 3, "specular_reflection": 4, "mask": 5}
WRAP = {"": 0, "velocity": 1, "pressure": 2, "force_corrected": 3}
LOC = {"left": 0, "right": 1, "bottom": 2, "top": 3, "back": 4, "front": 5}
DELTA = {"peskin3": 0, "peskin4": 1, "cosine4": 2, "hat2": 3}
CHAIN = {"auto": 0, "barrier": 1, "cluster": 2, "launches": 3, "cta": 4}
DIAG = {"velocity_magnitude": 0, "velocity_gradient": 1, "vorticity": 2, "vorticity_magnitude": 3, "divergence": 4,
        "strain_rate": 5, "strain_rate_magnitude": 6, "kinetic_energy": 7, "pressure": 8, "enstrophy": 9,
        "q_criterion": 10}
MG_DIR = {"left": 0, "right": 1, "up": 2, "down": 3}

EXPORTS = [
    "vsb_abi_version", "vsb_last_error", "vsb_streaming", "vsb_macroscopic", "vsb_equilibrium", "vsb_collision",
    "vsb_guo_forcing_term", "vsb_forcing", "vsb_post_op", "vsb_boundary_characteristic", "vsb_ib_delta",
    "vsb_ib_stencil", "vsb_ib_interpolate", "vsb_ib_spread", "vsb_ib_mdf", "vsb_step", "vsb_ib_window_moments",
    "vsb_body_newmark", "vsb_edge_fused", "vsb_edge_fused_supported", "vsb_ib_fused", "vsb_ib_fused_supported",
    "vsb_halo_push", "vsb_body_newmark_host", "vsb_halo_send", "vsb_halo_wait", "vsb_step_host_ode",
    "vsb_post_field", "vsb_post_mean", "vsb_mg_fine_to_coarse", "vsb_mg_coarse_to_fine", "vsb_run_host_ode",
    "vsb_run_host_ode_multi", "vsb_ibshard_chain", "vsb_ibshard_barrier", "vsb_enqueue_host_ode", "vsb_sync_status", "vsb_ib_window_moments_cells", "vsb_ib_spread_ordered", "vsb_ib_spread_ordered_workspace",
]


class VsbGrid(C.Structure):
    _fields_ = [("dim", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int)]


class VsbWallValue(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("value", C.c_float)]


class VsbPostOp(C.Structure):
    _fields_ = [("kind", C.c_int), ("wrap", C.c_int), ("loc", C.c_int), ("rho", VsbWallValue),
                ("u", VsbWallValue * 3), ("g", VsbWallValue * 3), ("mask", C.c_void_p)]


class VsbBodyState(C.Structure):
    _fields_ = [("d", C.c_float * 3), ("v", C.c_float * 3), ("a", C.c_float * 3), ("h", C.c_float * 3),
                ("force_sum", C.c_float * 3), ("origin2", (C.c_int * 3) * 2), ("ticket", C.c_int),
                ("step", C.c_int)]


class VsbHostMail(C.Structure):
    _fields_ = [("force", C.c_float * 3), ("seq", C.c_int), ("next", C.c_int)]


BODY_BYTES = C.sizeof(VsbBodyState)   # 15 fp32 + 8 int32 = 92 bytes


class VsbBodyParams(C.Structure):
    _fields_ = [("n_dof", C.c_int), ("follow", C.c_int), ("origin0", C.c_float * 3), ("grid_size", C.c_int * 3),
                ("win_size", C.c_int * 3), ("m", C.c_double), ("k", C.c_double), ("c", C.c_double),
                ("added_mass", C.c_double), ("history", C.c_void_p), ("history_capacity", C.c_int),
                ("rotation", C.c_int), ("center", C.c_float * 2), ("matrix_form", C.c_int),
                ("mat_m", C.c_double * 9), ("mat_k", C.c_double * 9), ("mat_c", C.c_double * 9),
                ("added_mass_v", C.c_double * 3)]


class VsbMdfArgs(C.Structure):
    _fields_ = [("dim", C.c_int), ("delta_kind", C.c_int), ("n_iter", C.c_int), ("parity", C.c_int),
                ("n_markers", C.c_int64), ("win_origin0", C.c_int * 3), ("win_size", C.c_int * 3),
                ("markers0", C.c_void_p), ("u_target", C.c_void_p),
                ("ds_ptr", C.c_void_p), ("ds_value", C.c_float), ("u_win", C.c_void_p), ("g_win", C.c_void_p),
                ("g_win_next", C.c_void_p),
                ("scratch", C.c_void_p), ("scratch_next", C.c_void_p), ("marker_u", C.c_void_p),
                ("marker_force", C.c_void_p), ("body", C.c_void_p), ("barrier", C.c_void_p),
                ("host_mail", C.c_void_p), ("mail_seq", C.c_int), ("chunk_offsets", C.c_void_p), ("n_chunks", C.c_int),
                ("rotation", C.c_int), ("center", C.c_float * 2), ("chain_mode", C.c_int),
                ("nbr_list", C.c_void_p), ("nbr_stride", C.c_int), ("reach_cells", C.c_void_p),
                ("n_reach_cells", C.c_int64)]


class VsbStepArgs(C.Structure):
    _fields_ = [("grid", VsbGrid), ("collision", C.c_int), ("forcing", C.c_int), ("omega", C.c_double),
                ("mrt_op_host", C.c_void_p), ("mrt_fop_host", C.c_void_p), ("do_stream", C.c_int),
                ("do_collide", C.c_int), ("row_begin", C.c_int), ("row_end", C.c_int), ("f_in", C.c_void_p),
                ("f_out", C.c_void_p), ("g_uniform", C.c_float * 3), ("g_win", C.c_void_p),
                ("win_origin", C.c_int * 3), ("win_size", C.c_int * 3), ("body", C.c_void_p), ("parity", C.c_int),
                ("n_post", C.c_int), ("post", C.POINTER(VsbPostOp)), ("vec", C.c_int), ("band", C.c_int),
                ("edges", C.c_int), ("sub_begin", C.c_int), ("sub_end", C.c_int), ("edge_rows_only", C.c_int),
                ("win_shift", C.c_int * 3), ("halo", C.c_void_p), ("halo_mode", C.c_int), ("early_launch", C.c_int)]


MAX_RANKS = 8


class VsbIbShard(C.Structure):
    _fields_ = [("n_ranks", C.c_int), ("rank", C.c_int), ("fields", C.c_void_p * MAX_RANKS),
                ("flags", C.c_void_p * MAX_RANKS), ("sums", C.c_void_p * MAX_RANKS),
                ("need_lo", (C.c_int * 3) * MAX_RANKS), ("need_hi", (C.c_int * 3) * MAX_RANKS),
                ("x_lo", C.c_int * MAX_RANKS), ("x_hi", C.c_int * MAX_RANKS),
                ("marker_begin", C.c_int64), ("marker_end", C.c_int64), ("chunk_begin", C.c_int), ("chunk_end", C.c_int),
                ("counter", C.c_void_p), ("staging", C.c_void_p * MAX_RANKS), ("force_field", C.c_void_p),
                ("cells", C.c_void_p), ("n_cells", C.c_int64), ("ev_window_done", C.c_void_p), ("trace", C.c_void_p)]


class VsbHostPlan(C.Structure):
    _fields_ = [("main", C.c_void_p), ("ib", C.c_void_p), ("edge", C.c_void_p), ("ev_fork", C.c_void_p),
                ("ev_ib", C.c_void_p), ("ev_edge", C.c_void_p)]


class VsbHaloArgs(C.Structure):
    _fields_ = [("grid", VsbGrid), ("state", C.c_void_p), ("left_state", C.c_void_p), ("right_state", C.c_void_p),
                ("my_flags", C.c_void_p), ("left_flags", C.c_void_p), ("right_flags", C.c_void_p), ("counter", C.c_void_p)]


_lib = None


class VsbError(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VsbError(f"{LIB_PATH} is missing: run `python -m vivsim_b200._build` (there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.vsb_last_error.restype = C.c_char_p
        for name in EXPORTS:
            getattr(_lib, name)  # AttributeError if the ABI and this binding diverge
    return _lib


def check(rc):
    if rc != OK:
        raise VsbError(lib().vsb_last_error().decode())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dev(t, dtype=torch.float32, name="tensor"):
    """Validate a device operand and return it contiguous."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor on a CUDA device, got {type(t).__name__}")
    if not t.is_cuda:
        raise VsbError(f"{name}: tensor is on {t.device}; vivsim_b200 has no CPU path")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def host_matrix(m, q):
    """Q x Q operator as a contiguous fp32 HOST array (accepts torch / numpy / nested lists)."""
    if isinstance(m, torch.Tensor):
        m = m.detach().cpu().numpy()
    a = np.ascontiguousarray(np.asarray(m, dtype=np.float32))
    if a.shape != (q, q):
        raise ValueError(f"expected a ({q}, {q}) matrix, got {a.shape}")
    return a


def grid_of(shape):
    if len(shape) == 2:
        return VsbGrid(2, int(shape[0]), int(shape[1]), 1)
    if len(shape) == 3:
        return VsbGrid(3, int(shape[0]), int(shape[1]), int(shape[2]))
    raise ValueError(f"spatial shape must have 2 or 3 dims, got {tuple(shape)}")


def wall_value(v, face_shape, keep, name):
    """Scalar or face-shaped array -> VsbWallValue (reference broadcast_wall_values)."""
    if isinstance(v, torch.Tensor) and v.ndim > 0:
        t = dev(v, name=name)
        if tuple(t.shape) != tuple(face_shape):
            raise ValueError(f"{name}: expected face shape {tuple(face_shape)}, got {tuple(t.shape)}")
        keep.append(t)
        return VsbWallValue(t.data_ptr(), 0.0)
    if isinstance(v, np.ndarray) and v.ndim > 0:
        raise TypeError(f"{name}: pass wall arrays as CUDA tensors")
    return VsbWallValue(None, float(v))
