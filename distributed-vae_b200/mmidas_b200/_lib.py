"""ctypes binding of libmixvae_b200.so (C ABI declared in include/mixvae_b200.h).

The library is hand-written sm_100a CUDA; there is no CPU or PyTorch fallback: if the shared
object is missing or does not load, importing the compute path raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmixvae_b200.so")

N_PARAM_TENSORS = 28
MAX_ARMS = 16

PARAM_ORDER = ("fc1", "fc2", "fc3", "fc4", "fc5", "fcc", "fc_mu", "fc_sigma",
               "fc6", "fc7", "fc8", "fc9", "fc10", "fc11")
BN_ORDER = ("batch_l1", "batch_l2", "batch_l3", "batch_l4", "batch_l5", "batch_s")


class Dims(C.Structure):
    _fields_ = [("n_arm", C.c_int32), ("batch", C.c_int32), ("input_dim", C.c_int32), ("fc_dim", C.c_int32),
                ("lowD_dim", C.c_int32), ("n_categories", C.c_int32), ("state_dim", C.c_int32),
                ("n_arm_total", C.c_int32), ("arm_offset", C.c_int32)]


class HParams(C.Structure):
    _fields_ = [("tau", C.c_float), ("temp", C.c_float), ("beta", C.c_float), ("lam", C.c_float),
                ("eps", C.c_float), ("momentum", C.c_float), ("x_drop", C.c_float), ("s_drop", C.c_float),
                ("hard", C.c_int32), ("precision", C.c_int32)]


class Layout(C.Structure):
    _fields_ = [("offset", C.c_int64 * N_PARAM_TENSORS), ("numel", C.c_int64 * N_PARAM_TENSORS),
                ("arm_stride", C.c_int64), ("bn_stride", C.c_int64), ("bn_offset", C.c_int64 * 6),
                ("work_floats", C.c_int64)]


class State(C.Structure):
    _fields_ = [("params", C.c_void_p), ("grads", C.c_void_p), ("adam_m", C.c_void_p), ("adam_v", C.c_void_p),
                ("bn_running", C.c_void_p), ("bn_batches", C.c_void_p), ("work", C.c_void_p)]


class Inputs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("x_arm_stride", C.c_int64), ("x_row_stride", C.c_int64),
                ("U", C.c_void_p), ("E", C.c_void_p), ("keep_x", C.c_void_p), ("keep_s", C.c_void_p), ("cat_mask", C.c_void_p),
                ("seed", C.c_uint64), ("step", C.c_uint64), ("counters", C.c_void_p), ("training", C.c_int32)]


class Outputs(C.Structure):
    _fields_ = [("x_low", C.c_void_p), ("c_prob", C.c_void_p), ("qc", C.c_void_p), ("c_smp", C.c_void_p),
                ("s_mean", C.c_void_p), ("s_logvar", C.c_void_p), ("s_smp", C.c_void_p), ("x_rec", C.c_void_p)]


PRECISIONS = {"tf32x3_fc1": 0, "tf32x3": 1, "tf32": 2, "fp32_simt": 3}

# every symbol include/mixvae_b200.h declares
EXPORTS = ("mvae_last_error", "mvae_abi_version", "mvae_compute_layout", "mvae_forward", "mvae_loss",
           "mvae_backward", "mvae_adam", "mvae_train_step", "mvae_grad_step", "mvae_argmax", "mvae_dropout_mask", "mvae_launch_count",
           "mvae_timing_enable", "mvae_timing_read", "mvae_debug_tc_gemm", "mvae_confmat", "mvae_fold_affine",
           "mvae_linear_act", "mvae_fma_rows", "mvae_unpack_rows", "mvae_adam_peer", "mvae_pdl_enable")

_lib = None


def load():
    """Load the shared library once; raise (never fall back) if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C distributed-vae_b200/csrc`). There is no CPU/PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.mvae_last_error.restype = C.c_char_p
    lib.mvae_abi_version.restype = C.c_int
    lib.mvae_launch_count.restype = C.c_int64
    P = C.POINTER
    lib.mvae_compute_layout.argtypes = [P(Dims), P(Layout)]
    lib.mvae_forward.argtypes = [P(Dims), P(HParams), P(State), P(Inputs), P(Outputs), C.c_void_p]
    lib.mvae_loss.argtypes = [P(Dims), P(HParams), P(State), P(Inputs), P(Outputs), C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_int, C.c_void_p]
    lib.mvae_backward.argtypes = [P(Dims), P(HParams), P(State), P(Inputs), P(Outputs), C.c_void_p, C.c_void_p]
    lib.mvae_adam.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float,
                              C.c_float, C.c_float, C.c_float, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]
    lib.mvae_train_step.argtypes = [P(Dims), P(HParams), P(State), P(Inputs), P(Outputs), C.c_void_p, C.c_float,
                                    C.c_float, C.c_float, C.c_float, C.c_int64, C.c_void_p]
    lib.mvae_grad_step.argtypes = [P(Dims), P(HParams), P(State), P(Inputs), P(Outputs), C.c_void_p, C.c_void_p]
    lib.mvae_argmax.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    lib.mvae_dropout_mask.argtypes = [P(Dims), P(HParams), P(Inputs), C.c_void_p, C.c_void_p]
    lib.mvae_confmat.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.mvae_confmat.restype = C.c_int
    lib.mvae_fold_affine.argtypes = [C.c_void_p] * 5 + [C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mvae_fold_affine.restype = C.c_int
    lib.mvae_linear_act.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                                    C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.mvae_linear_act.restype = C.c_int
    lib.mvae_fma_rows.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                  C.c_int64, C.c_int32, C.c_float, C.c_void_p]
    lib.mvae_fma_rows.restype = C.c_int
    lib.mvae_adam_peer.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_float,
                                   C.c_float, C.c_float, C.c_float, C.c_int64, C.c_void_p, C.c_void_p]
    lib.mvae_adam_peer.restype = C.c_int
    lib.mvae_unpack_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
    lib.mvae_unpack_rows.restype = C.c_int
    lib.mvae_debug_tc_gemm.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    lib.mvae_debug_tc_gemm.restype = C.c_int
    lib.mvae_timing_enable.argtypes = [C.c_int]
    lib.mvae_timing_enable.restype = C.c_int
    lib.mvae_pdl_enable.argtypes = [C.c_int]
    lib.mvae_pdl_enable.restype = C.c_int
    lib.mvae_timing_read.argtypes = [P(C.c_float), P(C.c_int32), C.c_int32]
    lib.mvae_timing_read.restype = C.c_int
    for name in ("mvae_compute_layout", "mvae_forward", "mvae_loss", "mvae_backward", "mvae_adam",
                 "mvae_train_step", "mvae_grad_step", "mvae_argmax", "mvae_dropout_mask"):
        getattr(lib, name).restype = C.c_int
    if lib.mvae_abi_version() != 2:
        raise RuntimeError("libmixvae_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().mvae_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def compute_layout(dims: Dims) -> Layout:
    lay = Layout()
    check(load().mvae_compute_layout(C.byref(dims), C.byref(lay)), "mvae_compute_layout")
    return lay


TIMING_GROUPS = ("fc1_fwd", "fc11_loss_grad", "fc1_wgrad", "narrow_fwd", "narrow_bwd", "coupling", "narrow_wgrad", "adam")


def pdl_enable(on: bool) -> bool:
    """Programmatic dependent launch between the library's kernels (default on); returns the previous setting."""
    return bool(load().mvae_pdl_enable(int(on)))


def timing_enable(on: bool):
    check(load().mvae_timing_enable(int(on)), "mvae_timing_enable")


def timing_read():
    """{group: (total ms, spans)} of the spans recorded since timing_enable(True)."""
    n = len(TIMING_GROUPS)
    ms = (C.c_float * n)()
    cnt = (C.c_int32 * n)()
    check(load().mvae_timing_read(ms, cnt, n), "mvae_timing_read")
    return {g: (float(ms[i]), int(cnt[i])) for i, g in enumerate(TIMING_GROUPS)}


_replayed_launches = 0     # kernels run by CUDA-graph replays (the library's counter only sees launch CALLS)
_captured_launches = 0     # launch calls that were recorded into a graph instead of running


def note_capture(n: int):
    global _captured_launches
    _captured_launches += n


def note_replay(n: int):
    global _replayed_launches
    _replayed_launches += n


def launch_count() -> int:
    """Kernels of this library that ran on the device in this process: direct launches + graph replays."""
    return int(load().mvae_launch_count()) - _captured_launches + _replayed_launches
