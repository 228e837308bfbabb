"""Process-group bring-up, mirror of mmidas/_dist_utils.py (init_dist_env :12, destroy_dist_env :20,
set_print :54, find_addr :58, find_port :62).  One process per GPU, NCCL over NVLink 5 / NVSwitch;
``backend="gloo"`` is accepted for the CPU tests of the host-side logic."""
from __future__ import annotations

import builtins
import os
import signal
import socket
from datetime import timedelta
from functools import partial

import torch
import torch.distributed as dist

_ORIG_PRINT = builtins.print


def init_dist_env(rank, world_size, addr=None, port=None, backend="nccl"):
    """Same call as the reference; addr/port default to MASTER_ADDR/MASTER_PORT or 127.0.0.1/free port."""
    _init_dist_flags(addr, port)
    if backend == "nccl":
        _init_gpu_flags()
        torch.cuda.set_device(rank % max(torch.cuda.device_count(), 1))
    init_pg(rank, world_size, backend)
    set_print(rank)


def destroy_dist_env():
    destroy_pg()
    builtins.print = _ORIG_PRINT


def _init_dist_flags(addr, port):
    if addr is not None:
        os.environ["MASTER_ADDR"] = str(addr)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if port is not None:
        os.environ["MASTER_PORT"] = str(port)
    os.environ.setdefault("MASTER_PORT", str(find_port(os.environ["MASTER_ADDR"])))


def _init_gpu_flags():
    # reference: _dist_utils.py:30-40 (the A100-only TF32 / NCCL_P2P_LEVEL switches do not apply:
    # on B200 every peer is one NVSwitch hop away and precision is chosen per GEMM in the kernels)
    os.environ["TORCH_SHOW_CPP_STACKTRACES"] = str(1)
    os.environ["TORCH_NCCL_ASYNC_ERROR_HANDLING"] = str(1)


def init_pg(rank, world_size, backend="nccl"):
    dist.init_process_group(backend, rank=rank, world_size=world_size, timeout=timedelta(seconds=300))
    try:
        signal.signal(signal.SIGINT, lambda _, __: destroy_pg())
    except ValueError:
        pass  # not in the main thread


def destroy_pg():
    if dist.is_initialized():
        dist.destroy_process_group()


def set_print(rank):
    builtins.print = partial(_ORIG_PRINT, f"[R{rank}]")


def find_addr():
    try:
        return socket.gethostbyname_ex(socket.gethostname())[2][0]
    except OSError:
        return "127.0.0.1"


def find_port(addr):
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind((addr, 0))
        s.listen(1)
        port = s.getsockname()[1]
    return port
