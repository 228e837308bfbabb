"""Process-group bring-up with the call surface of mmidas/_dist_utils.py (init_dist_env :12, destroy_dist_env :20,
set_print :54, find_addr :58, find_port :62).  One process per GPU, NCCL over NVLink 5 / NVSwitch;
``backend="gloo"`` is accepted for the CPU tests of the host-side logic.

Differences from the reference, all deliberate: the rendezvous address / port fall back to the environment and then to
127.0.0.1 / a free port (the reference requires both); the A100-only switches of its ``_init_gpu_flags``
(``NCCL_P2P_LEVEL``, global TF32) are not set — on B200 every peer is one NVSwitch hop away and precision is chosen per
GEMM inside the kernels; ``destroy_dist_env`` restores ``print`` and tolerates an uninitialised group."""
from __future__ import annotations

import builtins
import contextlib
import datetime
import functools
import os
import signal
import socket

import torch
import torch.distributed as dist

_PLAIN_PRINT = builtins.print
_PG_TIMEOUT = datetime.timedelta(minutes=5)


def find_addr() -> str:
    """First address the host name resolves to (loopback when it does not resolve, e.g. in containers)."""
    with contextlib.suppress(OSError):
        addresses = socket.gethostbyname_ex(socket.gethostname())[2]
        if addresses:
            return addresses[0]
    return "127.0.0.1"


def find_port(addr) -> int:
    """A TCP port that is free on ``addr`` right now (the kernel picks it)."""
    probe = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    try:
        probe.bind((addr, 0))
        return probe.getsockname()[1]
    finally:
        probe.close()


def set_print(rank) -> None:
    """Prefix everything this process prints with its rank."""
    builtins.print = functools.partial(_PLAIN_PRINT, f"[R{rank}]")


def _teardown(*_signal_args) -> None:
    if dist.is_available() and dist.is_initialized():
        from .nn_model import release_all_graphs
        release_all_graphs()          # graphs that captured NCCL collectives pin the communicator: destroy would block
        dist.destroy_process_group()


def init_dist_env(rank, world_size, addr=None, port=None, backend="nccl") -> None:
    rendezvous = {"MASTER_ADDR": addr, "MASTER_PORT": port}
    for key, value in rendezvous.items():
        if value is not None:
            os.environ[key] = str(value)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if "MASTER_PORT" not in os.environ:
        os.environ["MASTER_PORT"] = str(find_port(os.environ["MASTER_ADDR"]))
    if backend == "nccl":
        for flag in ("TORCH_SHOW_CPP_STACKTRACES", "TORCH_NCCL_ASYNC_ERROR_HANDLING"):
            os.environ[flag] = "1"
        torch.cuda.set_device(rank % max(torch.cuda.device_count(), 1))
    dist.init_process_group(backend, rank=rank, world_size=world_size, timeout=_PG_TIMEOUT)
    with contextlib.suppress(ValueError):          # signal handlers can only be installed from the main thread
        signal.signal(signal.SIGINT, _teardown)
    set_print(rank)


def destroy_dist_env() -> None:
    _teardown()
    builtins.print = _PLAIN_PRINT
