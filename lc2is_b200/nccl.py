"""A minimal NCCL communicator bound with ctypes (the libnccl.so.2 that torch already loaded), for the data-parallel
exchange steps of the head (north_star stage 4).

Why not ``torch.distributed`` for these: ProcessGroupNCCL runs every collective on its own internal stream and watches
it from a host thread, which made the collectives impossible to capture into the step's CUDA graph on this stack (the
watchdog hung).  ``ncclAllReduce`` itself is an ordinary stream-ordered launch: issued on a stream the caller picks it
can be forked / joined with events, overlapped with the head kernels and captured into a CUDA graph like any kernel.
``torch.distributed`` stays the bootstrap (it carries the 128-byte unique id to the ranks) and the fallback.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.distributed as dist

_DTYPES = {torch.int8: 0, torch.uint8: 1, torch.int32: 2, torch.int64: 4, torch.float16: 6, torch.float32: 7,
           torch.float64: 8, torch.bfloat16: 9}
_SUM = 0


class _UniqueId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_byte * 128)]


_lib = None


def _nccl():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL("libnccl.so.2")
        lib.ncclGetErrorString.restype = ctypes.c_char_p
        lib.ncclGetErrorString.argtypes = [ctypes.c_int]
        lib.ncclGetUniqueId.argtypes = [ctypes.POINTER(_UniqueId)]
        lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _UniqueId, ctypes.c_int]
        lib.ncclAllReduce.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p]
        lib.ncclCommDestroy.argtypes = [ctypes.c_void_p]
        lib.ncclGroupStart.argtypes = []
        lib.ncclGroupEnd.argtypes = []
        _lib = lib
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed: {_nccl().ncclGetErrorString(rc).decode()} ({rc})")


class Comm:
    """One communicator over all ranks of the default torch process group (which must exist: it is the bootstrap)."""

    def __init__(self, device: torch.device) -> None:
        lib = _nccl()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        uid = _UniqueId()
        if self.rank == 0:
            _check(lib.ncclGetUniqueId(ctypes.byref(uid)), "ncclGetUniqueId")
        box = [bytes(bytearray(uid.internal))] if self.rank == 0 else [None]
        dist.broadcast_object_list(box, src=0)
        ctypes.memmove(ctypes.byref(uid), box[0], 128)
        self.device = device
        self.comm = ctypes.c_void_p()
        with torch.cuda.device(device):
            _check(lib.ncclCommInitRank(ctypes.byref(self.comm), self.world, uid, self.rank), "ncclCommInitRank")

    def all_reduce_(self, t: torch.Tensor, stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        """In-place SUM all-reduce of a contiguous CUDA tensor, enqueued on `stream` (default: the current stream)."""
        assert t.is_cuda and t.is_contiguous()
        s = (stream or torch.cuda.current_stream()).cuda_stream
        _check(_nccl().ncclAllReduce(t.data_ptr(), t.data_ptr(), t.numel(), _DTYPES[t.dtype], _SUM, self.comm, s),
               "ncclAllReduce")
        return t

    def all_reduce_many_(self, tensors, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Several all-reduces as one NCCL group (one launch)."""
        lib = _nccl()
        _check(lib.ncclGroupStart(), "ncclGroupStart")
        try:
            for t in tensors:
                self.all_reduce_(t, stream)
        finally:
            _check(lib.ncclGroupEnd(), "ncclGroupEnd")

    def close(self) -> None:
        if self.comm:
            _nccl().ncclCommDestroy(self.comm)
            self.comm = ctypes.c_void_p()


_default: Optional[Comm] = None


def default_comm(device: torch.device) -> Optional[Comm]:
    """The process-wide communicator (created on first use, collectively), or None without an NCCL process group."""
    global _default
    if _default is None:
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
                and dist.get_backend() == "nccl"):
            return None
        _default = Comm(device)
    return _default


def close_default() -> None:
    global _default
    if _default is not None:
        _default.close()
        _default = None
