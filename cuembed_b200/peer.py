"""Peer (NVLink / NVSwitch) exchange buffers for the row-sharded mode.

A `PeerBuffer` is one symmetric allocation: every rank allocates the same
number of bytes with cuembed_peer_alloc, exports a CUDA-IPC handle, the handles
travel through `torch.distributed.all_gather_object` (plumbing only) and every
rank maps every other rank's buffer (cuembed_peer_open).  `ptrs[o]` is rank o's
buffer as seen from this process, which is what the fused kernels of
csrc/sharded_p2p.cu store to / the copy engines write to.

`LocalPeerGroup` builds the same pointer tables for several *virtual* ranks
living in ONE process on one GPU; the kernels cannot tell the difference, so
the whole exchange protocol can be checked on a single-GPU box.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib

MAX_WORLD = 16
CHANNELS = 4
FLAG_BYTES = 512
HANDLE_BYTES = 64

_TYPESTR = {torch.float32: "<f4", torch.float16: "<f2", torch.int32: "<i4",
            torch.int64: "<i8", torch.uint8: "|u1", torch.int16: "<i2"}


def _check(rc: int) -> None:
    if rc != 0:
        msg = _lib.load().cuembed_error_string(rc).decode()
        raise RuntimeError(f"cuembed_b200 peer memory error {rc}: {msg}")


class _CudaArray:
    """__cuda_array_interface__ carrier so torch can view raw device memory."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": typestr, "data": (ptr, False),
            "version": 3, "strides": None}


def view(ptr: int, shape: Sequence[int], dtype: torch.dtype,
         device: torch.device) -> torch.Tensor:
    """A torch tensor over device memory owned by the library (no copy)."""
    if dtype == torch.bfloat16:
        t = torch.as_tensor(_CudaArray(ptr, shape, "<i2"), device=device)
        return t.view(torch.bfloat16)
    return torch.as_tensor(_CudaArray(ptr, shape, _TYPESTR[dtype]), device=device)


def ptr_array(ptrs: Sequence[int], offset: int = 0):
    arr = (ctypes.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p + offset
    return arr


class PeerBuffer:
    """`nbytes` of device memory on every rank of `group`, mapped everywhere."""

    def __init__(self, nbytes: int, group: Optional[dist.ProcessGroup] = None,
                 device: Optional[torch.device] = None):
        lib = _lib.load()
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > MAX_WORLD:
            raise ValueError(f"at most {MAX_WORLD} ranks")
        self.device = device if device is not None else \
            torch.device("cuda", torch.cuda.current_device())
        self.nbytes = int(nbytes)
        with torch.cuda.device(self.device):
            p = ctypes.c_void_p()
            _check(lib.cuembed_peer_alloc(self.nbytes, ctypes.byref(p)))
            self.local = int(p.value)
            handle = ctypes.create_string_buffer(HANDLE_BYTES)
            if self.world > 1:
                _check(lib.cuembed_peer_export(self.local, handle))
            handles: List[Optional[bytes]] = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(handles, bytes(handle.raw), group=group)
            self.ptrs: List[int] = []
            self._opened: List[int] = []
            for o in range(self.world):
                if o == self.rank:
                    self.ptrs.append(self.local)
                    continue
                q = ctypes.c_void_p()
                _check(lib.cuembed_peer_open(handles[o], ctypes.byref(q)))
                self.ptrs.append(int(q.value))
                self._opened.append(int(q.value))

    def tensor(self, offset: int, shape, dtype: torch.dtype) -> torch.Tensor:
        return view(self.local + offset, shape, dtype, self.device)

    def close(self) -> None:
        lib = _lib.load()
        if self.local is None:
            return
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)  # nobody still writes to a peer
        for q in self._opened:
            lib.cuembed_peer_close(q)
        self._opened = []
        if self.world > 1:
            dist.barrier(group=self.group)  # everyone unmapped before the free
        lib.cuembed_peer_free(self.local)
        self.local = None


class LocalPeerGroup:
    """`world` virtual ranks in this process: buffer tables without IPC."""

    def __init__(self, world: int, device: torch.device):
        self.world = world
        self.device = device
        self._allocs: List[List[int]] = []

    def alloc(self, nbytes: int) -> List["LocalPeerView"]:
        lib = _lib.load()
        ptrs = []
        with torch.cuda.device(self.device):
            for _ in range(self.world):
                p = ctypes.c_void_p()
                _check(lib.cuembed_peer_alloc(int(nbytes), ctypes.byref(p)))
                ptrs.append(int(p.value))
        self._allocs.append(ptrs)
        return [LocalPeerView(r, self.world, ptrs, int(nbytes), self.device)
                for r in range(self.world)]

    def close(self) -> None:
        torch.cuda.synchronize(self.device)
        lib = _lib.load()
        for ptrs in self._allocs:
            for p in ptrs:
                lib.cuembed_peer_free(p)
        self._allocs = []


class LocalPeerView:
    """What a PeerBuffer looks like to virtual rank `rank`."""

    def __init__(self, rank, world, ptrs, nbytes, device):
        self.rank, self.world, self.ptrs = rank, world, list(ptrs)
        self.local = ptrs[rank]
        self.nbytes = nbytes
        self.device = device

    def tensor(self, offset: int, shape, dtype: torch.dtype) -> torch.Tensor:
        return view(self.local + offset, shape, dtype, self.device)

    def close(self) -> None:
        pass
