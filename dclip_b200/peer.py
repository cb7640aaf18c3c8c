"""Symmetric peer-memory buffers for the data-parallel objective: push-based all-gather over NVLink.

One process per GPU (torch.distributed supplies the rendezvous only): every rank allocates one buffer through
the C ABI (``dclip_p2p_alloc``: cudaMalloc + cudaIpc handle), the 64-byte handles are all-gathered as Python
objects, and each rank maps its peers' buffers.  The data path afterwards is CUDA kernels only
(``dclip_p2p_push`` / ``dclip_p2p_wait``): no NCCL collective, no host synchronisation, graph-capturable.
"""
import ctypes

import torch

from . import _lib
from ._lib import check, stream_ptr

KIND_OPERANDS, KIND_STATS, KIND_DONE = 0, 1, 2


class PeerBuffer:
    """``nbytes`` of symmetric device memory (+ the flag block) shared by the ranks of ``group``."""

    def __init__(self, nbytes, device, group=None):
        import torch.distributed as dist
        self.lib = _lib.load_library()
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.flags_offset = (int(nbytes) + 255) // 256 * 256
        self.nbytes = self.flags_offset + self.lib.dclip_p2p_flags_bytes()
        with torch.cuda.device(device):
            base = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(64)
            check(self.lib.dclip_p2p_alloc(self.nbytes, ctypes.byref(base), handle), "p2p_alloc")
            self.base = base.value
            check(self.lib.dclip_p2p_init_flags(ctypes.c_void_p(self.base), self.flags_offset, KIND_DONE), "p2p_init_flags")
            handles = [None] * self.world
            dist.all_gather_object(handles, (self.rank, handle.raw), group=group)
            self.peers = [None] * self.world
            for r, raw in handles:
                if r == self.rank:
                    self.peers[r] = self.base
                else:
                    p = ctypes.c_void_p()
                    check(self.lib.dclip_p2p_open(raw, ctypes.byref(p)), "p2p_open")
                    self.peers[r] = p.value
            dist.barrier(group=group)            # every rank has initialised its flags and mapped its peers
        self._peer_arr = (ctypes.c_void_p * self.world)(*self.peers)
        self._closed = False

    def ptr(self, offset=0):
        return ctypes.c_void_p(self.base + offset)

    def push(self, ranges, wait_kind, signal_kind, wait_after=False):
        """ranges: [(offset, nbytes), ...] of the local buffer, written to the same offsets of every peer.
        wait_after: the same launch also waits for every peer's push of `signal_kind` (= an all-gather)."""
        n = len(ranges)
        offs = (ctypes.c_size_t * max(n, 1))(*[o for o, _ in ranges])
        lens = (ctypes.c_size_t * max(n, 1))(*[b for _, b in ranges])
        check(self.lib.dclip_p2p_push(self._peer_arr, self.world, self.rank, offs, lens, n, self.flags_offset, wait_kind, signal_kind,
                                      1 if wait_after else 0, stream_ptr(self.device)), "p2p_push")

    def wait(self, kind):
        check(self.lib.dclip_p2p_wait(ctypes.c_void_p(self.base), self.flags_offset, self.world, self.rank, kind, stream_ptr(self.device)),
              "p2p_wait")

    def close(self):
        if self._closed:
            return
        self._closed = True
        torch.cuda.synchronize(self.device)
        for r, p in enumerate(self.peers):
            if r != self.rank and p:
                self.lib.dclip_p2p_close(ctypes.c_void_p(p))
        import torch.distributed as dist
        if dist.is_initialized():
            dist.barrier(group=self.group)       # nobody frees while a peer still has the mapping open
        self.lib.dclip_p2p_free(ctypes.c_void_p(self.base))
