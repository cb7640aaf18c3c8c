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


class _Peers(ctypes.Structure):
    """dclip_peers (include/dclip_b200.h): the ranks' buffers as mapped into this process."""
    _fields_ = [("bases", ctypes.POINTER(ctypes.c_void_p)), ("world", ctypes.c_int), ("rank", ctypes.c_int), ("flags_offset", ctypes.c_size_t)]


class PeerSetupError(RuntimeError):
    """The symmetric buffer could not be set up; raised on EVERY rank of the group together."""


class PeerBuffer:
    """``nbytes`` of symmetric device memory (+ the flag block) shared by the ranks of ``group``.

    Construction is collective and fails collectively: after each phase (allocate + export, map the peers) the
    ranks all-reduce a status, so either every rank returns a usable buffer or every rank raises PeerSetupError
    after releasing what it had allocated / mapped - no rank is left waiting in a collective the others skipped."""

    def __init__(self, nbytes, device, group=None):
        import torch.distributed as dist
        self.lib = _lib.load_library()
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.flags_offset = (int(nbytes) + 255) // 256 * 256
        self.nbytes = self.flags_offset + self.lib.dclip_p2p_flags_bytes()
        self.base = None
        self.peers = [None] * self.world
        self._closed = True

        def agree(ok, err):
            flag = torch.tensor([1 if ok else 0], device=device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag) == 0:
                self._release()
                raise PeerSetupError(err or "a peer rank could not set up its symmetric buffer")

        with torch.cuda.device(device):
            # phase 1: allocate, initialise the flags, export the handle
            err, handle = None, ctypes.create_string_buffer(64)
            try:
                base = ctypes.c_void_p()
                check(self.lib.dclip_p2p_alloc(self.nbytes, ctypes.byref(base), handle), "p2p_alloc")
                self.base = base.value
                check(self.lib.dclip_p2p_init_flags(ctypes.c_void_p(self.base), self.flags_offset, KIND_DONE), "p2p_init_flags")
            except RuntimeError as e:
                err = str(e)
            agree(err is None, err)
            # phase 2: exchange the handles (every rank takes part: phase 1 succeeded everywhere), map the peers
            handles = [None] * self.world
            dist.all_gather_object(handles, (self.rank, handle.raw), group=group)
            try:
                for r, raw in handles:
                    if r == self.rank:
                        self.peers[r] = self.base
                    else:
                        p = ctypes.c_void_p()
                        check(self.lib.dclip_p2p_open(raw, ctypes.byref(p)), "p2p_open")
                        self.peers[r] = p.value
            except RuntimeError as e:
                err = str(e)
            agree(err is None, err)          # doubles as the barrier: every rank has initialised its flags and mapped its peers
        self._peer_arr = (ctypes.c_void_p * self.world)(*self.peers)
        self._peers_struct = _Peers(ctypes.cast(self._peer_arr, ctypes.POINTER(ctypes.c_void_p)), self.world, self.rank, self.flags_offset)
        self._closed = False

    def _release(self):
        """Local clean-up of a partially constructed buffer (no collective)."""
        for r, p in enumerate(self.peers):
            if r != self.rank and p:
                self.lib.dclip_p2p_close(ctypes.c_void_p(p))
        self.peers = [None] * self.world
        if self.base:
            self.lib.dclip_p2p_free(ctypes.c_void_p(self.base))
            self.base = None

    def ptr(self, offset=0):
        return ctypes.c_void_p(self.base + offset)

    def peers_arg(self):
        """Pointer to the dclip_peers struct of this buffer (for the entry points that do their exchange in-kernel)."""
        return ctypes.cast(ctypes.pointer(self._peers_struct), ctypes.c_void_p)

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
