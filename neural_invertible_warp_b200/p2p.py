"""Peer-memory gradient all-reduce (csrc/p2p.cu) -- the exchange step of the data-parallel training step.

``P2PChannel`` owns one exchange block per rank (CUDA IPC: allocated by ``niw_p2p_alloc``, the 64-byte handles travel
through ``torch.distributed.all_gather_object``, every rank maps its peers' blocks with ``niw_p2p_open``) and sums a flat
fp32 tensor in place over the ranks of ONE node with two kernel launches of our own (``niw_allreduce_p2p``): no library
collective on the data path.  ``torch.distributed`` is plumbing only (rendezvous, the handle exchange).

One channel serves one stream of calls: concurrent reductions (the NeRF segment on the side stream, the pose / warp
segment on the main stream) use one channel each.
"""
import ctypes as _c
import os
import socket

import torch
import torch.distributed as dist

from . import _lib


def available(group=None):
    """All ranks of ``group`` are CUDA ranks of one host, at most 8, and NIW_P2P_ALLREDUCE is not 0."""
    if os.environ.get("NIW_P2P_ALLREDUCE", "1") == "0":
        return False
    if not (dist.is_available() and dist.is_initialized() and torch.cuda.is_available()):
        return False
    world = dist.get_world_size(group)
    if world < 2 or world > 8 or dist.get_backend(group) != "nccl":
        return False
    hosts = [None] * world
    dist.all_gather_object(hosts, (socket.gethostname(), torch.cuda.current_device()), group=group)
    same_host = len({h for h, _ in hosts}) == 1
    distinct = len({d for _, d in hosts}) == world
    ok = same_host and distinct
    if ok:
        me = torch.cuda.current_device()
        ok = all(d == me or torch.cuda.can_device_access_peer(me, d) for _, d in hosts)
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok), group=group)
    return all(flags)


class P2PChannel:
    """In-place sum of fp32 tensors of up to ``max_floats`` elements over the ranks of ``group`` (one node)."""

    def __init__(self, max_floats, group=None):
        """Collective: every rank of ``group`` constructs its channel at the same point.  Raises ``RuntimeError`` on ALL ranks
        when any rank could not allocate or map a block (the outcome of each stage is agreed on before the next), so a
        caller can fall back to NCCL consistently."""
        self.lib = _lib.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.half = (int(max_floats) + 3) // 4 * 4
        nbytes = 256 + 4 * self.half * 4        # flag block, two exchange halves, two result halves (two-shot)
        mine, handle, err = _c.c_void_p(), _c.create_string_buffer(64), None
        rc = self.lib.niw_p2p_alloc(nbytes, _c.byref(mine), handle)
        if rc:
            mine, err = None, "niw_p2p_alloc: %s" % self.lib.niw_error_string(rc).decode()
        elif os.environ.get("NIW_P2P_TEST_FAIL") == str(self.rank):      # tests: this rank pretends its allocation failed
            self.lib.niw_p2p_free(mine)
            mine, err = None, "simulated allocation failure (NIW_P2P_TEST_FAIL)"
        self.mine = mine
        self._opened = []
        stage = [None] * self.world
        dist.all_gather_object(stage, (err, handle.raw), group=group)
        if any(e is not None for e, _ in stage):
            self._release()
            raise RuntimeError("P2PChannel: " + "; ".join("rank %d: %s" % (r, e) for r, (e, _) in enumerate(stage) if e))
        self.blocks = (_c.c_void_p * self.world)()
        for r, (_, h) in enumerate(stage):
            if r == self.rank:
                self.blocks[r] = mine.value
                continue
            p = _c.c_void_p()
            rc = self.lib.niw_p2p_open(_c.create_string_buffer(h, 64), _c.byref(p))
            if rc:
                err = "niw_p2p_open(rank %d): %s" % (r, self.lib.niw_error_string(rc).decode())
                break
            self.blocks[r] = p.value
            self._opened.append(p)
        stage = [None] * self.world
        dist.all_gather_object(stage, err, group=group)      # (also: every rank has mapped every block before the first flag is raised)
        if any(e is not None for e in stage):
            self._release()
            raise RuntimeError("P2PChannel: " + "; ".join("rank %d: %s" % (r, e) for r, e in enumerate(stage) if e))

    def _release(self):
        for p in self._opened:
            self.lib.niw_p2p_close(p)
        self._opened = []
        if self.mine is not None:
            self.lib.niw_p2p_free(self.mine)
            self.mine = None

    def allreduce_(self, t):
        """Sum ``t`` (contiguous CUDA fp32, numel % 4 == 0, 16-byte aligned) over the ranks, in place, on the current stream."""
        if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous() or t.numel() > self.half:
            raise RuntimeError("P2PChannel.allreduce_: contiguous CUDA float32 tensor of at most %d elements" % self.half)
        _lib.check(self.lib.niw_allreduce_p2p(_c.c_void_p(t.data_ptr()), t.numel(), self.blocks, self.rank, self.world, self.half,
                                              _c.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return t

    def error(self):
        """0, or 1 + the rank whose flag did not arrive within the kernel's time-out (synchronises)."""
        e = _c.c_uint(0)
        _lib.check(self.lib.niw_p2p_error(self.mine, _c.byref(e)))
        return int(e.value)

    def close(self):
        if self.mine is None:
            return
        torch.cuda.synchronize()
        dist.barrier(group=self.group)   # nobody unmaps while a peer may still read
        for p in self._opened:
            self.lib.niw_p2p_close(p)
        self._opened = []
        dist.barrier(group=self.group)
        self._release()
