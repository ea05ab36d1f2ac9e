"""Peer-memory transport of the sharded loss: symmetric-memory arena + the dvlp_p2p_* kernels.

One ``PeerExchange`` per (process group, device).  The arena is allocated with torch's symmetric memory
(every rank maps every other rank's arena over NVLink / NVSwitch); the exchanges themselves are single kernel
launches of libdemovlp_b200.so (csrc/p2p.cu) that store into / read from the peers' arenas directly, so a step
costs three launches and three flag round trips instead of six latency-bound NCCL collectives.

Arena layout (identical on every rank; offsets 256-byte aligned):
    [control block][slot 0: gathered tensors][slot 1: gathered tensors][fp32 text-gradient partial sums]
Two gather slots let a forward run while the previous forward's gathered texts are still needed by its
backward (e.g. two graph instances in flight); a third forward before the first backward is refused.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib

N_SLOTS = 2


def _align(x, a=256):
    return (x + a - 1) // a * a


def available(t) -> bool:
    """Whether the peer-memory transport may be tried for tensor ``t`` (CUDA, not disabled by DVLP_P2P=0)."""
    if not t.is_cuda or os.environ.get("DVLP_P2P", "1") == "0":
        return False
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
    except Exception:
        return False
    return True


class PeerExchange:
    def __init__(self, group, device):
        self.group = group if group is not None else dist.group.WORLD
        self.device = torch.device(device)
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.key = None
        self.calls = 0
        self.slot_call = [-1] * N_SLOTS
        self.last_reduce_call = -1

    # ------------------------------------------------------------------ arena
    def ensure(self, shard_specs, reduce_shard_shape):
        """(Re)build the arena for per-rank shards ``shard_specs`` = [(shape, dtype), ...] (gathered along dim 0)
        and an fp32 reduce-scatter whose per-rank result has ``reduce_shard_shape``.  Collective."""
        key = (tuple((tuple(s), d) for s, d in shard_specs), tuple(reduce_shard_shape))
        if key == self.key:
            return
        import torch.distributed._symmetric_memory as symm
        lib = _lib.load()
        ctrl = int(lib.dvlp_p2p_control_bytes())
        self.shard_bytes, self.gshapes, self.gdtypes = [], [], []
        off = 0
        self.rel_off = []
        for shape, dtype in shard_specs:
            nb = int(torch.empty((), dtype=dtype).element_size())
            for d in shape:
                nb *= int(d)
            if nb % 16:
                raise RuntimeError(f"peer transport needs shard sizes that are multiples of 16 bytes, got {nb}")
            self.shard_bytes.append(nb)
            self.gshapes.append((self.world * shape[0],) + tuple(shape[1:]))
            self.gdtypes.append(dtype)
            self.rel_off.append(off)
            off = _align(off + nb * self.world)
        self.slot_bytes = off
        self.slot_off = [_align(ctrl) + i * self.slot_bytes for i in range(N_SLOTS)]
        self.red_off = _align(ctrl) + N_SLOTS * self.slot_bytes
        self.red_shard_shape = tuple(reduce_shard_shape)
        self.red_shard_floats = 1
        for d in reduce_shard_shape:
            self.red_shard_floats *= int(d)
        if self.red_shard_floats % 4:
            raise RuntimeError("peer transport needs a reduce-scatter shard that is a multiple of 4 floats")
        total = self.red_off + 4 * self.red_shard_floats * self.world
        torch.cuda.synchronize(self.device)
        self.arena = symm.empty(total, dtype=torch.uint8, device=self.device)
        self.hdl = symm.rendezvous(self.arena, self.group.group_name)
        self.arena.zero_()
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)                       # every control block is zero before anybody signals
        self.peers = torch.tensor([int(p) for p in self.hdl.buffer_ptrs], dtype=torch.int64, device=self.device)
        self.key = key
        self.calls = 0
        self.slot_call = [-1] * N_SLOTS
        self.last_reduce_call = -1

    def _view(self, off, nbytes, dtype, shape):
        return self.arena[off:off + nbytes].view(dtype).view(shape)

    # -------------------------------------------------------------- exchanges
    def allgather(self, shards):
        """Gathered (world * b, ...) views of the next arena slot, plus a ticket for ``check``."""
        lib = _lib.load()
        slot = self.calls % N_SLOTS
        self.slot_call[slot] = self.calls
        ticket = (slot, self.calls)
        self.calls += 1
        n = len(shards)
        shards = [t.contiguous() for t in shards]
        src = (C.c_void_p * n)(*[t.data_ptr() for t in shards])
        nb = (C.c_size_t * n)(*self.shard_bytes)
        offs = (C.c_size_t * n)(*[self.slot_off[slot] + r for r in self.rel_off])
        for t, b in zip(shards, self.shard_bytes):
            if t.numel() * t.element_size() != b:
                raise RuntimeError("peer transport: shard size differs from the arena layout")
        with torch.cuda.device(self.device):
            rc = lib.dvlp_p2p_allgather(_lib.ptr(self.peers), self.world, self.rank, n, src, nb, offs, _lib.stream_ptr())
        _lib.check(rc, "p2p_allgather")
        outs = [self._view(self.slot_off[slot] + r, b * self.world, dt, shp)
                for r, b, dt, shp in zip(self.rel_off, self.shard_bytes, self.gdtypes, self.gshapes)]
        return outs, ticket

    def check(self, ticket):
        slot, call = ticket
        if self.slot_call[slot] != call:
            raise RuntimeError("demovlp_b200: the gathered text embeddings of this forward were overwritten "
                               f"(more than {N_SLOTS} sharded forwards in flight before their backward); "
                               "set DVLP_P2P=0 to use the NCCL transport")

    def sum_scalar(self, value):
        lib = _lib.load()
        value = value.to(torch.float32).contiguous()
        out = torch.empty((), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = lib.dvlp_p2p_sum_scalar(_lib.ptr(self.peers), self.world, self.rank, _lib.ptr(value), _lib.ptr(out),
                                         _lib.stream_ptr())
        _lib.check(rc, "p2p_sum_scalar")
        return out

    def check_errors(self):
        """Raise if a peer-memory wait of this rank ever timed out (DVLP_P2P_TIMEOUT_S).  Synchronises the stream:
        call it at a point where the host waits anyway (e.g. where the trainer reads ``loss.item()``)."""
        with torch.cuda.device(self.device):
            rc = _lib.load().dvlp_p2p_check(_lib.ptr(self.arena), _lib.stream_ptr())
        _lib.check(rc, "p2p_check")

    def reduce_buffer(self):
        """fp32 (world * b, ...) region of the arena the backward writes its text-gradient partial sums into.
        The region is single-buffered: a second backward before another forward (gather) has ordered the peers'
        reads of the previous reduce-scatter could overwrite it, so that sequence is refused."""
        if self.last_reduce_call == self.calls:
            raise RuntimeError("demovlp_b200: two sharded backwards without a forward in between share the single "
                               "reduce-scatter region of the peer arena; set DVLP_P2P=0 to use the NCCL transport")
        self.last_reduce_call = self.calls
        shp = (self.world * self.red_shard_shape[0],) + self.red_shard_shape[1:]
        return self._view(self.red_off, 4 * self.red_shard_floats * self.world, torch.float32, shp)

    def reduce_scatter(self, out_dtype=torch.float32):
        lib = _lib.load()
        out = torch.empty(self.red_shard_shape, dtype=out_dtype, device=self.device)
        with torch.cuda.device(self.device):
            rc = lib.dvlp_p2p_reduce_scatter(_lib.ptr(self.peers), self.world, self.rank, self.red_off,
                                             self.red_shard_floats, _lib.ptr(out),
                                             _lib.BF16 if out_dtype == torch.bfloat16 else _lib.F32, _lib.stream_ptr())
        _lib.check(rc, "p2p_reduce_scatter")
        return out
