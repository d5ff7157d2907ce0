"""Multi-GPU plumbing: one process per GPU, torch.distributed for rendezvous / setup exchange,
NCCL (inside libb200fe.so) for the data path.

Stands in for deal.II's Utilities::MPI::Partitioner setup: every rank tells
the owners which of their DoFs it ghosts; owners turn that into their send (import) lists.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from ._lib import check, lib


class _HaloDesc(C.Structure):
    _fields_ = [("rank", C.c_int), ("n_ranks", C.c_int), ("nccl_unique_id", C.c_void_p),
                ("n_owned", C.c_uint32), ("n_ghost", C.c_uint32), ("n_peers", C.c_int), ("peers", C.c_void_p),
                ("recv_offset", C.c_void_p), ("recv_count", C.c_void_p), ("send_offset", C.c_void_p),
                ("send_count", C.c_void_p), ("h_send_indices", C.c_void_p), ("n_send", C.c_uint32)]


def exchange_lists(mesh, group=None):
    """Peer tables of one rank.  Works with any torch.distributed backend (gloo on CPU).

    Returns dict(peers, recv_offset, recv_count, send_offset, send_count, send_indices): the ghost
    segment is grouped by owner (sorted by global index), the send list holds for each peer the
    owned local indices that peer ghosts, in that peer's ghost order."""
    rank, world = mesh.rank, mesh.n_ranks
    recv = {}
    if mesh.n_ghost:
        owners = mesh.ghost_owner
        starts = np.nonzero(np.diff(owners, prepend=owners[0] - 1))[0]
        ends = np.append(starts[1:], len(owners))
        for s, e in zip(starts, ends):
            recv[int(owners[s])] = (int(s), int(e - s))
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (mesh.ghost_global, mesh.ghost_owner), group=group)
    else:
        gathered = [(mesh.ghost_global, mesh.ghost_owner)]
    send = {}
    for t, (gg, go) in enumerate(gathered):
        if t == rank:
            continue
        sel = gg[go == rank]
        if len(sel):
            send[t] = (sel - np.uint64(mesh.owned_begin)).astype(np.uint32)
    peers = sorted(set(recv) | set(send))
    recv_offset = np.array([recv.get(t, (0, 0))[0] for t in peers], dtype=np.uint32)
    recv_count = np.array([recv.get(t, (0, 0))[1] for t in peers], dtype=np.uint32)
    send_count = np.array([len(send.get(t, ())) for t in peers], dtype=np.uint32)
    send_offset = (np.cumsum(send_count) - send_count).astype(np.uint32)
    send_indices = np.concatenate([send[t] for t in peers if t in send]).astype(np.uint32) if send else np.zeros(0, np.uint32)
    return dict(peers=np.array(peers, dtype=np.int32), recv_offset=recv_offset, recv_count=recv_count,
                send_offset=send_offset, send_count=send_count, send_indices=send_indices)


def exchange_lists_local(mesh):
    """The same peer tables from the C ABI (b200fe_exchange_*): every rank replays the other ranks' mesh views, no
    communication at all.  Same dict as exchange_lists()."""
    h = C.c_void_p()
    check(mesh._exchange_create(C.byref(mesh._desc), C.byref(h)))
    try:
        n_peers, n_send = C.c_int(), C.c_uint32()
        check(lib.b200fe_exchange_info(h, C.byref(n_peers), C.byref(n_send), None, None))
        peers = np.empty(n_peers.value, dtype=np.int32)
        arrs = [np.empty(n_peers.value, dtype=np.uint32) for _ in range(4)]
        send_indices = np.empty(n_send.value, dtype=np.uint32)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib.b200fe_exchange_fill(h, ptr(peers), *[ptr(a) for a in arrs], ptr(send_indices)))
    finally:
        lib.b200fe_exchange_destroy(h)
    return dict(peers=peers, recv_offset=arrs[0], recv_count=arrs[1], send_offset=arrs[2], send_count=arrs[3], send_indices=send_indices)


class Halo:
    """Ghost exchange object of one rank (NCCL communicator + pack lists on the device)."""

    def __init__(self, mesh, group=None):
        self.mesh = mesh
        self.lists = exchange_lists(mesh, group)
        L = self.lists
        uid = [None]
        if mesh.n_ranks > 1:
            if mesh.rank == 0:
                buf = C.create_string_buffer(128)
                check(lib.b200fe_comm_unique_id(buf))
                uid[0] = buf.raw
            dist.broadcast_object_list(uid, src=0, group=group)
        self._uid = C.create_string_buffer(uid[0], 128) if uid[0] is not None else None
        d = _HaloDesc()
        d.rank, d.n_ranks = mesh.rank, mesh.n_ranks
        d.nccl_unique_id = C.cast(self._uid, C.c_void_p) if self._uid is not None else None
        d.n_owned, d.n_ghost = mesh.n_owned, mesh.n_ghost
        d.n_peers = len(L["peers"])
        self._keep = [np.ascontiguousarray(L[k]) for k in ("peers", "recv_offset", "recv_count", "send_offset", "send_count", "send_indices")]
        ptr = lambda a: a.ctypes.data if a.size else None
        d.peers, d.recv_offset, d.recv_count, d.send_offset, d.send_count, d.h_send_indices = [ptr(a) for a in self._keep]
        d.n_send = len(L["send_indices"])
        self._h = C.c_void_p()
        check(lib.b200fe_halo_create(C.byref(d), C.byref(self._h)))

    def _sp(self, stream=None):
        s = torch.cuda.current_stream() if stream is None else stream
        return C.c_void_p(s.cuda_stream)

    def update_ghost_values(self, v: torch.Tensor, stream=None):
        check(lib.b200fe_halo_update_ghosts(self._h, C.c_void_p(v.data_ptr()), self._sp(stream)))

    def compress_add(self, v: torch.Tensor, stream=None):
        check(lib.b200fe_halo_compress_add(self._h, C.c_void_p(v.data_ptr()), self._sp(stream)))

    def zero_out_ghost_values(self, v: torch.Tensor, stream=None):
        check(lib.b200fe_halo_zero_ghosts(self._h, C.c_void_p(v.data_ptr()), self._sp(stream)))

    def transport(self) -> str:
        """'p2p' (peer stores over CUDA-IPC windows) or 'nccl' (grouped send/recv)."""
        avail, used = C.c_int(), C.c_int()
        check(lib.b200fe_halo_transport(self._h, C.byref(avail), C.byref(used)))
        return "p2p" if used.value else "nccl"

    def p2p_available(self) -> bool:
        avail, used = C.c_int(), C.c_int()
        check(lib.b200fe_halo_transport(self._h, C.byref(avail), C.byref(used)))
        return bool(avail.value)

    def set_transport(self, name: str):
        """Collective: every rank must switch before the next exchange."""
        check(lib.b200fe_halo_set_transport(self._h, int(name == "p2p")))

    def status(self):
        """Synchronises; raises if a bounded wait of the P2P transport expired."""
        check(lib.b200fe_halo_status(self._h))

    def allreduce_sum(self, vals: torch.Tensor, stream=None):
        check(lib.b200fe_halo_allreduce_sum(self._h, C.c_void_p(vals.data_ptr()), vals.numel(), self._sp(stream)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.b200fe_halo_destroy(h)
            self._h = None
