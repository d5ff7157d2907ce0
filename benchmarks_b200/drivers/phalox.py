#!/usr/bin/env python
"""phalox.py -- the reference's halo-exchange microbenchmark (p-halox/phalox.cc) over NVLink: NCCL send/recv and
CUDA P2P (peer stores into CUDA-IPC windows + flags, csrc/halo_p2p.cu), side by side.

  torchrun --nproc-per-node N --master-addr 127.0.0.1 --master-port P benchmarks_b200/drivers/phalox.py \
           [dim=2] [KB=64] [nMsg=2] [is_periodic=1] [warmup=30] [print_topo=0]
  PHALOX_TRANSPORT=nccl|p2p|both (default both) selects the transports; every line ends with `transport= ... payload= ok`.
  Every message carries a recognisable payload (sender rank, sender's message slot, position) that the receiver checks
  after the timed rounds -- the reference only moves zeros.

Same Cartesian decomposition (MPI_Dims_create / MPI_Cart_shift semantics, phalox.cc:49-88: self and
non-periodic boundaries dropped, duplicates kept), same message schedule ((warmup + nMsg) rounds of
"receive from all, send to all, complete", clock after a barrier at msg == warmup) and the same output line
(:148-154).  Two timings are printed: `sync` completes every round on the host like MPI_Waitall,
`stream` enqueues all rounds on the CUDA stream and times them with CUDA events (the GPU-native way); with the P2P
transport `launch` runs the nMsg rounds inside one kernel launch (b200fe_halo_exchange_raw_rounds; small messages)."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from benchmarks_b200._lib import check, lib  # noqa: E402
from benchmarks_b200.dist import _HaloDesc  # noqa: E402


def dims_create(n, dim):
    """MPI_Dims_create: balanced factors, non-increasing."""
    dims = [1] * dim
    f, m = [], n
    d = 2
    while d * d <= m:
        while m % d == 0:
            f.append(d)
            m //= d
        d += 1
    if m > 1:
        f.append(m)
    for p in sorted(f, reverse=True):
        dims[int(np.argmin(dims))] *= p
    return sorted(dims, reverse=True)


def main():
    a = sys.argv[1:]
    dim, KB, nMsg, periodic, warmup, print_topo = [int(a[i]) if len(a) > i else v for i, v in enumerate((2, 64, 2, 1, 30, 0))]
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gloo = dist.new_group(backend="gloo")
    # PHALOX_SWEEP_DIMS / PHALOX_SWEEP_KB (comma lists): the README's sweep (p-halox/README.md:62-68) inside one launch
    dims_list = [int(x) for x in os.environ.get("PHALOX_SWEEP_DIMS", str(dim)).split(",")]
    kb_list = [int(x) for x in os.environ.get("PHALOX_SWEEP_KB", str(KB)).split(",")]
    for dd in dims_list:
        for kb in kb_list:
            run_config(dd, kb, nMsg, periodic, warmup, print_topo, rank, world, gloo)
    dist.destroy_process_group()


def run_config(dim, KB, nMsg, periodic, warmup, print_topo, rank, world, gloo):
    dims = dims_create(world, dim)
    coords = list(np.unravel_index(rank, dims))  # row-major like MPI_Cart
    neighbors = []
    for d in range(dim):
        for step in (-1, +1):
            c = list(coords)
            c[d] += step
            if periodic:
                c[d] %= dims[d]
            elif not (0 <= c[d] < dims[d]):
                continue
            nb = int(np.ravel_multi_index(c, dims))
            if nb != rank:
                neighbors.append(nb)
    nneigh = len(neighbors)
    n_doubles = KB * 1024 // 8
    send = torch.zeros(max(nneigh * n_doubles, 1), dtype=torch.float64, device="cuda")
    recv = torch.zeros_like(send)
    pos = (torch.arange(n_doubles, device="cuda") % 7).to(torch.float64) * 0.125
    for j in range(nneigh):   # payload of my j-th message
        send[j * n_doubles:(j + 1) * n_doubles] = rank * 1000.0 + j + pos
    all_neighbors = [None] * world
    dist.all_gather_object(all_neighbors, neighbors, group=gloo)

    def expected_sender_slot(k):
        """My k-th receive pairs with the sender's messages to me in posting order (MPI / NCCL matching rule)."""
        B = neighbors[k]
        occ = sum(1 for q in range(k) if neighbors[q] == B)
        slots = [q for q, t in enumerate(all_neighbors[B]) if t == rank]
        return slots[occ]

    def payload_ok():
        ok = True
        for k in range(nneigh):
            want = neighbors[k] * 1000.0 + expected_sender_slot(k) + pos
            ok &= bool(torch.equal(recv[k * n_doubles:(k + 1) * n_doubles], want))
        return ok
    uid = [None]
    if rank == 0:
        buf = C.create_string_buffer(128)
        check(lib.b200fe_comm_unique_id(buf))
        uid[0] = buf.raw
    dist.broadcast_object_list(uid, src=0, group=gloo)
    d = _HaloDesc()
    d.rank, d.n_ranks = rank, world
    uid_buf = C.create_string_buffer(uid[0], 128)
    d.nccl_unique_id = C.cast(uid_buf, C.c_void_p)
    d.n_owned = d.n_ghost = d.n_send = nneigh * n_doubles
    d.n_peers = nneigh
    peers = np.array(neighbors, dtype=np.int32)
    off = (np.arange(nneigh) * n_doubles).astype(np.uint32)
    cnt = np.full(nneigh, n_doubles, dtype=np.uint32)
    ptr = lambda x: x.ctypes.data if x.size else None
    d.peers, d.recv_offset, d.recv_count, d.send_offset, d.send_count = ptr(peers), ptr(off), ptr(cnt), ptr(off), ptr(cnt)
    d.h_send_indices = None
    h = C.c_void_p()
    check(lib.b200fe_halo_create(C.byref(d), C.byref(h)))
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    def one_round():
        check(lib.b200fe_halo_exchange_raw(h, C.c_void_p(send.data_ptr()), C.c_void_p(recv.data_ptr()), sp))

    def run_transport(transport):
        results = {}
        if transport == "p2p":
            # "launch": the nMsg rounds inside ONE kernel launch (b200fe_halo_exchange_raw_rounds) -- the exchange as a
            # device-resident loop issues it; no launch gap between rounds: the latency floor of the NVLink fabric
            recv.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            check(lib.b200fe_halo_exchange_raw_rounds(h, C.c_void_p(send.data_ptr()), C.c_void_p(recv.data_ptr()), max(warmup, 1), sp))
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(lib.b200fe_halo_exchange_raw_rounds(h, C.c_void_p(send.data_ptr()), C.c_void_p(recv.data_ptr()), nMsg, sp))
            e1.record()
            torch.cuda.synchronize()
            results["launch"] = (e0.elapsed_time(e1) * 1e-3, payload_ok() and lib.b200fe_halo_status(h) == 0)
        for mode in ("sync", "stream"):
            recv.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            for msg in range(nMsg + warmup):
                if msg == warmup:
                    torch.cuda.synchronize()
                    dist.barrier()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    e0 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                one_round()
                if mode == "sync":
                    stream.synchronize()
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            torch.cuda.synchronize()
            results[mode] = ((time.perf_counter() - t0) if mode == "sync" else e0.elapsed_time(e1) * 1e-3, payload_ok() and lib.b200fe_halo_status(h) == 0)
        for mode, (t, ok) in results.items():
            times = [None] * world
            dist.all_gather_object(times, (t, nneigh * nMsg * KB, ok), group=gloo)
            if rank == 0:
                ts = np.array([x[0] for x in times])
                kb = sum(x[1] for x in times)
                print(f"P= {world} dim= {dim} KB= {KB} nMsg= {nMsg} is_periodic= {periodic} warmup= {warmup} print_topo= {print_topo}"
                      f" min_time_s= {ts.min():.6g} min_Rank= {int(ts.argmin())} max_time_s= {ts.max():.6g} max_Rank= {int(ts.argmax())}"
                      f" avg_time_s= {ts.mean():.6g} agg_BW_GBps= {kb / ts.max() / (1024.0 * 1024.0):.6g} mode= {mode}"
                      f" transport= {transport} payload= {'ok' if all(x[2] for x in times) else 'WRONG'}", flush=True)

    avail, used = C.c_int(), C.c_int()
    check(lib.b200fe_halo_transport(h, C.byref(avail), C.byref(used)))
    want = os.environ.get("PHALOX_TRANSPORT", "both")
    transports = [t for t in ("nccl", "p2p") if want in (t, "both") and (t == "nccl" or avail.value)]
    for transport in transports:
        check(lib.b200fe_halo_set_transport(h, int(transport == "p2p")))
        run_transport(transport)
    if print_topo and rank == 0:
        print("dims =", dims)
    lib.b200fe_halo_destroy(h)
    del send, recv
    torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
