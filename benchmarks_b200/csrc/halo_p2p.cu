// halo_p2p.cu -- one-sided ghost exchange and scalar all-reduce over CUDA-IPC peer memory (NVLink 5 / NVSwitch).
//
// Why: on one 8xB200 box a ghost exchange moves a few hundred KB per neighbour -- its cost is latency, not bandwidth
// (profiles/r01b_phalox_8gpu.txt: 10-14 us per NCCL send/recv round; SCALE_r01: +0.16 ms per CG iteration at 8 GPUs from
// 2 ncclAllReduce + 2 grouped send/recv rounds).  NVSwitch gives every GPU load/store access to every peer, so the pack
// kernel can write straight into the receiver and raise a flag -- no proxy thread, no rendezvous, no second stream:
//
//   update_ghost_values  : owner's SEND kernel gathers v[import_indices] and stores them into each ghosting peer's receive
//                          window, then (last block) releases one flag per peer; the receiver's WAIT kernel acquires the
//                          flags, moves the window into the ghost segment of v and acknowledges.
//   compress(add)        : the same in the other direction (ghost segments -> owner's window -> atomic add), ghosts zeroed
//                          by the SEND kernel.
//   CG inner products    : every rank stores its partial sums into its slot of every peer's window (all-gather), then each
//                          rank adds the slots in rank order -- bitwise identical on all ranks, one 32-thread kernel.
//
// One-sided means "post early, wait late" on ONE stream: the 3-phase overlap schedule of
// bakeoff_problems_dealii/include/portable_laplace_operator.h:669-696 becomes  send | interior cells | wait | boundary cells.
// Replaces MPI_Isend/Irecv/Waitall of p-halox/phalox.cc:104-126 and deal.II's Partitioner exchange.
//
// Protocol: every (sender entry, receiver entry) pair owns one flag and one acknowledgement word, both monotonically
// increasing epochs kept on the device (robust under replay).  Windows are double-buffered by epoch parity: a sender only
// needs the acknowledgement of the message before the previous one, so consecutive rounds pipeline and no round trip sits on
// the critical path.  A SEND kernel never depends on anything in flight, so the two-kernel form cannot deadlock whatever
// the residency of its blocks; small messages use ONE single-block kernel per round (send, flag, wait, copy out, ack: a
// block that has sent everything before it waits cannot deadlock either) -- r02b: the two-kernel round cost 20 us against
// NCCL's 12 us at 1 KiB.  Ordering: the data stores of a block are ordered before its flag by bar.sync + ONE system-scope
// release (cumulativity), not by a fence in every thread.  All waits are bounded (kTimeoutNs): on expiry a sticky error flag
// is raised, later waits return at once, and the host reports B200FE_ERR_COMM at its next check instead of hanging.
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "common.h"
#include "halo.h"

namespace b200fe {

namespace {

constexpr int kMaxEntries = 32;   // peer-table entries per rank (26 neighbours of a block + duplicates of periodic rings)
constexpr int kMaxRanks = 32;
constexpr int kMaxComps = 4;      // components per exchange (BP6: 3)
constexpr unsigned long long kTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

struct RedSlot {
    double val[4];
    unsigned long long epoch;
    unsigned long long pad[3];
};

// head of every window; written by peers only (and zeroed once at creation)
struct Ctrl {
    unsigned long long upd_flag[kMaxEntries];  // [my entry k]  update data of epoch e from peers[k] has landed
    unsigned long long upd_ack[kMaxEntries];   // [my entry j]  peers[j] has consumed my update data of epoch e
    unsigned long long cmp_flag[kMaxEntries];  // [my entry j]  compress data of epoch e from peers[j] has landed
    unsigned long long cmp_ack[kMaxEntries];   // [my entry k]  peers[k] has consumed my compress data of epoch e
    RedSlot red[2][kMaxRanks];                 // [epoch parity][source rank]
};

struct EntryDev {
    // this entry as a sender of update data / receiver of compress data (send_cnt > 0)
    double *peer_recv;                  // peer's receive window at the matching entry's offset
    unsigned long long *peer_upd_flag;  // peer's upd_flag[k_peer]
    unsigned long long *peer_cmp_ack;   // peer's cmp_ack[k_peer]
    uint32_t peer_n_ghost;              // component stride of the peer's receive window
    unsigned long long peer_recv_half;  // doubles per half of the peer's receive window (double buffering by epoch parity)
    ulonglong2 *peer_ll_recv;           // the same for the low-latency (flag-in-data) window
    unsigned long long peer_ll_recv_half;
    // this entry as a receiver of update data / sender of compress data (recv_cnt > 0)
    double *peer_back;                  // peer's compress window at the matching entry's offset
    unsigned long long *peer_cmp_flag;  // peer's cmp_flag[j_peer]
    unsigned long long *peer_upd_ack;   // peer's upd_ack[j_peer]
    uint32_t peer_n_send;
    unsigned long long peer_back_half;
    ulonglong2 *peer_ll_back;
    unsigned long long peer_ll_back_half;
    uint32_t send_off, send_cnt, recv_off, recv_cnt;
};

struct DevState {  // local device memory, never touched by peers
    EntryDev e[kMaxEntries];
    RedSlot *peer_red[kMaxRanks];  // &window(r).ctrl.red[0][my_rank]; parity 1 is kMaxRanks slots further
    int n_entries, n_ranks, rank;
    unsigned long long recv_half, back_half;  // doubles per window half of THIS rank (receive / compress regions)
    unsigned long long ll_recv_half, ll_back_half;  // elements per half of the low-latency windows
    unsigned long long upd_send_epoch, upd_wait_epoch, cmp_send_epoch, cmp_wait_epoch, red_epoch;
    unsigned int ctr_send, ctr_wait;
    int error;
};

struct Meta {  // exchanged once at creation
    cudaIpcMemHandle_t handle;
    uint32_t n_ghost, n_send, n_entries, ok, comps, pad[3];
    uint32_t entry[kMaxEntries][5];  // peer, recv_off, recv_cnt, send_off, send_cnt
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// bounded wait for *flag >= target; false after a timeout or when another wait has already failed
__device__ bool spin_until(const unsigned long long *flag, unsigned long long target, int *error)
{
    // poll with relaxed loads (no fence per poll), acquire once the flag is there
    for (int i = 0; i < 4000; ++i)  // tight polling first: the common wait is NVLink latency
        if (ld_relaxed_sys(flag) >= target) return ld_acquire_sys(flag) >= target;
    const unsigned long long t0 = global_ns();
    while (ld_relaxed_sys(flag) < target) {
        if (*(volatile int *)error != 0) return false;
        if (global_ns() - t0 > kTimeoutNs) {
            atomicExch(error, 1);
            return false;
        }
        __nanosleep(100);
    }
    return ld_acquire_sys(flag) >= target;
}

enum : int { MODE_UPDATE = 0, MODE_COMPRESS = 1 };

// entry that holds position j of the packed send list (update) or of the ghost segment (compress)
__device__ __forceinline__ int find_entry(const uint32_t *off, const uint32_t *cnt, int n, uint32_t j)
{
    for (int k = 0; k < n; ++k)
        if (j - off[k] < cnt[k]) return k;  // unsigned: j >= off && j < off + cnt
    return -1;
}

// SEND: store this rank's data into the peers' windows, then release one flag per peer (last block).
//   update   : src = v[c*stride + send_idx[j]] (or raw_send[j]),        j < n_send,  -> peer_recv[c*peer_n_ghost + j - send_off]
//   compress : src = v[c*stride + n_owned + j], then zeroed,            j < n_ghost, -> peer_back[c*peer_n_send + j - recv_off]
__global__ void p2p_send_kernel(DevState *st, Ctrl *my, int mode, double *v, uint32_t n_owned, uint32_t n_items,
                                const uint32_t *__restrict__ send_idx, int ncomp, size_t stride, const double *__restrict__ raw_send)
{
    __shared__ uint32_t s_off[kMaxEntries], s_cnt[kMaxEntries];
    __shared__ bool s_last;
    const int n = st->n_entries;
    const unsigned long long epoch = (mode == MODE_UPDATE ? st->upd_send_epoch : st->cmp_send_epoch) + 1;
    if ((int)threadIdx.x < n) {
        const EntryDev &e = st->e[threadIdx.x];
        s_off[threadIdx.x] = mode == MODE_UPDATE ? e.send_off : e.recv_off;
        s_cnt[threadIdx.x] = mode == MODE_UPDATE ? e.send_cnt : e.recv_cnt;
        // the peer must have consumed the message that used this half of its window (two epochs ago)
        if (s_cnt[threadIdx.x] && epoch > 2)
            spin_until(mode == MODE_UPDATE ? &my->upd_ack[threadIdx.x] : &my->cmp_ack[threadIdx.x], epoch - 2, &st->error);
    }
    __syncthreads();
    const size_t total = (size_t)n_items * ncomp;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t)(i / n_items), j = (uint32_t)(i - (size_t)c * n_items);
        const int k = find_entry(s_off, s_cnt, n, j);
        if (k < 0) continue;
        const EntryDev &e = st->e[k];
        if (mode == MODE_UPDATE) {
            const double x = raw_send ? raw_send[j] : v[c * stride + send_idx[j]];
            e.peer_recv[(epoch & 1ull) * e.peer_recv_half + (size_t)c * e.peer_n_ghost + (j - s_off[k])] = x;
        } else {
            double *g = v + c * stride + n_owned + j;
            e.peer_back[(epoch & 1ull) * e.peer_back_half + (size_t)c * e.peer_n_send + (j - s_off[k])] = *g;
            *g = 0.0;  // compress(add) leaves zeroed ghosts
        }
    }
    __syncthreads();  // the block's stores happen-before thread 0's fence (bar.sync), which is cumulative
    if (threadIdx.x == 0) {
        __threadfence_system();
        s_last = atomicAdd(&st->ctr_send, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    if ((int)threadIdx.x < n && s_cnt[threadIdx.x]) {
        const EntryDev &e = st->e[threadIdx.x];
        st_release_sys(mode == MODE_UPDATE ? e.peer_upd_flag : e.peer_cmp_flag, epoch);
    }
    if (threadIdx.x == 0) {
        st->ctr_send = 0;
        if (mode == MODE_UPDATE) st->upd_send_epoch = epoch;
        else st->cmp_send_epoch = epoch;
    }
}

// WAIT: acquire the peers' flags, move the window into place, acknowledge (last block).
//   update   : v[c*stride + n_owned + j] = recv[c*n_ghost + j]  (or raw_recv[j]),   j < n_ghost
//   compress : v[c*stride + send_idx[j]] += back[c*n_send + j],                      j < n_send
__global__ void p2p_wait_kernel(DevState *st, Ctrl *my, int mode, double *v, uint32_t n_owned, uint32_t n_items,
                                const uint32_t *__restrict__ send_idx, int ncomp, size_t stride, const double *win, double *raw_recv)
{
    __shared__ bool s_last;
    const unsigned long long half = mode == MODE_UPDATE ? st->recv_half : st->back_half;
    const int n = st->n_entries;
    const unsigned long long epoch = (mode == MODE_UPDATE ? st->upd_wait_epoch : st->cmp_wait_epoch) + 1;
    if ((int)threadIdx.x < n) {
        const EntryDev &e = st->e[threadIdx.x];
        if (mode == MODE_UPDATE ? e.recv_cnt : e.send_cnt)
            spin_until(mode == MODE_UPDATE ? &my->upd_flag[threadIdx.x] : &my->cmp_flag[threadIdx.x], epoch, &st->error);
    }
    __syncthreads();
    const size_t total = (size_t)n_items * ncomp;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t)(i / n_items), j = (uint32_t)(i - (size_t)c * n_items);
        const double x = __ldcg(win + (epoch & 1ull) * half + i);  // written by a peer: never through L1
        if (mode == MODE_UPDATE) {
            if (raw_recv) raw_recv[j] = x;
            else v[c * stride + n_owned + j] = x;
        } else {
            atomicAdd(v + c * stride + send_idx[j], x);  // the same owned DoF may be ghosted by several peers
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(&st->ctr_wait, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    if ((int)threadIdx.x < n) {
        const EntryDev &e = st->e[threadIdx.x];
        if (mode == MODE_UPDATE ? e.recv_cnt : e.send_cnt) st_release_sys(mode == MODE_UPDATE ? e.peer_upd_ack : e.peer_cmp_ack, epoch);
    }
    if (threadIdx.x == 0) {
        st->ctr_wait = 0;
        if (mode == MODE_UPDATE) st->upd_wait_epoch = epoch;
        else st->cmp_wait_epoch = epoch;
    }
}

// One round in ONE single-block kernel (small messages): send, flag, wait, copy out, acknowledge.  The block has sent
// everything before it waits, and waits only for peers' sends, so it cannot deadlock; no second launch, no grid-wide counter.
__global__ void __launch_bounds__(1024) p2p_round_kernel(DevState *st, Ctrl *my, int mode, double *v, uint32_t n_owned, uint32_t n_out, uint32_t n_in,
                                                          const uint32_t *__restrict__ send_idx, int ncomp, size_t stride,
                                                          const double *__restrict__ raw_send, const double *win, double *raw_recv)
{
    __shared__ uint32_t s_off[kMaxEntries], s_cnt[kMaxEntries];
    const int n = st->n_entries, tid = threadIdx.x;
    const unsigned long long epoch = (mode == MODE_UPDATE ? st->upd_send_epoch : st->cmp_send_epoch) + 1;
    const unsigned long long par = epoch & 1ull;
    bool out_k = false, in_k = false;
    if (tid < n) {
        const EntryDev &e = st->e[tid];
        s_off[tid] = mode == MODE_UPDATE ? e.send_off : e.recv_off;
        s_cnt[tid] = mode == MODE_UPDATE ? e.send_cnt : e.recv_cnt;
        out_k = s_cnt[tid] != 0;
        in_k = (mode == MODE_UPDATE ? e.recv_cnt : e.send_cnt) != 0;
        if (out_k && epoch > 2) spin_until(mode == MODE_UPDATE ? &my->upd_ack[tid] : &my->cmp_ack[tid], epoch - 2, &st->error);
    }
    __syncthreads();
    const uint32_t total_out = n_out * (uint32_t)ncomp;
    for (uint32_t i = tid; i < total_out; i += blockDim.x) {
        const uint32_t c = i / n_out, j = i - c * n_out;
        const int k = find_entry(s_off, s_cnt, n, j);
        if (k < 0) continue;
        const EntryDev &e = st->e[k];
        if (mode == MODE_UPDATE) {
            const double x = raw_send ? raw_send[j] : v[c * stride + send_idx[j]];
            e.peer_recv[par * e.peer_recv_half + (size_t)c * e.peer_n_ghost + (j - s_off[k])] = x;
        } else {
            double *g = v + c * stride + n_owned + j;
            e.peer_back[par * e.peer_back_half + (size_t)c * e.peer_n_send + (j - s_off[k])] = *g;
            *g = 0.0;
        }
    }
    __syncthreads();
    if (tid < n) {
        const EntryDev &e = st->e[tid];
        if (out_k) st_release_sys(mode == MODE_UPDATE ? e.peer_upd_flag : e.peer_cmp_flag, epoch);  // release: cumulative over the block's stores
        if (in_k) spin_until(mode == MODE_UPDATE ? &my->upd_flag[tid] : &my->cmp_flag[tid], epoch, &st->error);
    }
    __syncthreads();
    const unsigned long long half = mode == MODE_UPDATE ? st->recv_half : st->back_half;
    const uint32_t total_in = n_in * (uint32_t)ncomp;
    for (uint32_t i = tid; i < total_in; i += blockDim.x) {
        const uint32_t c = i / n_in, j = i - c * n_in;
        const double x = __ldcg(win + par * half + i);
        if (mode == MODE_UPDATE) {
            if (raw_recv) raw_recv[j] = x;
            else v[c * stride + n_owned + j] = x;
        } else {
            atomicAdd(v + c * stride + send_idx[j], x);
        }
    }
    __syncthreads();
    if (tid < n && in_k) {
        const EntryDev &e = st->e[tid];
        st_release_sys(mode == MODE_UPDATE ? e.peer_upd_ack : e.peer_cmp_ack, epoch);
    }
    if (tid == 0) {  // both halves of the split protocol advance together
        if (mode == MODE_UPDATE) { st->upd_send_epoch = epoch; st->upd_wait_epoch = epoch; }
        else { st->cmp_send_epoch = epoch; st->cmp_wait_epoch = epoch; }
    }
}

// Low-latency round (small messages, every rank below kLLMax doubles): NCCL-LL style "flag in the data".  A double travels as
// two 8-byte words {epoch tag, 32-bit half}; 8-byte stores are single-copy atomic, so the receiver simply polls every element
// until both tags carry this epoch -- no fence between data and flag, hence no NVLink round trip on the sender's side
// (the release in p2p_round_kernel costs ~3 us; r02d: 10.8 us per 1 KiB round, NCCL 13 us, reference MPI ~6 us).
__device__ __forceinline__ void st_relaxed_sys_v2(ulonglong2 *p, unsigned long long a, unsigned long long b)
{
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_relaxed_sys_v2(const ulonglong2 *p)
{
    ulonglong2 v;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// n_rounds > 1 (raw exchanges only): that many consecutive rounds of the same buffers inside ONE launch -- the exchange as a
// device-resident loop would issue it (no launch between rounds; a round then costs the NVLink hop and the polling, not the
// 3 us launch-to-launch gap of dependent kernels).  Epochs, window halves and acknowledgements advance exactly as over launches.
__global__ void __launch_bounds__(1024) p2p_round_ll_kernel(DevState *st, Ctrl *my, int mode, double *v, uint32_t n_owned, uint32_t n_out, uint32_t n_in,
                                                             const uint32_t *__restrict__ send_idx, int ncomp, size_t stride,
                                                             const double *__restrict__ raw_send, const ulonglong2 *win, double *raw_recv,
                                                             int n_rounds)
{
    // the peer table of this launch in shared memory: read from device memory once, not once per element and round
    __shared__ uint32_t s_off[kMaxEntries], s_cnt[kMaxEntries], s_cstride[kMaxEntries];
    __shared__ ulonglong2 *s_dst[kMaxEntries];
    __shared__ unsigned long long s_dhalf[kMaxEntries];
    const int n = st->n_entries, tid = threadIdx.x;
    const unsigned long long epoch0 = (mode == MODE_UPDATE ? st->upd_send_epoch : st->cmp_send_epoch);
    const unsigned long long half = mode == MODE_UPDATE ? st->ll_recv_half : st->ll_back_half;
    bool in_k = false;
    unsigned long long *ack_out = nullptr;
    if (tid < n) {
        const EntryDev &e = st->e[tid];
        s_off[tid] = mode == MODE_UPDATE ? e.send_off : e.recv_off;
        s_cnt[tid] = mode == MODE_UPDATE ? e.send_cnt : e.recv_cnt;
        s_dst[tid] = mode == MODE_UPDATE ? e.peer_ll_recv : e.peer_ll_back;
        s_dhalf[tid] = mode == MODE_UPDATE ? e.peer_ll_recv_half : e.peer_ll_back_half;
        s_cstride[tid] = mode == MODE_UPDATE ? e.peer_n_ghost : e.peer_n_send;
        in_k = (mode == MODE_UPDATE ? e.recv_cnt : e.send_cnt) != 0;
        ack_out = mode == MODE_UPDATE ? e.peer_upd_ack : e.peer_cmp_ack;
    }
  for (int round = 0; round < n_rounds; ++round) {
    const unsigned long long epoch = epoch0 + 1 + (unsigned long long)round;
    const unsigned long long par = epoch & 1ull, tag = (epoch & 0xFFFFFFFFull) << 32;
    // the peer must have consumed the message that used this half of its window (two epochs ago)
    if (tid < n && s_cnt[tid] && epoch > 2) spin_until(mode == MODE_UPDATE ? &my->upd_ack[tid] : &my->cmp_ack[tid], epoch - 2, &st->error);
    __syncthreads();
    const uint32_t total_out = n_out * (uint32_t)ncomp;
    for (uint32_t i = tid; i < total_out; i += blockDim.x) {
        const uint32_t c = i / n_out, j = i - c * n_out;
        const int k = find_entry(s_off, s_cnt, n, j);
        if (k < 0) continue;
        double x;
        if (mode == MODE_UPDATE) {
            x = raw_send ? raw_send[j] : v[c * stride + send_idx[j]];
        } else {
            double *g = v + c * stride + n_owned + j;
            x = *g;
            *g = 0.0;
        }
        ulonglong2 *dst = s_dst[k] + par * s_dhalf[k] + (size_t)c * s_cstride[k] + (j - s_off[k]);
        const unsigned long long bits = (unsigned long long)__double_as_longlong(x);
        st_relaxed_sys_v2(dst, tag | (bits & 0xFFFFFFFFull), tag | (bits >> 32));
    }
    // receive: every element carries its own flag
    const uint32_t total_in = n_in * (uint32_t)ncomp;
    for (uint32_t i = tid; i < total_in; i += blockDim.x) {
        const uint32_t c = i / n_in, j = i - c * n_in;
        const ulonglong2 *src = win + par * half + i;
        ulonglong2 w = ld_relaxed_sys_v2(src);
        if ((w.x & 0xFFFFFFFF00000000ull) != tag || (w.y & 0xFFFFFFFF00000000ull) != tag) {
            const unsigned long long t0 = global_ns();
            for (;;) {
                w = ld_relaxed_sys_v2(src);
                if ((w.x & 0xFFFFFFFF00000000ull) == tag && (w.y & 0xFFFFFFFF00000000ull) == tag) break;
                if (*(volatile int *)&st->error != 0) break;
                if (global_ns() - t0 > kTimeoutNs) { atomicExch(&st->error, 1); break; }
            }
        }
        const double x = __longlong_as_double((long long)((w.x & 0xFFFFFFFFull) | (w.y << 32)));
        if (mode == MODE_UPDATE) {
            if (raw_recv) raw_recv[j] = x;
            else v[c * stride + n_owned + j] = x;
        } else {
            atomicAdd(v + c * stride + send_idx[j], x);
        }
    }
    __syncthreads();  // every element of this round has been read: the senders may reuse this half two rounds from now
    if (tid < n && in_k) st_relaxed_sys(ack_out, epoch);
  }  // rounds
    if (tid == 0) {
        const unsigned long long epoch = epoch0 + (unsigned long long)n_rounds;
        if (mode == MODE_UPDATE) { st->upd_send_epoch = epoch; st->upd_wait_epoch = epoch; }
        else { st->cmp_send_epoch = epoch; st->cmp_wait_epoch = epoch; }
    }
}

// all-gather of the partial sums into every rank's window, then the sum in rank order (identical bits on every rank)
__global__ void p2p_allreduce_kernel(DevState *st, Ctrl *my, double *vals, int count)
{
    const int lane = threadIdx.x, R = st->n_ranks;
    const unsigned long long epoch = st->red_epoch + 1;
    const int par = (int)(epoch & 1ull);
    if (lane < R) {
        RedSlot *dst = st->peer_red[lane] + par * kMaxRanks;
        for (int i = 0; i < count; ++i) dst->val[i] = vals[i];
        st_release_sys(&dst->epoch, epoch);  // release: the payload stores above are visible before the epoch
        spin_until(&my->red[par][lane].epoch, epoch, &st->error);
    }
    __syncwarp();
    if (lane < count) {
        double s = 0.0;
        for (int r = 0; r < R; ++r) s += __ldcg(&my->red[par][r].val[lane]);
        vals[lane] = s;
    }
    if (lane == 0) st->red_epoch = epoch;
}

inline unsigned blocks_for(size_t n) { return n == 0 ? 1u : (unsigned)std::min<size_t>((n + 255) / 256, 148u * 4u); }

}  // namespace

struct P2P {
    void *window = nullptr;               // my window (cudaMalloc, exported through CUDA IPC)
    std::vector<void *> peer_base;        // mapped windows of the other ranks (null for my own rank)
    DevState *d_state = nullptr;
    double *recv = nullptr, *back = nullptr;  // regions of my window
    ulonglong2 *ll_recv = nullptr, *ll_back = nullptr;
    int comps = 1;
    int ll_max_comps = 0;  // components per exchange for which EVERY rank stays below kLLMax doubles in both directions
    ~P2P()
    {
        for (void *p : peer_base)
            if (p) cudaIpcCloseMemHandle(p);
        cudaFree(d_state);
        cudaFree(window);
    }
};

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
static size_t recv_offset_bytes() { return align256(sizeof(Ctrl)); }
// every region holds two halves (epoch parity); a half is a multiple of 32 doubles
static size_t half_doubles(uint32_t n, int comps) { return ((size_t)n * comps + 31) & ~size_t(31); }
static size_t back_offset_bytes(uint32_t n_ghost, int comps) { return recv_offset_bytes() + 2 * sizeof(double) * half_doubles(n_ghost, comps); }
constexpr size_t kLLMax = 4096;  // doubles per direction and round that may use the low-latency windows
static size_t ll_half(uint32_t n, int comps) { return std::min(half_doubles(n, comps), (kLLMax + 31) & ~size_t(31)); }
static size_t ll_recv_offset_bytes(uint32_t n_ghost, uint32_t n_send, int comps)
{
    return back_offset_bytes(n_ghost, comps) + 2 * sizeof(double) * half_doubles(n_send, comps);
}
static size_t ll_back_offset_bytes(uint32_t n_ghost, uint32_t n_send, int comps)
{
    return ll_recv_offset_bytes(n_ghost, n_send, comps) + 2 * sizeof(ulonglong2) * ll_half(n_ghost, comps);
}
static size_t window_bytes(uint32_t n_ghost, uint32_t n_send, int comps)
{
    return ll_back_offset_bytes(n_ghost, n_send, comps) + 2 * sizeof(ulonglong2) * ll_half(n_send, comps);
}

int p2p_max_components() { return kMaxComps; }

void p2p_destroy(Halo &h)
{
    delete h.p2p;
    h.p2p = nullptr;
    h.use_p2p = false;
}

// Collective over the communicator of h.  On any failure on any rank every rank ends with h.p2p == nullptr (NCCL transport).
int p2p_setup(Halo &h, bool raw_mode, NcclAllGatherFn all_gather, NcclAllReduceMinFn all_reduce_min)
{
    const char *env = std::getenv("B200FE_HALO_P2P");
    const bool wanted = !env || std::atoi(env) != 0;
    const int R = h.n_ranks, me = h.rank;
    auto p = std::make_unique<P2P>();
    p->comps = raw_mode ? 1 : kMaxComps;
    Meta mine;
    std::memset(&mine, 0, sizeof(mine));
    mine.n_ghost = h.n_ghost; mine.n_send = raw_mode ? 0u : h.n_send; mine.n_entries = (uint32_t)h.peers.size(); mine.comps = (uint32_t)p->comps;
    bool ok = wanted && R <= kMaxRanks && (int)h.peers.size() <= kMaxEntries;
    if (ok) {
        for (size_t k = 0; k < h.peers.size(); ++k) {
            mine.entry[k][0] = (uint32_t)h.peers[k];
            mine.entry[k][1] = h.recv_off[k]; mine.entry[k][2] = h.recv_cnt[k];
            mine.entry[k][3] = h.send_off[k]; mine.entry[k][4] = h.send_cnt[k];
        }
        const size_t bytes = window_bytes(mine.n_ghost, mine.n_send, p->comps);
        ok = cudaMalloc(&p->window, bytes) == cudaSuccess && cudaMemset(p->window, 0, bytes) == cudaSuccess &&
             cudaIpcGetMemHandle(&mine.handle, p->window) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
    mine.ok = ok ? 1u : 0u;
    // all-gather of the descriptors through NCCL (device staging buffer)
    std::vector<Meta> all(R);
    {
        Meta *d_all = nullptr;
        B200FE_CUDA_TRY(cudaMalloc(&d_all, sizeof(Meta) * R));
        B200FE_CUDA_TRY(cudaMemcpy(d_all + me, &mine, sizeof(Meta), cudaMemcpyHostToDevice));
        int rc = all_gather(h, d_all + me, d_all, sizeof(Meta));
        if (rc == B200FE_OK && cudaMemcpy(all.data(), d_all, sizeof(Meta) * R, cudaMemcpyDeviceToHost) != cudaSuccess) rc = B200FE_ERR_CUDA;
        cudaFree(d_all);
        if (rc != B200FE_OK) return rc;
    }
    for (int r = 0; r < R; ++r) ok = ok && all[r].ok;
    p->peer_base.assign(R, nullptr);
    if (ok) {
        for (int r = 0; r < R && ok; ++r) {
            if (r == me) continue;
            ok = cudaIpcOpenMemHandle(&p->peer_base[r], all[r].handle, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            if (!ok) { p->peer_base[r] = nullptr; cudaGetLastError(); }
        }
    }
    DevState hs;
    std::memset(&hs, 0, sizeof(hs));
    if (ok) {
        auto base_of = [&](int r) { return r == me ? (char *)p->window : (char *)p->peer_base[r]; };
        hs.n_entries = (int)h.peers.size(); hs.n_ranks = R; hs.rank = me;
        hs.recv_half = half_doubles(mine.n_ghost, p->comps); hs.back_half = half_doubles(mine.n_send, p->comps);
        hs.ll_recv_half = ll_half(mine.n_ghost, p->comps); hs.ll_back_half = ll_half(mine.n_send, p->comps);
        // low-latency rounds need every rank below the cap (a sender picks the window the receiver polls)
        p->ll_max_comps = p->comps;
        for (int r = 0; r < R; ++r) {
            const size_t big = std::max<size_t>(std::max(all[r].n_ghost, raw_mode ? h.n_send : all[r].n_send), 1);
            p->ll_max_comps = (int)std::min<size_t>((size_t)p->ll_max_comps, kLLMax / big);
        }
        for (int r = 0; r < R; ++r) hs.peer_red[r] = &reinterpret_cast<Ctrl *>(base_of(r))->red[0][me];
        // k-th message from me to B pairs with B's k-th receive from me, in table order (MPI / NCCL matching order)
        for (int k = 0; k < hs.n_entries && ok; ++k) {
            EntryDev &e = hs.e[k];
            const int B = h.peers[k];
            const Meta &mb = all[B];
            e.send_off = h.send_off[k]; e.send_cnt = h.send_cnt[k]; e.recv_off = h.recv_off[k]; e.recv_cnt = h.recv_cnt[k];
            Ctrl *cb = reinterpret_cast<Ctrl *>(base_of(B));
            if (e.send_cnt) {
                int occ = 0;
                for (int q = 0; q < k; ++q) occ += (h.peers[q] == B && h.send_cnt[q]) ? 1 : 0;
                int kb = -1;
                for (uint32_t q = 0; q < mb.n_entries; ++q)
                    if ((int)mb.entry[q][0] == me && mb.entry[q][2] && occ-- == 0) { kb = (int)q; break; }
                if (kb < 0 || mb.entry[kb][2] != e.send_cnt) { ok = false; break; }
                e.peer_recv = reinterpret_cast<double *>(base_of(B) + recv_offset_bytes()) + mb.entry[kb][1];
                e.peer_upd_flag = &cb->upd_flag[kb];
                e.peer_cmp_ack = &cb->cmp_ack[kb];
                e.peer_n_ghost = mb.n_ghost;
                e.peer_recv_half = half_doubles(mb.n_ghost, (int)mb.comps);
                e.peer_ll_recv = reinterpret_cast<ulonglong2 *>(base_of(B) + ll_recv_offset_bytes(mb.n_ghost, mb.n_send, (int)mb.comps)) + mb.entry[kb][1];
                e.peer_ll_recv_half = ll_half(mb.n_ghost, (int)mb.comps);
            }
            if (e.recv_cnt) {
                int occ = 0;
                for (int q = 0; q < k; ++q) occ += (h.peers[q] == B && h.recv_cnt[q]) ? 1 : 0;
                int jb = -1;
                for (uint32_t q = 0; q < mb.n_entries; ++q)
                    if ((int)mb.entry[q][0] == me && mb.entry[q][4] && occ-- == 0) { jb = (int)q; break; }
                if (jb < 0 || mb.entry[jb][4] != e.recv_cnt) { ok = false; break; }
                e.peer_back = reinterpret_cast<double *>(base_of(B) + back_offset_bytes(mb.n_ghost, (int)mb.comps)) + mb.entry[jb][3];
                e.peer_cmp_flag = &cb->cmp_flag[jb];
                e.peer_upd_ack = &cb->upd_ack[jb];
                e.peer_n_send = mb.n_send;
                e.peer_back_half = half_doubles(mb.n_send, (int)mb.comps);
                e.peer_ll_back = reinterpret_cast<ulonglong2 *>(base_of(B) + ll_back_offset_bytes(mb.n_ghost, mb.n_send, (int)mb.comps)) + mb.entry[jb][3];
                e.peer_ll_back_half = ll_half(mb.n_send, (int)mb.comps);
            }
        }
    }
    if (ok) {
        ok = cudaMalloc(&p->d_state, sizeof(DevState)) == cudaSuccess &&
             cudaMemcpy(p->d_state, &hs, sizeof(DevState), cudaMemcpyHostToDevice) == cudaSuccess;
        if (!ok) cudaGetLastError();
        p->recv = reinterpret_cast<double *>((char *)p->window + recv_offset_bytes());
        p->back = reinterpret_cast<double *>((char *)p->window + back_offset_bytes(mine.n_ghost, p->comps));
        p->ll_recv = reinterpret_cast<ulonglong2 *>((char *)p->window + ll_recv_offset_bytes(mine.n_ghost, mine.n_send, p->comps));
        p->ll_back = reinterpret_cast<ulonglong2 *>((char *)p->window + ll_back_offset_bytes(mine.n_ghost, mine.n_send, p->comps));
    }
    // agreement: P2P only if every rank got every mapping
    int agreed = ok ? 1 : 0;
    if (int rc = all_reduce_min(h, &agreed)) return rc;
    if (agreed) {
        h.p2p = p.release();
        h.use_p2p = true;
    }
    return B200FE_OK;
}

int p2p_update_send(Halo &h, double *v, int ncomp, size_t stride, const double *raw_send, cudaStream_t s)
{
    P2P &p = *h.p2p;
    if (ncomp > p.comps) return fail(B200FE_ERR_UNSUPPORTED, "P2P halo: %d components exceed the window (%d)", ncomp, p.comps);
    p2p_send_kernel<<<blocks_for((size_t)h.n_send * ncomp), 256, 0, s>>>(p.d_state, (Ctrl *)p.window, MODE_UPDATE, v, h.n_owned, h.n_send, h.d_send_idx,
                                                                         ncomp, stride, raw_send);
    B200FE_CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

int p2p_update_wait(Halo &h, double *v, int ncomp, size_t stride, double *raw_recv, cudaStream_t s)
{
    P2P &p = *h.p2p;
    p2p_wait_kernel<<<blocks_for((size_t)h.n_ghost * ncomp), 256, 0, s>>>(p.d_state, (Ctrl *)p.window, MODE_UPDATE, v, h.n_owned, h.n_ghost, h.d_send_idx,
                                                                          ncomp, stride, p.recv, raw_recv);
    B200FE_CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

int p2p_compress_send(Halo &h, double *v, int ncomp, size_t stride, cudaStream_t s)
{
    P2P &p = *h.p2p;
    if (ncomp > p.comps) return fail(B200FE_ERR_UNSUPPORTED, "P2P halo: %d components exceed the window (%d)", ncomp, p.comps);
    p2p_send_kernel<<<blocks_for((size_t)h.n_ghost * ncomp), 256, 0, s>>>(p.d_state, (Ctrl *)p.window, MODE_COMPRESS, v, h.n_owned, h.n_ghost, h.d_send_idx,
                                                                          ncomp, stride, nullptr);
    B200FE_CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

int p2p_compress_wait(Halo &h, double *v, int ncomp, size_t stride, cudaStream_t s)
{
    P2P &p = *h.p2p;
    p2p_wait_kernel<<<blocks_for((size_t)h.n_send * ncomp), 256, 0, s>>>(p.d_state, (Ctrl *)p.window, MODE_COMPRESS, v, h.n_owned, h.n_send, h.d_send_idx,
                                                                         ncomp, stride, p.back, nullptr);
    B200FE_CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

// whole rounds (post + complete back to back): one single-block kernel when the messages are small, else SEND + WAIT
static size_t fused_max()
{
    static const size_t n = [] { const char *e = std::getenv("B200FE_P2P_FUSED_MAX"); return e ? (size_t)std::atoll(e) : (size_t)4096; }();  // r02d: one block moves 32 KB in ~13 us, 256 KB in 35 us
    return n;
}

// a CTA no larger than the message needs (launch and barrier cost of 32 warps shows at 1 KiB)
static unsigned ll_threads(uint32_t n, int ncomp)
{
    const size_t total = (size_t)n * ncomp;
    return total <= 512 ? 128u : total <= 2048 ? 256u : 1024u;
}

static bool ll_enabled()
{
    static const bool on = [] { const char *e = std::getenv("B200FE_P2P_LL"); return !e || std::atoi(e) != 0; }();
    return on;
}

int p2p_update(Halo &h, double *v, int ncomp, size_t stride, const double *raw_send, double *raw_recv, cudaStream_t s)
{
    P2P &p = *h.p2p;
    if (ncomp > p.comps) return fail(B200FE_ERR_UNSUPPORTED, "P2P halo: %d components exceed the window (%d)", ncomp, p.comps);
    if (ncomp <= p.ll_max_comps && ll_enabled()) {
        p2p_round_ll_kernel<<<1, ll_threads(std::max(h.n_send, h.n_ghost), ncomp), 0, s>>>(p.d_state, (Ctrl *)p.window, MODE_UPDATE, v, h.n_owned, h.n_send, h.n_ghost, h.d_send_idx, ncomp, stride,
                                               raw_send, p.ll_recv, raw_recv, 1);
        B200FE_CUDA_TRY(cudaGetLastError());
        return B200FE_OK;
    }
    if ((size_t)h.n_send * ncomp <= fused_max() && (size_t)h.n_ghost * ncomp <= fused_max()) {
        p2p_round_kernel<<<1, 1024, 0, s>>>(p.d_state, (Ctrl *)p.window, MODE_UPDATE, v, h.n_owned, h.n_send, h.n_ghost, h.d_send_idx, ncomp, stride,
                                            raw_send, p.recv, raw_recv);
        B200FE_CUDA_TRY(cudaGetLastError());
        return B200FE_OK;
    }
    if (int rc = p2p_update_send(h, v, ncomp, stride, raw_send, s)) return rc;
    return p2p_update_wait(h, v, ncomp, stride, raw_recv, s);
}

// n_rounds raw exchange rounds in one launch when the low-latency round applies, else launch by launch
int p2p_update_rounds(Halo &h, const double *raw_send, double *raw_recv, int n_rounds, cudaStream_t s)
{
    P2P &p = *h.p2p;
    if (1 <= p.ll_max_comps && ll_enabled()) {
        p2p_round_ll_kernel<<<1, ll_threads(std::max(h.n_send, h.n_ghost), 1), 0, s>>>(p.d_state, (Ctrl *)p.window, MODE_UPDATE, nullptr, h.n_owned, h.n_send, h.n_ghost,
                                                                                      h.d_send_idx, 1, 0, raw_send, p.ll_recv, raw_recv, n_rounds);
        B200FE_CUDA_TRY(cudaGetLastError());
        return B200FE_OK;
    }
    for (int r = 0; r < n_rounds; ++r)
        if (int rc = p2p_update(h, nullptr, 1, 0, raw_send, raw_recv, s)) return rc;
    return B200FE_OK;
}

int p2p_compress(Halo &h, double *v, int ncomp, size_t stride, cudaStream_t s)
{
    P2P &p = *h.p2p;
    if (ncomp > p.comps) return fail(B200FE_ERR_UNSUPPORTED, "P2P halo: %d components exceed the window (%d)", ncomp, p.comps);
    if (ncomp <= p.ll_max_comps && ll_enabled()) {
        p2p_round_ll_kernel<<<1, ll_threads(std::max(h.n_send, h.n_ghost), ncomp), 0, s>>>(p.d_state, (Ctrl *)p.window, MODE_COMPRESS, v, h.n_owned, h.n_ghost, h.n_send, h.d_send_idx, ncomp, stride,
                                               nullptr, p.ll_back, nullptr, 1);
        B200FE_CUDA_TRY(cudaGetLastError());
        return B200FE_OK;
    }
    if ((size_t)h.n_send * ncomp <= fused_max() && (size_t)h.n_ghost * ncomp <= fused_max()) {
        p2p_round_kernel<<<1, 1024, 0, s>>>(p.d_state, (Ctrl *)p.window, MODE_COMPRESS, v, h.n_owned, h.n_ghost, h.n_send, h.d_send_idx, ncomp, stride,
                                            nullptr, p.back, nullptr);
        B200FE_CUDA_TRY(cudaGetLastError());
        return B200FE_OK;
    }
    if (int rc = p2p_compress_send(h, v, ncomp, stride, s)) return rc;
    return p2p_compress_wait(h, v, ncomp, stride, s);
}

int p2p_allreduce(Halo &h, double *d_vals, int count, cudaStream_t s)
{
    if (count > 4) return fail(B200FE_ERR_UNSUPPORTED, "P2P all-reduce: at most 4 values per call");
    P2P &p = *h.p2p;
    p2p_allreduce_kernel<<<1, 32, 0, s>>>(p.d_state, (Ctrl *)p.window, d_vals, count);
    B200FE_CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

// Synchronises the device: 0 = healthy, B200FE_ERR_COMM after a wait has timed out.
int p2p_status(Halo &h)
{
    if (!h.p2p) return B200FE_OK;
    int err = 0;
    B200FE_CUDA_TRY(cudaMemcpy(&err, &h.p2p->d_state->error, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) return fail(B200FE_ERR_COMM, "P2P halo: a peer did not answer within %llu s (rank %d of %d)", kTimeoutNs / 1000000000ull, h.rank, h.n_ranks);
    return B200FE_OK;
}

}  // namespace b200fe
