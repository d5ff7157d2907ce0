// halo.h -- ghost-DoF exchange between the ranks of one NVSwitch domain (NCCL send/recv).
//
// Replaces deal.II's LinearAlgebra::distributed::Vector::update_ghost_values / compress(add) /
// zero_out_ghost_values as used by CEED_bp/include/portable_laplace_operator.h:133,169,170 and the
// split-phase variants of bakeoff_problems_dealii/include/portable_laplace_operator.h:669-695,
// and the raw MPI_Isend/Irecv exchange of p-halox/phalox.cc:104-126.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

namespace b200fe {

struct P2P;  // halo.cu: CUDA-IPC windows of the NVSwitch domain (direct peer stores + flags), optional

struct Halo {
    P2P *p2p = nullptr;    // owned; non-null once every rank has mapped every peer's window (b200fe_halo_create)
    bool use_p2p = false;  // transport of the exchanges and of the scalar all-reduce: peer stores (true) or NCCL (false)
    void *comm = nullptr;  // ncclComm_t (owned when created from a unique id)
    bool owns_comm = false;
    int rank = 0, n_ranks = 1;
    uint32_t n_owned = 0, n_ghost = 0;
    std::vector<int> peers;
    std::vector<uint32_t> recv_off, recv_cnt;  // slices of the ghost segment (offset relative to n_owned)
    std::vector<uint32_t> send_off, send_cnt;  // slices of the packed send list
    uint32_t n_send = 0;
    uint32_t *d_send_idx = nullptr;  // owned local indices to pack, grouped by peer
    double *d_pack = nullptr;        // [n_send] packed owner values (update) / incoming contributions (compress)
    double *d_pack_multi = nullptr;  // [components][n_send], allocated on first use by the *_components calls
    int pack_multi_comps = 0;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
    double *pending_v = nullptr;  // P2P split-phase update: the vector whose ghost segment *_finish fills
    bool d_send_idx_ok() const { return n_send == 0 || d_send_idx != nullptr; }
    Halo() = default;
    Halo(const Halo &) = delete;
    Halo &operator=(const Halo &) = delete;
    ~Halo();  // halo.cu: frees the communicator (if owned), pack lists, stream and events
};

// ---- halo_p2p.cu: one-sided transport over CUDA-IPC peer memory (see the file header) --------------------------------
using NcclAllGatherFn = int (*)(Halo &h, const void *d_send, void *d_recv, size_t bytes_per_rank);
using NcclAllReduceMinFn = int (*)(Halo &h, int *host_value);
// collective; on success on EVERY rank h.p2p is set and h.use_p2p = true, otherwise the halo keeps the NCCL transport
int p2p_setup(Halo &h, bool raw_mode, NcclAllGatherFn all_gather, NcclAllReduceMinFn all_reduce_min);
void p2p_destroy(Halo &h);
// post (stores into the peers' windows + flags) and complete (wait for the peers' flags, move into place, acknowledge);
// both on stream s, nothing else needed in between: "post early, wait late" is the overlap schedule
int p2p_update_send(Halo &h, double *d_v, int ncomp, size_t stride, const double *d_raw_send, cudaStream_t s);
int p2p_update_wait(Halo &h, double *d_v, int ncomp, size_t stride, double *d_raw_recv, cudaStream_t s);
int p2p_compress_send(Halo &h, double *d_v, int ncomp, size_t stride, cudaStream_t s);  // also zeroes the ghost entries of v
int p2p_compress_wait(Halo &h, double *d_v, int ncomp, size_t stride, cudaStream_t s);
// whole rounds (post + complete back to back; one single-block kernel for small messages)
int p2p_update(Halo &h, double *d_v, int ncomp, size_t stride, const double *d_raw_send, double *d_raw_recv, cudaStream_t s);
int p2p_update_rounds(Halo &h, const double *d_raw_send, double *d_raw_recv, int n_rounds, cudaStream_t s);
int p2p_compress(Halo &h, double *d_v, int ncomp, size_t stride, cudaStream_t s);
int p2p_allreduce(Halo &h, double *d_vals, int count, cudaStream_t s);
int p2p_status(Halo &h);
int p2p_max_components();  // synchronising health check: B200FE_ERR_COMM after a bounded wait has expired

// owner -> ghost copy of v (all on stream s)
int halo_update_ghosts(Halo &h, double *d_v, cudaStream_t s);
// ghost -> owner additive reduction of v, ghosts zeroed afterwards
int halo_compress_add(Halo &h, double *d_v, cudaStream_t s);
// the same for `ncomp` component-blocked vectors (`stride` apart) in ONE pack kernel and ONE NCCL group per direction
// (vector-valued problems: three exchanges per apply would triple the launch / group latency, which is what a
// strong-scaled halo costs); B200FE_HALO_BATCH=0 falls back to one exchange per component
int halo_update_ghosts_components(Halo &h, double *d_v, int ncomp, size_t stride, cudaStream_t s);
int halo_compress_add_components(Halo &h, double *d_v, int ncomp, size_t stride, cudaStream_t s);
// split-phase versions for the 3-phase overlap schedule: work is issued on h.comm_stream after
// everything already queued on s; *_finish makes s wait for it.
int halo_update_ghosts_start(Halo &h, double *d_v, cudaStream_t s);
int halo_update_ghosts_finish(Halo &h, cudaStream_t s);
int halo_compress_start(Halo &h, double *d_v, cudaStream_t s);
int halo_compress_finish(Halo &h, double *d_v, cudaStream_t s);
int halo_zero_ghosts(Halo &h, double *d_v, cudaStream_t s);
// sum over ranks of n doubles in place (CG inner products)
int halo_allreduce_sum(Halo &h, double *d_vals, int n, cudaStream_t s);

}  // namespace b200fe
