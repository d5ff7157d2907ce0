// halo.cu -- C ABI section 6: ghost-DoF exchange and scalar all-reduce over NCCL (NVLink 5 / NVSwitch).
//
// One process per GPU.  NCCL is bound lazily with dlopen("libnccl.so.2") so that the library (and
// its symbol table) loads on machines without NCCL; inside a PyTorch process this resolves to the
// NCCL build torch already loaded.  Exchange pattern = deal.II Partitioner:
//   update_ghosts : owners pack import_indices, peers receive straight into their ghost segment
//   compress_add  : ghost segments travel back, owners add them into import_indices, ghosts zeroed
// which is also p-halox's Irecv/Isend/Waitall round (p-halox/phalox.cc:104-126) with NCCL grouping.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <memory>

#include "common.h"
#include "halo.h"

namespace b200fe {

namespace {

struct Nccl {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    bool ok = false;
};

Nccl &nccl()
{
    static Nccl n = [] {
        Nccl x;
        x.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!x.h) x.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!x.h) return x;
        auto sym = [&](const char *name) { return dlsym(x.h, name); };
        x.GetUniqueId = (decltype(x.GetUniqueId))sym("ncclGetUniqueId");
        x.CommInitRank = (decltype(x.CommInitRank))sym("ncclCommInitRank");
        x.CommDestroy = (decltype(x.CommDestroy))sym("ncclCommDestroy");
        x.Send = (decltype(x.Send))sym("ncclSend");
        x.Recv = (decltype(x.Recv))sym("ncclRecv");
        x.AllReduce = (decltype(x.AllReduce))sym("ncclAllReduce");
        x.AllGather = (decltype(x.AllGather))sym("ncclAllGather");
        x.GroupStart = (decltype(x.GroupStart))sym("ncclGroupStart");
        x.GroupEnd = (decltype(x.GroupEnd))sym("ncclGroupEnd");
        x.GetErrorString = (decltype(x.GetErrorString))sym("ncclGetErrorString");
        x.GetVersion = (decltype(x.GetVersion))sym("ncclGetVersion");
        x.ok = x.GetUniqueId && x.CommInitRank && x.CommDestroy && x.Send && x.Recv && x.AllReduce && x.AllGather && x.GroupStart &&
               x.GroupEnd && x.GetErrorString;
        return x;
    }();
    return n;
}

int fail_nccl(ncclResult_t r, const char *what)
{
    return fail(B200FE_ERR_COMM, "NCCL error %d (%s) in %s", (int)r, nccl().GetErrorString ? nccl().GetErrorString(r) : "?", what);
}

#define B200FE_NCCL_TRY(expr)                                  \
    do {                                                       \
        ncclResult_t _r = (expr);                              \
        if (_r != ncclSuccess) return fail_nccl(_r, #expr);    \
    } while (0)

__global__ void pack_kernel(uint32_t n, const uint32_t *__restrict__ idx, const double *__restrict__ v,
                            double *__restrict__ buf)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) buf[i] = v[idx[i]];
}

// the same owned DoF may be ghosted by several peers -> atomic accumulate
__global__ void unpack_add_kernel(uint32_t n, const uint32_t *__restrict__ idx, const double *__restrict__ buf,
                                  double *__restrict__ v)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(v + idx[i], buf[i]);
}

// component-blocked versions: buf[c][i], v + c*stride
__global__ void pack_components_kernel(uint32_t n, int ncomp, size_t stride, const uint32_t *__restrict__ idx, const double *__restrict__ v,
                                       double *__restrict__ buf)
{
    const size_t total = (size_t)n * ncomp;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t c = i / n;
        buf[i] = v[c * stride + idx[i - c * n]];
    }
}

__global__ void unpack_add_components_kernel(uint32_t n, int ncomp, size_t stride, const uint32_t *__restrict__ idx,
                                             const double *__restrict__ buf, double *__restrict__ v)
{
    const size_t total = (size_t)n * ncomp;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t c = i / n;
        atomicAdd(v + c * stride + idx[i - c * n], buf[i]);
    }
}

inline unsigned blocks_for(uint32_t n) { return n == 0 ? 1u : std::min<unsigned>((n + 255) / 256, 148u * 8u); }

int exchange_update(Halo &h, double *v, cudaStream_t s)
{
    if (h.use_p2p) {
        return p2p_update(h, v, 1, 0, nullptr, nullptr, s);
    }
    Nccl &n = nccl();
    if (h.n_send) {
        pack_kernel<<<blocks_for(h.n_send), 256, 0, s>>>(h.n_send, h.d_send_idx, v, h.d_pack);
        B200FE_CUDA_TRY(cudaGetLastError());
    }
    B200FE_NCCL_TRY(n.GroupStart());
    for (size_t k = 0; k < h.peers.size(); ++k) {
        if (h.recv_cnt[k]) B200FE_NCCL_TRY(n.Recv(v + h.n_owned + h.recv_off[k], h.recv_cnt[k], ncclDouble, h.peers[k], (ncclComm_t)h.comm, s));
        if (h.send_cnt[k]) B200FE_NCCL_TRY(n.Send(h.d_pack + h.send_off[k], h.send_cnt[k], ncclDouble, h.peers[k], (ncclComm_t)h.comm, s));
    }
    B200FE_NCCL_TRY(n.GroupEnd());
    return B200FE_OK;
}

int exchange_compress(Halo &h, double *v, cudaStream_t s)
{
    Nccl &n = nccl();
    B200FE_NCCL_TRY(n.GroupStart());
    for (size_t k = 0; k < h.peers.size(); ++k) {
        if (h.send_cnt[k]) B200FE_NCCL_TRY(n.Recv(h.d_pack + h.send_off[k], h.send_cnt[k], ncclDouble, h.peers[k], (ncclComm_t)h.comm, s));
        if (h.recv_cnt[k]) B200FE_NCCL_TRY(n.Send(v + h.n_owned + h.recv_off[k], h.recv_cnt[k], ncclDouble, h.peers[k], (ncclComm_t)h.comm, s));
    }
    B200FE_NCCL_TRY(n.GroupEnd());
    return B200FE_OK;
}

int unpack_and_zero(Halo &h, double *v, cudaStream_t s)
{
    if (h.n_send) {
        unpack_add_kernel<<<blocks_for(h.n_send), 256, 0, s>>>(h.n_send, h.d_send_idx, h.d_pack, v);
        B200FE_CUDA_TRY(cudaGetLastError());
    }
    return halo_zero_ghosts(h, v, s);
}

// setup-time collectives handed to p2p_setup (halo_p2p.cu does not bind NCCL itself)
int setup_all_gather(Halo &h, const void *d_send, void *d_recv, size_t bytes)
{
    B200FE_NCCL_TRY(nccl().AllGather(d_send, d_recv, bytes, ncclChar, (ncclComm_t)h.comm, nullptr));
    B200FE_CUDA_TRY(cudaStreamSynchronize(nullptr));
    return B200FE_OK;
}
int setup_all_reduce_min(Halo &h, int *value)
{
    int *d = nullptr;
    B200FE_CUDA_TRY(cudaMalloc(&d, sizeof(int)));
    cudaError_t e = cudaMemcpy(d, value, sizeof(int), cudaMemcpyHostToDevice);
    ncclResult_t r = e == cudaSuccess ? nccl().AllReduce(d, d, 1, ncclInt, ncclMin, (ncclComm_t)h.comm, nullptr) : ncclSuccess;
    if (e == cudaSuccess && r == ncclSuccess) e = cudaMemcpy(value, d, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (r != ncclSuccess) return fail_nccl(r, "ncclAllReduce(min)");
    if (e != cudaSuccess) return fail_cuda(e, "setup_all_reduce_min");
    return B200FE_OK;
}

}  // namespace

Halo::~Halo()
{
    p2p_destroy(*this);
    if (comm && owns_comm) nccl().CommDestroy((ncclComm_t)comm);
    cudaFree(d_send_idx);
    cudaFree(d_pack);
    cudaFree(d_pack_multi);
    if (comm_stream) cudaStreamDestroy(comm_stream);
    if (ev_ready) cudaEventDestroy(ev_ready);
    if (ev_done) cudaEventDestroy(ev_done);
}

int halo_zero_ghosts(Halo &h, double *v, cudaStream_t s)
{
    if (h.n_ghost) B200FE_CUDA_TRY(cudaMemsetAsync(v + h.n_owned, 0, sizeof(double) * h.n_ghost, s));
    return B200FE_OK;
}

int halo_update_ghosts(Halo &h, double *v, cudaStream_t s)
{
    if (h.n_ranks == 1) return B200FE_OK;
    NvtxRange range("update_ghost_values");
    if (h.n_send && !h.d_send_idx) return fail(B200FE_ERR_INVALID_ARG, "halo was created in raw mode (no send indices)");
    return exchange_update(h, v, s);
}

int halo_compress_add(Halo &h, double *v, cudaStream_t s)
{
    if (h.n_ranks == 1) return B200FE_OK;
    NvtxRange range("compress_add");
    if (h.n_send && !h.d_send_idx) return fail(B200FE_ERR_INVALID_ARG, "halo was created in raw mode (no send indices)");
    if (h.use_p2p) {
        return p2p_compress(h, v, 1, 0, s);
    }
    if (int rc = exchange_compress(h, v, s)) return rc;
    return unpack_and_zero(h, v, s);
}

namespace {
bool halo_batching()
{
    // NCCL transport only (the P2P transport always moves all components in one kernel).  On by default since the
    // 2-GPU check of round 2 (profiles/r02b_dist_check_2gpu.log: batched == per-component, bitwise); B200FE_HALO_BATCH=0
    // restores one exchange per component.
    const char *e = std::getenv("B200FE_HALO_BATCH");  // read per call: tools/dist_check.py flips it to compare the two forms
    return !e || std::atoi(e) != 0;
}
int ensure_pack_multi(Halo &h, int ncomp)
{
    if (ncomp <= h.pack_multi_comps) return B200FE_OK;
    cudaFree(h.d_pack_multi);
    h.d_pack_multi = nullptr;
    h.pack_multi_comps = 0;
    B200FE_CUDA_TRY(cudaMalloc(&h.d_pack_multi, std::max<size_t>((size_t)h.n_send * ncomp, 1) * sizeof(double)));
    h.pack_multi_comps = ncomp;
    return B200FE_OK;
}
}  // namespace

int halo_update_ghosts_components(Halo &h, double *v, int ncomp, size_t stride, cudaStream_t s)
{
    if (h.n_ranks == 1) return B200FE_OK;
    if (h.use_p2p && h.d_send_idx_ok() && ncomp <= p2p_max_components()) {
        NvtxRange range("update_ghost_values");
        return p2p_update(h, v, ncomp, stride, nullptr, nullptr, s);
    }
    if (ncomp == 1 || !halo_batching() || (h.n_send && !h.d_send_idx)) {
        for (int c = 0; c < ncomp; ++c)
            if (int rc = halo_update_ghosts(h, v + c * stride, s)) return rc;
        return B200FE_OK;
    }
    NvtxRange range("update_ghost_values");
    if (int rc = ensure_pack_multi(h, ncomp)) return rc;
    Nccl &n = nccl();
    if (h.n_send) {
        pack_components_kernel<<<blocks_for(h.n_send * (uint32_t)ncomp), 256, 0, s>>>(h.n_send, ncomp, stride, h.d_send_idx, v, h.d_pack_multi);
        B200FE_CUDA_TRY(cudaGetLastError());
    }
    B200FE_NCCL_TRY(n.GroupStart());
    for (int c = 0; c < ncomp; ++c)
        for (size_t k = 0; k < h.peers.size(); ++k) {
            if (h.recv_cnt[k]) B200FE_NCCL_TRY(n.Recv(v + c * stride + h.n_owned + h.recv_off[k], h.recv_cnt[k], ncclDouble, h.peers[k], (ncclComm_t)h.comm, s));
            if (h.send_cnt[k]) B200FE_NCCL_TRY(n.Send(h.d_pack_multi + (size_t)c * h.n_send + h.send_off[k], h.send_cnt[k], ncclDouble, h.peers[k], (ncclComm_t)h.comm, s));
        }
    B200FE_NCCL_TRY(n.GroupEnd());
    return B200FE_OK;
}

int halo_compress_add_components(Halo &h, double *v, int ncomp, size_t stride, cudaStream_t s)
{
    if (h.n_ranks == 1) return B200FE_OK;
    if (h.use_p2p && h.d_send_idx_ok() && ncomp <= p2p_max_components()) {
        NvtxRange range("compress_add");
        return p2p_compress(h, v, ncomp, stride, s);
    }
    if (ncomp == 1 || !halo_batching() || (h.n_send && !h.d_send_idx)) {
        for (int c = 0; c < ncomp; ++c)
            if (int rc = halo_compress_add(h, v + c * stride, s)) return rc;
        return B200FE_OK;
    }
    NvtxRange range("compress_add");
    if (int rc = ensure_pack_multi(h, ncomp)) return rc;
    Nccl &n = nccl();
    B200FE_NCCL_TRY(n.GroupStart());
    for (int c = 0; c < ncomp; ++c)
        for (size_t k = 0; k < h.peers.size(); ++k) {
            if (h.send_cnt[k]) B200FE_NCCL_TRY(n.Recv(h.d_pack_multi + (size_t)c * h.n_send + h.send_off[k], h.send_cnt[k], ncclDouble, h.peers[k], (ncclComm_t)h.comm, s));
            if (h.recv_cnt[k]) B200FE_NCCL_TRY(n.Send(v + c * stride + h.n_owned + h.recv_off[k], h.recv_cnt[k], ncclDouble, h.peers[k], (ncclComm_t)h.comm, s));
        }
    B200FE_NCCL_TRY(n.GroupEnd());
    if (h.n_send) {
        unpack_add_components_kernel<<<blocks_for(h.n_send * (uint32_t)ncomp), 256, 0, s>>>(h.n_send, ncomp, stride, h.d_send_idx, h.d_pack_multi, v);
        B200FE_CUDA_TRY(cudaGetLastError());
    }
    for (int c = 0; c < ncomp; ++c)
        if (int rc = halo_zero_ghosts(h, v + c * stride, s)) return rc;
    return B200FE_OK;
}

int halo_update_ghosts_start(Halo &h, double *v, cudaStream_t s)
{
    if (h.n_ranks == 1) return B200FE_OK;
    if (h.use_p2p) {  // one-sided: post now on s, complete in *_finish on s (v is kept for the wait)
        h.pending_v = v;
        return p2p_update_send(h, v, 1, 0, nullptr, s);
    }
    B200FE_CUDA_TRY(cudaEventRecord(h.ev_ready, s));
    B200FE_CUDA_TRY(cudaStreamWaitEvent(h.comm_stream, h.ev_ready, 0));
    if (int rc = exchange_update(h, v, h.comm_stream)) return rc;
    B200FE_CUDA_TRY(cudaEventRecord(h.ev_done, h.comm_stream));
    return B200FE_OK;
}

int halo_update_ghosts_finish(Halo &h, cudaStream_t s)
{
    if (h.n_ranks == 1) return B200FE_OK;
    if (h.use_p2p) return p2p_update_wait(h, h.pending_v, 1, 0, nullptr, s);
    B200FE_CUDA_TRY(cudaStreamWaitEvent(s, h.ev_done, 0));
    return B200FE_OK;
}

int halo_compress_start(Halo &h, double *v, cudaStream_t s)
{
    if (h.n_ranks == 1) return B200FE_OK;
    if (h.use_p2p) return p2p_compress_send(h, v, 1, 0, s);
    B200FE_CUDA_TRY(cudaEventRecord(h.ev_ready, s));
    B200FE_CUDA_TRY(cudaStreamWaitEvent(h.comm_stream, h.ev_ready, 0));
    if (int rc = exchange_compress(h, v, h.comm_stream)) return rc;
    B200FE_CUDA_TRY(cudaEventRecord(h.ev_done, h.comm_stream));
    return B200FE_OK;
}

int halo_compress_finish(Halo &h, double *v, cudaStream_t s)
{
    if (h.n_ranks == 1) return B200FE_OK;
    if (h.use_p2p) return p2p_compress_wait(h, v, 1, 0, s);
    B200FE_CUDA_TRY(cudaStreamWaitEvent(s, h.ev_done, 0));
    return unpack_and_zero(h, v, s);
}

int halo_allreduce_sum(Halo &h, double *d_vals, int count, cudaStream_t s)
{
    if (h.n_ranks == 1) return B200FE_OK;
    if (h.use_p2p && count <= 4) return p2p_allreduce(h, d_vals, count, s);
    B200FE_NCCL_TRY(nccl().AllReduce(d_vals, d_vals, count, ncclDouble, ncclSum, (ncclComm_t)h.comm, s));
    return B200FE_OK;
}

}  // namespace b200fe

using namespace b200fe;

extern "C" {

int b200fe_comm_available(void) { return nccl().ok ? 1 : 0; }

int b200fe_comm_unique_id(char *id128)
{
    B200FE_REQUIRE(id128, "b200fe_comm_unique_id: null pointer");
    if (!nccl().ok) return fail(B200FE_ERR_COMM, "libnccl.so.2 not found");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    B200FE_NCCL_TRY(nccl().GetUniqueId(&id));
    std::memcpy(id128, &id, 128);
    return B200FE_OK;
}

int b200fe_halo_create(const b200fe_halo_desc *d, b200fe_halo **out)
{
    B200FE_REQUIRE(d && out, "b200fe_halo_create: null pointer");
    B200FE_REQUIRE(d->n_ranks >= 1 && d->rank >= 0 && d->rank < d->n_ranks, "b200fe_halo_create: bad rank %d of %d", d->rank, d->n_ranks);
    B200FE_REQUIRE(d->n_peers >= 0 && (d->n_peers == 0 || (d->peers && d->recv_offset && d->recv_count && d->send_offset && d->send_count)),
                   "b200fe_halo_create: peer tables missing");
    auto h = std::make_unique<Halo>();
    h->rank = d->rank; h->n_ranks = d->n_ranks; h->n_owned = d->n_owned; h->n_ghost = d->n_ghost; h->n_send = d->n_send;
    for (int k = 0; k < d->n_peers; ++k) {
        B200FE_REQUIRE(d->peers[k] >= 0 && d->peers[k] < d->n_ranks && d->peers[k] != d->rank, "b200fe_halo_create: bad peer %d", d->peers[k]);
        B200FE_REQUIRE((uint64_t)d->recv_offset[k] + d->recv_count[k] <= d->n_ghost, "b200fe_halo_create: recv slice of peer %d exceeds the ghost segment", d->peers[k]);
        B200FE_REQUIRE((uint64_t)d->send_offset[k] + d->send_count[k] <= d->n_send, "b200fe_halo_create: send slice of peer %d exceeds the send list", d->peers[k]);
        h->peers.push_back(d->peers[k]);
        h->recv_off.push_back(d->recv_offset[k]); h->recv_cnt.push_back(d->recv_count[k]);
        h->send_off.push_back(d->send_offset[k]); h->send_cnt.push_back(d->send_count[k]);
    }
    // h_send_indices == NULL with n_send > 0: raw mode (b200fe_halo_exchange_raw only, no pack list)
    if (d->h_send_indices)
        for (uint32_t i = 0; i < d->n_send; ++i)
            B200FE_REQUIRE(d->h_send_indices[i] < d->n_owned, "b200fe_halo_create: send index %u is not an owned DoF", d->h_send_indices[i]);
    if (d->n_ranks > 1) {
        B200FE_REQUIRE(d->nccl_unique_id, "b200fe_halo_create: nccl_unique_id missing");
        if (!nccl().ok) return fail(B200FE_ERR_COMM, "libnccl.so.2 not found");
        ncclUniqueId id;
        std::memcpy(&id, d->nccl_unique_id, 128);
        ncclComm_t comm;
        B200FE_NCCL_TRY(nccl().CommInitRank(&comm, d->n_ranks, id, d->rank));
        h->comm = comm; h->owns_comm = true;
        if (d->n_send && d->h_send_indices) {
            B200FE_CUDA_TRY(cudaMalloc(&h->d_send_idx, d->n_send * sizeof(uint32_t)));
            B200FE_CUDA_TRY(cudaMemcpy(h->d_send_idx, d->h_send_indices, d->n_send * sizeof(uint32_t), cudaMemcpyHostToDevice));
            B200FE_CUDA_TRY(cudaMalloc(&h->d_pack, d->n_send * sizeof(double)));
        }
        B200FE_CUDA_TRY(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
        B200FE_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
        B200FE_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
        // one-sided transport over CUDA-IPC windows where every rank can map every peer (one NVSwitch domain); collective
        const bool raw_mode = d->n_send > 0 && d->h_send_indices == nullptr;
        if (int rc = p2p_setup(*h, raw_mode, setup_all_gather, setup_all_reduce_min)) return rc;
    }
    *out = reinterpret_cast<b200fe_halo *>(h.release());
    return B200FE_OK;
}

void b200fe_halo_destroy(b200fe_halo *halo) { delete reinterpret_cast<Halo *>(halo); }

int b200fe_halo_update_ghosts(b200fe_halo *halo, double *d_v, void *stream)
{
    B200FE_REQUIRE(halo && d_v, "b200fe_halo_update_ghosts: null pointer");
    return halo_update_ghosts(*reinterpret_cast<Halo *>(halo), d_v, (cudaStream_t)stream);
}

int b200fe_halo_compress_add(b200fe_halo *halo, double *d_v, void *stream)
{
    B200FE_REQUIRE(halo && d_v, "b200fe_halo_compress_add: null pointer");
    return halo_compress_add(*reinterpret_cast<Halo *>(halo), d_v, (cudaStream_t)stream);
}

int b200fe_halo_zero_ghosts(b200fe_halo *halo, double *d_v, void *stream)
{
    B200FE_REQUIRE(halo && d_v, "b200fe_halo_zero_ghosts: null pointer");
    return halo_zero_ghosts(*reinterpret_cast<Halo *>(halo), d_v, (cudaStream_t)stream);
}

int b200fe_halo_exchange_raw(b200fe_halo *halo, const double *d_send, double *d_recv, void *stream)
{
    B200FE_REQUIRE(halo && d_send && d_recv, "b200fe_halo_exchange_raw: null pointer");
    Halo &h = *reinterpret_cast<Halo *>(halo);
    if (h.n_ranks == 1) return B200FE_OK;
    cudaStream_t s = (cudaStream_t)stream;
    // peer stores into the receivers' windows + flags, then wait + copy out (halo_p2p.cu).  Beyond ~1 GiB per rank and round
    // the staged copy loses to NCCL's direct receive (r02h, 8 GPUs, 512 MiB messages: 2.7-3.4 vs 4.4-4.8 TB/s aggregate;
    // at 64 MiB and below P2P is 1.2-2.6x faster): those rounds stay on NCCL.
    if (h.use_p2p && (size_t)h.n_ghost * sizeof(double) < (size_t)1 << 30) return p2p_update(h, nullptr, 1, 0, d_send, d_recv, s);
    Nccl &n = nccl();
    // one round of p-halox: post all receives, all sends, complete together (phalox.cc:111-125)
    B200FE_NCCL_TRY(n.GroupStart());
    for (size_t k = 0; k < h.peers.size(); ++k)
        if (h.recv_cnt[k]) B200FE_NCCL_TRY(n.Recv(d_recv + h.recv_off[k], h.recv_cnt[k], ncclDouble, h.peers[k], (ncclComm_t)h.comm, s));
    for (size_t k = 0; k < h.peers.size(); ++k)
        if (h.send_cnt[k]) B200FE_NCCL_TRY(n.Send(d_send + h.send_off[k], h.send_cnt[k], ncclDouble, h.peers[k], (ncclComm_t)h.comm, s));
    B200FE_NCCL_TRY(n.GroupEnd());
    return B200FE_OK;
}

int b200fe_halo_exchange_raw_rounds(b200fe_halo *halo, const double *d_send, double *d_recv, int n_rounds, void *stream)
{
    B200FE_REQUIRE(halo && d_send && d_recv && n_rounds >= 1, "b200fe_halo_exchange_raw_rounds: bad arguments");
    Halo &h = *reinterpret_cast<Halo *>(halo);
    if (h.n_ranks == 1) return B200FE_OK;
    if (h.use_p2p && (size_t)h.n_ghost * sizeof(double) < (size_t)1 << 30) return p2p_update_rounds(h, d_send, d_recv, n_rounds, (cudaStream_t)stream);
    for (int r = 0; r < n_rounds; ++r)
        if (int rc = b200fe_halo_exchange_raw(halo, d_send, d_recv, stream)) return rc;
    return B200FE_OK;
}

int b200fe_halo_transport(b200fe_halo *halo, int *p2p_available, int *p2p_in_use)
{
    B200FE_REQUIRE(halo, "b200fe_halo_transport: null pointer");
    Halo &h = *reinterpret_cast<Halo *>(halo);
    if (p2p_available) *p2p_available = h.p2p != nullptr;
    if (p2p_in_use) *p2p_in_use = h.use_p2p;
    return B200FE_OK;
}

int b200fe_halo_set_transport(b200fe_halo *halo, int use_p2p)
{
    B200FE_REQUIRE(halo, "b200fe_halo_set_transport: null pointer");
    Halo &h = *reinterpret_cast<Halo *>(halo);
    if (use_p2p && !h.p2p) return fail(B200FE_ERR_UNSUPPORTED, "b200fe_halo_set_transport: no peer windows on this halo (B200FE_HALO_P2P=0, or a peer could not be mapped)");
    h.use_p2p = use_p2p != 0;
    return B200FE_OK;
}

int b200fe_halo_status(b200fe_halo *halo)
{
    B200FE_REQUIRE(halo, "b200fe_halo_status: null pointer");
    return p2p_status(*reinterpret_cast<Halo *>(halo));
}

int b200fe_halo_allreduce_sum(b200fe_halo *halo, double *d_vals, int count, void *stream)
{
    B200FE_REQUIRE(halo && d_vals && count >= 0, "b200fe_halo_allreduce_sum: bad arguments");
    return halo_allreduce_sum(*reinterpret_cast<Halo *>(halo), d_vals, count, (cudaStream_t)stream);
}

}  // extern "C"
