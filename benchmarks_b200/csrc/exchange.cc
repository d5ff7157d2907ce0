// exchange.cc -- ghost-exchange lists of one rank, computed WITHOUT communication.
//
// deal.II's Utilities::MPI::Partitioner finds the import (send) lists with a round of point-to-point messages
// ("which of my DoFs do you ghost?").  The meshes of this library are closed-form / replayable on
// every rank, so a rank can simply build the view of every other rank and read off who ghosts which of its DoFs:
// b200fe_halo_create gets its peer tables from here with no MPI, gloo or file rendezvous (only the NCCL id has to travel).
//   ghost segment: sorted by global index => grouped by owner rank; recv slice of peer t = its run in that segment
//   send list of peer t = my owned local indices that t ghosts, in t's ghost order (= ascending global index)
#include <algorithm>
#include <memory>
#include <vector>

#include "common.h"
#include "mesh_common.h"

namespace b200fe {

struct Exchange {
    uint32_t n_owned = 0, n_ghost = 0;
    std::vector<int32_t> peers;
    std::vector<uint32_t> recv_off, recv_cnt, send_off, send_cnt, send_idx;
};

namespace {

struct RankView {
    uint64_t owned_begin = 0;
    uint32_t n_owned = 0;
    std::vector<uint64_t> ghost_global;
    std::vector<int32_t> ghost_owner;
};

// view(t, &v): fills the partition view of rank t, returns a status code
template <class ViewFn>
int build_exchange(int n_ranks, int rank, ViewFn view, Exchange &ex)
{
    RankView mine;
    if (int rc = view(rank, mine)) return rc;
    ex.n_owned = mine.n_owned;
    ex.n_ghost = (uint32_t)mine.ghost_global.size();
    std::vector<std::vector<uint32_t>> send(n_ranks);
    std::vector<uint32_t> r_off(n_ranks, 0), r_cnt(n_ranks, 0);
    for (size_t k = 0; k < mine.ghost_owner.size(); ++k) {
        const int t = mine.ghost_owner[k];
        if (t < 0 || t >= n_ranks || t == rank) return fail(B200FE_ERR_INVALID_ARG, "exchange lists: ghost %zu has owner %d", k, t);
        if (r_cnt[t] == 0) r_off[t] = (uint32_t)k;
        else if (r_off[t] + r_cnt[t] != k) return fail(B200FE_ERR_INVALID_ARG, "exchange lists: ghosts of rank %d are not contiguous", t);
        ++r_cnt[t];
    }
    for (int t = 0; t < n_ranks; ++t) {
        if (t == rank) continue;
        RankView v;
        if (int rc = view(t, v)) return rc;
        for (size_t k = 0; k < v.ghost_global.size(); ++k)
            if (v.ghost_owner[k] == rank) {
                const uint64_t g = v.ghost_global[k];
                if (g < mine.owned_begin || g >= mine.owned_begin + mine.n_owned)
                    return fail(B200FE_ERR_INVALID_ARG, "exchange lists: rank %d ghosts DoF %llu of rank %d outside its owned range", t, (unsigned long long)g, rank);
                send[t].push_back((uint32_t)(g - mine.owned_begin));
            }
    }
    for (int t = 0; t < n_ranks; ++t) {
        if (r_cnt[t] == 0 && send[t].empty()) continue;
        ex.peers.push_back(t);
        ex.recv_off.push_back(r_off[t]);
        ex.recv_cnt.push_back(r_cnt[t]);
        ex.send_off.push_back((uint32_t)ex.send_idx.size());
        ex.send_cnt.push_back((uint32_t)send[t].size());
        ex.send_idx.insert(ex.send_idx.end(), send[t].begin(), send[t].end());
    }
    return B200FE_OK;
}

}  // namespace
}  // namespace b200fe

using namespace b200fe;

extern "C" {

int b200fe_exchange_create_box(const b200fe_boxmesh_desc *desc, b200fe_exchange **out)
{
    B200FE_REQUIRE(desc && out, "b200fe_exchange_create_box: null pointer");
    auto ex = std::make_unique<Exchange>();
    if (desc->ghosts == B200FE_GHOSTS_MINIMAL) {
        // from this rank's own view only (mesh.cc: boxmesh_send_lists) -- no replay of the other ranks' meshes
        b200fe_boxmesh *m = nullptr;
        if (int rc = b200fe_boxmesh_create(desc, &m)) return rc;
        struct Guard { b200fe_boxmesh *m; ~Guard() { b200fe_boxmesh_destroy(m); } } guard{m};
        b200fe_boxmesh_info_t info;
        if (int rc = b200fe_boxmesh_info(m, &info)) return rc;
        std::vector<uint64_t> gg(info.n_ghost);
        std::vector<int32_t> go(info.n_ghost);
        if (int rc = b200fe_boxmesh_fill(m, nullptr, nullptr, gg.data(), go.data(), nullptr, nullptr)) return rc;
        std::vector<std::vector<uint32_t>> send;
        if (int rc = boxmesh_send_lists(m, send)) return rc;
        const int R = desc->n_ranks, me = desc->rank;
        ex->n_owned = info.n_owned; ex->n_ghost = info.n_ghost;
        std::vector<uint32_t> r_off(R, 0), r_cnt(R, 0);
        for (size_t k = 0; k < go.size(); ++k) {
            const int t = go[k];
            if (t < 0 || t >= R || t == me) return fail(B200FE_ERR_INVALID_ARG, "exchange lists: ghost %zu has owner %d", k, t);
            if (r_cnt[t] == 0) r_off[t] = (uint32_t)k;
            else if (r_off[t] + r_cnt[t] != k) return fail(B200FE_ERR_INVALID_ARG, "exchange lists: ghosts of rank %d are not contiguous", t);
            ++r_cnt[t];
        }
        for (int t = 0; t < R; ++t) {
            if (r_cnt[t] == 0 && send[t].empty()) continue;
            ex->peers.push_back(t);
            ex->recv_off.push_back(r_off[t]);
            ex->recv_cnt.push_back(r_cnt[t]);
            ex->send_off.push_back((uint32_t)ex->send_idx.size());
            ex->send_cnt.push_back((uint32_t)send[t].size());
            ex->send_idx.insert(ex->send_idx.end(), send[t].begin(), send[t].end());
        }
        *out = reinterpret_cast<b200fe_exchange *>(ex.release());
        return B200FE_OK;
    }
    auto view = [&](int t, RankView &v) -> int {
        b200fe_boxmesh_desc d = *desc;
        d.rank = t;
        b200fe_boxmesh *m = nullptr;
        if (int rc = b200fe_boxmesh_create(&d, &m)) return rc;
        b200fe_boxmesh_info_t info;
        int rc = b200fe_boxmesh_info(m, &info);
        if (rc == B200FE_OK) {
            v.owned_begin = info.owned_begin; v.n_owned = info.n_owned;
            v.ghost_global.resize(info.n_ghost); v.ghost_owner.resize(info.n_ghost);
            rc = b200fe_boxmesh_fill(m, nullptr, nullptr, v.ghost_global.data(), v.ghost_owner.data(), nullptr, nullptr);
        }
        b200fe_boxmesh_destroy(m);
        return rc;
    };
    if (int rc = build_exchange(desc->n_ranks, desc->rank, view, *ex)) return rc;
    *out = reinterpret_cast<b200fe_exchange *>(ex.release());
    return B200FE_OK;
}

int b200fe_exchange_create_hang(const b200fe_hangmesh_desc *desc, b200fe_exchange **out)
{
    B200FE_REQUIRE(desc && out, "b200fe_exchange_create_hang: null pointer");
    auto ex = std::make_unique<Exchange>();
    auto view = [&](int t, RankView &v) -> int {
        b200fe_hangmesh_desc d = *desc;
        d.box.rank = t;
        b200fe_hangmesh *m = nullptr;
        if (int rc = b200fe_hangmesh_create(&d, &m)) return rc;
        b200fe_hangmesh_info_t info;
        int rc = b200fe_hangmesh_info(m, &info);
        if (rc == B200FE_OK) {
            v.owned_begin = info.owned_begin; v.n_owned = info.n_owned;
            v.ghost_global.resize(info.n_ghost); v.ghost_owner.resize(info.n_ghost);
            rc = b200fe_hangmesh_fill(m, nullptr, nullptr, v.ghost_global.data(), v.ghost_owner.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        }
        b200fe_hangmesh_destroy(m);
        return rc;
    };
    if (int rc = build_exchange(desc->box.n_ranks, desc->box.rank, view, *ex)) return rc;
    *out = reinterpret_cast<b200fe_exchange *>(ex.release());
    return B200FE_OK;
}

void b200fe_exchange_destroy(b200fe_exchange *ex) { delete reinterpret_cast<Exchange *>(ex); }

int b200fe_exchange_info(const b200fe_exchange *ex, int *n_peers, uint32_t *n_send, uint32_t *n_owned, uint32_t *n_ghost)
{
    B200FE_REQUIRE(ex, "b200fe_exchange_info: null pointer");
    const Exchange *e = reinterpret_cast<const Exchange *>(ex);
    if (n_peers) *n_peers = (int)e->peers.size();
    if (n_send) *n_send = (uint32_t)e->send_idx.size();
    if (n_owned) *n_owned = e->n_owned;
    if (n_ghost) *n_ghost = e->n_ghost;
    return B200FE_OK;
}

int b200fe_exchange_fill(const b200fe_exchange *ex, int32_t *h_peers, uint32_t *h_recv_offset, uint32_t *h_recv_count,
                         uint32_t *h_send_offset, uint32_t *h_send_count, uint32_t *h_send_indices)
{
    B200FE_REQUIRE(ex, "b200fe_exchange_fill: null pointer");
    const Exchange *e = reinterpret_cast<const Exchange *>(ex);
    auto put = [](auto *dst, const auto &v) {
        if (dst && !v.empty()) std::copy(v.begin(), v.end(), dst);
    };
    put(h_peers, e->peers); put(h_recv_offset, e->recv_off); put(h_recv_count, e->recv_cnt);
    put(h_send_offset, e->send_off); put(h_send_count, e->send_cnt); put(h_send_indices, e->send_idx);
    return B200FE_OK;
}

}  // extern "C"
