// mesh.cc -- box mesh, FE_Q DoF numbering, partition, ghost lists (host only, OpenMP).
//
// The reference takes all of this from deal.II (un-vendored): GridGenerator::subdivided_hyper_rectangle
// + refine_global (CEED_bp/src/bp3.cc:452-488), DoFHandler::distribute_dofs (bp3.cc:135-136),
// Dirichlet constraints on boundary id 0 (bp3.cc:147-151), the per-cell lexicographic index table
// with invalid_unsigned_int on constrained DoFs (CEED_bp/include/portable_laplace_operator.h:304-394)
// and the Partitioner's owned/ghost layout.  Here the same objects
// are produced in closed form for structured boxes -- no cell-by-cell "first touch" sweep:
//
//   * active-cell order  = coarse cell (lexicographic, x fastest) * 8^n + Morton(local x,y,z)
//   * owner of a mesh entity (vertex/line/quad/hex) = lowest rank among the cells sharing it
//   * within its owner, an entity is numbered by the first (lowest active index) cell sharing it,
//     in the order vertices, lines, quads, interior of that cell
//   => global index = prefix[first cell] + (new entities of that cell before it) + index in entity,
//      with one prefix sum over all cells in active order (ranks own contiguous cell ranges).
//
// The result is bit-identical to the literal first-touch simulation in oracle/fe_oracle.py
// (tests/test_mesh.py).
#include <sched.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>
#include <vector>

#include "common.h"
#include "mesh_common.h"

namespace b200fe {

namespace meshdetail {
void use_setup_threads()
{
#ifdef _OPENMP
    int want = 0;
    if (const char *e = std::getenv("B200FE_SETUP_THREADS")) want = std::atoi(e);
    if (want < 1) {
        int cores = 0;
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
        if (cores < 1) cores = omp_get_num_procs();
        int local_ranks = 1;
        if (const char *e = std::getenv("LOCAL_WORLD_SIZE")) local_ranks = std::max(1, std::atoi(e));
        want = std::max(1, std::min(cores / local_ranks, 32));
    }
    omp_set_num_threads(want);
#endif
}
}  // namespace meshdetail

using meshdetail::kEnt;
using meshdetail::morton3;

struct BoxMesh {
    int sub[3], nref, p, nranks, rank, scheme, ghost_mode, dirichlet;
    int64_t cells[3];
    double p1[3], p2[3], h[3];
    int64_t n_cells_global = 0;
    uint64_t n_dofs_global = 0;
    std::vector<uint32_t> newmask;   // by active index: which of the 27 entities the cell numbers
    std::vector<uint64_t> gprefix;   // by active index: first global DoF numbered by the cell
    std::vector<int64_t> rank_cell_begin;
    std::vector<uint64_t> rank_dof_begin;
    // local data of `rank`
    int64_t cell_begin = 0, cell_end = 0;
    uint64_t owned_begin = 0, owned_end = 0;
    std::vector<uint32_t> lbase;         // [n_local_cells][27] local index of the first DoF of each entity of the cell
    std::vector<uint64_t> ghost_global;  // sorted
    std::vector<int32_t> ghost_owner;
    std::vector<uint32_t> constrained;   // owned local indices on the Dirichlet boundary
    std::vector<int32_t> cell_xyz;       // [n_local_cells][3]
    // lexicographic local dof -> (entity, index in entity)
    std::vector<int> l_ent, l_idx;
    int ent_size[27];

    int64_t n_local_cells() const { return cell_end - cell_begin; }

    // active index of cell (x,y,z): table lookup once build() has filled it (the Morton bit loop was the whole cost of the
    // numbering passes: 125 neighbour look-ups per cell), closed form otherwise
    std::vector<uint32_t> pos_tab;
    uint64_t pos_of(int64_t x, int64_t y, int64_t z) const
    {
        if (!pos_tab.empty()) return pos_tab[(size_t)((z * cells[1] + y) * cells[0] + x)];
        return pos_of_closed(x, y, z);
    }
    uint64_t pos_of_closed(int64_t x, int64_t y, int64_t z) const
    {
        const uint32_t f = (1u << nref) - 1;
        const uint64_t coarse = (uint64_t)(x >> nref) + (uint64_t)sub[0] * ((uint64_t)(y >> nref) + (uint64_t)sub[1] * (uint64_t)(z >> nref));
        return (coarse << (3 * nref)) | morton3((uint32_t)x & f, (uint32_t)y & f, (uint32_t)z & f, nref);
    }
    void xyz_of(uint64_t pos, int64_t &x, int64_t &y, int64_t &z) const
    {
        const uint64_t per = 1ull << (3 * nref);
        const uint64_t coarse = pos >> (3 * nref), loc = pos & (per - 1);
        uint32_t lx = 0, ly = 0, lz = 0;
        for (int b = 0; b < nref; ++b) {
            lx |= (uint32_t)((loc >> (3 * b)) & 1) << b;
            ly |= (uint32_t)((loc >> (3 * b + 1)) & 1) << b;
            lz |= (uint32_t)((loc >> (3 * b + 2)) & 1) << b;
        }
        const int64_t X = coarse % sub[0], Y = (coarse / sub[0]) % sub[1], Z = coarse / ((uint64_t)sub[0] * sub[1]);
        x = (X << nref) | lx; y = (Y << nref) | ly; z = (Z << nref) | lz;
    }
    int rank_of_pos(uint64_t pos) const
    {
        const uint64_t N = (uint64_t)n_cells_global, P = (uint64_t)nranks;
        if (scheme == B200FE_PARTITION_P4EST) return (int)(((pos + 1) * P - 1) / N);
        const uint64_t per = (N + P - 1) / P;  // create_triangulation.h:44-51
        return (int)(pos / per);
    }

    // owner rank and first-touch cell of entity `e` of cell (x,y,z); returns the entity's id
    // relative to that first cell in *ef.  Ranks own contiguous ranges of the active-cell order (rank_of_pos is
    // monotone), so "lowest rank, then lowest cell" is simply the lowest active index among the sharing cells: the
    // inner loop needs no rank computation (two 64-bit divisions per candidate in the first version of this file --
    // 8.4 s single-threaded for the 2 M cells of the 8-GPU headline mesh; SCALE_r01 setup_s 11 s).
    void entity_owner(const int64_t c[3], int e, int &owner, uint64_t &first, int *ef, bool want_owner = true) const
    {
        int64_t lo[3], hi[3];
        for (int d = 0; d < 3; ++d) {
            const int t = kEnt.code[e][d];
            lo[d] = t == 0 ? c[d] - 1 : c[d];
            hi[d] = t == 2 ? c[d] + 1 : c[d];
            lo[d] = std::max<int64_t>(lo[d], 0);
            hi[d] = std::min<int64_t>(hi[d], cells[d] - 1);
        }
        owner = nranks;
        first = ~0ull;
        int64_t fx = 0, fy = 0, fz = 0;
        for (int64_t z = lo[2]; z <= hi[2]; ++z)
            for (int64_t y = lo[1]; y <= hi[1]; ++y)
                for (int64_t x = lo[0]; x <= hi[0]; ++x) {
                    const uint64_t ps = pos_of(x, y, z);
                    if (ps < first) {
                        first = ps; fx = x; fy = y; fz = z;
                    }
                }
        if (want_owner) owner = rank_of_pos(first);
        if (ef) {
            const int64_t f[3] = {fx, fy, fz};
            int rc[3];
            for (int d = 0; d < 3; ++d) {
                const int t = kEnt.code[e][d];
                if (t == 1) rc[d] = 1;
                else {
                    const int64_t plane = c[d] + (t == 2 ? 1 : 0);
                    rc[d] = plane == f[d] ? 0 : 2;
                }
            }
            *ef = kEnt.id_of_code[rc[0]][rc[1]][rc[2]];
        }
    }

    // 3x3x3 neighbourhood of cell c: active index of every neighbour, ~0 where the mesh ends; slot = (dx+1) + 3 (dy+1) + 9 (dz+1)
    void neighbourhood(const int64_t c[3], uint64_t nb[27]) const
    {
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int64_t x = c[0] + dx, y = c[1] + dy, z = c[2] + dz;
                    const bool in = x >= 0 && y >= 0 && z >= 0 && x < cells[0] && y < cells[1] && z < cells[2];
                    nb[(dx + 1) + 3 * (dy + 1) + 9 * (dz + 1)] = in ? pos_of(x, y, z) : ~0ull;
                }
    }
    // per entity: the neighbourhood slots of the cells sharing it, and the entity's id as seen from each of them
    struct NbTables {
        uint32_t sharers[27];
        int n_sharers[27], sharer_slot[27][8], ef_of[27][27];
        NbTables()
        {
            for (int e = 0; e < 27; ++e) {
                sharers[e] = 0;
                n_sharers[e] = 0;
                for (int k = 0; k < 27; ++k) ef_of[e][k] = -1;
                int lo[3], hi[3];
                for (int d = 0; d < 3; ++d) {
                    const int t = kEnt.code[e][d];
                    lo[d] = t == 0 ? -1 : 0;
                    hi[d] = t == 2 ? 1 : 0;
                }
                for (int dz = lo[2]; dz <= hi[2]; ++dz)
                    for (int dy = lo[1]; dy <= hi[1]; ++dy)
                        for (int dx = lo[0]; dx <= hi[0]; ++dx) {
                            const int slot = (dx + 1) + 3 * (dy + 1) + 9 * (dz + 1), off[3] = {dx, dy, dz};
                            int rc[3];
                            for (int d = 0; d < 3; ++d) {
                                const int t = kEnt.code[e][d];
                                rc[d] = t == 1 ? 1 : (off[d] == (t == 2 ? 1 : 0) ? 0 : 2);  // the entity's plane as seen from that cell
                            }
                            sharers[e] |= 1u << slot;
                            sharer_slot[e][n_sharers[e]++] = slot;
                            ef_of[e][slot] = kEnt.id_of_code[rc[0]][rc[1]][rc[2]];
                        }
            }
        }
    };
    static const NbTables &nb_tables()
    {
        static const NbTables t;
        return t;
    }

    bool cell_at_boundary(const int32_t *c) const
    {
        bool at = false;
        for (int d = 0; d < 3; ++d) at |= c[d] == 0 || c[d] == cells[d] - 1;
        return at;
    }
    // the index table [n_local_cells][nm^3] (lexicographic, local numbering, invalid on the Dirichlet boundary), written
    // straight into the caller's buffer: its pages are first touched by the threads that fill them
    void expand_indices(uint32_t *out) const
    {
        meshdetail::use_setup_threads();
        const int nm = p + 1, nm3 = nm * nm * nm;
        const int64_t nloc = n_local_cells();
        const int64_t dimsL[3] = {cells[0] * p, cells[1] * p, cells[2] * p};
#pragma omp parallel for schedule(static)
        for (int64_t ci = 0; ci < nloc; ++ci) {
            const uint32_t *lb = &lbase[(size_t)ci * 27];
            const int32_t *c = &cell_xyz[ci * 3];
            uint32_t *row = out + (size_t)ci * nm3;
            for (int l = 0; l < nm3; ++l) row[l] = lb[l_ent[l]] + (uint32_t)l_idx[l];
            if (!(dirichlet && cell_at_boundary(c))) continue;
            for (int l = 0; l < nm3; ++l) {
                const int a = l % nm, b = (l / nm) % nm, cc = l / (nm * nm);
                const int64_t X = (int64_t)c[0] * p + a, Y = (int64_t)c[1] * p + b, Z = (int64_t)c[2] * p + cc;
                if (X == 0 || Y == 0 || Z == 0 || X == dimsL[0] || Y == dimsL[1] || Z == dimsL[2]) row[l] = B200FE_INVALID_INDEX;
            }
        }
    }

    uint64_t within_prefix(uint32_t mask, int e) const
    {
        uint64_t s = 0;
        for (int k = 0; k < e; ++k)
            if (mask >> k & 1) s += ent_size[k];
        return s;
    }

    int build();
};

int BoxMesh::build()
{
    meshdetail::use_setup_threads();
    // B200FE_SETUP_TRACE=1: wall time of the build phases on stderr
    static const bool trace = [] { const char *e = std::getenv("B200FE_SETUP_TRACE"); return e && std::atoi(e) != 0; }();
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[b200fe setup] box mesh rank %d: %-28s %7.1f ms\n", rank, what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    const int nm = p + 1, nm3 = nm * nm * nm;
    for (int e = 0; e < 27; ++e) ent_size[e] = meshdetail::entity_size(p, e);
    for (int d = 0; d < 3; ++d) {
        cells[d] = (int64_t)sub[d] << nref;
        h[d] = (p2[d] - p1[d]) / (double)cells[d];
    }
    n_cells_global = cells[0] * cells[1] * cells[2];
    n_dofs_global = (uint64_t)(cells[0] * p + 1) * (uint64_t)(cells[1] * p + 1) * (uint64_t)(cells[2] * p + 1);
    if (n_cells_global < nranks) return fail(B200FE_ERR_INVALID_ARG, "box mesh: fewer cells (%lld) than ranks (%d)", (long long)n_cells_global, nranks);

    meshdetail::lexicographic_entities(p, l_ent, l_idx);  // lexicographic local dof -> entity and index in entity
    if (n_cells_global < 0xFFFFFFFFll) {
        pos_tab.clear();
        std::vector<uint32_t> tab((size_t)n_cells_global);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n_cells_global; ++i) {
            const int64_t x = i % cells[0], y = (i / cells[0]) % cells[1], z = i / (cells[0] * cells[1]);
            tab[i] = (uint32_t)pos_of_closed(x, y, z);
        }
        pos_tab.swap(tab);
    }

    lap("position table");
    // pass 1 over ALL cells: which entities does each cell number, and how many DoFs.  One look-up per neighbour (26 per
    // cell) instead of one per (entity, sharing cell) pair (up to 216): a cell numbers an entity iff none of the cells
    // sharing it comes earlier in the active order.
    const NbTables &nt = nb_tables();
    newmask.assign(n_cells_global, 0);
    gprefix.assign(n_cells_global + 1, 0);
    lap("pass 1 allocation");
#pragma omp parallel for schedule(static)
    for (int64_t ps = 0; ps < n_cells_global; ++ps) {
        int64_t c[3];
        xyz_of((uint64_t)ps, c[0], c[1], c[2]);
        uint64_t nb[27];
        neighbourhood(c, nb);
        uint32_t lower = 0;  // neighbours that exist and come before this cell
        for (int k = 0; k < 27; ++k) lower |= (uint32_t)(nb[k] < (uint64_t)ps) << k;
        uint32_t mask = 0;
        uint64_t cnt = 0;
        for (int e = 0; e < 27; ++e)
            if (ent_size[e] != 0 && (lower & nt.sharers[e]) == 0) { mask |= 1u << e; cnt += ent_size[e]; }
        newmask[ps] = mask;
        gprefix[ps + 1] = cnt;
    }
    lap("pass 1 (all cells)");
    for (int64_t ps = 0; ps < n_cells_global; ++ps) gprefix[ps + 1] += gprefix[ps];
    if (gprefix[n_cells_global] != n_dofs_global)
        return fail(B200FE_ERR_INVALID_ARG, "box mesh: internal numbering error (%llu != %llu)",
                    (unsigned long long)gprefix[n_cells_global], (unsigned long long)n_dofs_global);

    lap("pass 1 (all cells) + prefix");
    rank_cell_begin.resize(nranks + 1);
    rank_dof_begin.resize(nranks + 1);
    {
        int64_t ps = 0;
        for (int r = 0; r <= nranks; ++r) {
            while (ps < n_cells_global && rank_of_pos((uint64_t)ps) < r) ++ps;  // monotone
            rank_cell_begin[r] = r == nranks ? n_cells_global : ps;
        }
        for (int r = 0; r <= nranks; ++r) rank_dof_begin[r] = gprefix[rank_cell_begin[r]];
    }
    cell_begin = rank_cell_begin[rank];
    cell_end = rank_cell_begin[rank + 1];
    owned_begin = rank_dof_begin[rank];
    owned_end = rank_dof_begin[rank + 1];
    const int64_t nloc = n_local_cells();
    const uint64_t n_owned = owned_end - owned_begin;
    if (n_owned >= 0xFFFFFFFFull) return fail(B200FE_ERR_UNSUPPORTED, "box mesh: more than 2^32-2 owned DoFs on one rank");

    // pass 2 over own cells: global index of the first DoF of each of the cell's 27 entities (the DoFs of an entity are
    // contiguous in the global numbering, so everything below works per entity, not per DoF); entities of other ranks
    // are the ghost candidates
    std::vector<uint64_t> gbase((size_t)nloc * 27);
    cell_xyz.resize((size_t)nloc * 3);
    struct Ent { uint64_t base; uint32_t size; };
    std::vector<Ent> cand;
#pragma omp parallel
    {
        std::vector<Ent> mine;
#pragma omp for schedule(static) nowait
        for (int64_t ci = 0; ci < nloc; ++ci) {
            int64_t c[3];
            xyz_of((uint64_t)(cell_begin + ci), c[0], c[1], c[2]);
            for (int d = 0; d < 3; ++d) cell_xyz[ci * 3 + d] = (int32_t)c[d];
            uint64_t nb[27];
            neighbourhood(c, nb);
            for (int e = 0; e < 27; ++e) {
                if (ent_size[e] == 0) { gbase[ci * 27 + e] = 0; continue; }
                uint64_t first = ~0ull;
                int slot = 13;
                for (int k = 0; k < nt.n_sharers[e]; ++k) {
                    const int sl = nt.sharer_slot[e][k];
                    if (nb[sl] < first) { first = nb[sl]; slot = sl; }
                }
                const uint64_t b0 = gprefix[first] + within_prefix(newmask[first], nt.ef_of[e][slot]);
                gbase[ci * 27 + e] = b0;
                if (b0 < owned_begin || b0 >= owned_end) mine.push_back(Ent{b0, (uint32_t)ent_size[e]});
            }
        }
#pragma omp critical
        cand.insert(cand.end(), mine.begin(), mine.end());
    }
    lap("pass 2 (own cells)");
    // deal.II "relevant" ghosts: also the DoFs on the vertex-neighbour cells of own cells
    if (ghost_mode == B200FE_GHOSTS_RELEVANT && nranks > 1) {
        std::vector<uint64_t> layer;
        for (int64_t ci = 0; ci < nloc; ++ci) {
            const int32_t *c = &cell_xyz[ci * 3];
            for (int dz = -1; dz <= 1; ++dz)
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int64_t x = c[0] + dx, y = c[1] + dy, z = c[2] + dz;
                        if (x < 0 || y < 0 || z < 0 || x >= cells[0] || y >= cells[1] || z >= cells[2]) continue;
                        const uint64_t ps = pos_of(x, y, z);
                        if ((int64_t)ps < cell_begin || (int64_t)ps >= cell_end) layer.push_back(ps);
                    }
        }
        std::sort(layer.begin(), layer.end());
        layer.erase(std::unique(layer.begin(), layer.end()), layer.end());
        for (uint64_t ps : layer) {
            int64_t c[3];
            xyz_of(ps, c[0], c[1], c[2]);
            for (int e = 0; e < 27; ++e) {
                if (ent_size[e] == 0) continue;
                int owner, ef; uint64_t first;
                entity_owner(c, e, owner, first, &ef);
                if (owner == rank) continue;
                cand.push_back(Ent{gprefix[first] + within_prefix(newmask[first], ef), (uint32_t)ent_size[e]});
            }
        }
    }
    std::sort(cand.begin(), cand.end(), [](const Ent &a, const Ent &b) { return a.base < b.base; });
    cand.erase(std::unique(cand.begin(), cand.end(), [](const Ent &a, const Ent &b) { return a.base == b.base; }), cand.end());
    // ghost entities sorted by global index: their DoFs in that order are the (sorted) ghost list
    std::vector<uint64_t> gent_base(cand.size());
    std::vector<uint32_t> gent_local(cand.size() + 1, 0);  // position of the entity's first DoF in the ghost segment
    for (size_t i = 0; i < cand.size(); ++i) {
        gent_base[i] = cand[i].base;
        gent_local[i + 1] = gent_local[i] + cand[i].size;
    }
    const size_t n_ghost = gent_local[cand.size()];
    if (n_owned + n_ghost >= 0xFFFFFFFFull) return fail(B200FE_ERR_UNSUPPORTED, "box mesh: local vector too long for 32-bit indices");
    ghost_global.resize(n_ghost);
    ghost_owner.resize(n_ghost);
    for (size_t i = 0; i < cand.size(); ++i) {
        const int32_t owner = (int32_t)(std::upper_bound(rank_dof_begin.begin(), rank_dof_begin.end(), cand[i].base) - rank_dof_begin.begin() - 1);
        for (uint32_t k = 0; k < cand[i].size; ++k) {
            ghost_global[gent_local[i] + k] = cand[i].base + k;
            ghost_owner[gent_local[i] + k] = owner;
        }
    }

    lap("ghost list");
    // local index of the first DoF of every entity of every own cell (the index table itself -- nm^3 entries per cell, 360 MB
    // for the 64^3-cell headline mesh -- is expanded from this straight into the caller's buffer by b200fe_boxmesh_fill)
    lbase.resize((size_t)nloc * 27);
#pragma omp parallel for schedule(static)
    for (int64_t ci = 0; ci < nloc; ++ci)
        for (int e = 0; e < 27; ++e) {
            const uint64_t g = gbase[ci * 27 + e];
            uint32_t lb;
            if (ent_size[e] == 0) lb = 0;
            else if (g >= owned_begin && g < owned_end) lb = (uint32_t)(g - owned_begin);
            else lb = (uint32_t)(n_owned + gent_local[std::lower_bound(gent_base.begin(), gent_base.end(), g) - gent_base.begin()]);
            lbase[ci * 27 + e] = lb;
        }
    // owned constrained list: DoFs of own boundary cells on the Dirichlet boundary
    std::vector<uint32_t> cons;
    if (dirichlet) {
        const int64_t dimsL[3] = {cells[0] * p, cells[1] * p, cells[2] * p};
#pragma omp parallel
        {
            std::vector<uint32_t> mine;
#pragma omp for schedule(static) nowait
            for (int64_t ci = 0; ci < nloc; ++ci) {
                const int32_t *c = &cell_xyz[ci * 3];
                if (!cell_at_boundary(c)) continue;
                for (int l = 0; l < nm3; ++l) {
                    const int a = l % nm, b = (l / nm) % nm, cc = l / (nm * nm);
                    const int64_t X = (int64_t)c[0] * p + a, Y = (int64_t)c[1] * p + b, Z = (int64_t)c[2] * p + cc;
                    if (X == 0 || Y == 0 || Z == 0 || X == dimsL[0] || Y == dimsL[1] || Z == dimsL[2]) {
                        const uint32_t loc = lbase[ci * 27 + l_ent[l]] + (uint32_t)l_idx[l];
                        if (loc < n_owned) mine.push_back(loc);
                    }
                }
            }
#pragma omp critical
            cons.insert(cons.end(), mine.begin(), mine.end());
        }
    }
    std::sort(cons.begin(), cons.end());
    cons.erase(std::unique(cons.begin(), cons.end()), cons.end());
    constrained.swap(cons);
    lap("local entity bases, constrained");
    // owned boundary DoFs that only OTHER ranks' cells touch cannot exist: the owner is the lowest
    // rank among the cells sharing the entity, so it has a cell there.
    return B200FE_OK;
}

// Owned local indices that every other rank ghosts (GHOSTS_MINIMAL), ascending = in that rank's ghost order, from this
// rank's own view: rank t ghosts the DoFs of an entity E that I own exactly when one of t's cells shares E, and the owner
// of E always has a cell on it -- so a sweep over my cells' entities and their sharing cells finds every (E, t) pair.
// Replaces the replay of all n_ranks meshes (O(n_ranks) host work per rank).
int boxmesh_send_lists(const b200fe_boxmesh *mesh, std::vector<std::vector<uint32_t>> &send)
{
    const BoxMesh &m = *reinterpret_cast<const BoxMesh *>(mesh);
    send.assign(m.nranks, {});
    if (m.nranks == 1) return B200FE_OK;
    if (m.ghost_mode != B200FE_GHOSTS_MINIMAL) return fail(B200FE_ERR_UNSUPPORTED, "boxmesh_send_lists: minimal ghost sets only");
    const int64_t nloc = m.n_local_cells();
#pragma omp parallel
    {
        std::vector<std::vector<uint32_t>> mine(m.nranks);
#pragma omp for schedule(static) nowait
        for (int64_t ci = 0; ci < nloc; ++ci) {
            const int64_t c[3] = {m.cell_xyz[ci * 3], m.cell_xyz[ci * 3 + 1], m.cell_xyz[ci * 3 + 2]};
            for (int e = 0; e < 27; ++e) {
                if (m.ent_size[e] == 0) continue;
                int64_t lo[3], hi[3];
                for (int d = 0; d < 3; ++d) {
                    const int t = kEnt.code[e][d];
                    lo[d] = std::max<int64_t>(t == 0 ? c[d] - 1 : c[d], 0);
                    hi[d] = std::min<int64_t>(t == 2 ? c[d] + 1 : c[d], m.cells[d] - 1);
                }
                if (lo[0] == hi[0] && lo[1] == hi[1] && lo[2] == hi[2]) continue;  // not shared
                // foreign ranks among the sharing cells (at most 7)
                int foreign[8], nf = 0;
                bool any_lower = false;
                for (int64_t z = lo[2]; z <= hi[2]; ++z)
                    for (int64_t y = lo[1]; y <= hi[1]; ++y)
                        for (int64_t x = lo[0]; x <= hi[0]; ++x) {
                            const uint64_t ps = m.pos_of(x, y, z);
                            if ((int64_t)ps >= m.cell_begin && (int64_t)ps < m.cell_end) continue;
                            if ((int64_t)ps < m.cell_begin) { any_lower = true; continue; }  // a lower rank owns E
                            const int r = m.rank_of_pos(ps);
                            bool seen = false;
                            for (int k = 0; k < nf; ++k) seen |= foreign[k] == r;
                            if (!seen) foreign[nf++] = r;
                        }
                if (any_lower || nf == 0) continue;
                int owner, ef; uint64_t first;
                m.entity_owner(c, e, owner, first, &ef, false);
                const uint64_t b0 = m.gprefix[first] + m.within_prefix(m.newmask[first], ef);
                for (int k = 0; k < nf; ++k)
                    for (int i = 0; i < m.ent_size[e]; ++i) mine[foreign[k]].push_back((uint32_t)(b0 + i - m.owned_begin));
            }
        }
#pragma omp critical
        for (int t = 0; t < m.nranks; ++t) send[t].insert(send[t].end(), mine[t].begin(), mine[t].end());
    }
    for (auto &v : send) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
    }
    return B200FE_OK;
}

int boxmesh_tables(const b200fe_boxmesh *mesh, BoxMeshTables *t)
{
    const BoxMesh &m = *reinterpret_cast<const BoxMesh *>(mesh);
    t->p = m.p; t->dirichlet = m.dirichlet;
    t->n_cells_local = m.n_local_cells();
    for (int d = 0; d < 3; ++d) t->cells[d] = m.cells[d];
    t->lbase = m.lbase.data(); t->cell_xyz = m.cell_xyz.data();
    t->l_ent = m.l_ent.data(); t->l_idx = m.l_idx.data();
    return B200FE_OK;
}

}  // namespace b200fe

using namespace b200fe;

extern "C" {

int b200fe_boxmesh_create(const b200fe_boxmesh_desc *d, b200fe_boxmesh **out)
{
    B200FE_REQUIRE(d && out, "b200fe_boxmesh_create: null pointer");
    if (d->p < 1 || d->p > 8) return fail(B200FE_ERR_UNSUPPORTED, "box mesh: degree p=%d outside 1..8", d->p);
    B200FE_REQUIRE(d->n_refine >= 0 && d->n_refine <= 10, "box mesh: n_refine out of range");
    B200FE_REQUIRE(d->n_ranks >= 1 && d->rank >= 0 && d->rank < d->n_ranks, "box mesh: bad rank %d of %d", d->rank, d->n_ranks);
    B200FE_REQUIRE(d->partition == B200FE_PARTITION_P4EST || d->partition == B200FE_PARTITION_BLOCKS, "box mesh: bad partition scheme");
    B200FE_REQUIRE(d->ghosts == B200FE_GHOSTS_MINIMAL || d->ghosts == B200FE_GHOSTS_RELEVANT, "box mesh: bad ghost mode");
    auto m = std::make_unique<BoxMesh>();
    for (int k = 0; k < 3; ++k) {
        B200FE_REQUIRE(d->subdivisions[k] >= 1, "box mesh: subdivisions must be >= 1");
        B200FE_REQUIRE(d->p2[k] > d->p1[k], "box mesh: p2 must exceed p1");
        m->sub[k] = d->subdivisions[k];
        m->p1[k] = d->p1[k];
        m->p2[k] = d->p2[k];
    }
    m->nref = d->n_refine; m->p = d->p; m->nranks = d->n_ranks; m->rank = d->rank;
    m->scheme = d->partition; m->ghost_mode = d->ghosts; m->dirichlet = d->dirichlet;
    if (int rc = m->build()) return rc;
    *out = reinterpret_cast<b200fe_boxmesh *>(m.release());
    return B200FE_OK;
}

void b200fe_boxmesh_destroy(b200fe_boxmesh *mesh) { delete reinterpret_cast<BoxMesh *>(mesh); }

int b200fe_boxmesh_info(const b200fe_boxmesh *mesh, b200fe_boxmesh_info_t *info)
{
    B200FE_REQUIRE(mesh && info, "b200fe_boxmesh_info: null pointer");
    const BoxMesh *m = reinterpret_cast<const BoxMesh *>(mesh);
    info->n_cells_global = (uint64_t)m->n_cells_global;
    info->n_dofs_global = m->n_dofs_global;
    info->n_cells_local = (uint32_t)m->n_local_cells();
    info->first_cell = (uint64_t)m->cell_begin;
    info->owned_begin = m->owned_begin;
    info->n_owned = (uint32_t)(m->owned_end - m->owned_begin);
    info->n_ghost = (uint32_t)m->ghost_global.size();
    info->n_constrained = (uint32_t)m->constrained.size();
    for (int d = 0; d < 3; ++d) { info->cells[d] = (uint32_t)m->cells[d]; info->h[d] = m->h[d]; info->origin[d] = m->p1[d]; }
    return B200FE_OK;
}

int b200fe_boxmesh_fill(const b200fe_boxmesh *mesh, uint32_t *h_dof_indices, uint32_t *h_constrained,
                        uint64_t *h_ghost_global, int32_t *h_ghost_owner, int32_t *h_cell_xyz,
                        uint64_t *h_rank_dof_begin)
{
    B200FE_REQUIRE(mesh, "b200fe_boxmesh_fill: null mesh");
    const BoxMesh *m = reinterpret_cast<const BoxMesh *>(mesh);
    if (h_dof_indices) m->expand_indices(h_dof_indices);
    if (h_constrained) std::memcpy(h_constrained, m->constrained.data(), m->constrained.size() * sizeof(uint32_t));
    if (h_ghost_global) std::memcpy(h_ghost_global, m->ghost_global.data(), m->ghost_global.size() * sizeof(uint64_t));
    if (h_ghost_owner) std::memcpy(h_ghost_owner, m->ghost_owner.data(), m->ghost_owner.size() * sizeof(int32_t));
    if (h_cell_xyz) std::memcpy(h_cell_xyz, m->cell_xyz.data(), m->cell_xyz.size() * sizeof(int32_t));
    if (h_rank_dof_begin) std::memcpy(h_rank_dof_begin, m->rank_dof_begin.data(), m->rank_dof_begin.size() * sizeof(uint64_t));
    return B200FE_OK;
}

}  // extern "C"
