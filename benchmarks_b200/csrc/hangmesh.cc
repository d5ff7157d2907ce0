// hangmesh.cc -- box mesh with ONE extra level of local refinement: FE_Q DoF numbering across the two levels,
// hanging-node constraint rows, partition and ghost lists (host only).
//
// BASELINE config C5 ("... on a deformed mesh with hanging nodes") and the north star's subsystem (1)
// ("element-to-DoF gather/scatter with hanging-node/constraint handling").  The reference has no such code
// (SURVEY.md section 8c); it would obtain these objects from deal.II: Triangulation::execute_coarsening_and_refinement,
// DoFHandler::distribute_dofs, DoFTools::make_hanging_node_constraints, AffineConstraints, and the per-cell index
// table of CEED_bp/include/portable_laplace_operator.h:304-394.  Conventions restated (same list H1-H5 as
// oracle/hanging_oracle.py, which this file is bit-identical to, tests/test_hanging.py):
//   H1 DoFs live on mesh objects; only the vertices of the coarse mesh are shared between the two levels.
//   H2 a fine-level DoF on the closure of an unrefined active cell K is constrained to the trace of K.
//   H3 cells are visited level by level (unrefined cells in z-order index order, then children grouped by parent);
//      first-touch numbering, vertices -> lines -> quads -> interior inside a cell.
//   H4 p4est curve (depth first) cut at floor(N r / P), families of 8 children kept whole (majority rank, ties low;
//      families of unrefined siblings are not corrected -- irrelevant for 1/2/4/8 ranks on the benchmark meshes);
//      a DoF belongs to the lowest rank among the active cells that have it.
//   H5 Dirichlet on the whole boundary takes precedence over a hanging constraint.
//
// Method: DoF identities are positions on two dense integer lattices (coarse level: cells*p+1 points per axis;
// fine level: the refined box with 2p intervals per coarse cell), so the first-touch sweep is a plain array walk
// (no hashing): ~2e8 lookups per second, 0.5 GB for the 2 x 67 M lattices of config C5.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#include "common.h"
#include "mesh_common.h"

namespace b200fe {

using meshdetail::kEnt;
using meshdetail::morton3;

namespace {
constexpr uint32_t kNone = 0xFFFFFFFFu;

struct CurveCell {
    int8_t level;
    int32_t x, y, z;     // coordinates at the cell's own level
    int32_t family;      // base position of the refined parent, -1 for unrefined cells
};
}  // namespace

struct HangMesh {
    int sub[3], nref, p, nranks, rank, dirichlet;
    int lo[3], hi[3];  // refined box, base-cell coordinates [lo, hi)
    int64_t cells[3];
    double p1[3], p2[3], h[3];
    uint64_t n_cells_global = 0, n_dofs_global = 0;
    std::vector<uint64_t> rank_dof_begin;
    uint64_t owned_begin = 0, owned_end = 0;
    // local data of `rank`
    std::vector<int32_t> cell_lxyz;      // [n_local_cells][4] level, x, y, z
    std::vector<uint32_t> dof_indices;   // [n_local_cells][nm^3]
    std::vector<uint64_t> ghost_global;
    std::vector<int32_t> ghost_owner;
    std::vector<uint32_t> constrained;   // owned local: Dirichlet + hanging
    std::vector<uint32_t> hang_dof, hang_row_ptr, hang_col;
    std::vector<double> hang_w;
    // the same rows grouped by coarse face (tensor-product trace interpolation): [n_face_blocks][(p+1)^2], [..][(2p+1)^2]
    std::vector<uint32_t> face_parents, face_children;
    uint32_t n_face_blocks = 0;
    uint64_t first_cell = 0;  // position of the rank's first cell on the p4est curve

    bool refined(int64_t x, int64_t y, int64_t z) const
    {
        return x >= lo[0] && x < hi[0] && y >= lo[1] && y < hi[1] && z >= lo[2] && z < hi[2];
    }
    void base_xyz(uint64_t pos, int64_t &x, int64_t &y, int64_t &z) const
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
    int build();
};

namespace {

// 1-D Lagrange weights of the coarse basis (GLL nodes t[0..p]) at the fine lattice positions rel = 0..2p of a coarse
// cell (rel = half*p + a  <->  xi = (half + t[a]) / 2).  Positions that coincide with a coarse node give exact unit
// vectors (rel = 0, 2p; rel = p for even p).
std::vector<double> trace_weights(int p, const std::vector<double> &t)
{
    const int nm = p + 1;
    std::vector<double> W((size_t)(2 * p + 1) * nm, 0.0);
    for (int rel = 0; rel <= 2 * p; ++rel) {
        double *w = &W[(size_t)rel * nm];
        if (rel == 0) { w[0] = 1.0; continue; }
        if (rel == 2 * p) { w[p] = 1.0; continue; }
        if (rel == p && p % 2 == 0) { w[p / 2] = 1.0; continue; }
        const int half = rel / p, a = rel % p;
        const double xi = 0.5 * (half + t[a]);
        for (int j = 0; j < nm; ++j) {
            double v = 1.0;
            for (int m = 0; m < nm; ++m)
                if (m != j) v *= (xi - t[m]) / (t[j] - t[m]);
            w[j] = v;
        }
    }
    return W;
}

}  // namespace

int HangMesh::build()
{
    meshdetail::use_setup_threads();
    // B200FE_SETUP_TRACE=1: wall time of the build phases on stderr
    static const bool trace = [] { const char *e = std::getenv("B200FE_SETUP_TRACE"); return e && std::atoi(e) != 0; }();
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[b200fe setup] hanging mesh rank %d: %-28s %7.1f ms\n", rank, what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    const int nm = p + 1, nm3 = nm * nm * nm;
    for (int d = 0; d < 3; ++d) {
        cells[d] = (int64_t)sub[d] << nref;
        h[d] = (p2[d] - p1[d]) / (double)cells[d];
        if (lo[d] < 0 || hi[d] > cells[d] || lo[d] >= hi[d])
            return fail(B200FE_ERR_INVALID_ARG, "hanging mesh: refine box [%d,%d) outside the %lld cells of axis %d", lo[d], hi[d], (long long)cells[d], d);
    }
    const int64_t n_base = cells[0] * cells[1] * cells[2];
    const int64_t n_ref = (int64_t)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
    n_cells_global = (uint64_t)(n_base + 7 * n_ref);
    if (n_cells_global < (uint64_t)nranks) return fail(B200FE_ERR_INVALID_ARG, "hanging mesh: fewer cells than ranks");
    if (n_cells_global >= 0x7FFFFFFFull) return fail(B200FE_ERR_UNSUPPORTED, "hanging mesh: too many cells");

    // ---- p4est curve and partition (H4) -------------------------------------------------------------------------
    std::vector<CurveCell> curve;
    curve.reserve(n_cells_global);
    for (int64_t ps = 0; ps < n_base; ++ps) {
        int64_t x, y, z;
        base_xyz((uint64_t)ps, x, y, z);
        if (refined(x, y, z)) {
            for (int ch = 0; ch < 8; ++ch)
                curve.push_back(CurveCell{1, (int32_t)(2 * x + (ch & 1)), (int32_t)(2 * y + ((ch >> 1) & 1)), (int32_t)(2 * z + ((ch >> 2) & 1)), (int32_t)ps});
        } else
            curve.push_back(CurveCell{0, (int32_t)x, (int32_t)y, (int32_t)z, -1});
    }
    const uint64_t N = n_cells_global, P = (uint64_t)nranks;
    auto raw_rank = [&](uint64_t i) { return (int)(((i + 1) * P - 1) / N); };  // first cell of rank r = floor(N r / P)
    std::vector<int32_t> subdomain(N);
    for (uint64_t i = 0; i < N;) {
        if (curve[i].family < 0) { subdomain[i] = raw_rank(i); ++i; continue; }
        // family of 8 consecutive cells: majority rank, ties to the lower rank
        int best = raw_rank(i), best_cnt = 0, cur = best, cnt = 0;
        for (int k = 0; k < 8; ++k) {
            const int r = raw_rank(i + k);
            if (r != cur) { if (cnt > best_cnt) { best = cur; best_cnt = cnt; } cur = r; cnt = 0; }
            ++cnt;
        }
        if (cnt > best_cnt) best = cur;
        for (int k = 0; k < 8; ++k) subdomain[i + k] = best;
        i += 8;
    }

    lap("curve + partition");
    // ---- dense lattices (H1) ------------------------------------------------------------------------------------
    const int64_t d0[3] = {cells[0] * p + 1, cells[1] * p + 1, cells[2] * p + 1};
    const int64_t d1[3] = {2 * (int64_t)(hi[0] - lo[0]) * p + 1, 2 * (int64_t)(hi[1] - lo[1]) * p + 1, 2 * (int64_t)(hi[2] - lo[2]) * p + 1};
    const uint64_t n0 = (uint64_t)d0[0] * d0[1] * d0[2], n1 = (uint64_t)d1[0] * d1[1] * d1[2];
    if (n0 + n1 >= 0xFFFFFFFFull) return fail(B200FE_ERR_UNSUPPORTED, "hanging mesh: more than 2^32-2 lattice points");
    std::vector<uint32_t> L0(n0, kNone), L1(n1, kNone);
    const int twop = 2 * p;
    // slot of local node (a,b,c) of a cell
    auto slot = [&](const CurveCell &c, int a, int b, int cc) -> uint32_t * {
        if (c.level == 0)
            return &L0[(uint64_t)(c.x * (int64_t)p + a) + d0[0] * ((uint64_t)(c.y * (int64_t)p + b) + d0[1] * (uint64_t)(c.z * (int64_t)p + cc))];
        const int64_t F[3] = {c.x * (int64_t)p + a, c.y * (int64_t)p + b, c.z * (int64_t)p + cc};  // absolute fine lattice
        if (F[0] % twop == 0 && F[1] % twop == 0 && F[2] % twop == 0)                                 // coarse vertex
            return &L0[(uint64_t)(F[0] / 2) + d0[0] * ((uint64_t)(F[1] / 2) + d0[1] * (uint64_t)(F[2] / 2))];
        return &L1[(uint64_t)(F[0] - (int64_t)lo[0] * twop) + d1[0] * ((uint64_t)(F[1] - (int64_t)lo[1] * twop) + d1[1] * (uint64_t)(F[2] - (int64_t)lo[2] * twop))];
    };

    // hierarchical -> lexicographic local order (H3 / SURVEY A2)
    std::vector<int> l_ent, l_idx, h2l(nm3);
    meshdetail::lexicographic_entities(p, l_ent, l_idx);
    for (int l = 0; l < nm3; ++l) h2l[l] = l;
    std::stable_sort(h2l.begin(), h2l.end(), [&](int a, int b) { return l_ent[a] != l_ent[b] ? l_ent[a] < l_ent[b] : l_idx[a] < l_idx[b]; });
    std::vector<int> la(nm3), lb(nm3), lc(nm3);
    for (int hh = 0; hh < nm3; ++hh) { const int l = h2l[hh]; la[hh] = l % nm; lb[hh] = (l / nm) % nm; lc[hh] = l / (nm * nm); }

    lap("lattices");
    // ---- first-touch numbering: ranks one after the other, inside a rank level by level (H3, H4) ---------------
    // rank cell ranges on the curve are contiguous; iterate each twice (level 0, then level 1)
    std::vector<uint64_t> rank_cell_begin(nranks + 1, N);
    {
        uint64_t i = 0;
        for (int r = 0; r <= nranks; ++r) {
            while (i < N && subdomain[i] < r) ++i;
            rank_cell_begin[r] = r == nranks ? N : i;
        }
    }
    rank_dof_begin.assign(nranks + 1, 0);
    uint32_t next = 0;
    for (int r = 0; r < nranks; ++r) {
        rank_dof_begin[r] = next;
        for (int lvl = 0; lvl <= 1; ++lvl)
            for (uint64_t i = rank_cell_begin[r]; i < rank_cell_begin[r + 1]; ++i) {
                const CurveCell &c = curve[i];
                if (c.level != lvl) continue;
                for (int hh = 0; hh < nm3; ++hh) {
                    uint32_t *s = slot(c, la[hh], lb[hh], lc[hh]);
                    if (*s == kNone) *s = next++;
                }
            }
    }
    rank_dof_begin[nranks] = next;
    n_dofs_global = next;
    owned_begin = rank_dof_begin[rank];
    owned_end = rank_dof_begin[rank + 1];
    first_cell = rank_cell_begin[rank];

    lap("first-touch numbering");
    // ---- local cells in iterator order --------------------------------------------------------------------------
    std::vector<uint64_t> mine;
    for (int lvl = 0; lvl <= 1; ++lvl)
        for (uint64_t i = rank_cell_begin[rank]; i < rank_cell_begin[rank + 1]; ++i)
            if (curve[i].level == lvl) mine.push_back(i);
    const int64_t nloc = (int64_t)mine.size();
    cell_lxyz.resize((size_t)nloc * 4);
    std::vector<uint32_t> gidx((size_t)nloc * nm3);
    std::vector<uint8_t> flag((size_t)nloc * nm3);  // 1 = Dirichlet, 2 = hanging
    const int64_t fmax[3] = {cells[0] * twop, cells[1] * twop, cells[2] * twop};
    // hanging test for an absolute fine lattice point: is it on the closure of an unrefined base cell?  Returns that cell.
    auto unrefined_neighbour = [&](const int64_t F[3], int64_t K[3]) -> bool {
        // a point strictly inside the refined box has refined cells all around it (the refined region is a box)
        if (F[0] > (int64_t)lo[0] * twop && F[0] < (int64_t)hi[0] * twop && F[1] > (int64_t)lo[1] * twop && F[1] < (int64_t)hi[1] * twop &&
            F[2] > (int64_t)lo[2] * twop && F[2] < (int64_t)hi[2] * twop)
            return false;
        int64_t cand[3][2];
        int nc[3];
        for (int d = 0; d < 3; ++d) {
            nc[d] = 0;
            const int64_t c = F[d] / twop;
            if (c < cells[d]) cand[d][nc[d]++] = c;
            if (F[d] % twop == 0 && c - 1 >= 0) cand[d][nc[d]++] = c - 1;
        }
        for (int i0 = 0; i0 < nc[0]; ++i0)
            for (int i1 = 0; i1 < nc[1]; ++i1)
                for (int i2 = 0; i2 < nc[2]; ++i2)
                    if (!refined(cand[0][i0], cand[1][i1], cand[2][i2])) {
                        K[0] = cand[0][i0]; K[1] = cand[1][i1]; K[2] = cand[2][i2];
                        return true;
                    }
        return false;
    };
#pragma omp parallel for schedule(static)
    for (int64_t ci = 0; ci < nloc; ++ci) {
        const CurveCell &c = curve[mine[ci]];
        cell_lxyz[ci * 4 + 0] = c.level; cell_lxyz[ci * 4 + 1] = c.x; cell_lxyz[ci * 4 + 2] = c.y; cell_lxyz[ci * 4 + 3] = c.z;
        for (int l = 0; l < nm3; ++l) {
            const int a = l % nm, b = (l / nm) % nm, cc = l / (nm * nm);
            gidx[ci * nm3 + l] = *slot(c, a, b, cc);
            const int scale = c.level == 0 ? 2 : 1;  // to the absolute fine lattice
            const int64_t F[3] = {scale * (c.x * (int64_t)p + a), scale * (c.y * (int64_t)p + b), scale * (c.z * (int64_t)p + cc)};
            uint8_t f = 0;
            if (dirichlet && (F[0] == 0 || F[1] == 0 || F[2] == 0 || F[0] == fmax[0] || F[1] == fmax[1] || F[2] == fmax[2])) f = 1;
            else if (c.level == 1 && !(F[0] % twop == 0 && F[1] % twop == 0 && F[2] % twop == 0)) {
                int64_t K[3];
                if (unrefined_neighbour(F, K)) f = 2;
            }
            flag[ci * nm3 + l] = f;
        }
    }

    lap("local indices + flags");
    // ---- hanging rows needed by this rank (H2): unique hanging DoFs of the local cells, sorted by global index ----
    struct Row { uint32_t g; int64_t F[3]; };
    std::vector<Row> rows;
    for (int64_t ci = 0; ci < nloc; ++ci) {
        const CurveCell &c = curve[mine[ci]];
        if (c.level == 0) continue;
        for (int l = 0; l < nm3; ++l)
            if (flag[ci * nm3 + l] == 2)
                rows.push_back(Row{gidx[ci * nm3 + l], {c.x * (int64_t)p + l % nm, c.y * (int64_t)p + (l / nm) % nm, c.z * (int64_t)p + l / (nm * nm)}});
    }
    std::sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) { return a.g < b.g; });
    rows.erase(std::unique(rows.begin(), rows.end(), [](const Row &a, const Row &b) { return a.g == b.g; }), rows.end());

    std::vector<double> t(nm), wt(nm);
    if (int rc = b200fe_basis_1d(p, nm, B200FE_QUAD_GLL, nullptr, nullptr, nullptr, t.data(), wt.data())) return rc;
    const std::vector<double> W = trace_weights(p, t);
    std::vector<uint32_t> row_ptr_g(1, 0), col_g;  // parents as GLOBAL indices first
    std::vector<double> wgt;
    for (const Row &r : rows) {
        int64_t K[3];
        if (!unrefined_neighbour(r.F, K)) return fail(B200FE_ERR_INVALID_ARG, "hanging mesh: internal error (row without an unrefined neighbour)");
        const double *w0 = &W[(size_t)(r.F[0] - K[0] * twop) * nm], *w1 = &W[(size_t)(r.F[1] - K[1] * twop) * nm], *w2 = &W[(size_t)(r.F[2] - K[2] * twop) * nm];
        std::vector<std::pair<uint32_t, double>> ent;
        for (int cc = 0; cc < nm; ++cc) {
            if (w2[cc] == 0.0) continue;  // (unit rows of W on a face / edge: a zero factor makes the product vanish)
            for (int b = 0; b < nm; ++b) {
                if (w1[b] == 0.0) continue;
                for (int a = 0; a < nm; ++a) {
                    const double w = w0[a] * w1[b] * w2[cc];
                    if (!(std::fabs(w) > 1e-14)) continue;
                    const int64_t X = K[0] * p + a, Y = K[1] * p + b, Z = K[2] * p + cc;
                    if (dirichlet && (X == 0 || Y == 0 || Z == 0 || X == d0[0] - 1 || Y == d0[1] - 1 || Z == d0[2] - 1)) continue;  // value 0
                    const uint32_t g = L0[(uint64_t)X + d0[0] * ((uint64_t)Y + d0[1] * (uint64_t)Z)];
                    if (g == kNone) return fail(B200FE_ERR_INVALID_ARG, "hanging mesh: internal error (unnumbered parent)");
                    ent.emplace_back(g, w);
                }
            }
        }
        std::sort(ent.begin(), ent.end());
        for (auto &e : ent) { col_g.push_back(e.first); wgt.push_back(e.second); }
        row_ptr_g.push_back((uint32_t)col_g.size());
    }

    lap("hanging rows");
    // ---- ghosts: non-owned DoFs of the local cells and non-owned parents of the rows ----------------------------
    std::vector<uint64_t> cand;
    for (uint32_t g : gidx)
        if (g < owned_begin || g >= owned_end) cand.push_back(g);
    for (uint32_t g : col_g)
        if (g < owned_begin || g >= owned_end) cand.push_back(g);
    std::sort(cand.begin(), cand.end());
    cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
    ghost_global.swap(cand);
    ghost_owner.resize(ghost_global.size());
    for (size_t i = 0; i < ghost_global.size(); ++i)
        ghost_owner[i] = (int32_t)(std::upper_bound(rank_dof_begin.begin(), rank_dof_begin.end(), ghost_global[i]) - rank_dof_begin.begin() - 1);
    const uint64_t n_owned = owned_end - owned_begin;
    if (n_owned + ghost_global.size() >= 0xFFFFFFFFull) return fail(B200FE_ERR_UNSUPPORTED, "hanging mesh: local vector too long for 32-bit indices");
    auto local_of = [&](uint64_t g) -> uint32_t {
        if (g >= owned_begin && g < owned_end) return (uint32_t)(g - owned_begin);
        return (uint32_t)(n_owned + (std::lower_bound(ghost_global.begin(), ghost_global.end(), g) - ghost_global.begin()));
    };

    dof_indices.resize(gidx.size());
    std::vector<uint32_t> cons;
#pragma omp parallel
    {
        std::vector<uint32_t> my_cons;
#pragma omp for schedule(static) nowait
        for (int64_t i = 0; i < (int64_t)gidx.size(); ++i) {
            const uint32_t loc = local_of(gidx[i]);
            dof_indices[i] = flag[i] == 1 ? B200FE_INVALID_INDEX : loc;
            if (flag[i] != 0 && loc < n_owned) my_cons.push_back(loc);
        }
#pragma omp critical
        cons.insert(cons.end(), my_cons.begin(), my_cons.end());
    }
    std::sort(cons.begin(), cons.end());
    cons.erase(std::unique(cons.begin(), cons.end()), cons.end());
    constrained.swap(cons);
    hang_dof.resize(rows.size());
    for (size_t r = 0; r < rows.size(); ++r) hang_dof[r] = local_of(rows[r].g);
    hang_row_ptr = row_ptr_g;
    hang_col.resize(col_g.size());
    for (size_t i = 0; i < col_g.size(); ++i) hang_col[i] = local_of(col_g[i]);
    hang_w.swap(wgt);

    lap("ghosts + local table");
    // ---- face-structured form of the same rows ------------------------------------------------------------------
    // Every hanging DoF of a box-refined mesh lies on a face shared by an unrefined cell K and a refined cell; its
    // (2p+1)^2 fine nodes are the tensor-product interpolation W (x) W of the (p+1)^2 coarse face nodes.  Blocks in
    // (K position, axis, side) order; a DoF on several faces belongs to the first block that needs it.
    {
        const int nf = twop + 1;
        std::vector<uint8_t> claimed(rows.size(), 0);
        auto row_of_global = [&](uint32_t g) -> int64_t {
            auto it = std::lower_bound(rows.begin(), rows.end(), g, [](const Row &r, uint32_t v) { return r.g < v; });
            return (it != rows.end() && it->g == g) ? (int64_t)(it - rows.begin()) : -1;
        };
        auto local_or_invalid = [&](uint64_t g) -> uint32_t {
            if (g >= owned_begin && g < owned_end) return (uint32_t)(g - owned_begin);
            auto it = std::lower_bound(ghost_global.begin(), ghost_global.end(), g);
            return (it != ghost_global.end() && *it == g) ? (uint32_t)(n_owned + (it - ghost_global.begin())) : B200FE_INVALID_INDEX;
        };
        std::vector<uint32_t> par((size_t)nm * nm), chi((size_t)nf * nf);
        size_t n_claimed = 0;
        for (int64_t ps = 0; ps < n_base && !rows.empty(); ++ps) {
            int64_t K[3];
            base_xyz((uint64_t)ps, K[0], K[1], K[2]);
            if (refined(K[0], K[1], K[2])) continue;
            for (int axis = 0; axis < 3; ++axis)
                for (int side = 0; side < 2; ++side) {
                    int64_t N[3] = {K[0], K[1], K[2]};
                    N[axis] += side ? 1 : -1;
                    if (N[axis] < 0 || N[axis] >= cells[axis] || !refined(N[0], N[1], N[2])) continue;
                    const int u1 = axis == 0 ? 1 : 0, u2 = axis == 2 ? 1 : 2;  // in-face axes, u1 < u2
                    bool any_child = false;
                    for (int bb = 0; bb < nf; ++bb)
                        for (int a = 0; a < nf; ++a) {
                            chi[a + (size_t)nf * bb] = B200FE_INVALID_INDEX;
                            int64_t F[3];
                            F[axis] = (K[axis] + side) * twop; F[u1] = K[u1] * twop + a; F[u2] = K[u2] * twop + bb;
                            if (F[0] % twop == 0 && F[1] % twop == 0 && F[2] % twop == 0) continue;  // coarse vertex
                            const uint32_t g = L1[(uint64_t)(F[0] - (int64_t)lo[0] * twop) + d1[0] * ((uint64_t)(F[1] - (int64_t)lo[1] * twop) + d1[1] * (uint64_t)(F[2] - (int64_t)lo[2] * twop))];
                            if (g == kNone) continue;
                            const int64_t r = row_of_global(g);
                            if (r < 0 || claimed[r]) continue;
                            claimed[r] = 1;
                            ++n_claimed;
                            chi[a + (size_t)nf * bb] = hang_dof[r];
                            any_child = true;
                        }
                    if (!any_child) continue;
                    for (int bb = 0; bb < nm; ++bb)
                        for (int a = 0; a < nm; ++a) {
                            int64_t X[3];
                            X[axis] = (K[axis] + side) * p; X[u1] = K[u1] * p + a; X[u2] = K[u2] * p + bb;
                            uint32_t v = B200FE_INVALID_INDEX;
                            const bool bnd = dirichlet && (X[0] == 0 || X[1] == 0 || X[2] == 0 || X[0] == d0[0] - 1 || X[1] == d0[1] - 1 || X[2] == d0[2] - 1);
                            if (!bnd) {
                                const uint32_t g = L0[(uint64_t)X[0] + d0[0] * ((uint64_t)X[1] + d0[1] * (uint64_t)X[2])];
                                if (g != kNone) v = local_or_invalid(g);
                            }
                            par[a + (size_t)nm * bb] = v;
                        }
                    face_parents.insert(face_parents.end(), par.begin(), par.end());
                    face_children.insert(face_children.end(), chi.begin(), chi.end());
                    ++n_face_blocks;
                }
        }
        if (n_claimed != rows.size()) return fail(B200FE_ERR_INVALID_ARG, "hanging mesh: internal error (%zu of %zu hanging DoFs lie on a hanging face)", n_claimed, rows.size());
    }
    lap("face blocks");
    return B200FE_OK;
}

}  // namespace b200fe

using namespace b200fe;

extern "C" {

int b200fe_hangmesh_create(const b200fe_hangmesh_desc *d, b200fe_hangmesh **out)
{
    B200FE_REQUIRE(d && out, "b200fe_hangmesh_create: null pointer");
    const b200fe_boxmesh_desc &b = d->box;
    if (b.p < 1 || b.p > 8) return fail(B200FE_ERR_UNSUPPORTED, "hanging mesh: degree p=%d outside 1..8", b.p);
    B200FE_REQUIRE(b.n_refine >= 0 && b.n_refine <= 9, "hanging mesh: n_refine out of range");
    B200FE_REQUIRE(b.n_ranks >= 1 && b.rank >= 0 && b.rank < b.n_ranks, "hanging mesh: bad rank %d of %d", b.rank, b.n_ranks);
    if (b.partition != B200FE_PARTITION_P4EST || b.ghosts != B200FE_GHOSTS_MINIMAL)
        return fail(B200FE_ERR_UNSUPPORTED, "hanging mesh: only the p4est partition and the minimal ghost set are built");
    auto m = std::make_unique<HangMesh>();
    for (int k = 0; k < 3; ++k) {
        B200FE_REQUIRE(b.subdivisions[k] >= 1, "hanging mesh: subdivisions must be >= 1");
        B200FE_REQUIRE(b.p2[k] > b.p1[k], "hanging mesh: p2 must exceed p1");
        m->sub[k] = b.subdivisions[k];
        m->p1[k] = b.p1[k];
        m->p2[k] = b.p2[k];
        m->lo[k] = d->refine_lo[k];
        m->hi[k] = d->refine_hi[k];
    }
    m->nref = b.n_refine; m->p = b.p; m->nranks = b.n_ranks; m->rank = b.rank; m->dirichlet = b.dirichlet;
    if (int rc = m->build()) return rc;
    *out = reinterpret_cast<b200fe_hangmesh *>(m.release());
    return B200FE_OK;
}

void b200fe_hangmesh_destroy(b200fe_hangmesh *mesh) { delete reinterpret_cast<HangMesh *>(mesh); }

int b200fe_hangmesh_info(const b200fe_hangmesh *mesh, b200fe_hangmesh_info_t *info)
{
    B200FE_REQUIRE(mesh && info, "b200fe_hangmesh_info: null pointer");
    const HangMesh *m = reinterpret_cast<const HangMesh *>(mesh);
    info->n_cells_global = m->n_cells_global;
    info->n_dofs_global = m->n_dofs_global;
    info->first_cell = m->first_cell;
    info->owned_begin = m->owned_begin;
    info->n_cells_local = (uint32_t)(m->cell_lxyz.size() / 4);
    info->n_owned = (uint32_t)(m->owned_end - m->owned_begin);
    info->n_ghost = (uint32_t)m->ghost_global.size();
    info->n_constrained = (uint32_t)m->constrained.size();
    info->n_hanging_rows = (uint32_t)m->hang_dof.size();
    info->n_hanging_entries = (uint32_t)m->hang_col.size();
    info->n_face_blocks = m->n_face_blocks;
    for (int d = 0; d < 3; ++d) { info->cells[d] = (uint32_t)m->cells[d]; info->h[d] = m->h[d]; info->origin[d] = m->p1[d]; }
    return B200FE_OK;
}

int b200fe_hangmesh_fill(const b200fe_hangmesh *mesh, uint32_t *h_dof_indices, uint32_t *h_constrained, uint64_t *h_ghost_global,
                         int32_t *h_ghost_owner, int32_t *h_cell_lxyz, uint64_t *h_rank_dof_begin, uint32_t *h_hang_dof,
                         uint32_t *h_hang_row_ptr, uint32_t *h_hang_col, double *h_hang_w)
{
    B200FE_REQUIRE(mesh, "b200fe_hangmesh_fill: null mesh");
    const HangMesh *m = reinterpret_cast<const HangMesh *>(mesh);
    auto put = [](void *dst, const auto &v) {
        if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0]));
    };
    put(h_dof_indices, m->dof_indices);
    put(h_constrained, m->constrained);
    put(h_ghost_global, m->ghost_global);
    put(h_ghost_owner, m->ghost_owner);
    put(h_cell_lxyz, m->cell_lxyz);
    put(h_rank_dof_begin, m->rank_dof_begin);
    put(h_hang_dof, m->hang_dof);
    put(h_hang_row_ptr, m->hang_row_ptr);
    put(h_hang_col, m->hang_col);
    put(h_hang_w, m->hang_w);
    return B200FE_OK;
}

int b200fe_hangmesh_fill_faces(const b200fe_hangmesh *mesh, uint32_t *h_face_parents, uint32_t *h_face_children)
{
    B200FE_REQUIRE(mesh, "b200fe_hangmesh_fill_faces: null mesh");
    const HangMesh *m = reinterpret_cast<const HangMesh *>(mesh);
    if (h_face_parents && !m->face_parents.empty()) std::memcpy(h_face_parents, m->face_parents.data(), m->face_parents.size() * sizeof(uint32_t));
    if (h_face_children && !m->face_children.empty()) std::memcpy(h_face_children, m->face_children.data(), m->face_children.size() * sizeof(uint32_t));
    return B200FE_OK;
}

int b200fe_trace_weights(int p, double *h_W)
{
    B200FE_REQUIRE(h_W, "b200fe_trace_weights: null pointer");
    if (p < 1 || p > 8) return fail(B200FE_ERR_UNSUPPORTED, "b200fe_trace_weights: degree p=%d outside 1..8", p);
    std::vector<double> t(p + 1);
    if (int rc = b200fe_basis_1d(p, p + 1, B200FE_QUAD_GLL, nullptr, nullptr, nullptr, t.data(), nullptr)) return rc;
    const std::vector<double> W = trace_weights(p, t);
    std::memcpy(h_W, W.data(), W.size() * sizeof(double));
    return B200FE_OK;
}

}  // extern "C"
