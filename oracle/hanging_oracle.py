"""oracle/hanging_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Literal (loop-based) construction of a box mesh with ONE extra level of local refinement, the conforming
FE_Q(p) space on it with hanging-node constraints, its partition / ghost lists, and the constrained operator
A = C^T Ahat C -- the "element-to-DoF gather/scatter with hanging-node/constraint handling" of BASELINE.json's
north star (config C5).

PARITY UNPINNED.  The reference contains no hanging-node code at all (grep `hanging|FESystem|GaussLobatto` in
/root/reference finds nothing); what deal.II 9.7/9.8 (un-vendored dependency) does is
restated from the library's documented behaviour:
  H1  DoFs live on mesh objects.  A refined cell is replaced by 8 children (child = x + 2y + 4z) with their own
      lines / quads / interiors and new vertices; only the VERTICES of the coarse mesh are shared between the
      two levels.  A fine-level DoF whose support point coincides with a coarse support point that is not a
      coarse vertex (edge / face midpoints for even p) is a separate DoF (constrained with weight 1).
  H2  make_hanging_node_constraints: every fine-level DoF on the closure of an unrefined active cell K is
      constrained to the trace of K:  u_h = sum_j phi_j^K(x_h) u_j  (weights = tensor Lagrange values).
  H3  active_cell_iterators() visits level by level: the unrefined cells in their (z-order) index order, then the
      children, grouped by parent in parent order.  distribute_dofs numbers by first touch in that order,
      vertices -> lines -> quads -> interior inside a cell.
  H4  parallel::distributed: the p4est curve (depth first: a refined cell is replaced in place by its children)
      is cut into pieces of floor(N r / P) cells, then corrected so that no family of 8 siblings is split
      (p4est_partition with partition_for_coarsening: the family goes to the rank holding most of it, ties to
      the lower rank).  Only the families of children are treated that way; p4est would also keep a complete family
      of eight unrefined siblings whole, which the uniform-mesh rule of SURVEY A4 (fe_oracle.BoxMesh.partition) does not
      model -- no difference for 1/2/4/8 ranks on the benchmark meshes, where the cuts fall on family boundaries.
      A DoF belongs to the lowest rank among the active cells that have it; ranks number their DoFs one after the other.
  H5  Dirichlet: every DoF on the domain boundary (boundary id 0, bp3.cc:147-151); it takes precedence over a
      hanging constraint (all parents of such a DoF are boundary DoFs, so both give the value 0).
      Constrained rows of the operator act as identity (portable_laplace_operator.h:171).
Correctness of the constraint weights / index handling is established by patch tests (tests/test_hanging.py):
the constrained space reproduces polynomials of degree <= p exactly, so  A u = -int phi_i Laplace(u)  on the
unconstrained rows -- false for any wrong weight or index.
"""
from __future__ import annotations

import numpy as np

from . import fe_oracle as fe

INVALID = fe.INVALID


class TwoLevelMesh:
    def __init__(self, subdivisions, n_refine, refine_box, p1=(-1.0, -1.0, -1.0), p2=None):
        """BoxMesh(subdivisions, n_refine); its cells (x,y,z) with lo <= (x,y,z) < hi of
        refine_box = ((x0,y0,z0),(x1,y1,z1)) are refined once more."""
        self.base = fe.BoxMesh(subdivisions, n_refine, p1=p1, p2=p2)
        lo, hi = refine_box
        self.lo, self.hi = tuple(int(v) for v in lo), tuple(int(v) for v in hi)
        self.refined = np.zeros(self.base.cells, dtype=bool)
        self.refined[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = True
        # p4est curve: (level, x, y, z) with coordinates at the cell's own level; family id (or -1) alongside
        self.curve, self.family = [], []
        for c in range(self.base.n_cells):
            x, y, z = (int(v) for v in self.base.cell_xyz[c])
            if self.refined[x, y, z]:
                for ch in range(8):
                    self.curve.append((1, 2 * x + (ch & 1), 2 * y + ((ch >> 1) & 1), 2 * z + ((ch >> 2) & 1)))
                    self.family.append(c)
            else:
                self.curve.append((0, x, y, z))
                self.family.append(-1)
        self.n_cells = len(self.curve)

    def partition(self, nranks):
        """H4: subdomain id of every cell of the curve."""
        N = self.n_cells
        starts = [(N * r) // nranks for r in range(nranks + 1)]
        raw = np.searchsorted(np.array(starts[1:]), np.arange(N), side="right").astype(np.int64)
        sub = raw.copy()
        fam = np.array(self.family)
        for f in np.unique(fam[fam >= 0]):
            members = np.nonzero(fam == f)[0]
            ranks, counts = np.unique(raw[members], return_counts=True)
            sub[members] = ranks[np.argmax(counts)]  # argmax returns the first (= lowest rank) maximum
        assert (np.diff(sub) >= 0).all()
        return sub

    def rank_cells(self, subdomain, r):
        """H3: curve indices of rank r's cells in iterator order (level by level, stable)."""
        mine = [i for i in range(self.n_cells) if subdomain[i] == r]
        return [i for i in mine if self.curve[i][0] == 0] + [i for i in mine if self.curve[i][0] == 1]

    def cell_origin_size(self, cell):
        lvl, x, y, z = cell
        h = self.base.h / (2 ** lvl)
        return self.base.p1 + np.array([x, y, z]) * h, h


def node_key(cell, a, b, c, p):
    """H1: identity of the DoF at local lattice position (a,b,c) of `cell`."""
    lvl, x, y, z = cell
    if lvl == 0:
        return (0, x * p + a, y * p + b, z * p + c)
    F = (x * p + a, y * p + b, z * p + c)  # fine lattice, 2p intervals per coarse cell
    if all(f % (2 * p) == 0 for f in F):    # a vertex of the coarse mesh
        return (0, F[0] // 2, F[1] // 2, F[2] // 2)
    return (1,) + F


def build_space(mesh: TwoLevelMesh, p: int, nranks: int = 1, dirichlet: bool = True):
    """Global numbering (H3, H4), hanging rows (H2), Dirichlet set (H5)."""
    n = p + 1
    t, _ = fe.gll_01(n)
    h2l = fe.hierarchic_to_lexicographic(p)
    subdomain = mesh.partition(nranks)
    dof_of, owner, pos = {}, [], []
    owned_range = []
    cell_global = {}
    for r in range(nranks):
        begin = len(pos)
        for ci in mesh.rank_cells(subdomain, r):
            cell = mesh.curve[ci]
            o, h = mesh.cell_origin_size(cell)
            g = np.empty(n ** 3, dtype=np.int64)
            for hh in range(n ** 3):
                l = int(h2l[hh])
                a, b, c = l % n, (l // n) % n, l // (n * n)
                k = node_key(cell, a, b, c, p)
                if k not in dof_of:
                    dof_of[k] = len(pos)
                    pos.append(o + h * np.array([t[a], t[b], t[c]]))
                    owner.append(r)
                g[l] = dof_of[k]
            cell_global[ci] = g
        owned_range.append((begin, len(pos)))
    pos = np.array(pos)
    n_dofs = len(pos)
    B = mesh.base
    on_bnd = np.zeros(n_dofs, dtype=bool)
    if dirichlet:
        on_bnd = (np.abs(pos - B.p1) < 1e-12).any(axis=1) | (np.abs(pos - B.p2) < 1e-12).any(axis=1)
    # H2: fine-level DoFs on the closure of an unrefined cell
    rows = {}
    fine = {k: d for k, d in dof_of.items() if k[0] == 1}
    coarse_cells = {tuple(mesh.curve[ci][1:]): ci for ci in range(mesh.n_cells) if mesh.curve[ci][0] == 0}
    for k, d in fine.items():
        F = k[1:]
        # coarse cells whose closure contains the point: c*2p <= f <= (c+1)*2p per axis
        cand = [[c for c in ([f // (2 * p), f // (2 * p) - 1] if f % (2 * p) == 0 else [f // (2 * p)]) if 0 <= c < B.cells[dd]]
                for dd, f in enumerate(F)]
        for cx in cand[0]:
            for cy in cand[1]:
                for cz in cand[2]:
                    if (cx, cy, cz) not in coarse_cells:
                        continue
                    K = mesh.curve[coarse_cells[(cx, cy, cz)]]
                    o, h = mesh.cell_origin_size(K)
                    xi = (pos[d] - o) / h
                    assert (xi > -1e-12).all() and (xi < 1 + 1e-12).all()
                    w1 = [fe.lagrange_values(t, np.array([xi[dd]]))[0] for dd in range(3)]
                    gK = cell_global[coarse_cells[(cx, cy, cz)]]
                    entries = {}
                    for l in range(n ** 3):
                        a, b, c = l % n, (l // n) % n, l // (n * n)
                        w = w1[0][a] * w1[1][b] * w1[2][c]
                        if abs(w) > 1e-14:
                            entries[int(gK[l])] = w
                    if d in rows:  # conformity: every unrefined neighbour gives the same trace
                        assert set(rows[d]) == set(entries) and all(abs(rows[d][j] - entries[j]) < 1e-12 for j in entries)
                    else:
                        rows[d] = entries
    hanging = np.array(sorted(d for d in rows if not on_bnd[d]), dtype=np.int64)  # H5: Dirichlet wins
    parents = {j for d in hanging for j in rows[int(d)]}
    assert not parents & set(rows), "constraint chains"
    keys = [None] * n_dofs  # H1 identity of every global DoF: numbering-independent name
    for k, d in dof_of.items():
        keys[d] = k
    return dict(p=p, n_dofs=n_dofs, pos=pos, keys=keys, owner=np.array(owner), owned_range=owned_range, subdomain=subdomain,
                cell_global=cell_global, on_bnd=on_bnd, hanging=hanging,
                rows={int(d): {j: w for j, w in sorted(rows[int(d)].items()) if not on_bnd[j]} for d in hanging})


def rank_data(mesh: TwoLevelMesh, sp, rank: int):
    """What rank `rank` hands to the operator: local index table (Dirichlet DoFs masked, hanging DoFs are
    ordinary entries), ghost lists, owned constrained list, local CSR of the hanging rows it needs."""
    begin, end = sp["owned_range"][rank]
    cells = mesh.rank_cells(sp["subdomain"], rank)
    n3 = (sp["p"] + 1) ** 3
    G = np.stack([sp["cell_global"][ci] for ci in cells]) if cells else np.zeros((0, n3), np.int64)
    on_bnd = sp["on_bnd"]
    hang_set = set(sp["hanging"].tolist())
    needed = sorted({int(g) for g in G.ravel() if int(g) in hang_set})
    touched = set(G.ravel().tolist())
    for d in needed:
        touched.update(sp["rows"][d])
    ghosts = np.array(sorted(g for g in touched if not begin <= g < end), dtype=np.int64)
    n_owned = end - begin
    g2l = {int(g): n_owned + i for i, g in enumerate(ghosts)}
    loc = lambda g: g - begin if begin <= g < end else g2l[g]
    idx = np.empty(G.shape, dtype=np.uint32)
    for a in range(G.shape[0]):
        for b in range(G.shape[1]):
            g = int(G[a, b])
            idx[a, b] = INVALID if on_bnd[g] else loc(g)
    local_set = set(G.ravel().tolist())
    constrained = np.array(sorted(g - begin for g in range(begin, end)
                                  if (on_bnd[g] and g in local_set) or g in hang_set), dtype=np.uint32)
    row_ptr, col, wgt = [0], [], []
    for d in needed:
        for j, w in sp["rows"][d].items():
            col.append(loc(j))
            wgt.append(w)
        row_ptr.append(len(col))
    return dict(rank=rank, cells=[mesh.curve[ci] for ci in cells], n_owned=n_owned, n_ghost=len(ghosts), owned_begin=begin,
                ghost_global=ghosts, ghost_owner=sp["owner"][ghosts] if len(ghosts) else np.zeros(0, np.int64),
                dof_indices=idx, cell_global=G, constrained=constrained,
                hang_dof=np.array([loc(d) for d in needed], dtype=np.uint32), hang_row_ptr=np.array(row_ptr, dtype=np.uint32),
                hang_col=np.array(col, dtype=np.uint32), hang_w=np.array(wgt, dtype=np.float64))


def cell_nodes(mesh: TwoLevelMesh, cells, p_geo: int = 1, deform=None):
    """Mapping support points X[c, d, z, y, x] of the given (level,x,y,z) cells (see fe.cell_nodes)."""
    t, _ = fe.gll_01(p_geo + 1)
    out = np.empty((len(cells), 3, p_geo + 1, p_geo + 1, p_geo + 1))
    for c, cell in enumerate(cells):
        o, h = mesh.cell_origin_size(cell)
        Z, Y, X = np.meshgrid(o[2] + h[2] * t, o[1] + h[1] * t, o[0] + h[0] * t, indexing="ij")
        pts = np.stack([X, Y, Z], axis=-1)
        if deform is not None:
            pts = deform(pts)
        out[c] = np.moveaxis(pts, -1, 0)
    return out


def distribute(rd, u):
    """u_hat = C u on a local vector: hanging entries from their parents (Dirichlet parents count as 0)."""
    v = u.copy()
    for r, d in enumerate(rd["hang_dof"]):
        a, b = rd["hang_row_ptr"][r], rd["hang_row_ptr"][r + 1]
        v[d] = np.dot(rd["hang_w"][a:b], u[rd["hang_col"][a:b]])
    return v


def condense(rd, v):
    """C^T v on a local vector: hanging rows added to their parents, then zeroed."""
    out = v.copy()
    for r, d in enumerate(rd["hang_dof"]):
        a, b = rd["hang_row_ptr"][r], rd["hang_row_ptr"][r + 1]
        np.add.at(out, rd["hang_col"][a:b], rd["hang_w"][a:b] * v[d])
        out[d] = 0.0
    return out


def op_apply(rd, bas, G, u, JxW=None, laplace=True, mass=False):
    """Single-rank constrained operator: dst = C^T Ahat C u, constrained rows (hanging + Dirichlet) = identity."""
    uh = distribute(rd, u)
    vh = fe.op_apply(uh, rd["dof_indices"], bas, G, JxW, laplace=laplace, mass=mass, n_local=len(u), constrained=np.zeros(0, np.uint32))
    v = condense(rd, vh)
    v[rd["constrained"]] = u[rd["constrained"]]
    return v


def rhs_one(rd, bas, JxW):
    """b = C^T int phi_i, constrained rows 0."""
    b = condense(rd, fe.rhs_one(rd, bas, JxW))
    b[rd["constrained"]] = 0.0
    return b


def distributed_apply(rds, bas, Gs, u_global, JxWs=None, laplace=True, mass=False):
    """The multi-rank vmult sequence on one process (what b200fe_op_vmult does with a halo attached):
    update_ghost_values -> distribute -> cells -> condense -> compress(add) -> identity on constrained rows.
    rds: per-rank dicts (rank_data(), or the same fields taken from the product's mesh objects)."""
    dst = np.zeros_like(u_global)
    for r, rd in enumerate(rds):
        b0, n_own = rd["owned_begin"], rd["n_owned"]
        gg = np.asarray(rd["ghost_global"], dtype=np.int64)
        u = np.concatenate([u_global[b0:b0 + n_own], u_global[gg]])
        uh = distribute(rd, u)
        vh = fe.op_apply(uh, rd["dof_indices"], bas, Gs[r], None if JxWs is None else JxWs[r], laplace=laplace, mass=mass,
                         n_local=len(u), constrained=np.zeros(0, np.uint32))
        v = condense(rd, vh)
        dst[b0:b0 + n_own] += v[:n_own]
        np.add.at(dst, gg, v[n_own:])
    for rd in rds:
        c = rd["owned_begin"] + np.asarray(rd["constrained"], dtype=np.int64)
        dst[c] = u_global[c]
    return dst


# --------------------------------------------------------------------------
# Face-structured form of the same constraints (tensor-product trace interpolation)
# --------------------------------------------------------------------------
def trace_weights(p: int):
    """W[rel, j] = l_j((half + t_a) / 2) for the fine lattice positions rel = half*p + a = 0..2p of a coarse interval:
    1-D Lagrange values of the coarse GLL basis; exact unit rows where a fine node coincides with a coarse one."""
    t, _ = fe.gll_01(p + 1)
    W = np.zeros((2 * p + 1, p + 1))
    for rel in range(2 * p + 1):
        if rel == 0:
            W[rel, 0] = 1.0
        elif rel == 2 * p:
            W[rel, p] = 1.0
        elif rel == p and p % 2 == 0:
            W[rel, p // 2] = 1.0
        else:
            half, a = divmod(rel, p)
            W[rel] = fe.lagrange_values(t, np.array([0.5 * (half + t[a])]))[0]
    return W


def face_blocks(mesh: TwoLevelMesh, sp, rd):
    """The hanging rows of rank data `rd` grouped by coarse face: every hanging DoF lies on a face shared by an
    unrefined cell K and a refined cell; the (2p+1)^2 fine nodes of that face are the tensor-product interpolation
    u_f(a', b') = sum_ab W[a', a] W[b', b] u_K(a, b) of the (p+1)^2 coarse face nodes.
    Blocks are listed in (K position on the base mesh, axis, side) order; a hanging DoF shared by several faces is
    assigned to the first block that needs it.  Returns parents [n, (p+1)^2] and children [n, (2p+1)^2] in local
    indices (a / a' fastest along the lower in-face axis), INVALID = Dirichlet or absent parent (value 0) / not a child of this
    block (coarse vertex, Dirichlet, not needed by this rank, or assigned to an earlier block)."""
    p = sp["p"]
    nm, nf = p + 1, 2 * p + 1
    B = mesh.base
    key_to_global = {k: d for d, k in enumerate(sp["keys"])}
    begin, n_own = rd["owned_begin"], rd["n_owned"]
    gl = {int(g): n_own + i for i, g in enumerate(rd["ghost_global"])}
    loc = lambda g: g - begin if begin <= g < begin + n_own else gl.get(g)
    needed = {}  # global hanging dof -> local index, for the rows this rank holds
    inv = {v: k for k, v in gl.items()}
    for h in rd["hang_dof"]:
        h = int(h)
        needed[h + begin if h < n_own else inv[h]] = h
    claimed = set()
    parents, children = [], []
    for c in range(B.n_cells):
        K = tuple(int(v) for v in B.cell_xyz[c])
        if mesh.refined[K]:
            continue
        for axis in range(3):
            for side in (0, 1):
                N = list(K)
                N[axis] += 1 if side else -1
                if not (0 <= N[axis] < B.cells[axis]) or not mesh.refined[tuple(N)]:
                    continue
                d1, d2 = [d for d in range(3) if d != axis]
                par = np.full(nm * nm, INVALID, dtype=np.uint32)
                chi = np.full(nf * nf, INVALID, dtype=np.uint32)
                for b in range(nm):
                    for a in range(nm):
                        X = [0, 0, 0]
                        X[axis], X[d1], X[d2] = (K[axis] + side) * p, K[d1] * p + a, K[d2] * p + b
                        g = key_to_global[(0, X[0], X[1], X[2])]
                        if not sp["on_bnd"][g] and loc(g) is not None:
                            par[a + nm * b] = loc(g)
                any_child = False
                for b in range(nf):
                    for a in range(nf):
                        F = [0, 0, 0]
                        F[axis], F[d1], F[d2] = (K[axis] + side) * 2 * p, K[d1] * 2 * p + a, K[d2] * 2 * p + b
                        if all(f % (2 * p) == 0 for f in F):
                            continue
                        g = key_to_global.get((1, F[0], F[1], F[2]))
                        if g is None or g not in needed or g in claimed:
                            continue
                        claimed.add(g)
                        chi[a + nf * b] = needed[g]
                        any_child = True
                if any_child:
                    parents.append(par)
                    children.append(chi)
    assert claimed == set(needed), "every hanging DoF of a box-refined mesh lies on a hanging face"
    n = len(parents)
    return (np.array(parents, dtype=np.uint32).reshape(n, nm * nm), np.array(children, dtype=np.uint32).reshape(n, nf * nf))


def distribute_faces(p, parents, children, u):
    """u_hat = C u through the face blocks (what the face kernels do): two 1-D interpolations per face."""
    W = trace_weights(p)
    nm, nf = p + 1, 2 * p + 1
    v = u.copy()
    for par, chi in zip(parents, children):
        P = np.where(par != INVALID, u[np.where(par != INVALID, par, 0)], 0.0).reshape(nm, nm)  # [b, a]
        F = W @ P @ W.T                                                                        # [b', a']
        ok = chi != INVALID
        v[chi[ok]] = F.ravel()[ok]
    return v


def condense_faces(p, parents, children, v):
    """C^T v through the face blocks: children gathered and zeroed, transposed interpolation added to the parents."""
    W = trace_weights(p)
    nm, nf = p + 1, 2 * p + 1
    out = v.copy()
    for par, chi in zip(parents, children):
        ok = chi != INVALID
        C = np.zeros(nf * nf)
        C[ok] = v[chi[ok]]
        out[chi[ok]] = 0.0
        P = W.T @ C.reshape(nf, nf) @ W
        okp = par != INVALID
        np.add.at(out, par[okp], P.ravel()[okp])
    return out
