"""oracle/fe_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

numpy restatement of everything the reference's BP ("L-vector") path relies on:
1-D bases, box mesh + FE_Q DoF numbering + partition + ghost lists, geometric
factors, the operator apply with masked gather / additive scatter, and the
deal.II CG loop.  Small cases only (pure-Python loops in the numbering).

Reference files followed (paths relative to /root/reference):
  * cell kernel + gather/scatter   CEED_bp/include/bk3_kokkos_kernel.h:115-141, 147-353, 357-381
  * operator wrapper (vmult)       CEED_bp/include/portable_laplace_operator.h:124-172
  * geometric factors              CEED_bp/include/portable_laplace_operator.h:239-302 with the
                                   authoritative math of bakeoff_problems_dealii/include/
                                   portable_laplace_operator.h:227-258
  * Dirichlet masks                CEED_bp/include/portable_laplace_operator.h:304-394
  * mesh sweep, rhs, CG protocol   CEED_bp/src/bp3.cc:184-239, 246-329, 433-488
  * Helmholtz variant, src vector  bp5_kokkos/benchmark.cc:62-137, 341-347, 355
  * partition by equal blocks      bp5_kokkos/create_triangulation.h:44-52

Everything deal.II supplies (FE_Q numbering, GLL support points, active-cell
order, ownership rule, SolverCG) is an un-vendored dependency (deal.II 9.7/9.8);
it is restated here from the published library behaviour listed in SURVEY.md
Appendix A (A1-A9).  PARITY PINS: the CG iteration counts and reduction rates
of CEED_bp/results/1xGH200_P4.txt:636-640 (tests/test_oracle_pins.py) pin the
1-D bases, the operator, the Dirichlet treatment, the right-hand side and the
CG loop.  The DoF numbering / ghost lists (A1-A6) are "parity unpinned" against
deal.II itself: no reference test holds them; they are pinned only by global
DoF counts and by consistency checks.
"""
from __future__ import annotations

import numpy as np

INVALID = np.uint32(0xFFFFFFFF)  # numbers::invalid_unsigned_int (portable_laplace_operator.h:380-384)


# --------------------------------------------------------------------------
# 1-D bases
# --------------------------------------------------------------------------
def gauss_legendre_01(n: int):
    """Gauss-Legendre points/weights on [0,1], ascending, sum(w)=1 (QGauss<1>(n))."""
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def gll_01(n: int):
    """n Gauss-Lobatto-Legendre points/weights on [0,1] (FE_Q support points, QGaussLobatto)."""
    if n == 1:
        return np.array([0.5]), np.array([1.0])
    if n == 2:
        return np.array([0.0, 1.0]), np.array([0.5, 0.5])
    N = n - 1
    PN = np.polynomial.legendre.Legendre.basis(N)
    xi = np.sort(np.real(PN.deriv().roots()))
    # Newton polish on P_N'(x) = 0
    d1, d2 = PN.deriv(1), PN.deriv(2)
    for _ in range(3):
        xi = xi - d1(xi) / d2(xi)
    x = np.concatenate([[-1.0], xi, [1.0]])
    w = 2.0 / (N * (N + 1) * PN(x) ** 2)
    x = 0.5 * (x - x[::-1])  # enforce symmetry
    return 0.5 * (x + 1.0), 0.5 * w


def lagrange_values(nodes, x):
    """V[q, i] = l_i(x_q) for the Lagrange basis through `nodes`."""
    nodes = np.asarray(nodes, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    n = len(nodes)
    V = np.ones((len(x), n))
    for i in range(n):
        for m in range(n):
            if m != i:
                V[:, i] *= (x - nodes[m]) / (nodes[i] - nodes[m])
    return V


def lagrange_derivs(nodes, x):
    """Dv[q, i] = l_i'(x_q)."""
    nodes = np.asarray(nodes, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    n = len(nodes)
    Dv = np.zeros((len(x), n))
    for i in range(n):
        for k in range(n):
            if k == i:
                continue
            term = np.full(len(x), 1.0 / (nodes[i] - nodes[k]))
            for m in range(n):
                if m != i and m != k:
                    term *= (x - nodes[m]) / (nodes[i] - nodes[m])
            Dv[:, i] += term
    return Dv


def basis_1d(p: int, nq: int, quad: str = "gauss"):
    """1-D data of one operator.

    Returns dict with
      xq, wq     quadrature points / weights on [0,1]
      B[q, i]    value of FE_Q(p) shape function i (GLL nodes) at point q
      D[q, n]    derivative at point q of the collocation Lagrange function through point n
      Bg[q, i]   derivative of shape function i at point q (used by geometry and diagonals)
    deal.II arrays (bk3_kokkos_kernel.h:161,243) are the transposes:
      shape_values[i*nq+q] = B[q,i],  co_shape_gradients[n*nq+q] = D[q,n].
    quad = 'gauss' (QGauss(nq)) or 'gll' (collocated BP5: nq = p+1 and B = identity).
    """
    nodes, _ = gll_01(p + 1)
    if quad == "gauss":
        xq, wq = gauss_legendre_01(nq)
    elif quad == "gll":
        xq, wq = gll_01(nq)
    else:
        raise ValueError(quad)
    B = lagrange_values(nodes, xq)
    if quad == "gll" and nq == p + 1:
        B = np.eye(nq)
    return dict(p=p, nq=nq, quad=quad, nodes=nodes, xq=xq, wq=wq, B=B,
                D=lagrange_derivs(xq, xq), Bg=lagrange_derivs(nodes, xq))


# --------------------------------------------------------------------------
# FE_Q local numbering
# --------------------------------------------------------------------------
def hierarchic_to_lexicographic(p: int):
    """h2l[h] = lexicographic local index (x fastest) of hierarchical local DoF h of FE_Q<3>(p)."""
    n = p + 1
    m = p - 1  # dofs per line
    h2l = []
    lex = lambda x, y, z: x + n * (y + n * z)
    # vertices v = x + 2y + 4z
    for v in range(8):
        h2l.append(lex(p * (v & 1), p * ((v >> 1) & 1), p * ((v >> 2) & 1)))
    # lines: 0:(v0-v2) 1:(v1-v3) 2:(v0-v1) 3:(v2-v3); 4-7 same at z=1; 8-11 vertical
    for z in (0, p):
        for i in range(m):
            h2l.append(lex(0, i + 1, z))
        for i in range(m):
            h2l.append(lex(p, i + 1, z))
        for i in range(m):
            h2l.append(lex(i + 1, 0, z))
        for i in range(m):
            h2l.append(lex(i + 1, p, z))
    for (x, y) in ((0, 0), (p, 0), (0, p), (p, p)):
        for i in range(m):
            h2l.append(lex(x, y, i + 1))
    # quads: x-faces run (y fastest, z), y-faces (z fastest, x), z-faces (x fastest, y)
    for x in (0, p):
        for i in range(m):
            for j in range(m):
                h2l.append(lex(x, j + 1, i + 1))
    for y in (0, p):
        for i in range(m):
            for j in range(m):
                h2l.append(lex(i + 1, y, j + 1))
    for z in (0, p):
        for i in range(m):
            for j in range(m):
                h2l.append(lex(j + 1, i + 1, z))
    # interior, x fastest
    for i in range(m):
        for j in range(m):
            for k in range(m):
                h2l.append(lex(k + 1, j + 1, i + 1))
    h2l = np.array(h2l, dtype=np.int64)
    assert len(h2l) == n ** 3 and len(set(h2l.tolist())) == n ** 3
    return h2l


# --------------------------------------------------------------------------
# Box mesh (bp3.cc:452-488, create_triangulation.h:17-29)
# --------------------------------------------------------------------------
def _morton3(x, y, z, nbits):
    code = 0
    for b in range(nbits):
        code |= ((x >> b) & 1) << (3 * b) | ((y >> b) & 1) << (3 * b + 1) | ((z >> b) & 1) << (3 * b + 2)
    return code


class BoxMesh:
    """subdivided_hyper_rectangle(subdivisions, p1, p2) + refine_global(n_refine).

    Active cells are ordered coarse-cell-lexicographic (x fastest) and, inside each
    coarse cell, along the z-order curve (child = x + 2y + 4z), recursively.
    """

    def __init__(self, subdivisions, n_refine, p1=(-1.0, -1.0, -1.0), p2=None):
        self.sub = tuple(int(s) for s in subdivisions)
        self.n_refine = int(n_refine)
        f = 1 << self.n_refine
        self.cells = tuple(s * f for s in self.sub)
        self.p1 = np.array(p1, dtype=np.float64)
        if p2 is None:  # bp3.cc: side 1.9 per coarse cell
            p2 = [a + 1.9 * s for a, s in zip(p1, self.sub)]
        self.p2 = np.array(p2, dtype=np.float64)
        self.h = (self.p2 - self.p1) / np.array(self.cells)
        self.n_cells = int(np.prod(self.cells))
        cx, cy, cz = self.cells
        order = np.empty((cx, cy, cz), dtype=np.int64)
        per_coarse = 8 ** self.n_refine
        for x in range(cx):
            for y in range(cy):
                for z in range(cz):
                    X, Y, Z = x >> self.n_refine, y >> self.n_refine, z >> self.n_refine
                    coarse = X + self.sub[0] * (Y + self.sub[1] * Z)
                    order[x, y, z] = coarse * per_coarse + _morton3(x & (f - 1), y & (f - 1), z & (f - 1), self.n_refine)
        self.pos_of_cell = order  # active index of cell (x,y,z)
        self.cell_xyz = np.empty((self.n_cells, 3), dtype=np.int64)
        for x in range(cx):
            for y in range(cy):
                for z in range(cz):
                    self.cell_xyz[order[x, y, z]] = (x, y, z)

    @staticmethod
    def bp3_cycle(cycle):
        """Mesh of bp3.cc:443-473 for sweep index `cycle`."""
        n_refine, rem = cycle // 3, cycle % 3
        sub = [2 if d < rem else 1 for d in range(3)]
        return BoxMesh(sub, n_refine)

    def partition(self, nranks, scheme="p4est"):
        """subdomain id of every active cell.  p4est: first cell of rank r = floor(N r / P)
        (SURVEY A4); 'blocks': active_cell_index / ceil(N/P) (create_triangulation.h:44-51)."""
        idx = np.arange(self.n_cells)
        if scheme == "p4est":
            starts = [(self.n_cells * r) // nranks for r in range(nranks + 1)]
            return np.searchsorted(np.array(starts[1:]), idx, side="right").astype(np.int32)
        per = (self.n_cells + nranks - 1) // nranks
        return (idx // per).astype(np.int32)


# --------------------------------------------------------------------------
# DoF numbering + ownership + ghost lists
# --------------------------------------------------------------------------
def distribute_dofs(mesh: BoxMesh, p: int, nranks: int = 1, scheme: str = "p4est"):
    """Literal simulation of DoFHandler::distribute_dofs on a distributed triangulation.

    Returns dict:
      lattice_of_global  (n_dofs,3) lattice coordinates of every global DoF
      global_of_lattice  array [X,Y,Z] -> global DoF
      owner              rank owning each global DoF
      owned_range        list of (begin,end) per rank (contiguous, rank order)
      subdomain          subdomain id per active cell
    """
    n = p + 1
    h2l = hierarchic_to_lexicographic(p)
    subdomain = mesh.partition(nranks, scheme)
    dims = tuple(c * p + 1 for c in mesh.cells)
    # lowest subdomain id touching each lattice point
    min_rank = np.full(dims, nranks, dtype=np.int64)
    for c in range(mesh.n_cells):
        x, y, z = mesh.cell_xyz[c]
        sl = (slice(x * p, x * p + n), slice(y * p, y * p + n), slice(z * p, z * p + n))
        min_rank[sl] = np.minimum(min_rank[sl], subdomain[c])
    glob = np.full(dims, -1, dtype=np.int64)
    owned_range = []
    next_free = 0
    for r in range(nranks):
        # step 1: first-touch enumeration over the cells of rank r (vertices, lines, quads, interior)
        local = {}
        for c in np.nonzero(subdomain == r)[0]:
            x, y, z = mesh.cell_xyz[c]
            for h in range(n ** 3):
                l = int(h2l[h])
                key = (x * p + l % n, y * p + (l // n) % n, z * p + l // (n * n))
                if key not in local:
                    local[key] = len(local)
        # step 2: drop interface DoFs that a lower rank owns, compact, shift
        begin = next_free
        for key in sorted(local, key=local.get):
            if min_rank[key] == r:
                glob[key] = next_free
                next_free += 1
        owned_range.append((begin, next_free))
    assert (glob >= 0).all() and next_free == int(np.prod(dims))
    lattice = np.empty((next_free, 3), dtype=np.int64)
    X, Y, Z = np.meshgrid(*[np.arange(d) for d in dims], indexing="ij")
    lattice[glob.ravel()] = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    return dict(p=p, lattice_of_global=lattice, global_of_lattice=glob, owner=min_rank.ravel()[np.argsort(glob.ravel())],
                owned_range=owned_range, subdomain=subdomain, dims=dims)


def rank_data_single_fast(mesh: BoxMesh, p: int, dirichlet: bool = True):
    """Vectorised equivalent of rank_data(mesh, distribute_dofs(mesh, p, 1), 0) for ONE rank (no Python loop over
    DoFs): the first-touch number of a lattice point is the rank of its smallest key (active cell index, hierarchical
    local index) among all lattice points.  Checked against the literal simulation in tests/test_mesh.py; used by
    bench.py's CPU arm so that the reference arm stands on oracle/ alone."""
    n = p + 1
    n3 = n ** 3
    h2l = hierarchic_to_lexicographic(p)
    h_of_l = np.empty(n3, dtype=np.int64)
    h_of_l[h2l] = np.arange(n3)
    dims = tuple(c * p + 1 for c in mesh.cells)
    l = np.arange(n3)
    lx, ly, lz = l % n, (l // n) % n, l // (n * n)
    cx, cy, cz = mesh.cell_xyz[:, 0], mesh.cell_xyz[:, 1], mesh.cell_xyz[:, 2]  # active order
    X = cx[:, None] * p + lx[None, :]
    Y = cy[:, None] * p + ly[None, :]
    Z = cz[:, None] * p + lz[None, :]
    lat = (X * dims[1] + Y) * dims[2] + Z                       # lattice id, [cell][lexicographic local]
    key = np.arange(mesh.n_cells, dtype=np.int64)[:, None] * n3 + h_of_l[None, :]
    n_lat = int(np.prod(dims))
    first = np.full(n_lat, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(first, lat.ravel(), key.ravel())
    glob = np.empty(n_lat, dtype=np.int64)
    glob[np.argsort(first, kind="stable")] = np.arange(n_lat)
    G = glob[lat]
    on_bdry = ((X == 0) | (Y == 0) | (Z == 0) | (X == dims[0] - 1) | (Y == dims[1] - 1) | (Z == dims[2] - 1)) & dirichlet
    idx = np.where(on_bdry, INVALID, G.astype(np.uint32)).astype(np.uint32)
    constrained = np.unique(G[on_bdry]).astype(np.uint32)
    return dict(rank=0, cells=np.arange(mesh.n_cells), n_owned=n_lat, n_ghost=0, owned_begin=0, ghost_global=np.zeros(0, np.int64),
                ghost_owner=np.zeros(0, np.int64), dof_indices=idx, cell_global=G, constrained=constrained)


def cell_dofs_global(mesh: BoxMesh, dofs, c: int):
    """Global DoF indices of active cell c in lexicographic local order (k = x fastest)."""
    p = dofs["p"]
    n = p + 1
    x, y, z = mesh.cell_xyz[c]
    blk = dofs["global_of_lattice"][x * p:x * p + n, y * p:y * p + n, z * p:z * p + n]
    return np.transpose(blk, (2, 1, 0)).ravel()  # [z][y][x]


def rank_data(mesh: BoxMesh, dofs, rank: int, ghost_set: str = "minimal", dirichlet: bool = True):
    """Everything rank `rank` hands to the operator (portable_laplace_operator.h:304-394).

    dof_indices[cell, local] : partitioner-local index (owned first, then ghosts sorted by global
                               index) or INVALID on constrained DoFs.
    ghost_set 'minimal'  = DoFs gathered by owned cells but owned elsewhere;
              'relevant' = every DoF of the one-cell ghost layer (deal.II's locally relevant set, A5).
    """
    p = dofs["p"]
    n = p + 1
    begin, end = dofs["owned_range"][rank]
    my_cells = np.nonzero(dofs["subdomain"] == rank)[0]
    G = np.stack([cell_dofs_global(mesh, dofs, c) for c in my_cells]) if len(my_cells) else np.zeros((0, n ** 3), np.int64)
    touched = np.unique(G)
    if ghost_set == "relevant":
        glat = dofs["global_of_lattice"]
        mask = np.zeros(mesh.cells, dtype=bool)
        for c in my_cells:
            x, y, z = mesh.cell_xyz[c]
            mask[max(x - 1, 0):x + 2, max(y - 1, 0):y + 2, max(z - 1, 0):z + 2] = True
        rel = []
        for c in range(mesh.n_cells):
            x, y, z = mesh.cell_xyz[c]
            if mask[x, y, z]:
                rel.append(glat[x * p:x * p + n, y * p:y * p + n, z * p:z * p + n].ravel())
        touched = np.unique(np.concatenate(rel)) if rel else touched
    ghosts = touched[(touched < begin) | (touched >= end)]  # sorted by global index
    n_owned = end - begin
    g2l = {int(g): n_owned + i for i, g in enumerate(ghosts)}

    lat = dofs["lattice_of_global"]
    dims = dofs["dims"]

    def constrained(g):
        X, Y, Z = lat[g]
        return dirichlet and (X == 0 or Y == 0 or Z == 0 or X == dims[0] - 1 or Y == dims[1] - 1 or Z == dims[2] - 1)

    idx = np.empty(G.shape, dtype=np.uint32)
    for a in range(G.shape[0]):
        for b in range(G.shape[1]):
            g = int(G[a, b])
            if constrained(g):
                idx[a, b] = INVALID
            else:
                idx[a, b] = g - begin if begin <= g < end else g2l[g]
    owned_constrained = np.array([g - begin for g in range(begin, end) if constrained(g)], dtype=np.uint32)
    ghost_owner = dofs["owner"][ghosts] if len(ghosts) else np.zeros(0, np.int64)
    return dict(rank=rank, cells=my_cells, n_owned=n_owned, n_ghost=len(ghosts), owned_begin=begin,
                ghost_global=ghosts, ghost_owner=ghost_owner, dof_indices=idx, cell_global=G,
                constrained=owned_constrained)


def exchange_lists(all_rank_data):
    """Per rank: recv[(peer)] = slice of ghost segment owned by peer; send[(peer)] = owned local
    indices the peer ghosts, sorted by global index (deal.II Partitioner::import_indices, A5)."""
    out = []
    for rd in all_rank_data:
        recv, send = {}, {}
        for peer in np.unique(rd["ghost_owner"]):
            sel = np.nonzero(rd["ghost_owner"] == peer)[0]
            recv[int(peer)] = (int(sel[0]), int(sel[-1]) + 1)
            assert (np.diff(sel) == 1).all()
        out.append(dict(recv=recv, send=send))
    for rd, ex in zip(all_rank_data, out):
        for peer, (a, b) in ex["recv"].items():
            owner = all_rank_data[peer]
            out[peer]["send"][rd["rank"]] = (rd["ghost_global"][a:b] - owner["owned_begin"]).astype(np.uint32)
    return out


# --------------------------------------------------------------------------
# Geometry (MappingQ) -> node coordinates -> G, JxW   
# --------------------------------------------------------------------------
def cell_nodes(mesh: BoxMesh, cells, p_geo: int, deform=None):
    """Mapping support points X[c, d, z, y, x] (GLL lattice of degree p_geo), optionally deformed
    by a smooth map deform(xyz[...,3]) -> xyz (check_bk3.cc:50-52 style)."""
    t, _ = gll_01(p_geo + 1)
    out = np.empty((len(cells), 3, p_geo + 1, p_geo + 1, p_geo + 1))
    for a, c in enumerate(cells):
        x, y, z = mesh.cell_xyz[c]
        xs = mesh.p1[0] + (x + t) * mesh.h[0]
        ys = mesh.p1[1] + (y + t) * mesh.h[1]
        zs = mesh.p1[2] + (z + t) * mesh.h[2]
        Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
        pts = np.stack([X, Y, Z], axis=-1)
        if deform is not None:
            pts = deform(pts)
        out[a] = np.moveaxis(pts, -1, 0)
    return out


def geometric_factors(nodes, p_geo: int, bas):
    """From mapping support points to the kernel's geometric data at the nq^3 points.

    J[a][b] = d x_a / d xi_b with xi = (x^, y^, z^).  The kernel's reference directions are
    (r, s, t) = (slowest, middle, fastest local index) = (z^, y^, x^) (bk3_kokkos_kernel.h:212-253),
    so G_c for c = (rr, rs, rt, ss, st, tt) is JxW * (K K^T)[a,b] with K = J^{-1} and rows taken in
    the order (z^, y^, x^).  This is the mathematically correct pairing
    (bakeoff_problems_dealii/include/portable_laplace_operator.h:227-258); on the reference's cube
    cells it coincides with CEED_bp/include/portable_laplace_operator.h:281-290.
    Returns G[c, 6, nq^3] (point index p*nq^2+q*nq+r, p<->z), JxW[c, nq^3].
    """
    tg, _ = gll_01(p_geo + 1)
    xq, wq = bas["xq"], bas["wq"]
    V = lagrange_values(tg, xq)   # [q, node]
    dV = lagrange_derivs(tg, xq)
    # nodes[c, d, z, y, x]
    dx = np.einsum("cdzyx,rz,qy,px->cdrqp", nodes, V, V, dV)  # d/dx^ ; point index [r=z][q=y][p=x]
    dy = np.einsum("cdzyx,rz,qy,px->cdrqp", nodes, V, dV, V)
    dz = np.einsum("cdzyx,rz,qy,px->cdrqp", nodes, dV, V, V)
    J = np.stack([dx, dy, dz], axis=2)  # [c, a(real), b(ref), Z, Y, X]
    J = np.moveaxis(J, (1, 2), (-2, -1))  # [c, Z, Y, X, a, b]
    det = np.linalg.det(J)
    K = np.linalg.inv(J)  # K[b(ref), a(real)]
    W = np.einsum("r,q,p->rqp", wq, wq, wq)
    JxW = det * W
    KKt = np.einsum("...ba,...ca->...bc", K, K)  # (ref b, ref c), order (x^, y^, z^)
    perm = [2, 1, 0]  # (r,s,t) = (z^, y^, x^)
    comps = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
    nc = nodes.shape[0]
    G = np.empty((nc, 6) + det.shape[1:])
    for ci, (a, b) in enumerate(comps):
        G[:, ci] = JxW * KKt[..., perm[a], perm[b]]
    nq3 = len(xq) ** 3
    return G.reshape(nc, 6, nq3), JxW.reshape(nc, nq3)


# --------------------------------------------------------------------------
# Operator apply (masked gather -> cell kernel -> additive scatter)
# --------------------------------------------------------------------------
def cell_kernel(u, bas, G=None, JxW=None, laplace=True, mass=False):
    """u[c, nm, nm, nm] (z,y,x) -> out[c, nm, nm, nm].  bk3_kokkos_kernel.h:147-353."""
    B, D = bas["B"], bas["D"]
    nq = bas["nq"]
    v = np.einsum("cijk,pi,qj,rk->cpqr", u, B, B, B, optimize=True)
    w = np.zeros_like(v)
    if laplace:
        g = G.reshape(-1, 6, nq, nq, nq)
        qr = np.einsum("pn,cnqr->cpqr", D, v)
        qs = np.einsum("qn,cpnr->cpqr", D, v)
        qt = np.einsum("rn,cpqn->cpqr", D, v)
        rr = g[:, 0] * qr + g[:, 1] * qs + g[:, 2] * qt
        rs = g[:, 1] * qr + g[:, 3] * qs + g[:, 4] * qt
        rt = g[:, 2] * qr + g[:, 4] * qs + g[:, 5] * qt
        w += np.einsum("np,cnqr->cpqr", D, rr) + np.einsum("nq,cpnr->cpqr", D, rs) + np.einsum("nr,cpqn->cpqr", D, rt)
    if mass:
        w += JxW.reshape(-1, nq, nq, nq) * v
    return np.einsum("cpqr,pi,qj,rk->cijk", w, B, B, B, optimize=True)


def op_apply(src, rd_or_idx, bas, G=None, JxW=None, laplace=True, mass=False, n_local=None, constrained=None):
    """dst = A src on one rank's local vector (owned + ghosts), no exchange.
    portable_laplace_operator.h:131 (dst=0), bk3_kokkos_kernel.h:130-138 (masked gather),
    :370-379 (additive scatter), portable_laplace_operator.h:171 (constrained rows = identity)."""
    idx = rd_or_idx["dof_indices"] if isinstance(rd_or_idx, dict) else rd_or_idx
    if constrained is None and isinstance(rd_or_idx, dict):
        constrained = rd_or_idx["constrained"]
    nm = bas["p"] + 1
    valid = idx != INVALID
    safe = np.where(valid, idx, 0).astype(np.int64)
    u = np.where(valid, src[safe], 0.0).reshape(-1, nm, nm, nm)
    out = cell_kernel(u, bas, G, JxW, laplace, mass).reshape(idx.shape)
    n_dst = len(src) if n_local is None else n_local
    dst = np.bincount(safe[valid], weights=out[valid], minlength=n_dst).astype(np.float64)
    if constrained is not None and len(constrained):
        dst[constrained] = src[constrained]
    return dst


def op_diagonal(rd, bas, G=None, JxW=None, laplace=True, mass=False):
    """Matrix diagonal by applying the cell kernel to unit vectors (bp5_kokkos/benchmark.cc:218-251);
    constrained rows get 1."""
    idx = rd["dof_indices"]
    nm3 = idx.shape[1]
    nm = bas["p"] + 1
    diag = np.zeros(rd["n_owned"] + rd["n_ghost"])
    valid = idx != INVALID
    for l in range(nm3):
        e = np.zeros((idx.shape[0], nm3))
        e[:, l] = 1.0
        col = cell_kernel(e.reshape(-1, nm, nm, nm), bas, G, JxW, laplace, mass).reshape(idx.shape)[:, l]
        np.add.at(diag, idx[valid[:, l], l].astype(np.int64), col[valid[:, l]])
    diag[rd["constrained"]] = 1.0
    return diag


def rhs_one(rd, bas, JxW):
    """b_i = int phi_i * 1 with constrained rows dropped (bp3.cc:208-224)."""
    idx = rd["dof_indices"]
    nq = bas["nq"]
    B = bas["B"]
    loc = np.einsum("cpqr,pi,qj,rk->cijk", JxW.reshape(-1, nq, nq, nq), B, B, B).reshape(idx.shape)
    valid = idx != INVALID
    return np.bincount(idx[valid].astype(np.int64), weights=loc[valid],
                       minlength=rd["n_owned"] + rd["n_ghost"]).astype(np.float64)


# --------------------------------------------------------------------------
# Element-free second oracle on uniform box meshes (SURVEY "Validated during this survey")
# --------------------------------------------------------------------------
def kron_1d(cells, h, bas):
    """Assembled 1-D stiffness K and mass M (Dirichlet rows/cols removed) and load vector."""
    p, nq = bas["p"], bas["nq"]
    B, Bg, wq = bas["B"], bas["Bg"], bas["wq"]
    n = cells * p + 1
    K = np.zeros((n, n))
    M = np.zeros((n, n))
    b = np.zeros(n)
    Ke = (Bg.T * wq) @ Bg / h
    Me = (B.T * wq) @ B * h
    be = B.T @ wq * h
    for e in range(cells):
        s = slice(e * p, e * p + p + 1)
        K[s, s] += Ke
        M[s, s] += Me
        b[s] += be
    return K[1:-1, 1:-1], M[1:-1, 1:-1], b[1:-1]


def kron_apply(mesh, bas, u):
    """A u for u[z,y,x] on the interior DoFs of a uniform box mesh:
    (Kx (x) My (x) Mz + Mx (x) Ky (x) Mz + Mx (x) My (x) Kz) u."""
    (Kx, Mx, _), (Ky, My, _), (Kz, Mz, _) = [kron_1d(mesh.cells[d], mesh.h[d], bas) for d in range(3)]
    t = lambda A, Bm, C: np.einsum("ai,bj,ck,ijk->abc", A, Bm, C, u, optimize=True)
    return t(Mz, My, Kx) + t(Mz, Ky, Mx) + t(Kz, My, Mx)


# --------------------------------------------------------------------------
# deal.II SolverCG + ReductionControl (bp3.cc:268-285)
# --------------------------------------------------------------------------
def estimate_max_eigenvalue(apply, inv_diag, n_owned, n_iterations, rank=0):
    """Power iteration on D^-1 A from the product's fixed pseudo-random start vector (64-bit LCG), times deal.II's safety
    factor 1.2 (PreconditionChebyshev::estimate_eigenvalues uses a CG/Lanczos estimate with the same factor)."""
    g = (1 + rank * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    u = np.empty(n_owned)
    for i in range(n_owned):
        g = (g * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        u[i] = (g >> 11) * (1.0 / 9007199254740992.0) - 0.5
    lam = 0.0
    for _ in range(n_iterations):
        y = u / np.sqrt(u @ u)
        u = inv_diag * apply(y)
        lam = y @ u
    return 1.2 * lam


def chebyshev_preconditioner(apply, inv_diag, degree, lambda_max, smoothing_range):
    """z = p_k(D^-1 A) D^-1 r: `degree` terms of the Chebyshev polynomial that is optimal on
    [lambda_max / smoothing_range, lambda_max] (three-term recurrence; dealii::PreconditionChebyshev)."""
    lmin = lambda_max / smoothing_range
    theta, delta = 0.5 * (lambda_max + lmin), 0.5 * (lambda_max - lmin)
    sigma1 = theta / delta

    def M(r):
        rho = 1.0 / sigma1
        d = inv_diag * r / theta
        z = d.copy()
        for _ in range(1, degree):
            rho_new = 1.0 / (2.0 * sigma1 - rho)
            d = rho_new * rho * d + (2.0 * rho_new / delta) * inv_diag * (r - apply(z))
            z = z + d
            rho = rho_new
        return z
    return M


def p_transfer(idx_fine, idx_coarse, p_fine, p_coarse, n_fine, n_coarse):
    """MGTransferGlobalCoarsening for polynomial coarsening on the same cells: returns (prolongate, restrict) with
    prolongate(u_c) = P u_c (cell-wise tensor-product embedding, contributions weighted by 1 / valence of the fine DoF) and
    restrict(r_f) = P^T r_f.  Index tables [cell][(p+1)^3] lexicographic, INVALID = masked."""
    nf, nc = p_fine + 1, p_coarse + 1
    P1 = lagrange_values(gll_01(nc)[0], gll_01(nf)[0])  # [jf, ic]
    vf, vc = idx_fine != INVALID, idx_coarse != INVALID
    sf, sc = np.where(vf, idx_fine, 0).astype(np.int64), np.where(vc, idx_coarse, 0).astype(np.int64)
    w = np.bincount(sf[vf], minlength=n_fine).astype(np.float64)
    w = np.where(w > 0, 1.0 / np.maximum(w, 1), 0.0)

    def prolongate(uc):
        loc = np.where(vc, uc[sc], 0.0).reshape(-1, nc, nc, nc)
        uf = np.einsum("ck,bj,ai,zkji->zcba", P1, P1, P1, loc, optimize=True).reshape(idx_fine.shape)
        return np.bincount(sf[vf], weights=(w[sf] * uf)[vf], minlength=n_fine)

    def restrict(rf):
        loc = np.where(vf, w[sf] * rf[sf], 0.0).reshape(-1, nf, nf, nf)
        rc = np.einsum("ck,bj,ai,zcba->zkji", P1, P1, P1, loc, optimize=True).reshape(idx_coarse.shape)
        return np.bincount(sc[vc], weights=rc[vc], minlength=n_coarse)
    return prolongate, restrict


def pmg_vcycle(levels, transfers, degree, smoothing_range, coarse_degree):
    """z = V(r): levels = [(apply, inv_diag, lambda_max), ...] fine to coarse, transfers = [(prolongate, restrict), ...];
    Chebyshev pre- and post-smoothing with `degree` terms, `coarse_degree` terms on the coarsest level."""
    def cycle(l, b):
        apply, inv_diag, lam = levels[l]
        last = l + 1 == len(levels)
        S = chebyshev_preconditioner(apply, inv_diag, coarse_degree if last else degree, lam, smoothing_range)
        x = S(b)
        if last:
            return x
        prolongate, restrict = transfers[l]
        x = x + prolongate(cycle(l + 1, restrict(b - apply(x))))
        return x + S(b - apply(x))
    return lambda r: cycle(0, r)


def solver_cg(apply, b, max_it, abs_tol, rel_tol, precond_inv_diag=None, dot=np.dot, precond=None):
    """Returns (x, its, res0, resn, converged).  x0 = 0.  precond: callable z = M(r) (overrides precond_inv_diag)."""
    x = np.zeros_like(b)
    r = b.copy()
    res0 = res = np.sqrt(dot(r, r))
    if res <= abs_tol:
        return x, 0, res0, res, True
    pvec = None
    rho = 0.0
    it = 0
    while True:
        it += 1
        if precond is not None:
            z = precond(r)
            rho_old, rho = rho, dot(r, z)
        elif precond_inv_diag is None:
            z = r
            rho_old, rho = rho, res * res
        else:
            z = precond_inv_diag * r
            rho_old, rho = rho, dot(r, z)
        pvec = z.copy() if it == 1 else z + (rho / rho_old) * pvec
        v = apply(pvec)
        alpha = rho / dot(pvec, v)
        x += alpha * pvec
        r -= alpha * v
        res = np.sqrt(abs(dot(r, r)))
        if res <= abs_tol or res <= rel_tol * res0:
            return x, it, res0, res, True
        if it >= max_it:
            return x, it, res0, res, False
