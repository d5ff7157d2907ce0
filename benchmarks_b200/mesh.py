"""Host-side mesh / basis objects over the C ABI (sections 2 and 3 of include/b200fe.h).

Stands in for the deal.II objects the reference drivers build before they construct the operator
(CEED_bp/src/bp3.cc:129-182, 452-488): Triangulation + DoFHandler + AffineConstraints + Partitioner.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib

QUAD_GAUSS, QUAD_GLL = 0, 1
PARTITION_P4EST, PARTITION_BLOCKS = 0, 1
GHOSTS_MINIMAL, GHOSTS_RELEVANT = 0, 1


class _Desc(C.Structure):
    _fields_ = [("subdivisions", C.c_int * 3), ("n_refine", C.c_int), ("p1", C.c_double * 3), ("p2", C.c_double * 3),
                ("p", C.c_int), ("n_ranks", C.c_int), ("rank", C.c_int), ("partition", C.c_int), ("ghosts", C.c_int),
                ("dirichlet", C.c_int)]


class _Info(C.Structure):
    _fields_ = [("n_cells_global", C.c_uint64), ("n_dofs_global", C.c_uint64), ("first_cell", C.c_uint64),
                ("owned_begin", C.c_uint64), ("n_cells_local", C.c_uint32), ("n_owned", C.c_uint32),
                ("n_ghost", C.c_uint32), ("n_constrained", C.c_uint32), ("cells", C.c_uint32 * 3), ("h", C.c_double * 3),
                ("origin", C.c_double * 3)]


def basis_1d(p: int, nq: int, quad: int = QUAD_GAUSS):
    """deal.II-layout 1-D arrays: shape_values[i*nq+q], co_shape_gradients[n*nq+q], shape_gradients[i*nq+q]."""
    nm = p + 1
    sv, cg, sg = np.empty(nm * nq), np.empty(nq * nq), np.empty(nm * nq)
    x, w = np.empty(nq), np.empty(nq)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    check(lib.b200fe_basis_1d(p, nq, quad, ptr(sv), ptr(cg), ptr(sg), ptr(x), ptr(w)))
    return dict(p=p, nq=nq, quad=quad, shape_values=sv, co_shape_gradients=cg, shape_gradients=sg, points=x, weights=w)


class BoxMesh:
    """subdivided_hyper_rectangle(subdivisions, p1, p2) + refine_global(n_refine), FE_Q(p) DoFs,
    Dirichlet constraints on the whole boundary, this rank's partition of it."""

    def __init__(self, subdivisions, n_refine, p, *, p1=(-1.0, -1.0, -1.0), p2=None, n_ranks=1, rank=0,
                 partition=PARTITION_P4EST, ghosts=GHOSTS_MINIMAL, dirichlet=True):
        if p2 is None:  # bp3.cc:452-463: coarse cells of side 1.9
            p2 = [a + 1.9 * s for a, s in zip(p1, subdivisions)]
        d = _Desc()
        d.subdivisions[:] = list(subdivisions)
        d.n_refine = n_refine
        d.p1[:] = list(p1)
        d.p2[:] = list(p2)
        d.p, d.n_ranks, d.rank, d.partition, d.ghosts, d.dirichlet = p, n_ranks, rank, partition, ghosts, int(dirichlet)
        self._desc, self._exchange_create = d, lib.b200fe_exchange_create_box
        self._h = C.c_void_p()
        check(lib.b200fe_boxmesh_create(C.byref(d), C.byref(self._h)))
        info = _Info()
        check(lib.b200fe_boxmesh_info(self._h, C.byref(info)))
        self.p, self.n_ranks, self.rank = p, n_ranks, rank
        self.p1, self.p2 = np.array(p1, float), np.array(p2, float)
        self.n_cells_global, self.n_dofs_global = info.n_cells_global, info.n_dofs_global
        self.first_cell, self.owned_begin = info.first_cell, info.owned_begin
        self.n_cells, self.n_owned, self.n_ghost = info.n_cells_local, info.n_owned, info.n_ghost
        self.cells, self.h = tuple(info.cells), np.array(list(info.h))
        self._dof_indices = None  # host copy of the index table: expanded on first access (see dof_indices)
        self.constrained = np.empty(info.n_constrained, dtype=np.uint32)
        self.ghost_global = np.empty(self.n_ghost, dtype=np.uint64)
        self.ghost_owner = np.empty(self.n_ghost, dtype=np.int32)
        self.cell_xyz = np.empty((self.n_cells, 3), dtype=np.int32)
        self.rank_dof_begin = np.empty(n_ranks + 1, dtype=np.uint64)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib.b200fe_boxmesh_fill(self._h, None, ptr(self.constrained), ptr(self.ghost_global),
                                      ptr(self.ghost_owner), ptr(self.cell_xyz), ptr(self.rank_dof_begin)))

    @property
    def dof_indices(self) -> np.ndarray:
        """[n_cells][(p+1)^3] index table on the HOST (b200fe_boxmesh_fill), built on first access and kept: the operator
        takes it from here once it exists (tests edit it), otherwise it lets the device expand its own copy."""
        if self._dof_indices is None:
            self._dof_indices = np.empty((self.n_cells, (self.p + 1) ** 3), dtype=np.uint32)
            check(lib.b200fe_boxmesh_fill(self._h, self._dof_indices.ctypes.data_as(C.c_void_p), None, None, None, None, None))
        return self._dof_indices

    def dof_indices_device(self, device):
        """The index table as a device tensor [n_cells][(p+1)^3] (uint32), expanded by a device kernel from 27 numbers per
        cell (b200fe_boxmesh_dof_indices_device); the host copy, if one was materialised, is uploaded instead."""
        import torch
        if self._dof_indices is not None:
            return torch.from_numpy(np.ascontiguousarray(self._dof_indices)).to(device)
        out = torch.empty((self.n_cells, (self.p + 1) ** 3), dtype=torch.uint32, device=device)
        with torch.cuda.device(device):
            check(lib.b200fe_boxmesh_dof_indices_device(self._h, C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out

    @classmethod
    def bp3_cycle(cls, cycle: int, p: int, **kw):
        """Mesh number `cycle` of the reference's sweep (CEED_bp/src/bp3.cc:443-473)."""
        n_refine, rem = cycle // 3, cycle % 3
        return cls([2 if d < rem else 1 for d in range(3)], n_refine, p, **kw)

    @property
    def n_local(self):
        return self.n_owned + self.n_ghost

    def fill_nodes(self, p_geo, deform_kind, amplitude, frequency, d_nodes_ptr, stream_ptr):
        """Mapping support points of the owned cells into a device array [cell][3][(p_geo+1)^3]."""
        check(lib.b200fe_boxmesh_nodes(self._h, p_geo, deform_kind, amplitude, frequency, d_nodes_ptr, stream_ptr))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.b200fe_boxmesh_destroy(h)
            self._h = None


class _HangDesc(C.Structure):
    _fields_ = [("box", _Desc), ("refine_lo", C.c_int * 3), ("refine_hi", C.c_int * 3)]


class _HangInfo(C.Structure):
    _fields_ = [("n_cells_global", C.c_uint64), ("n_dofs_global", C.c_uint64), ("first_cell", C.c_uint64),
                ("owned_begin", C.c_uint64), ("n_cells_local", C.c_uint32), ("n_owned", C.c_uint32),
                ("n_ghost", C.c_uint32), ("n_constrained", C.c_uint32), ("n_hanging_rows", C.c_uint32),
                ("n_hanging_entries", C.c_uint32), ("n_face_blocks", C.c_uint32), ("cells", C.c_uint32 * 3), ("h", C.c_double * 3),
                ("origin", C.c_double * 3)]


class HangingBoxMesh:
    """BoxMesh whose cells [refine_lo, refine_hi) are refined once more: two-level mesh with hanging nodes
    (BASELINE config C5).  Same attributes as BoxMesh plus the hanging-node rows (CSR, local indices):
    hang_dof, hang_row_ptr, hang_col, hang_w; `constrained` lists Dirichlet and hanging DoFs; cell_lxyz holds
    (level, x, y, z) per owned cell."""

    def __init__(self, subdivisions, n_refine, p, refine_lo, refine_hi, *, p1=(-1.0, -1.0, -1.0), p2=None, n_ranks=1,
                 rank=0, dirichlet=True):
        if p2 is None:
            p2 = [a + 1.9 * s for a, s in zip(p1, subdivisions)]
        d = _HangDesc()
        d.box.subdivisions[:] = list(subdivisions)
        d.box.n_refine = n_refine
        d.box.p1[:] = list(p1)
        d.box.p2[:] = list(p2)
        d.box.p, d.box.n_ranks, d.box.rank = p, n_ranks, rank
        d.box.partition, d.box.ghosts, d.box.dirichlet = PARTITION_P4EST, GHOSTS_MINIMAL, int(dirichlet)
        d.refine_lo[:] = list(refine_lo)
        d.refine_hi[:] = list(refine_hi)
        self._desc, self._exchange_create = d, lib.b200fe_exchange_create_hang
        self._h = C.c_void_p()
        check(lib.b200fe_hangmesh_create(C.byref(d), C.byref(self._h)))
        info = _HangInfo()
        check(lib.b200fe_hangmesh_info(self._h, C.byref(info)))
        self.p, self.n_ranks, self.rank = p, n_ranks, rank
        self.p1, self.p2 = np.array(p1, float), np.array(p2, float)
        self.n_cells_global, self.n_dofs_global = info.n_cells_global, info.n_dofs_global
        self.first_cell, self.owned_begin = info.first_cell, info.owned_begin
        self.n_cells, self.n_owned, self.n_ghost = info.n_cells_local, info.n_owned, info.n_ghost
        self.cells, self.h = tuple(info.cells), np.array(list(info.h))
        nm3 = (p + 1) ** 3
        self.dof_indices = np.empty((self.n_cells, nm3), dtype=np.uint32)
        self.constrained = np.empty(info.n_constrained, dtype=np.uint32)
        self.ghost_global = np.empty(self.n_ghost, dtype=np.uint64)
        self.ghost_owner = np.empty(self.n_ghost, dtype=np.int32)
        self.cell_lxyz = np.empty((self.n_cells, 4), dtype=np.int32)
        self.rank_dof_begin = np.empty(n_ranks + 1, dtype=np.uint64)
        self.hang_dof = np.empty(info.n_hanging_rows, dtype=np.uint32)
        self.hang_row_ptr = np.empty(info.n_hanging_rows + 1, dtype=np.uint32)
        self.hang_col = np.empty(info.n_hanging_entries, dtype=np.uint32)
        self.hang_w = np.empty(info.n_hanging_entries, dtype=np.float64)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib.b200fe_hangmesh_fill(self._h, ptr(self.dof_indices), ptr(self.constrained), ptr(self.ghost_global),
                                       ptr(self.ghost_owner), ptr(self.cell_lxyz), ptr(self.rank_dof_begin), ptr(self.hang_dof),
                                       ptr(self.hang_row_ptr), ptr(self.hang_col), ptr(self.hang_w)))
        # the same rows grouped by coarse face (tensor-product trace interpolation)
        self.face_parents = np.empty((info.n_face_blocks, (p + 1) ** 2), dtype=np.uint32)
        self.face_children = np.empty((info.n_face_blocks, (2 * p + 1) ** 2), dtype=np.uint32)
        check(lib.b200fe_hangmesh_fill_faces(self._h, ptr(self.face_parents), ptr(self.face_children)))
        self.trace_weights = np.empty((2 * p + 1, p + 1))
        check(lib.b200fe_trace_weights(p, ptr(self.trace_weights)))

    @property
    def n_local(self):
        return self.n_owned + self.n_ghost

    def fill_nodes(self, p_geo, deform_kind, amplitude, frequency, d_nodes_ptr, stream_ptr):
        check(lib.b200fe_hangmesh_nodes(self._h, p_geo, deform_kind, amplitude, frequency, d_nodes_ptr, stream_ptr))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.b200fe_hangmesh_destroy(h)
            self._h = None
