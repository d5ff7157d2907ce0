"""Host-side mirror of the reference's operator / solver interface over the C ABI.

  LaplaceOperator   <-> Portable::LaplaceOperator<dim,fe_degree,nq,double>
                        (CEED_bp/include/portable_laplace_operator.h:17-96): vmult, Tvmult, vmult_dummy,
                        initialize_dof_vector, m, n, compute_diagonal, get_matrix_diagonal_inverse
                        (bp5_kokkos/benchmark.cc:157-168, 218-251), compute_rhs (bp3.cc:184-239)
  ReductionControl, SolverCG <-> dealii::ReductionControl / SolverCG as called at bp3.cc:266-285
Vectors are float64 CUDA tensors of n_owned + n_ghost entries (owned first, then ghosts), the
layout of LinearAlgebra::distributed::Vector.  torch is only the allocator / stream provider.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check, lib
from .mesh import QUAD_GAUSS, QUAD_GLL, BoxMesh, basis_1d  # noqa: F401

OP_LAPLACE, OP_MASS, OP_HELMHOLTZ = 1, 2, 3
INVALID = 0xFFFFFFFF


class _OpDesc(C.Structure):
    _fields_ = [("p", C.c_int), ("nq", C.c_int), ("op_kind", C.c_int), ("collocated", C.c_int),
                ("n_cells", C.c_uint32), ("n_owned", C.c_uint32), ("n_ghost", C.c_uint32),
                ("h_shape_values", C.c_void_p), ("h_co_shape_gradients", C.c_void_p),
                ("d_dof_indices", C.c_void_p), ("d_G", C.c_void_p), ("d_JxW", C.c_void_p),
                ("h_constrained", C.c_void_p), ("n_constrained", C.c_uint32),
                ("n_phase0", C.c_uint32), ("n_phase1", C.c_uint32),
                ("d_cell_G", C.c_void_p), ("h_weights", C.c_void_p), ("d_cell_vertices", C.c_void_p), ("h_points", C.c_void_p)]


class _CgResult(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("initial_residual", C.c_double),
                ("final_residual", C.c_double)]


def _sp(stream=None):
    s = torch.cuda.current_stream() if stream is None else stream
    return C.c_void_p(s.cuda_stream)


def _dp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def overlap_permutation(dof_indices: np.ndarray, n_owned: int, one_sided: bool = False, touches=None):
    """Cell order [interior half | cells touching a ghost DoF | interior half] and the two sizes
    (deal.II's three colours with overlap_communication_computation).
    one_sided (P2P transport): [all interior cells | cells touching a ghost DoF] -- the ghost values are posted before the
    interior launch and only awaited before the boundary launch, so two launches instead of three hide the exchange."""
    if touches is None:  # (or given: the per-cell flags computed on the device)
        touches = ((dof_indices >= n_owned) & (dof_indices != INVALID)).any(axis=1)
    interior = np.nonzero(~touches)[0]
    boundary = np.nonzero(touches)[0]
    half = len(interior) if one_sided else len(interior) // 2
    perm = np.concatenate([interior[:half], boundary, interior[half:]])
    return perm, half, len(boundary)


class LaplaceOperator:
    """Matrix-free operator on one rank's partition of a BoxMesh.

    kind 'laplace' with nq = p+2 is BP3 (the reference's bp3), nq = p+1 Gauss is "bp35",
    quad='gll' (nq = p+1) is the collocated CEED BP5; 'mass' is BP1; 'helmholtz' is bp5_kokkos' operator.
    """

    def __init__(self, mesh: BoxMesh, nq: int | None = None, quad: str = "gauss", kind: str = "laplace",
                 p_geo: int = 1, deform=None, overlap: bool = False, halo=None, with_jxw: bool = True,
                 device=None, geometry: str = "stored", constraints: str = "faces", node_transform=None):
        self.mesh = mesh
        p = mesh.p
        self.p, self.nm = p, p + 1
        self.quad = QUAD_GLL if quad == "gll" else QUAD_GAUSS
        self.nq = nq if nq is not None else (p + 1 if quad == "gll" else p + 2)
        self.collocated = quad == "gll" and self.nq == p + 1
        self.kind = {"laplace": OP_LAPLACE, "mass": OP_MASS, "helmholtz": OP_HELMHOLTZ}[kind]
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.basis = basis_1d(p, self.nq, self.quad)
        self.perm = None
        n_phase0 = n_phase1 = 0
        want_split = overlap and mesh.n_ranks > 1 and not len(getattr(mesh, "hang_dof", ()))  # no split with hanging rows
        one_sided = halo is not None and halo.transport() == "p2p"
        if hasattr(mesh, "dof_indices_device") and mesh.n_local < 2 ** 31:
            # box meshes: the table is expanded on the device (27 numbers per cell go over PCIe instead of (p+1)^3)
            idx_dev = mesh.dof_indices_device(self.device)
            if want_split:
                signed = idx_dev.view(torch.int32)  # INVALID = -1, valid indices < 2^31
                touches = (signed >= mesh.n_owned).any(dim=1).cpu().numpy()
                self.perm, n_phase0, n_phase1 = overlap_permutation(None, mesh.n_owned, one_sided=one_sided, touches=touches)
                idx_dev = signed[torch.from_numpy(self.perm).to(self.device)].contiguous().view(torch.uint32)
            self.dof_indices = idx_dev
        else:
            idx = mesh.dof_indices
            if want_split:
                self.perm, n_phase0, n_phase1 = overlap_permutation(idx, mesh.n_owned, one_sided=one_sided)
                idx = idx[self.perm]
            self.dof_indices = torch.from_numpy(np.ascontiguousarray(idx)).to(self.device)
        # geometry: mapping support points -> G, JxW on the device
        ng3 = (p_geo + 1) ** 3
        nodes = torch.empty(mesh.n_cells * 3 * ng3, dtype=torch.float64, device=self.device)
        kind_d, amp, freq = (0, 0.0, 0.0) if deform is None else (1, float(deform[0]), float(deform[1]))
        mesh.fill_nodes(p_geo, kind_d, amp, freq, _dp(nodes), _sp())
        if node_transform is not None:
            # any further mapping of the support points (GridTools::transform): callable on the tensor [cell][3][(p_geo+1)^3]
            nodes = node_transform(nodes.view(mesh.n_cells, 3, ng3)).contiguous().view(-1)
        if self.perm is not None:
            nodes = nodes.view(mesh.n_cells, 3 * ng3)[torch.from_numpy(self.perm).to(self.device)].contiguous().view(-1)
        nq3 = self.nq ** 3
        self.geometry = geometry
        self.cell_G = None
        self.cell_vertices = None
        if geometry == "affine":  # on-the-fly geometric factors from six per-cell constants
            if p_geo != 1 or deform is not None:
                raise ValueError("geometry='affine' needs p_geo=1 and no deformation (an affine node_transform is fine); "
                                 "mass / Helmholtz operators additionally need axis-aligned cells")
            self.cell_G = torch.empty(mesh.n_cells * 8, dtype=torch.float64, device=self.device)
            check(lib.b200fe_geometry_affine_from_nodes(mesh.n_cells, _dp(nodes), _dp(self.cell_G), _sp()))
        elif geometry == "trilinear":  # general hexahedra: the 8 vertices per cell ARE the geometry (MappingQ1), G rebuilt per point
            if p_geo != 1 or self.kind != OP_LAPLACE:
                raise ValueError("geometry='trilinear' needs p_geo=1 and the Laplace operator")
            self.cell_vertices = nodes   # [cell][3][2][2][2], kept alive: borrowed by the operator
        elif geometry != "stored":
            raise ValueError(geometry)
        need_G = bool(self.kind & OP_LAPLACE) and geometry == "stored"
        need_J = (bool(self.kind & OP_MASS) and geometry == "stored") or with_jxw
        self.G = torch.empty(mesh.n_cells * 6 * nq3, dtype=torch.float64, device=self.device) if need_G else None
        self.JxW = torch.empty(mesh.n_cells * nq3, dtype=torch.float64, device=self.device) if need_J else None
        if self.G is not None or self.JxW is not None:
            check(lib.b200fe_geometry_from_nodes(p_geo, self.nq, self.quad, mesh.n_cells, _dp(nodes), _dp(self.G), _dp(self.JxW), _sp()))
        torch.cuda.current_stream().synchronize()
        del nodes
        d = _OpDesc()
        d.p, d.nq, d.op_kind, d.collocated = p, self.nq, self.kind, int(self.collocated)
        d.n_cells, d.n_owned, d.n_ghost = mesh.n_cells, mesh.n_owned, mesh.n_ghost
        self._sv = np.ascontiguousarray(self.basis["shape_values"])
        self._cg = np.ascontiguousarray(self.basis["co_shape_gradients"])
        self._con = np.ascontiguousarray(mesh.constrained)
        d.h_shape_values = self._sv.ctypes.data
        d.h_co_shape_gradients = self._cg.ctypes.data
        d.d_dof_indices = self.dof_indices.data_ptr()
        d.d_G = self.G.data_ptr() if self.G is not None else None
        d.d_JxW = self.JxW.data_ptr() if self.JxW is not None else None
        d.h_constrained = self._con.ctypes.data if len(self._con) else None
        d.n_constrained = len(self._con)
        d.n_phase0, d.n_phase1 = n_phase0, n_phase1
        self._w = np.ascontiguousarray(self.basis["weights"])
        d.d_cell_G = self.cell_G.data_ptr() if self.cell_G is not None else None
        d.h_weights = self._w.ctypes.data
        self._x = np.ascontiguousarray(self.basis["points"])
        d.d_cell_vertices = self.cell_vertices.data_ptr() if self.cell_vertices is not None else None
        d.h_points = self._x.ctypes.data
        self._h = C.c_void_p()
        check(lib.b200fe_op_create(C.byref(d), C.byref(self._h)))
        self.halo = halo
        if halo is not None:
            check(lib.b200fe_op_set_halo(self._h, halo._h))
        n_rows = len(getattr(mesh, "hang_dof", ()))
        if constraints not in ("rows", "faces"):
            raise ValueError(constraints)
        if n_rows:  # hanging-node rows of a HangingBoxMesh: the operator becomes C^T A C
            ptr = lambda a: a.ctypes.data if a.size else None
            # the CSR rows are always handed over (compute_diagonal of C^T A C reads them); the face-structured form, when
            # chosen, takes precedence in the apply: tensor-product trace interpolation, one CTA per coarse face
            check(lib.b200fe_op_set_constraints(self._h, n_rows, ptr(mesh.hang_dof), ptr(mesh.hang_row_ptr), ptr(mesh.hang_col),
                                                ptr(mesh.hang_w)))
            if constraints == "faces" and len(mesh.face_parents):
                check(lib.b200fe_op_set_face_constraints(self._h, p, len(mesh.face_parents), ptr(mesh.face_parents),
                                                         ptr(mesh.face_children), ptr(mesh.trace_weights)))
        self._inv_diag = None

    # --- the reference operator's interface ------------------------------------------------
    def initialize_dof_vector(self) -> torch.Tensor:
        return torch.zeros(self.mesh.n_owned + self.mesh.n_ghost, dtype=torch.float64, device=self.device)

    def m(self) -> int:
        return int(self.mesh.n_dofs_global)

    n = m

    def vmult(self, dst: torch.Tensor, src: torch.Tensor, stream=None) -> None:
        check(lib.b200fe_op_vmult(self._h, _dp(dst), _dp(src), _sp(stream)))

    Tvmult = vmult  # symmetric operator (portable_laplace_operator.h:396-409)

    def vmult_dummy(self, dst, src, ghost_exchange_on: bool, computation_on: bool, stream=None) -> None:
        check(lib.b200fe_op_vmult_dummy(self._h, _dp(dst), _dp(src), int(ghost_exchange_on), int(computation_on), _sp(stream)))

    def vmult_components(self, dst, src, n_components: int, stream=None) -> None:
        """Vector-valued apply (BP2/BP4/BP6): component-blocked vectors [component][n_owned + n_ghost]."""
        check(lib.b200fe_op_vmult_components(self._h, n_components, _dp(dst), _dp(src), _sp(stream)))

    def vmult_dot(self, dst, src, stream=None) -> torch.Tensor:
        dot = torch.empty(1, dtype=torch.float64, device=self.device)
        check(lib.b200fe_op_vmult_dot(self._h, _dp(dst), _dp(src), _dp(dot), _sp(stream)))
        return dot

    def vmult_host(self, h_dst: torch.Tensor, h_src: torch.Tensor, stream=None) -> None:
        """HOST tensors of n_owned doubles (pinned for full PCIe speed)."""
        check(lib.b200fe_op_vmult_host(self._h, _dp(h_dst), _dp(h_src), _sp(stream)))

    def compute_diagonal(self) -> torch.Tensor:
        diag = self.initialize_dof_vector()
        check(lib.b200fe_op_diagonal(self._h, _dp(diag), _sp()))
        # reciprocal as in bp5_kokkos/benchmark.cc:240-250
        d = diag[: self.mesh.n_owned]
        self._inv_diag = torch.where(d > 0, 1.0 / d, torch.ones_like(d))
        return diag

    def get_matrix_diagonal_inverse(self) -> torch.Tensor:
        if self._inv_diag is None:
            self.compute_diagonal()
        return self._inv_diag

    def compute_rhs(self) -> torch.Tensor:
        """b_i = int phi_i * 1 (bp3.cc:184-239)."""
        b = self.initialize_dof_vector()
        check(lib.b200fe_op_rhs_one(self._h, _dp(b), _sp()))
        return b

    def distribute(self, x: torch.Tensor, stream=None) -> None:
        """AffineConstraints::distribute: hanging entries of x from their parents (after a solve)."""
        check(lib.b200fe_op_distribute(self._h, _dp(x), _sp(stream)))

    def timing_enable(self, max_launches: int) -> None:
        check(lib.b200fe_op_timing_enable(self._h, max_launches))

    def timing_read(self):
        """(total kernel ms, launches) of the cell kernel since the last read (CUDA events in the library)."""
        ms, n = C.c_double(), C.c_int()
        check(lib.b200fe_op_timing_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launch_info(self):
        v = [C.c_int() for _ in range(6)]
        check(lib.b200fe_op_launch_info(self._h, *[C.byref(x) for x in v]))
        keys = ("elems_per_block", "num_blocks", "threads_per_block", "smem_bytes", "blocks_per_sm", "regs_per_thread")
        eo = C.c_int()
        check(lib.b200fe_op_kernel_variant(self._h, C.byref(eo)))
        ex = C.c_int()
        check(lib.b200fe_op_exclusive_interior(self._h, C.byref(ex)))
        ca = C.c_int()
        check(lib.b200fe_op_cartesian(self._h, C.byref(ca)))
        return dict(zip(keys, [x.value for x in v]), even_odd=eo.value, multi_component=int(self.multi_component_kernel()),
                    exclusive_interior=ex.value, cartesian=ca.value)

    def multi_component_kernel(self) -> bool:
        """Vector-valued applies (vmult_components, n_components CG) run ONE cell-kernel launch that fetches the geometric
        factors of a cell once for all components (csrc/operator.cu: op_has_mc_kernel): collocated Laplace, stored G."""
        import os
        return (self.collocated and self.kind == OP_LAPLACE and self.geometry == "stored"
                and os.environ.get("B200FE_MULTI_COMPONENT", "1") != "0")

    # algorithmic bytes of one apply: G + indices per cell, 32 B per local DoF
    def algorithmic_bytes(self) -> int:
        nq3, nm3 = self.nq ** 3, self.nm ** 3
        g_bytes = 64 if self.geometry == "affine" else 192 if self.geometry == "trilinear" else 48 * nq3
        jxw_bytes = 8 * nq3 if self.geometry == "stored" else (0 if self.kind & OP_LAPLACE else 64)  # on the fly: det J among the cell constants
        per_cell = 4 * nm3 + (g_bytes if self.kind & OP_LAPLACE else 0) + (jxw_bytes if self.kind & OP_MASS else 0)
        return self.mesh.n_cells * per_cell + 32 * self.mesh.n_owned

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.b200fe_op_destroy(h)
            self._h = None


class ReductionControl:
    """dealii::ReductionControl(max_steps, tolerance, reduction)."""

    def __init__(self, max_steps: int = 100, tolerance: float = 1e-10, reduction: float = 1e-2):
        self.max_steps, self.tolerance, self.reduction = max_steps, tolerance, reduction
        self._res = None

    def last_step(self) -> int:
        return self._res.iterations

    def last_value(self) -> float:
        return self._res.final_residual

    def initial_value(self) -> float:
        return self._res.initial_residual


class NoConvergence(RuntimeError):
    """dealii::SolverControl::NoConvergence."""


class PreconditionChebyshev:
    """dealii::PreconditionChebyshev<Operator, Vector, DiagonalMatrix> as a CG preconditioner: a polynomial of `degree` terms in
    D^-1 A, optimal on [lambda_max / smoothing_range, lambda_max]; lambda_max from `eig_iterations` power iterations times 1.2
    (estimate_eigenvalues) unless given (AdditionalData::max_eigenvalue)."""

    def __init__(self, A: "LaplaceOperator", degree: int = 5, smoothing_range: float = 20.0, eig_iterations: int = 10,
                 max_eigenvalue: float | None = None):
        self.A, self.degree, self.smoothing_range = A, int(degree), float(smoothing_range)
        self.inv_diag = A.get_matrix_diagonal_inverse()
        if max_eigenvalue is None:
            lam = C.c_double()
            check(lib.b200fe_op_estimate_max_eigenvalue(A._h, _dp(self.inv_diag), int(eig_iterations), C.byref(lam), _sp()))
            max_eigenvalue = lam.value
        self.max_eigenvalue = float(max_eigenvalue)


class PTransfer:
    """deal.II MGTransferGlobalCoarsening between FE_Q(p_fine) and FE_Q(p_coarse) on the same cells (polynomial coarsening):
    prolongate_and_add / restrict_and_add on L-vectors of the two operators (same mesh cells, different degree)."""

    def __init__(self, A_fine: "LaplaceOperator", A_coarse: "LaplaceOperator"):
        if A_fine.mesh.n_cells != A_coarse.mesh.n_cells or A_fine.perm is not None or A_coarse.perm is not None:
            raise ValueError("PTransfer: the two operators must live on the same cells in the same order")
        self.fine, self.coarse = A_fine, A_coarse
        self._h = C.c_void_p()
        check(lib.b200fe_ptransfer_create(A_fine.p, A_coarse.p, A_fine.mesh.n_cells, _dp(A_fine.dof_indices), _dp(A_coarse.dof_indices),
                                          A_fine.mesh.n_owned + A_fine.mesh.n_ghost, A_coarse.mesh.n_owned + A_coarse.mesh.n_ghost,
                                          C.byref(self._h)))

    def prolongate_and_add(self, fine: torch.Tensor, coarse: torch.Tensor, stream=None):
        check(lib.b200fe_ptransfer_prolongate_add(self._h, _dp(fine), _dp(coarse), _sp(stream)))

    def restrict_and_add(self, coarse: torch.Tensor, fine: torch.Tensor, stream=None):
        check(lib.b200fe_ptransfer_restrict_add(self._h, _dp(coarse), _dp(fine), _sp(stream)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.b200fe_ptransfer_destroy(h)
            self._h = None


class PreconditionPMG:
    """p-multigrid V-cycle (Chebyshev smoothers on every level, polynomial coarsening on the same cells) as a SolverCG
    preconditioner: operators from fine to coarse, e.g. degrees 8, 4, 2, 1 on one BoxMesh geometry."""

    def __init__(self, operators, smoother_degree: int = 3, smoothing_range: float = 20.0, coarse_degree: int = 8, eig_iterations: int = 12):
        self.operators = list(operators)
        self.transfers = [PTransfer(a, c) for a, c in zip(self.operators[:-1], self.operators[1:])]
        self.inv_diags = [A.get_matrix_diagonal_inverse() for A in self.operators]
        self.lambdas = []
        for A, d in zip(self.operators, self.inv_diags):
            lam = C.c_double()
            check(lib.b200fe_op_estimate_max_eigenvalue(A._h, _dp(d), int(eig_iterations), C.byref(lam), _sp()))
            self.lambdas.append(lam.value)
        n = len(self.operators)
        ops = (C.c_void_p * n)(*[A._h for A in self.operators])
        diags = (C.c_void_p * n)(*[d.data_ptr() for d in self.inv_diags])
        lams = (C.c_double * n)(*self.lambdas)
        trs = (C.c_void_p * max(n - 1, 1))(*[t._h for t in self.transfers])
        self._h = C.c_void_p()
        check(lib.b200fe_pmg_create(n, ops, diags, lams, trs, int(smoother_degree), float(smoothing_range), int(coarse_degree), C.byref(self._h)))

    def vmult(self, z: torch.Tensor, r: torch.Tensor, stream=None):
        check(lib.b200fe_pmg_vcycle(self._h, _dp(z), _dp(r), _sp(stream)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.b200fe_pmg_destroy(h)
            self._h = None


class SolverCG:
    """dealii::SolverCG: solve(A, x, b, preconditioner); preconditioner None = PreconditionIdentity,
    a tensor holding the inverse diagonal (DiagonalMatrix / Jacobi), a PreconditionChebyshev or a PreconditionPMG.

    One difference from deal.II: the solve always starts from x0 = 0 -- whatever x holds on entry is OVERWRITTEN, not used
    as an initial guess (dealii::SolverCG would start from r = b - A x).  Every reference driver zeroes the solution
    before each solve (bp3.cc:271, bp5_kokkos/benchmark.cc:364), so their iteration counts are the cold-start ones; a
    caller that wants a warm start must solve for the correction: A dx = b - A x0, x = x0 + dx."""

    def __init__(self, control: ReductionControl, check_every: int = 8):
        self.control = control
        self.check_every = check_every

    def solve(self, A: LaplaceOperator, x: torch.Tensor, b: torch.Tensor, preconditioner=None, stream=None,
              n_components: int = 1):
        """n_components > 1: vector-valued problem (BP2/BP4/BP6), x and b component-blocked
        [component][n_owned + n_ghost]."""
        res = _CgResult()
        if isinstance(preconditioner, PreconditionPMG):
            if n_components != 1 or preconditioner.operators[0] is not A:
                raise ValueError("PreconditionPMG: scalar problems, built on the operator being solved")
            rc = lib.b200fe_cg_solve_pmg(preconditioner._h, _dp(x), _dp(b), self.control.tolerance, self.control.reduction,
                                         self.control.max_steps, self.check_every, C.byref(res), _sp(stream))
        elif isinstance(preconditioner, PreconditionChebyshev):
            if n_components != 1:
                raise ValueError("PreconditionChebyshev: scalar problems only")
            c = preconditioner
            rc = lib.b200fe_cg_solve_chebyshev(A._h, _dp(x), _dp(b), _dp(c.inv_diag), c.degree, c.max_eigenvalue, c.smoothing_range,
                                               self.control.tolerance, self.control.reduction, self.control.max_steps, self.check_every,
                                               C.byref(res), _sp(stream))
        else:
            rc = lib.b200fe_cg_solve_components(A._h, n_components, _dp(x), _dp(b), _dp(preconditioner), self.control.tolerance,
                                                self.control.reduction, self.control.max_steps, self.check_every, C.byref(res),
                                                _sp(stream))
        self.control._res = res
        if rc == 5:
            raise NoConvergence(f"CG: {res.iterations} iterations, residual {res.final_residual:g}")
        check(rc)
        return res

    def solve_host(self, A: LaplaceOperator, h_x: torch.Tensor, h_b: torch.Tensor, preconditioner=None, stream=None):
        res = _CgResult()
        rc = lib.b200fe_cg_solve_host(A._h, _dp(h_x), _dp(h_b), _dp(preconditioner), self.control.tolerance,
                                      self.control.reduction, self.control.max_steps, self.check_every, C.byref(res), _sp(stream))
        self.control._res = res
        if rc == 5:
            raise NoConvergence(f"CG: {res.iterations} iterations, residual {res.final_residual:g}")
        check(rc)
        return res
