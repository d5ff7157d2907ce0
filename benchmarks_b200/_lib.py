"""ctypes binding of libb200fe.so (the C ABI of include/b200fe.h)."""
from __future__ import annotations

import ctypes as C
import os

import torch  # noqa: F401  (first: so that libnccl.so.2 resolves to the build torch ships)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200FE_LIB") or os.path.join(_HERE, "libb200fe.so")  # env override: tuning variants only


class B200feError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200fe error {code}: {msg}")
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C benchmarks_b200/csrc`).  b200fe has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_vp, _i, _u32, _u64 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64
_pi = C.POINTER(C.c_int)
_pd = C.POINTER(C.c_double)

_SIGNATURES = {
    "b200fe_version": (_i, []),
    "b200fe_last_error": (C.c_char_p, []),
    "b200fe_bk1_apply": (_i, [_i, _i, _u32, _vp, _vp, _vp, _vp, _vp]),
    "b200fe_bk3_apply": (_i, [_i, _i, _u32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200fe_bk5_apply": (_i, [_i, _u32, _vp, _vp, _vp, _vp, _vp]),
    "b200fe_sum_squares": (_i, [_u64, _vp, _vp, _vp]),
    "b200fe_bk_launch_info": (_i, [_i, _i, _i, _u32, _pi, _pi, _pi, _pi]),
    "b200fe_debug_poison_smem": (_i, [_vp]),
    "b200fe_basis_1d": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "b200fe_boxmesh_create": (_i, [_vp, C.POINTER(_vp)]),
    "b200fe_boxmesh_destroy": (None, [_vp]),
    "b200fe_boxmesh_info": (_i, [_vp, _vp]),
    "b200fe_boxmesh_fill": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200fe_boxmesh_dof_indices_device": (_i, [_vp, _vp, _vp]),
    "b200fe_boxmesh_nodes": (_i, [_vp, _i, _i, C.c_double, C.c_double, _vp, _vp]),
    "b200fe_hangmesh_create": (_i, [_vp, C.POINTER(_vp)]),
    "b200fe_hangmesh_destroy": (None, [_vp]),
    "b200fe_hangmesh_info": (_i, [_vp, _vp]),
    "b200fe_hangmesh_fill": (_i, [_vp] * 11),
    "b200fe_hangmesh_fill_faces": (_i, [_vp, _vp, _vp]),
    "b200fe_trace_weights": (_i, [_i, _vp]),
    "b200fe_hangmesh_nodes": (_i, [_vp, _i, _i, C.c_double, C.c_double, _vp, _vp]),
    "b200fe_geometry_from_nodes": (_i, [_i, _i, _i, _u32, _vp, _vp, _vp, _vp]),
    "b200fe_geometry_from_inv_jacobian": (_i, [_u32, _i, _vp, _vp, _vp, _vp]),
    "b200fe_geometry_affine_from_nodes": (_i, [_u32, _vp, _vp, _vp]),
    "b200fe_op_create": (_i, [_vp, C.POINTER(_vp)]),
    "b200fe_op_destroy": (None, [_vp]),
    "b200fe_op_set_halo": (_i, [_vp, _vp]),
    "b200fe_op_set_constraints": (_i, [_vp, _u32, _vp, _vp, _vp, _vp]),
    "b200fe_op_set_face_constraints": (_i, [_vp, _i, _u32, _vp, _vp, _vp]),
    "b200fe_op_distribute": (_i, [_vp, _vp, _vp]),
    "b200fe_op_vmult": (_i, [_vp, _vp, _vp, _vp]),
    "b200fe_op_vmult_components": (_i, [_vp, _i, _vp, _vp, _vp]),
    "b200fe_op_vmult_dot": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "b200fe_op_vmult_dummy": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "b200fe_op_vmult_host": (_i, [_vp, _vp, _vp, _vp]),
    "b200fe_op_diagonal": (_i, [_vp, _vp, _vp]),
    "b200fe_op_rhs_one": (_i, [_vp, _vp, _vp]),
    "b200fe_op_launch_info": (_i, [_vp, _pi, _pi, _pi, _pi, _pi, _pi]),
    "b200fe_op_kernel_variant": (_i, [_vp, _pi]),
    "b200fe_op_exclusive_interior": (_i, [_vp, _pi]),
    "b200fe_op_cartesian": (_i, [_vp, _pi]),
    "b200fe_op_timing_enable": (_i, [_vp, _i]),
    "b200fe_op_timing_read": (_i, [_vp, _pd, _pi]),
    "b200fe_launch_count": (C.c_ulonglong, []),
    "b200fe_cg_solve": (_i, [_vp, _vp, _vp, _vp, C.c_double, C.c_double, _i, _i, _vp, _vp]),
    "b200fe_cg_solve_components": (_i, [_vp, _i, _vp, _vp, _vp, C.c_double, C.c_double, _i, _i, _vp, _vp]),
    "b200fe_cg_solve_host": (_i, [_vp, _vp, _vp, _vp, C.c_double, C.c_double, _i, _i, _vp, _vp]),
    "b200fe_cg_solve_chebyshev": (_i, [_vp, _vp, _vp, _vp, _i, C.c_double, C.c_double, C.c_double, C.c_double, _i, _i, _vp, _vp]),
    "b200fe_op_estimate_max_eigenvalue": (_i, [_vp, _vp, _i, _pd, _vp]),
    "b200fe_ptransfer_create": (_i, [_i, _i, _u32, _vp, _vp, _u32, _u32, _vp]),
    "b200fe_ptransfer_destroy": (None, [_vp]),
    "b200fe_ptransfer_prolongate_add": (_i, [_vp, _vp, _vp, _vp]),
    "b200fe_ptransfer_restrict_add": (_i, [_vp, _vp, _vp, _vp]),
    "b200fe_pmg_create": (_i, [_i, _vp, _vp, _vp, _vp, _i, C.c_double, _i, _vp]),
    "b200fe_pmg_destroy": (None, [_vp]),
    "b200fe_pmg_vcycle": (_i, [_vp, _vp, _vp, _vp]),
    "b200fe_cg_solve_pmg": (_i, [_vp, _vp, _vp, C.c_double, C.c_double, _i, _i, _vp, _vp]),
    "b200fe_exchange_create_box": (_i, [_vp, C.POINTER(_vp)]),
    "b200fe_exchange_create_hang": (_i, [_vp, C.POINTER(_vp)]),
    "b200fe_exchange_destroy": (None, [_vp]),
    "b200fe_exchange_info": (_i, [_vp, _pi, C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_u32)]),
    "b200fe_exchange_fill": (_i, [_vp] * 7),
    "b200fe_comm_available": (_i, []),
    "b200fe_comm_unique_id": (_i, [_vp]),
    "b200fe_halo_create": (_i, [_vp, C.POINTER(_vp)]),
    "b200fe_halo_destroy": (None, [_vp]),
    "b200fe_halo_update_ghosts": (_i, [_vp, _vp, _vp]),
    "b200fe_halo_compress_add": (_i, [_vp, _vp, _vp]),
    "b200fe_halo_zero_ghosts": (_i, [_vp, _vp, _vp]),
    "b200fe_halo_allreduce_sum": (_i, [_vp, _vp, _i, _vp]),
    "b200fe_halo_transport": (_i, [_vp, _pi, _pi]),
    "b200fe_halo_set_transport": (_i, [_vp, _i]),
    "b200fe_halo_status": (_i, [_vp]),
    "b200fe_halo_exchange_raw": (_i, [_vp, _vp, _vp, _vp]),
    "b200fe_halo_exchange_raw_rounds": (_i, [_vp, _vp, _vp, _i, _vp]),
}
for _name, (_res, _args) in _SIGNATURES.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args


def check(code: int) -> None:
    if code != 0:
        raise B200feError(code, lib.b200fe_last_error().decode())
