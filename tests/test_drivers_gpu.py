"""The C++ host layer (include/b200fe/operator.hpp) through the ported reference drivers:
benchmarks_b200/drivers/bk_benchmark (CEED_BK/src/BK*/templated_cuda_benchmark.cc protocol) must print the
reference's known-answer norms, benchmarks_b200/drivers/bp3 (CEED_bp/src/bp3.cc protocol) the reference's
golden CG iteration counts for p = 4."""
import json
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "benchmarks_b200", "drivers")


def _run(args, timeout=600):
    exe = os.path.join(DRV, args[0])
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", DRV], check=True)
    r = subprocess.run([exe] + [str(a) for a in args[1:]], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr
    return r.stdout


@pytest.mark.parametrize("kind,p", [("bk1", 2), ("bk3", 2), ("bk5", 4), ("bk3", 8), ("bk1", 7), ("bk5", 1)])
def test_bk_benchmark_prints_reference_norms(golden_dir, kind, p):
    gold = json.load(open(os.path.join(golden_dir, "bk_norms.json")))["kat_norms_nelmt64"][str(p)][kind]
    out = _run(["bk_benchmark", kind, p, 64, 2])
    last = out.strip().splitlines()[-1].split()
    assert last[0] == kind.upper() and int(last[1]) == p and int(last[2]) == 64
    assert float(last[-1]) == pytest.approx(gold, rel=1e-7)  # printed with 8 significant digits


def test_bp3_driver_reproduces_reference_table(golden_dir):
    rows = json.load(open(os.path.join(golden_dir, "bp3_cg_p4.json")))["rows"]
    out = _run(["bp3", 4, 10000, 300000])
    # last printed convergence table
    block = out[out.rindex(" cells    dofs    matvec"):]
    block = block[:block.index("mv_ghost_and_compute")]
    table = [l.split() for l in block.splitlines()[1:] if re.match(r"^\s*\d+\s+\d+\s+\d\.\d+e", l)]
    assert len(table) >= 5
    for got, want in zip(table, rows):
        assert int(got[0]) == want[1] and int(got[1]) == want[2]
        assert abs(int(got[5]) - want[3]) <= 1
        assert float(got[6]) == pytest.approx(want[4], abs=2e-3)


@pytest.mark.parametrize("extra,kernel", [((), "cartesian"), ((1, "gll"), "cartesian")])
def test_bp3_driver_with_geometry_on_the_fly_gives_the_same_table(extra, kernel):
    """C++ host layer, Geometry::OnTheFly: same cells / DoFs / CG iteration counts (+-1) and reduction rates
    as the stored-G run of the same driver; the reference's cube cells land on the separable (cartesian) kernels: the nodal
    one for QGauss(p+2), the collocated one for GLL."""
    def table(env):
        exe = os.path.join(DRV, "bp3")
        if not os.path.exists(exe):
            subprocess.run(["make", "-C", DRV], check=True)
        r = subprocess.run([exe, "4", "10000", "150000"] + [str(a) for a in extra], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr
        out = r.stdout
        block = out[out.rindex(" cells    dofs    matvec"):]
        block = block[:block.index("mv_ghost_and_compute")]
        return out, [l.split() for l in block.splitlines()[1:] if re.match(r"^\s*\d+\s+\d+\s+\d\.\d+e", l)]
    _, stored = table({})
    out, otf = table({"B200FE_GEOMETRY": "onthefly"})
    assert f"Geometry on the fly ({kernel} kernel)" in out
    assert len(stored) >= 4 and len(stored) == len(otf)
    for a, b in zip(stored, otf):
        assert a[0] == b[0] and a[1] == b[1]
        assert abs(int(a[5]) - int(b[5])) <= 1
        assert float(a[6]) == pytest.approx(float(b[6]), abs=2e-3)


def test_bp3_driver_reports_errors_like_the_reference():
    exe = os.path.join(DRV, "bp3")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", DRV], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Expected at least one argument" in r.stdout


def test_bp5_driver_with_geometry_on_the_fly_gives_the_same_row():
    """bp5_kokkos protocol with Geometry::OnTheFly: the unit-cube mesh takes the separable Helmholtz kernel (no G, no JxW,
    diagonal and rhs from the 1-D matrices) -- same DoFs and Jacobi-CG iteration count as the stored-geometry run."""
    def row(env):
        exe = os.path.join(DRV, "bp5")
        if not os.path.exists(exe):
            subprocess.run(["make", "-C", DRV], check=True)
        r = subprocess.run([exe, "4", "14", "1"], capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr
        return [x.strip() for x in r.stdout.strip().splitlines()[-1].split("|")]
    a, b = row({}), row({"B200FE_GEOMETRY": "onthefly"})
    assert a[:4] == b[:4]                       # p, q, cells, DoFs
    assert abs(int(a[6]) - int(b[6])) <= 1      # CG iterations


def test_bp5_driver_helmholtz_jacobi_matches_c_oracle(oracle_mod):
    """bp5_kokkos protocol (Helmholtz, QGauss(p+1), Jacobi CG capped at 100 iterations, rhs = i % 8): the C++ driver's
    iteration count and DoF count against the C oracle on the same mesh (s = 3: 1 x 1 x 7 cells, p = 3)."""
    import numpy as np
    out = _run(["bp5", 3, 3, 1])
    row = [x.strip() for x in out.strip().splitlines()[-1].split("|")]
    p, q, n_el, n_dofs, its = int(row[0]), int(row[1]), int(row[2]), int(row[3]), int(row[6])
    assert (p, q, n_el) == (3, 4, 7) and n_dofs == 4 * 4 * 22
    fe = oracle_mod.fe
    om = fe.BoxMesh((1, 1, 7), 0, p1=(0.0, 0.0, 0.0), p2=(1.0, 1.0, 7.0))
    od = fe.distribute_dofs(om, 3, 1)
    rd = fe.rank_data(om, od, 0)
    bas = fe.basis_1d(3, 4)
    G, JxW = fe.geometric_factors(fe.cell_nodes(om, rd["cells"], 1), 1, bas)
    rhs = np.arange(rd["n_owned"], dtype=np.float64) % 8
    rhs[rd["constrained"]] = 0.0
    inv_diag = fe.op_diagonal(rd, bas, G, JxW, laplace=True, mass=True)
    inv_diag = np.where(inv_diag > 0, 1.0 / inv_diag, 1.0)
    _, its_o, _, _, _ = oracle_mod.port.cg_solve(rhs, nm=4, nq=4, collocated=False, flags=3, shape_values=bas["B"].T.copy(),
                                                 co_shape_gradients=bas["D"].T.copy(), G=G, JxW=JxW, dof_indices=rd["dof_indices"],
                                                 constrained=rd["constrained"], inv_diag=inv_diag, max_it=100, abs_tol=1e-15, rel_tol=1e-8)
    assert abs(its - its_o) <= 1
