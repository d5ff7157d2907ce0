"""Even-odd decomposition of the 1-D contractions (csrc/eo_contract.h, tuning variant -DB200FE_EVEN_ODD): the four
contractions against the plain sums for every (nm, nq) the library instantiates, with the real Gauss / Gauss-Lobatto
matrices.  Host-only: the header is written host/device so that its algebra can be checked without a GPU."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "benchmarks_b200", "csrc")


def test_even_odd_contractions_equal_plain_sums():
    cuda_inc = "/usr/local/cuda/include"
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "eo_selftest")
        subprocess.run(["g++", "-std=c++17", "-O1", "-fopenmp", "-I", CSRC, "-I", cuda_inc, os.path.join(CSRC, "eo_selftest.cc"),
                        os.path.join(CSRC, "basis.cc"), "-o", exe], check=True)
        r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "worst relative deviation" in r.stdout
