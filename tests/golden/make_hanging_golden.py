#!/usr/bin/env python
"""Regenerates tests/golden/hanging_meshes.json: digests and small tables of the two-level (hanging-node) meshes as the
ORACLE (oracle/hanging_oracle.py, conventions H1-H5) constructs them.  There is no reference implementation to generate
these from ("parity unpinned"); the fixture freezes the conventions, so that a later change of the
oracle or of the product's builder (csrc/hangmesh.cc) that alters numbering, ghost lists or constraint rows is noticed.
Run from the repo root:  python tests/golden/make_hanging_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import hanging_oracle as ho  # noqa: E402

CASES = [  # subdivisions, n_refine, p, refine_lo, refine_hi, n_ranks
    ((2, 2, 2), 0, 1, (0, 0, 0), (1, 1, 1), 1), ((2, 2, 2), 0, 2, (0, 0, 0), (1, 1, 1), 1),
    ((2, 1, 1), 1, 3, (1, 0, 0), (3, 1, 2), 3), ((1, 1, 1), 2, 2, (1, 1, 1), (3, 3, 3), 4),
    ((1, 1, 1), 1, 4, (0, 1, 0), (1, 2, 1), 2),
]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def tables(case):
    sub, nref, p, lo, hi, nranks = case
    om = ho.TwoLevelMesh(sub, nref, (lo, hi))
    sp = ho.build_space(om, p, nranks)
    ranks = []
    for r in range(nranks):
        rd = ho.rank_data(om, sp, r)
        par, chi = ho.face_blocks(om, sp, rd)
        ranks.append(dict(n_owned=int(rd["n_owned"]), n_ghost=int(rd["n_ghost"]), owned_begin=int(rd["owned_begin"]),
                          n_cells=len(rd["cells"]), n_hanging_rows=int(len(rd["hang_dof"])), n_face_blocks=int(len(par)),
                          dof_indices=digest(rd["dof_indices"].astype(np.uint32)), ghost_global=digest(rd["ghost_global"].astype(np.uint64)),
                          constrained=digest(rd["constrained"].astype(np.uint32)), hang_dof=digest(rd["hang_dof"].astype(np.uint32)),
                          hang_row_ptr=digest(rd["hang_row_ptr"].astype(np.uint32)), hang_col=digest(rd["hang_col"].astype(np.uint32)),
                          hang_w_rounded=digest(np.round(rd["hang_w"], 12) + 0.0), face_parents=digest(par), face_children=digest(chi)))
    out = dict(case=[list(sub), nref, p, list(lo), list(hi), nranks], n_cells=om.n_cells, n_dofs=int(sp["n_dofs"]),
               n_hanging=int(len(sp["hanging"])), ranks=ranks)
    if nranks == 1 and sp["n_dofs"] < 100:  # one table in the clear
        rd = ho.rank_data(om, sp, 0)
        out["dof_indices"] = rd["dof_indices"].astype(np.int64).tolist()
        out["hang_dof"] = rd["hang_dof"].tolist()
        out["hang_col"] = rd["hang_col"].tolist()
        out["hang_w"] = rd["hang_w"].tolist()
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hanging_meshes.json")
    json.dump({"generated_by": "tests/golden/make_hanging_golden.py (oracle/hanging_oracle.py, conventions H1-H5)",
               "cases": [tables(c) for c in CASES]}, open(path, "w"), indent=1)
    print("wrote", path)
