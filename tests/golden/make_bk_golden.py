#!/usr/bin/env python
"""Regenerates tests/golden/bk_norms.json by running the REFERENCE's own serial kernels
(oracle/_ref, compiled in place from /root/reference by oracle/Makefile) on the reference's
seedless synthetic inputs (CEED_BK/src/BK{1,3,5}/serial_verification.cc: in=3, JxW=1, G=2,
basis=cos(index)), FP64, nelmt = 64 and a few nelmt = 1000 spot values.

Also stores sha256 digests of the reference's OUTPUT VECTORS for a seeded random input set, so
that the bit-exactness of the C restatement (oracle/bk_oracle.c) against the reference can be
re-checked on machines where /root/reference (and oracle/_ref) is absent.
Run from the repo root:  python tests/golden/make_bk_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402

assert oracle.ref is not None, "oracle/_ref not built (needs /root/reference)"


def random_case(kind, p, nelmt, seed):
    rng = np.random.default_rng(seed)
    nm = p + 1
    nq = p + 1 if kind == "bk5" else p + 2
    return dict(nm=nm, nq=nq, basis=rng.uniform(-1, 1, nq * nm), dbasis=rng.uniform(-1, 1, nq * nq),
                u=rng.uniform(-1, 1, nelmt * (nq if kind == "bk5" else nm) ** 3),
                JxW=rng.uniform(0.5, 1.5, nelmt * nq ** 3), G=rng.uniform(-1, 1, (nelmt, 6, nq, nq, nq)))


def serial_G(G):  # [e][6][p][q][r] -> reference serial layout [e][p][q][6][r]
    return np.ascontiguousarray(np.transpose(G, (0, 2, 3, 1, 4)))


def main():
    out = {"source": "oracle/_ref (reference serial kernels), g++ -O3 -ffp-contract=off, FP64",
           "kat_norms_nelmt64": {}, "kat_norms_nelmt1000": {}, "random_sha256": {}}
    for p in range(1, 9):
        k = oracle.kat_inputs("bk1", p, 64)
        _, s1 = oracle.ref.bk1(k["nq"], k["basis"], k["JxW"], k["u"])
        _, s3 = oracle.ref.bk3(k["nq"], k["basis"], k["dbasis"], k["G"], k["u"])
        k5 = oracle.kat_inputs("bk5", p, 64)
        _, s5 = oracle.ref.bk5(k5["nq"], k5["dbasis"], k5["G"], k5["u"])
        out["kat_norms_nelmt64"][str(p)] = {"bk1": float(np.sqrt(s1)), "bk3": float(np.sqrt(s3)), "bk5": float(np.sqrt(s5))}
    k = oracle.kat_inputs("bk1", 2, 1000)
    out["kat_norms_nelmt1000"]["bk1_p2"] = float(np.sqrt(oracle.ref.bk1(k["nq"], k["basis"], k["JxW"], k["u"])[1]))
    out["kat_norms_nelmt1000"]["bk1_p2_direct"] = float(np.sqrt(oracle.ref.bk1(k["nq"], k["basis"], k["JxW"], k["u"], "direct")[1]))
    out["kat_norms_nelmt1000"]["bk3_p2"] = float(np.sqrt(oracle.ref.bk3(k["nq"], k["basis"], k["dbasis"], k["G"], k["u"])[1]))
    k5 = oracle.kat_inputs("bk5", 4, 1000)
    out["kat_norms_nelmt1000"]["bk5_p4"] = float(np.sqrt(oracle.ref.bk5(k5["nq"], k5["dbasis"], k5["G"], k5["u"])[1]))
    for p in range(1, 9):
        c = random_case("bk1", p, 5, 1000 + p)
        o1, _ = oracle.ref.bk1(c["nq"], c["basis"], c["JxW"], c["u"])
        o3, _ = oracle.ref.bk3(c["nq"], c["basis"], c["dbasis"], serial_G(c["G"]).ravel(), c["u"])
        c5 = random_case("bk5", p, 5, 2000 + p)
        o5, _ = oracle.ref.bk5(c5["nq"], c5["dbasis"], serial_G(c5["G"]).ravel(), c5["u"])
        out["random_sha256"][str(p)] = {"bk1": hashlib.sha256(o1.tobytes()).hexdigest(),
                                        "bk3": hashlib.sha256(o3.tobytes()).hexdigest(),
                                        "bk5": hashlib.sha256(o5.tobytes()).hexdigest()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bk_norms.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
