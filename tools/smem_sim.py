#!/usr/bin/env python
"""Shared-memory wavefront model of sumfact2_kernel (interpolated Laplacian, CTA-barrier mode).

Replays every shared-memory access site of one element batch with the kernel's index formulas and counts
64-bit wavefronts per half-warp (16 lanes; lanes hitting the same 8-byte word broadcast, distinct words in
the same slot (address mod 16) serialise).  Used to pick strides / element offsets offline; validated against
ncu (l1tex__data_pipe_lsu_wavefronts_mem_shared) for BK3 p=4: measured 580 wavefronts per element.

  python tools/smem_sim.py            # table for the E-vector BK3 shapes
  python tools/smem_sim.py --cart     # the nodal separable kernel (sumfact_cart.cuh): wavefronts and bank conflicts per degree
"""
from __future__ import annotations

import sys


def odd(n):
    return n | 1


def pad_plane_q(nq):
    ps = nq * nq
    while (ps - nq) % 16:
        ps += 1
    return ps


def best_strides(nm, nq):
    def hw(n_threads, n_active, div, s_hi, s_lo):
        tot = 0
        for h in range(0, n_threads, 16):
            cnt = [0] * 16
            worst = 0
            for t in range(h, min(h + 16, n_active)):
                s = ((t // div) * s_hi + (t % div) * s_lo) % 16
                cnt[s] += 1
                worst = max(worst, cnt[s])
            tot += worst
        return tot

    def cost(pa, ra):
        return nq * hw(nq * nq, nm * nm, nm, pa, ra) + nm * hw(nq * nq, nm * nq, nq, pa, 1)

    best = (odd(nq), nm * odd(nq))
    bc = cost(best[1], best[0])
    for ra in range(nq, nq + 5):
        for pa in range(nm * ra, nm * ra + 17):
            c = cost(pa, ra)
            if c < bc or (c == bc and nm * pa < nm * best[1]):
                bc, best = c, (ra, pa)
    return best


class Layout:
    def __init__(self, nm, nq, epb, elem_off=None, g_stride=None):
        self.nm, self.nq, self.epb = nm, nq, epb
        self.n2, self.n3, self.m3 = nq * nq, nq ** 3, nm ** 3
        self.psq = pad_plane_q(nq)
        self.rsr = odd(nq)
        self.psr = nq * self.rsr
        self.ra, self.pa = best_strides(nm, nq)
        self.pb = self.psq
        self.ru = odd(nm)
        sz_flux = max(nq * self.psq, nq * self.psr)
        sz_interp = max(nm * self.pa, nm * self.pb, nm * nm * self.ru)
        self.region = (max(sz_flux, sz_interp) + 1) & ~1
        w = 3 * self.region
        while (w - self.n2) % 16:
            w += 1
        self.wpe = w
        # per-view element offsets (default: one common element stride, as in the kernel)
        self.elem_off = elem_off or {}
        self.g_stride = g_stride if g_stride is not None else 6 * self.n3

    def base(self, view, region, el):
        if view in self.elem_off:
            return region * self.epb * self.region_block() + el * self.elem_off[view]
        return el * self.wpe + region * self.region

    def region_block(self):
        return max([self.region] + list(self.elem_off.values()))


def wavefronts(addrs):
    """addrs: list (by tid) of double-index or None.  Returns number of 64-bit wavefronts."""
    total = 0
    for h in range(0, len(addrs), 16):
        slots = {}
        for a in addrs[h:h + 16]:
            if a is None:
                continue
            slots.setdefault(a % 16, set()).add(a)
        if slots:
            total += max(len(v) for v in slots.values())
    return total


def simulate(L: Layout, verbose=False):
    nm, nq, epb, n2, m3 = L.nm, L.nq, L.epb, L.n2, L.m3
    T = epb * n2
    T32 = (T + 31) // 32 * 32
    el_of = [t // n2 if t < T else None for t in range(T32)]
    t2_of = [t % n2 if t < T else None for t in range(T32)]
    sites = {}

    def run(name, fn, count=1):
        """fn(el, t2) -> address or None; count = how many times the site executes per batch (already expanded loops)."""
        addrs = [None if el_of[t] is None else fn(el_of[t], t2_of[t]) for t in range(T32)]
        sites[name] = sites.get(name, 0) + count * wavefronts(addrs)

    nk = (m3 + n2 - 1) // n2
    # U staging
    for c in range(nk):
        run("U.write", lambda el, t2, c=c: (L.base("U", 0, el) + ((t2 + c * n2) // nm) * L.ru + (t2 + c * n2) % nm) if t2 + c * n2 < m3 else None)
    for k in range(nm):
        run("U.read", lambda el, t2, k=k: (L.base("U", 0, el) + t2 * L.ru + k) if t2 < nm * nm else None)
    for r in range(nq):
        run("A.write", lambda el, t2, r=r: (L.base("A", 1, el) + (t2 // nm) * L.pa + (t2 % nm) * L.ra + r) if t2 < nm * nm else None)
    for j in range(nm):
        run("A.read", lambda el, t2, j=j: (L.base("A", 1, el) + (t2 // nq) * L.pa + j * L.ra + t2 % nq) if t2 < nm * nq else None)
    for q in range(nq):
        run("B.write", lambda el, t2, q=q: (L.base("B", 2, el) + (t2 // nq) * L.pb + q * nq + t2 % nq) if t2 < nm * nq else None)
    for i in range(nm):
        run("B.read", lambda el, t2, i=i: L.base("B", 2, el) + i * L.pb + t2)
    # flux part: V writes, Q/R in place, flux reads/writes + G reads, Q/R, final reads
    for p in range(nq):
        run("RQ.P", lambda el, t2, p=p: L.base("RQ", 0, el) + p * L.psq + t2, count=4)     # write v, read qs, write fs, read ws
        run("RR.P", lambda el, t2, p=p: L.base("RR", 1, el) + p * L.psr + (t2 // nq) * L.rsr + t2 % nq, count=4)
        for c in range(6):
            run("G.read", lambda el, t2, p=p, c=c: 10 ** 6 + el * L.g_stride + c * L.n3 + p * n2 + t2)
    for q in range(nq):
        run("RQ.Q", lambda el, t2, q=q: L.base("RQ", 0, el) + (t2 // nq) * L.psq + q * nq + t2 % nq, count=4)
    for r in range(nq):
        run("RR.R", lambda el, t2, r=r: L.base("RR", 1, el) + (t2 // nq) * L.psr + (t2 % nq) * L.rsr + r, count=4)
    # backward
    for i in range(nm):
        run("X.write", lambda el, t2, i=i: L.base("B", 2, el) + i * L.pb + t2)
    for q in range(nq):
        run("X.read", lambda el, t2, q=q: (L.base("B", 2, el) + (t2 // nq) * L.pb + q * nq + t2 % nq) if t2 < nm * nq else None)
    for j in range(nm):
        run("Y.write", lambda el, t2, j=j: (L.base("Y", 0, el) + (t2 // nq) * L.pa + j * L.ra + t2 % nq) if t2 < nm * nq else None)
    for r in range(nq):
        run("Y.read", lambda el, t2, r=r: (L.base("Y", 0, el) + (t2 // nm) * L.pa + (t2 % nm) * L.ra + r) if t2 < nm * nm else None)
    for k in range(nm):
        run("Z.write", lambda el, t2, k=k: (L.base("Z", 1, el) + t2 * L.ru + k) if t2 < nm * nm else None)
    for c in range(nk):
        run("Z.read", lambda el, t2, c=c: (L.base("Z", 1, el) + ((t2 + c * n2) // nm) * L.ru + (t2 + c * n2) % nm) if t2 + c * n2 < m3 else None)
    tma = epb * 6 * L.n3 * 8 / 128.0
    total = sum(sites.values())
    ideal = 0
    if verbose:
        for k, v in sites.items():
            print(f"   {k:10s} {v / epb:8.1f} per element")
    return total / epb, tma / epb, sites


def simulate_cart(nm, epb):
    """The nodal separable kernel (csrc/sumfact_cart.cuh, Laplace variant): nm^2 threads per element, two staging arrays
    X1 / X2 with the strides of best_strides(nm, nm), accessed in three layouts -- P: thread (j,k), loop over i; R: thread
    (i,j), loop over k; Q: thread (i,k), loop over j.  Per batch and thread: P nm stores + 2 nm loads, R nm loads + 2 nm
    stores, Q 2 nm loads + 2 nm stores.  Returns (wavefronts, conflict-free wavefronts) per element batch of the CTA.
    Checked against ncu (profiles/r02zz_bp3_p4_cart_kernel_summary.txt, nm = 5, 5 elements per CTA): 35 % of the
    shared-memory wavefronts are bank conflicts; this model: 620 wavefronts against 400 conflict-free = 35.5 %."""
    ra, pa = best_strides(nm, nm)
    n2 = nm * nm
    arr = (nm * pa + 1) & ~1
    wpe = 2 * arr
    while (wpe - n2) % 16:
        wpe += 1
    threads = epb * n2

    def cost(addr):
        tot = ideal = 0
        for h in range(0, threads, 16):
            cnt = {}
            for t in range(h, min(h + 16, threads)):
                el, t2 = divmod(t, n2)
                ta, tb = divmod(t2, nm)
                a = addr(el, ta, tb) % 16
                cnt[a] = cnt.get(a, 0) + 1
            tot += max(cnt.values())
            ideal += 1
        return tot, ideal

    p_c, p_i = cost(lambda el, ta, tb: el * wpe + ta * ra + tb)
    r_c, r_i = cost(lambda el, ta, tb: el * wpe + ta * pa + tb * ra)
    q_c, q_i = cost(lambda el, ta, tb: el * wpe + ta * pa + tb)
    n_p, n_r, n_q = 3 * nm, 3 * nm, 4 * nm
    return n_p * p_c + n_r * r_c + n_q * q_c, n_p * p_i + n_r * r_i + n_q * q_i, (ra, pa, wpe), (p_c, r_c, q_c, p_i)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--cart":
        print(f"{'nm':>3}{'epb':>4}{'strides (row, plane, element)':>32}{'P/R/Q wavefronts per access (ideal)':>38}{'wavefronts/batch':>18}{'conflicts':>11}")
        for nm, epb in ((2, 32), (3, 14), (4, 8), (5, 5), (6, 3), (7, 2), (8, 2), (9, 1)):
            w, ideal, strides, per = simulate_cart(nm, epb)
            print(f"{nm:>3}{epb:>4}{str(strides):>32}{str(per[:3]) + ' (' + str(per[3]) + ')':>38}{w:>18}{100.0 * (w - ideal) / w:>10.1f}%")
        return
    print(f"{'nm':>3}{'nq':>3}{'epb':>4}{'wavefronts/elem':>17}{'+TMA fill':>10}{'ideal(no conflicts)':>21}")
    for nm, nq, epb in ((2, 3, 14), (3, 4, 8), (4, 5, 5), (5, 6, 3), (6, 7, 1), (7, 8, 1), (8, 9, 1), (9, 10, 1), (5, 6, 1), (4, 5, 1)):
        L = Layout(nm, nq, epb)
        w, tma, sites = simulate(L)
        # ideal: every access of an active thread costs 1/16 wavefront
        L1 = Layout(nm, nq, 1)
        w1, _, _ = simulate(L1)
        print(f"{nm:>3}{nq:>3}{epb:>4}{w:>17.1f}{tma:>10.1f}{w1:>21.1f}")
    if len(sys.argv) > 1:
        nm, nq, epb = map(int, sys.argv[1:4])
        simulate(Layout(nm, nq, epb), verbose=True)


if __name__ == "__main__":
    main()
