#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output of inst.cu into a table: kernel variant, regs, spills, smem."""
import re, sys
txt = sys.stdin.read()
cur = None
rows = []
for line in txt.splitlines():
    m = re.search(r"sumfact2?_kernelILi(\d+)ELi(\d+)ELb([01])ELi(\d)ELb([01])ELi(\d+)ELi(\d+)E", line)
    if m and "Compiling" in line:
        cur = dict(nm=int(m[1]), nq=int(m[2]), coll=int(m[3]), qop=int(m[4]), lvec=int(m[5]), epb=int(m[6]), minb=int(m[7]))
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        cur["spill"] = int(m[2]) + int(m[3])
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = int(m[1])
        rows.append(cur)
        cur = None
print(f"{'nm':>3}{'nq':>3}{'coll':>5}{'qop':>4}{'lvec':>5}{'epb':>4}{'thr':>5}{'regs':>5}{'spillB':>7}")
for r in sorted(rows, key=lambda r: (r['nm'], r['lvec'], r['qop'], r['coll'], r['nq'])):
    print(f"{r['nm']:>3}{r['nq']:>3}{r['coll']:>5}{r['qop']:>4}{r['lvec']:>5}{r['epb']:>4}{r['epb']*r['nq']**2:>5}{r['regs']:>5}{r.get('spill',0):>7}")
