#!/usr/bin/env python
"""Top stall sites of an `ncu --page source --csv --print-source sass` dump.  usage: ncu_hot.py <source.csv> [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) > iexec and r[ia].startswith("0x")]
total = sum(int(r[isamp] or 0) for r in body)
texec = sum(int(r[iexec] or 0) for r in body)
print(f"total samples {total}, warp instructions executed {texec}")
base = int(body[0][ia], 16)
order = sorted(range(len(body)), key=lambda k: -int(body[k][isamp] or 0))[:top]
for k in sorted(order):
    r = body[k]
    st = sorted(((int(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:2]
    print(f"{int(r[ia],16)-base:6x} {int(r[isamp]):7d} {100*int(r[isamp])/total:5.1f}%  exec {int(r[iexec]):9d}  {r[isrc].strip():60s} {st}")
