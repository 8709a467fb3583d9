#!/usr/bin/env python
"""Top SASS instructions of a kernel by warp-stall samples, from an .ncu-rep captured with --set full --import-source on.
usage: python tools/ncu_hot.py file.ncu-rep [top_n]"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
body = [r for r in rows[2:] if len(r) > iexec]
total = sum(int(r[isamp] or 0) for r in body)
texec = sum(int(r[iexec] or 0) for r in body)
print(f"{len(body)} SASS instructions, {total} samples, {texec} warp instructions executed")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
order = sorted(range(len(body)), key=lambda i: -int(body[i][isamp] or 0))[:top]
for i in sorted(order):
    r = body[i]
    print(f"{i:5d} {100.0 * int(r[isamp] or 0) / max(total, 1):5.1f}%  exec {int(r[iexec] or 0):>11d}  {r[isrc].strip()}")
