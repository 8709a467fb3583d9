#!/usr/bin/env python
"""Top SASS instructions of every kernel in an .ncu-rep by warp-stall samples (captured with --set full --import-source on).
usage: python tools/ncu_hot.py file.ncu-rep [top_n]"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'body': []}
        blocks.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = r
    elif cur is not None:
        cur['body'].append(r)
seen = set()
for b in blocks:
    hdr = b['hdr']
    sig = (b['name'], len(b['body']), tuple(tuple(r) for r in b['body'][:50]))
    if sig in seen:
        continue
    seen.add(sig)
    isrc, isamp, iexec = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    body = [r for r in b['body'] if len(r) > iexec]
    total = sum(int(r[isamp] or 0) for r in body)
    texec = sum(int(r[iexec] or 0) for r in body)
    print(f"== {b['name']}\n{len(body)} SASS instructions, {total} samples, {texec} warp instructions executed")
    order = sorted(range(len(body)), key=lambda i: -int(body[i][isamp] or 0))[:top]
    for i in sorted(order):
        r = body[i]
        print(f"{i:6d} {100.0 * int(r[isamp] or 0) / max(total, 1):5.1f}%  exec {int(r[iexec] or 0):>11d}  {r[isrc].strip()}")
