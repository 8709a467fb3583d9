#!/usr/bin/env python
"""
SASS excerpt of a kernel's hot loop, from the built library (cuobjdump -sass).

    python tools/sass_excerpt.py <function-regex> <key-mnemonic>[,<also-required>...] [lib.so]

Prints the instruction mix of the whole function (count per mnemonic) and, among the innermost loops (a backward branch
and its target, no other loop inside) that hold every listed mnemonic, the one with the highest density of
<key-mnemonic>: the loop the kernel spends its time in.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "mdcraft_b200", "_lib", "libmdcraft_b200.so")
pat, keys = re.compile(sys.argv[1]), sys.argv[2].split(",")
key = keys[0]
text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in text.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        funcs[cur].append(line)
names = [n for n in funcs if pat.search(n)]
if not names:
    raise SystemExit(f"no function matches {sys.argv[1]}")
name = max(names, key=lambda n: sum(key in l for l in funcs[n]))
demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
ins = []
for l in funcs[name]:
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
mix = collections.Counter()
for _, t in ins:
    op = t.split()[1] if t.startswith("@") else t.split()[0]
    mix[op.split(".")[0] + ("." + op.split(".")[1] if op.split(".")[0] in ("MUFU", "ATOMS", "REDG", "RED", "LDS", "LDG", "STG", "STS") and "." in op else "")] += 1
print(f"function: {demangled}\n{len(ins)} SASS instructions; mix: " + ", ".join(f"{k} {v}" for k, v in mix.most_common(24)))
addr = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) in addr and addr[int(m.group(1), 16)] <= i:
        loops.append((addr[int(m.group(1), 16)], i))
best = None
for j, i in loops:
    if any((j2, i2) != (j, i) and j <= j2 and i2 <= i for j2, i2 in loops):
        continue                                   # not innermost
    n = sum(key in ins[k][1] for k in range(j, i + 1))
    if not all(any(q in ins[k][1] for k in range(j, i + 1)) for q in keys):
        continue
    if n and (best is None or n / (i - j + 1) > best[0] / (best[2] - best[1] + 1)):
        best = (n, j, i)
if best is None:
    raise SystemExit("no loop holds the key mnemonic")
n, j, i = best
print(f"hot loop: {i - j + 1} instructions, {n} x {key}")
for k in range(j, i + 1):
    print(f"  /*{ins[k][0]:05x}*/  {ins[k][1]}")
