#!/usr/bin/env python
"""Histogram of the hottest loop (innermost backward branch spanning the most FFMA2/DFMA) of one kernel.
usage: sass_loop.py <obj-or-so> <substring of mangled kernel name>"""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, funcs = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and cur: funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name: continue
    addr = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\s+(?:U?P\d+,\s*)?0x([0-9a-f]+)", t)
        if m and "BRA.DIV" not in t:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                body = ins[addr[tgt]:i + 1]
                n = sum(1 for _, x in body if re.search(r"\b(FFMA2|FMUL2|DFMA|DMUL)\b", x.split()[0] if not x.startswith("@") else x.split()[1]))
                inner = not any(re.search(r"BRA\s+(?:U?P\d+,\s*)?0x([0-9a-f]+)", x) and "BRA.DIV" not in x and int(re.search(r"0x([0-9a-f]+)", x).group(1), 16) < aa and int(re.search(r"0x([0-9a-f]+)", x).group(1), 16) >= tgt for aa, x in body[:-1])
                if n and (best is None or n > best[0]) and inner: best = (n, tgt, a, body)
    if not best: print(name, "no loop"); continue
    n, tgt, a, body = best
    h = collections.Counter()
    for _, x in body:
        parts = x.split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        op = op.split(".")[0] + (".MOV" if ".MOV" in op else "")
        h[op] += 1
    print(f"{name}\n  loop 0x{tgt:x}..0x{a:x}: {len(body)} instrs:", " ".join(f"{k}:{v}" for k, v in h.most_common()))
