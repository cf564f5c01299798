#!/usr/bin/env python
"""Splits `cuobjdump -sass` output per function and prints opcode histograms (developer tool)."""
import collections, re, subprocess, sys
lib = sys.argv[1]; pat = sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = {}; cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: cur = m.group(1); funcs[cur] = []; continue
    if cur: funcs[cur].append(line)
for name, lines in funcs.items():
    if not re.search(pat, name): continue
    ops = collections.Counter()
    n = 0
    for l in lines:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", l)
        if m: ops[m.group(2) + ((m.group(3) or "") if m.group(2) in ("IMAD","MOV","LDG","LDS","STS","LD","ST","LDC","LDCU","ULDC") else "")] += 1; n += 1
    print(name, n, "instructions")
    print("  ", ", ".join(f"{k}:{v}" for k, v in ops.most_common(28)))
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write("\n".join(lines))
