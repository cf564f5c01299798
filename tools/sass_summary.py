#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of quantr_b200/libqsv.so (cuobjdump -sass; no GPU needed): shows which kernels move
tiles with TMA (UTMALDG / UTMASTG / UBLKCP + SYNCS mbarriers) and that none uses LDGSTS (cp.async).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections, os, re, subprocess
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "quantr_b200", "libqsv.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
cur, funcs = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = collections.Counter(); continue
    m = cur and re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", line)
    if m:
        funcs[cur][m.group(2)] += 1
names = list(funcs)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
keys = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "DFMA", "DMUL", "DADD"]
print("# cuobjdump -sass quantr_b200/libqsv.so (sm_100a), instruction counts per kernel; made by tools/sass_summary.py")
print(f"{'kernel':44s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in keys))
for n, d in sorted(zip(names, dem), key=lambda t: t[1]):
    c = funcs[n]
    d = re.sub(r"\(.*$", "", re.sub(r"^void ", "", d)).replace("qsv::", "")
    print(f"{d[:44]:44s} {sum(c.values()):6d} " + " ".join(f"{c.get(k, 0):7d}" for k in keys))
