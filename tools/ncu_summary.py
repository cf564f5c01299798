#!/usr/bin/env python
"""Summarises an .ncu-rep: key raw metrics per launch + stall/opcode mix of the first kernel (developer tool)."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.avg',
        'launch__grid_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__cycles_active.avg']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w} [{units[i]}]:", [r[i] for r in rows[2:]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = []; secs.append(cur); continue
    if cur is not None: cur.append(r)
sec = secs[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
h = sec[0]; data = sec[1:]; ci = {x: i for i, x in enumerate(h)}
tot = collections.Counter()
for r in data:
    for x in h:
        if x.startswith('stall_') and 'Not Issued' not in x:
            try: tot[x] += int(r[ci[x]])
            except: pass
s = sum(tot.values())
print("stalls:", ", ".join(f"{k[6:]}:{100*v/s:.1f}%" for k, v in tot.most_common(10)))
ops = collections.Counter(); samp = collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[ci['Source']])
    if not m: continue
    try: ops[m.group(2)] += int(r[ci['Instructions Executed']]); samp[m.group(2)] += int(r[ci['# Samples']])
    except: pass
t = sum(ops.values()); ts = sum(samp.values())
print("total warp-inst", t)
print("opcodes:", ", ".join(f"{k}:{100*v/t:.1f}%({100*samp[k]/ts:.0f}%s)" for k, v in ops.most_common(22)))
if len(sys.argv) > 3:
    # top sampled instructions
    top = sorted(data, key=lambda r: -int(r[ci['# Samples']] or 0))[:int(sys.argv[3])]
    for r in top: print(r[ci['# Samples']], r[ci['Source']][:90])
