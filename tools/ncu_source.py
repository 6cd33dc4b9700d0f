#!/usr/bin/env python
"""Per-SASS-line view of an .ncu-rep captured with --import-source on: instruction counts, active
threads and stall samples, grouped into contiguous address blocks.  Usage: ncu_source.py REP [launch_idx] [block]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 32
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# first kernel only
hdr = rows[1]
ia, isrc, isamp, iinst, ithr = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
body = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == 'Kernel Name' or r[0] == 'Address':
        if r and r[0] == 'Kernel Name':
            break
        continue
    body.append(r)
tot_inst = sum(int(r[iinst]) for r in body)
tot_samp = sum(int(r[isamp]) for r in body)
print(f"lines {len(body)} inst {tot_inst} samples {tot_samp}")
base = int(body[0][ia], 16)
for s in range(0, len(body), blk):
    chunk = body[s:s + blk]
    inst = sum(int(r[iinst]) for r in chunk)
    thr = sum(int(r[ithr]) for r in chunk)
    samp = sum(int(r[isamp]) for r in chunk)
    ops = {}
    for r in chunk:
        op = r[isrc].split()[0] if not r[isrc].strip().startswith('@') else r[isrc].split()[1]
        op = op.split('.')[0]
        ops[op] = ops.get(op, 0) + 1
    top = ' '.join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
    print(f"+0x{int(chunk[0][ia],16)-base:05x} inst {100*inst/tot_inst:5.1f}%  avg_thr {thr/max(inst,1):5.1f}  samples {100*samp/max(tot_samp,1):5.1f}%  {top}")

# top individual instructions by stall samples
if len(sys.argv) > 3:
    col = sys.argv[3]
    ic = hdr.index(col)
    top = sorted(body, key=lambda r: -int(r[ic]))[:25]
    print(f"--- top instructions by {col}")
    for r in top:
        print(f"+0x{int(r[ia],16)-base:05x} {int(r[ic]):7d} inst={r[iinst]:>10s} thr/inst={int(r[ithr])/max(int(r[iinst]),1):5.1f}  {r[isrc][:90]}")
