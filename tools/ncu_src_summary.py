#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: stall mix, per-opcode sample share, hottest SASS lines."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot, 'instrs', len(data))
stalls = ['stall_barrier', 'stall_branch_resolving', 'stall_dispatch', 'stall_lg', 'stall_long_sb', 'stall_math',
          'stall_mio', 'stall_no_inst', 'stall_not_selected', 'stall_selected', 'stall_short_sb', 'stall_wait',
          'stall_misc', 'stall_sleep', 'stall_membar']
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls if s in ix}
print({k: round(100 * v / tot, 1) for k, v in agg.items()})
op = defaultdict(lambda: [0, 0])
for r in data:
    parts = r[ix['Source']].strip().split()
    o = parts[0] if not parts[0].startswith('@') else parts[1]
    o = o.split('.')[0]
    op[o][0] += int(r[ix['# Samples']])
    op[o][1] += int(r[ix['Instructions Executed']])
for o, (s, n) in sorted(op.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f"{o:10s} samples {100 * s / tot:5.1f}%  executed {n:10d}  samples/kexec {s / max(n, 1) * 1000:.2f}")
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:ntop]:
    st = {s: int(r[ix[s]]) for s in stalls if s in ix and int(r[ix[s]]) > 0}
    print(r[ix['# Samples']], r[ix['Source']].strip()[:70], st)
