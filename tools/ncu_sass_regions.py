#!/usr/bin/env python
"""Split the SASS of a profiled kernel into regions delimited by barriers / cp.async waits and print, per region, executed
warp instructions, stall samples and the dominant stall reasons (run here on a .ncu-rep with source counters)."""
import csv, subprocess, sys, collections
f = sys.argv[1]
out = subprocess.run(['ncu', '-i', f, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
regions = []
cur = {'start': 0, 'n': 0, 'inst': 0, 'samples': 0, 'stalls': collections.Counter(), 'ops': collections.Counter(), 'label': 'entry'}
tot_inst = tot_s = 0
for k, r in enumerate(rows[2:]):
    if len(r) < len(hdr):
        continue
    src = r[ix['Source']].strip()
    op = src.split()[0] if not src.startswith('@') else src.split()[1]
    inst = int(r[ix['Instructions Executed']] or 0)
    smp = int(r[ix['# Samples']] or 0)
    cur['n'] += 1; cur['inst'] += inst; cur['samples'] += smp
    cur['ops'][op.split('.')[0]] += inst
    for c in stall_cols:
        v = int(r[ix[c]] or 0)
        if v: cur['stalls'][c] += v
    tot_inst += inst; tot_s += smp
    if op.startswith('BAR') or op.startswith('LDGDEPBAR') or op.startswith('DEPBAR') or op.startswith('BRA') and inst and False:
        cur['end'] = k; cur['endop'] = src
        regions.append(cur)
        cur = {'start': k + 1, 'n': 0, 'inst': 0, 'samples': 0, 'stalls': collections.Counter(), 'ops': collections.Counter(), 'label': ''}
cur['end'] = len(rows); cur['endop'] = 'end'
regions.append(cur)
print(f'total warp-instructions {tot_inst}, samples {tot_s}')
for g in regions:
    if g['inst'] == 0 and g['samples'] == 0: continue
    st = ', '.join(f'{k[6:]}={v}' for k, v in g['stalls'].most_common(4))
    ops = ', '.join(f'{k}={v}' for k, v in g['ops'].most_common(6))
    print(f"[{g['start']:5d}-{g['end']:5d}] sass={g['n']:5d} inst={g['inst']:9d} ({100*g['inst']/tot_inst:4.1f}%) samples={g['samples']:6d} ({100*g['samples']/max(tot_s,1):4.1f}%)  {st}\n        ops: {ops}   ends: {g['endop'][:50]}")
