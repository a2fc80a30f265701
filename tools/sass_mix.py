#!/usr/bin/env python
"""Opcode mix of a profiled kernel from an .ncu-rep source page: python tools_sass_mix.py file.ncu-rep [N]"""
import collections, csv, re, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iE, iSamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ops, samp, tot = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    if len(r) <= iE:
        continue
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[iS].strip())
    op = m.group(2) if m else r[iS][:10]
    base = op if op.split('.')[0] in ('F2F', 'I2F', 'F2I', 'LDS', 'LDG', 'STG', 'STS') else op.split('.')[0]
    n = int(r[iE]); ops[base] += n; samp[base] += int(r[iSamp]); tot += n
print('total warp-instructions', tot)
for k, v in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f'{k:24s} {v:10d} {v / tot:6.3f}  stall samples {samp[k]}')
