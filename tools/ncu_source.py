#!/usr/bin/env python3
"""tools/ncu_source.py <report.ncu-rep> [kernel-substring] -- per-instruction view of an ncu --set full report (run here, no GPU): shared-memory
instructions with their wavefronts (excess = bank conflicts), and the instructions with the most stall samples."""
import csv, subprocess, sys
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
want = sys.argv[2] if len(sys.argv) > 2 else ''
blocks, cur = [], None
for row in csv.reader(raw.splitlines()):
    if row and row[0] == 'Kernel Name':
        cur = {'name': row[1], 'rows': [], 'hdr': None}
        blocks.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = row
    elif cur is not None and row:
        cur['rows'].append(row)
for b in blocks:
    if want not in b['name']:
        continue
    h = {n: i for i, n in enumerate(b['hdr'])}
    print('=====', b['name'][:90])
    tot = sum(int(r[h['# Samples']]) for r in b['rows'])
    exe = sum(int(r[h['Instructions Executed']]) for r in b['rows'])
    print('samples', tot, 'warp instructions executed', exe)
    print('-- shared-memory instructions: executed, wavefronts, ideal, excessive')
    for r in b['rows']:
        if r[h['Address Space']] == 'Shared' or 'LDS' in r[h['Source']] or 'STS' in r[h['Source']]:
            print('   %-48s %9s %10s %10s %10s  samples %s' % (r[h['Source']].strip()[:48], r[h['Instructions Executed']], r[h['L1 Wavefronts Shared']],
                                                   r[h['L1 Wavefronts Shared Ideal']], r[h['L1 Wavefronts Shared Excessive']], r[h['# Samples']]))
    print('-- top 40 instructions by stall samples')
    stall_cols = [n for n in b['hdr'] if n.startswith('stall_') and 'Not Issued' not in n]
    for r in sorted(b['rows'], key=lambda r: -int(r[h['# Samples']]))[:40]:
        st = sorted(((int(r[h[c]]), c) for c in stall_cols), reverse=True)[:2]
        print('   %-60s %7s  %s' % (r[h['Source']].strip()[:60], r[h['# Samples']], ' '.join('%s=%d' % (c[6:], v) for v, c in st)))
