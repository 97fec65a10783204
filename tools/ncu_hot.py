#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from an ncu report (source page)."""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS = hdr.index('# Samples'); iSrc = hdr.index('Source'); iEx = hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[iS]) for r in body)
print('total samples', tot, ' instructions', len(body))
top = sorted(range(len(body)), key=lambda i: -int(body[i][iS]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i in sorted(top):
    r = body[i]
    st = sorted(((int(r[j]), hdr[j]) for j in stall if int(r[j]) > 0), reverse=True)[:2]
    print('%5d %5.1f%% ex=%-9s %-60s %s' % (i, 100.0 * int(r[iS]) / tot, r[iEx], r[iSrc].strip()[:60], st))
