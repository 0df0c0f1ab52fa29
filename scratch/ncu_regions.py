#!/usr/bin/env python3
"""Per-source-line instruction counts from an .ncu-rep, printed in line order (file:line count% text)."""
import csv, subprocess, sys
rep = sys.argv[1]; thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = None
for r in rows:
    if r and r[0] == 'Line No' and len(r) > 4: h = r; break
ci = h.index('Instructions Executed')
data = {}
for r in rows:
    if len(r) == len(h) and r[0].isdigit():
        try: n = int(r[ci].replace(',', '') or 0)
        except ValueError: continue
        k = (int(r[0]), r[1].strip()[:120])
        data[k] = data.get(k, 0) + n
tot = sum(data.values())
print('total', tot)
for (ln, src), n in sorted(data.items()):
    if 100 * n / tot >= thresh: print(f'{ln:5d} {100*n/tot:5.1f}%  {src}')
