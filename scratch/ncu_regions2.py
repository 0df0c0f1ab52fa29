#!/usr/bin/env python3
"""Per-source-line instruction counts for ONE kernel (by name substring) of a multi-kernel .ncu-rep."""
import csv, subprocess, sys
rep, kname = sys.argv[1], sys.argv[2]; thresh = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:' + kname], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = None
for r in rows:
    if r and r[0] == 'Line No' and len(r) > 4: h = r; break
if h is None: print(out[:2000]); sys.exit(1)
ci = h.index('Instructions Executed'); cs = h.index('# Samples') if '# Samples' in h else None
data = {}; samp = {}
for r in rows:
    if len(r) == len(h) and r[0].isdigit():
        try: n = int(r[ci].replace(',', '') or 0); sm = int(r[cs].replace(',', '') or 0) if cs is not None else 0
        except ValueError: continue
        k = (int(r[0]), r[1].strip()[:110])
        data[k] = data.get(k, 0) + n; samp[k] = samp.get(k, 0) + sm
tot = sum(data.values()); ts = max(1, sum(samp.values()))
print('total inst', tot, 'samples', ts)
for (ln, src), n in sorted(data.items()):
    if 100 * n / tot >= thresh or 100 * samp[(ln, src)] / ts >= thresh: print(f'{ln:5d} {100*n/tot:5.1f}% i {100*samp[(ln,src)]/ts:5.1f}% s  {src}')
