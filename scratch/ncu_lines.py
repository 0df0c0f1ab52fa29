#!/usr/bin/env python3
"""Per-source-line instruction counts and stall samples from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
lines = out.splitlines()
# the cuda,sass view: blocks of source lines, each followed by their SASS rows; find header rows
rows = list(csv.reader(lines))
agg = collections.OrderedDict()
cur_file = None; cur = None; hdr = None
tot_i = tot_s = 0
for r in rows:
    if not r: continue
    if r[0] == 'File Name': cur_file = r[1].split('/')[-1]; continue
    if r[0] in ('Line No', 'Address', '#'): hdr = r; continue
    if hdr is None: continue
    try:
        if 'Source' in hdr and hdr[0] == 'Line No':
            pass
    except Exception:
        pass
sys.stdout.write('')
# simpler: use the "cuda" correlated view columns if present
h = None
for r in rows:
    if r and r[0] == 'Line No' and len(r) > 4: h = r; break
if h is None:
    print('no correlated view; header sample:', [r for r in rows if r and r[0] in ('Line No', 'Address')][:2]); sys.exit(0)
ci = h.index('Instructions Executed'); cs = h.index('# Samples') if '# Samples' in h else None
data = []
f = None
for r in rows:
    if r and r[0] == 'File Name': f = r[1].split('/')[-1]; continue
    if len(r) == len(h) and r[0].isdigit():
        try:
            n = int(r[ci].replace(',', '') or 0); s = int(r[cs].replace(',', '') or 0) if cs is not None else 0
        except ValueError:
            continue
        if n or s: data.append((n, s, f, int(r[0]), r[1].strip()[:110]))
ti = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print(f'total inst {ti} samples {ts}')
for n, s, f, ln, src in sorted(data, reverse=True)[:top]:
    print(f'{100*n/ti:5.1f}% inst {100*s/max(ts,1):5.1f}% stall  {f}:{ln}  {src}')
