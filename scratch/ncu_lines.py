#!/usr/bin/env python3
"""Per-source-line instruction counts / stall samples from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; agg = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) > 8 and r[0].strip().isdigit():
        try:
            agg.append((cur, int(r[0]), r[1].strip()[:110], int(r[hdr.index('Instructions Executed')]), int(r[hdr.index('# Samples')])))
        except ValueError:
            pass
tot_i = sum(a[3] for a in agg); tot_s = sum(a[4] for a in agg)
print(f'total warp-instructions {tot_i}, samples {tot_s}')
for a in sorted(agg, key=lambda a: -a[3])[:topn]:
    print(f'{a[0]}:{a[1]:4d} inst {100*a[3]/tot_i:5.1f}%  samp {100*a[4]/max(tot_s,1):5.1f}%  {a[2]}')
