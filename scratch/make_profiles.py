#!/usr/bin/env python3
"""Turns gpurun_out/<tag>_* into the committed evidence under profiles/: bench lines, the ncu launch list with per-kernel
shares, one text summary per --set full capture, and profiles/ncu_traffic.json (DRAM bytes per launch, read by bench.py)."""
import csv, json, os, subprocess, sys, collections
tag = sys.argv[1] if len(sys.argv) > 1 else "r01b"
frames_per_launch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)
for f in ("bench.json", "bench_reference.json", "launches.csv", "bench_cfg3.json", "bench_cfg4.json", "bench_n2_replicas.json"):
    src = os.path.join(G, f"{tag}_{f}")
    if os.path.exists(src):
        open(os.path.join(P, f"{tag}_{f}"), "w").write(open(src).read())
# launch list shares
rows = [r for r in csv.reader(open(os.path.join(G, f"{tag}_launches.csv"))) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = collections.OrderedDict()
for r in rows[1:]:
    if r[ki] == "Kernel Name": continue
    try: tot[r[ki]] = tot.get(r[ki], [0, 0.0]); tot[r[ki]][0] += 1; tot[r[ki]][1] += float(r[vi].replace(",", ""))
    except ValueError: pass
mine = {k: v for k, v in tot.items() if k.startswith("k_") or k.startswith("void k_") or "vsb" in k}
s = sum(v[1] for v in mine.values())
with open(os.path.join(P, f"{tag}_launch_shares.txt"), "w") as fh:
    fh.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 2 --warmup 3` (cold-cache, serialised: compare SHARES)\n")
    for k, (n, t) in sorted(mine.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"{k[:60]:60s} launches {n:4d}  total {t/1000:10.1f} us  share {100*t/s:5.1f}%\n")
print(open(os.path.join(P, f"{tag}_launch_shares.txt")).read())
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem']
traffic = {"note": f"dram__bytes_read.sum + dram__bytes_write.sum of one launch from ncu --set full --clock-control none ({tag}); bench.py scales by frames per launch", "kernels": {}}
names = {"k_blend": "blend", "k_remap_stage1_tab": "remap_stage1", "k_remap_stage2_tab": "remap_stage2", "k_down2": "down2", "k_coarse": "coarse", "k_down_tail": "down_tail"}
with open(os.path.join(P, f"{tag}_ncu_full_summary.txt"), "w") as fh:
    for k, short in names.items():
        rep = os.path.join(G, f"{tag}_prof_{k}.ncu-rep")
        if not os.path.exists(rep): continue
        out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rr = list(csv.reader(out.splitlines()))
        if len(rr) < 3: continue
        h, u, r = rr[0], rr[1], rr[2]
        fh.write(f"## {k}  ({frames_per_launch} frames per launch)\n")
        vals = {}
        for w in WANT:
            if w in h:
                vals[w] = (r[h.index(w)], u[h.index(w)]); fh.write(f"  {w:72s} {r[h.index(w)]:>18s} {u[h.index(w)]}\n")
        def tobytes(v, unit):
            v = float(v.replace(",", "")); return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        try:
            b = tobytes(*vals['dram__bytes_read.sum']) + tobytes(*vals['dram__bytes_write.sum'])
            traffic["kernels"][short] = {"dram_bytes_per_launch": b, "frames_per_launch": frames_per_launch}
        except Exception as e:
            print("traffic", k, e)
json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(traffic["kernels"], indent=1))
