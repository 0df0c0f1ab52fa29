#!/usr/bin/env python3
"""Round 2: turns gpurun_out/r02_* (one multi-kernel `ncu --set full` report + the launch list) and the bench lines into profiles/."""
import csv, json, os, subprocess, collections, shutil, sys
G, P, tag = "gpurun_out", "profiles", "r02"
copies = {"r2m_bench.json": "r02_bench.json", "r2l_bench_n2.json": "r02_bench_n2_shard_cfg3.json", "r2j_bench_n4.json": "r02_bench_n4_shard_cfg3.json",
          "r2k_bench_n2_cfg4.json": "r02_bench_n2_shard_cfg4.json", "r2l_bench_n2_fail.json": "r02_bench_n2_injected_failure_fallback.json",
          "r02_launches.csv": "r02_launches.csv", "r2h_bench_n2_b4.json": "r02_bench_n2_shard_cfg3_batch4.json"}
for a, b in copies.items():
    if os.path.exists(os.path.join(G, a)):
        lines = [l for l in open(os.path.join(G, a)).read().splitlines() if l.strip()]
        keep = [l for l in lines if l.startswith("{")] if a.endswith(".json") else lines
        open(os.path.join(P, b), "w").write("\n".join(keep) + "\n")
rows = [r for r in csv.reader(open(os.path.join(G, "r02_launches.csv"))) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = collections.OrderedDict()
for r in rows[1:]:
    if r[ki] == "Kernel Name": continue
    try: tot.setdefault(r[ki], [0, 0.0]); tot[r[ki]][0] += 1; tot[r[ki]][1] += float(r[vi].replace(",", ""))
    except ValueError: pass
mine = {k: v for k, v in tot.items() if "k_" in k}
other = {k: v for k, v in tot.items() if "k_" not in k}
hot = {k: v for k, v in mine.items() if any(n in k for n in ("k_remap_stage", "k_down2", "k_down_tail", "k_coarse", "k_blend"))}
s = sum(v[1] for v in hot.values())
with open(os.path.join(P, "r02_launch_shares.txt"), "w") as fh:
    fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400, `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline`\n"
             "# (cold-cache, serialised: compare SHARES, not absolutes).  Per-frame kernels first (shares among them), then calibration-time kernels, then library kernels.\n")
    for k, (n, t) in sorted(hot.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"{k[:70]:70s} launches {n:4d}  total {t/1000:10.1f} us  share {100*t/s:5.1f}%\n")
    fh.write("# calibration / recalibration kernels of this repository (run before the timed region)\n")
    for k, (n, t) in sorted(mine.items(), key=lambda kv: -kv[1][1]):
        if k not in hot: fh.write(f"{k[:70]:70s} launches {n:4d}  total {t/1000:10.1f} us\n")
    fh.write("# library kernels (torch fills / copies of the harness)\n")
    for k, (n, t) in sorted(other.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"{k[:70]:70s} launches {n:4d}  total {t/1000:10.1f} us\n")
print(open(os.path.join(P, "r02_launch_shares.txt")).read())
out = subprocess.run(['ncu', '-i', os.path.join(G, 'r02_prof.ncu-rep'), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(out.splitlines()))
h, u = rr[0], rr[1]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem']
names = [("k_remap_stage1_tab", "remap_stage1", 16), ("k_remap_stage2_tab", "remap_stage2", 16), ("k_down2", "down2", 8), ("k_down_tail", "down_tail", 8),
         ("k_coarse", "coarse", 8), ("k_blend_seam", "blend_seam", 8), ("k_blend_int", "blend_int", 8)]
traffic = {"note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch from ncu --set full --clock-control none (r02: 16-frame submissions; the remap "
                   "kernels see all 16 frames, the back half runs as two 8-frame sub-batches); bench.py scales by frames per launch", "kernels": {}}
def tobytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
tot_traffic = 0.0
with open(os.path.join(P, "r02_ncu_full_summary.txt"), "w") as fh, open(os.path.join(P, "r02_stalls.txt"), "w") as fs:
    fh.write("# one `ncu --set full --clock-control none --import-source on` capture (gpurun_out/r02_prof.ncu-rep) of `bench.py --steps 2 --warmup 3`, one launch per kernel\n")
    for r in rr[2:]:
        kn = r[h.index('Kernel Name')]
        m = [x for x in names if x[0] in kn]
        if not m: continue
        k, short, fpl = m[0]
        fh.write(f"## {kn[:60]}  ({fpl} frames per launch)\n")
        vals = {}
        for w in WANT:
            if w in h:
                vals[w] = (r[h.index(w)], u[h.index(w)]); fh.write(f"  {w:72s} {r[h.index(w)]:>18s} {u[h.index(w)]}\n")
        b = tobytes(*vals['dram__bytes_read.sum']) + tobytes(*vals['dram__bytes_write.sum'])
        traffic["kernels"][short] = {"dram_bytes_per_launch": b, "frames_per_launch": fpl, "dram_MB_per_frame": round(b / fpl / 1e6, 2)}
        tot_traffic += b / fpl
        fs.write(f"## {kn[:60]}\n")
        for w in ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread']:
            fs.write(f"  {w:66s} {r[h.index(w)]:>16s}\n")
        st = [(float(r[i].replace(',', '')), hh) for i, hh in enumerate(h) if hh.startswith('smsp__average_warps_issue_stalled') and hh.endswith('_per_issue_active.ratio') and r[i] not in ('', 'n/a')]
        st.sort(reverse=True)
        for v, hh in st[:7]:
            fs.write(f"  stall {hh.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):40s} {v:8.2f}\n")
traffic["dram_MB_per_frame_whole_path"] = round(tot_traffic / 1e6, 1)
traffic["b_io_MB_per_frame"] = 51.77
traffic["traffic_over_b_io"] = round(tot_traffic / 51.767118e6, 2)
json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
