TAG=r01f
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python scratch/kernels_of.py gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 60 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
python bench.py --workload cfg3 --no-cpu-baseline --steps 40 > gpurun_out/${TAG}_bench_cfg3.json 2>/dev/null; python scratch/kernels_of.py gpurun_out/${TAG}_bench_cfg3.json | head -1
python bench.py --workload cfg4 --batch 2 --no-cpu-baseline --steps 20 > gpurun_out/${TAG}_bench_cfg4.json 2>/dev/null; python scratch/kernels_of.py gpurun_out/${TAG}_bench_cfg4.json | head -1
python bench.py --workload cfg5 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_cfg5.json 2>/dev/null; python scratch/kernels_of.py gpurun_out/${TAG}_bench_cfg5.json | head -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_launches.log 2>&1
tail -1 gpurun_out/${TAG}_launches.log | cut -c1-200
for k in k_blend k_remap_stage1_tab k_remap_stage2_tab k_down2 k_coarse k_down_tail; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_$k.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_$k.log | cut -c1-120
done
