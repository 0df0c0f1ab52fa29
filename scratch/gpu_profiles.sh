# usage: bash scratch/gpu_profiles.sh <tag>
TAG=${1:-r01b}
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python scratch/kernels_of.py gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 100 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
# launch list of the same (default) command, two steps after warm-up
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_launches.log 2>&1
tail -1 gpurun_out/${TAG}_launches.log | cut -c1-200
for k in k_blend k_remap_stage1 k_remap_stage2 k_down2 k_coarse k_down1; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o gpurun_out/${TAG}_prof_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_$k.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_$k.log
done
