set -x
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_f1.json 2> gpurun_out/bench_f1.err; tail -c 3000 gpurun_out/bench_f1.json; tail -5 gpurun_out/bench_f1.err
python bench.py --steps 60 --warmup 5 --batch 4 --no-cpu-baseline > gpurun_out/bench_f4.json 2> gpurun_out/bench_f4.err; tail -c 2500 gpurun_out/bench_f4.json; tail -5 gpurun_out/bench_f4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
for k in k_blend_collapse k_remap_stage1 k_remap_stage2 k_pyr_down; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 2 -f -o gpurun_out/prof_$k python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$k.log 2>&1
tail -2 gpurun_out/ncu_$k.log
done
ls -la gpurun_out
