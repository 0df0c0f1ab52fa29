set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_c_f1.json 2> gpurun_out/bench_c_f1.err; python scratch/kernels_of.py gpurun_out/bench_c_f1.json; tail -5 gpurun_out/bench_c_f1.err
python bench.py --steps 40 --warmup 5 --batch 8 --no-cpu-baseline --no-e2e > gpurun_out/bench_c_f8.json 2> gpurun_out/bench_c_f8.err; python scratch/kernels_of.py gpurun_out/bench_c_f8.json; tail -5 gpurun_out/bench_c_f8.err
for k in k_blend k_remap_stage2 k_coarse; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof_c_$k python bench.py --steps 3 --warmup 3 --batch 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c_$k.log 2>&1
tail -2 gpurun_out/ncu_c_$k.log
done
