mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for rep in 1 2; do
python bench.py --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err
echo "split"; python scratch/kernels_of.py gpurun_out/bench_split.json | head -1; tail -2 gpurun_out/bench_split.err
VSB_NO_SPLIT=1 python bench.py --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/bench_nosplit.json 2> gpurun_out/bench_nosplit.err
echo "no split"; python scratch/kernels_of.py gpurun_out/bench_nosplit.json | head -1
done
