set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_b_f1.json 2> gpurun_out/bench_b_f1.err; tail -c 2500 gpurun_out/bench_b_f1.json; tail -5 gpurun_out/bench_b_f1.err
python bench.py --steps 40 --warmup 5 --batch 8 --no-cpu-baseline --no-e2e > gpurun_out/bench_b_f8.json 2> gpurun_out/bench_b_f8.err; tail -c 2500 gpurun_out/bench_b_f8.json; tail -5 gpurun_out/bench_b_f8.err
