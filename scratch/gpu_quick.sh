set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; python scratch/kernels_of.py gpurun_out/bench_q.json; tail -5 gpurun_out/bench_q.err
