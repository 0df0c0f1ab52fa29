mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "host or demo or feed" 2>&1 | tail -3
python bench.py --no-cpu-baseline --steps 60 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; python scratch/kernels_of.py gpurun_out/bench_e.json; tail -3 gpurun_out/bench_e.err
