mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
echo "coarse v2"; python scratch/kernels_of.py gpurun_out/bench_b.json; tail -2 gpurun_out/bench_b.err
VSB_COARSE_V1=1 python bench.py --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
echo "coarse v1"; python scratch/kernels_of.py gpurun_out/bench_c1.json | grep -E "fps|coarse"
