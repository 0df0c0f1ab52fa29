mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
echo "interior flag"; python scratch/kernels_of.py gpurun_out/bench_b.json | grep -E "fps|blend"; tail -2 gpurun_out/bench_b.err
VSB200_LIB=$PWD/scratch/variants/libvsb200_bl5.so python bench.py --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/bench_bl5.json 2> gpurun_out/bench_bl5.err
echo "bl5"; python scratch/kernels_of.py gpurun_out/bench_bl5.json | grep -E "fps|blend"
