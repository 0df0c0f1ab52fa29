mkdir -p gpurun_out
python bench.py --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
echo "default"; python scratch/kernels_of.py gpurun_out/bench_b.json | head -3; tail -2 gpurun_out/bench_b.err
VSB200_LIB=$PWD/scratch/variants/libvsb200_u2.so python bench.py --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/bench_u2.json 2> gpurun_out/bench_u2.err
echo "unroll 2"; python scratch/kernels_of.py gpurun_out/bench_u2.json | head -3
