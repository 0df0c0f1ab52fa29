mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline --no-e2e --steps 60 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
python scratch/kernels_of.py gpurun_out/bench_b.json; tail -3 gpurun_out/bench_b.err
for v in bl3 minb5 minb6; do
VSB200_LIB=$PWD/scratch/variants/libvsb200_$v.so python bench.py --no-cpu-baseline --no-e2e --steps 60 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
echo "variant $v"; python scratch/kernels_of.py gpurun_out/bench_$v.json; tail -3 gpurun_out/bench_$v.err
done
