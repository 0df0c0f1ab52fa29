mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for v in 0 1 2 3; do
VSB_REMAP_VARIANT=$v python bench.py --no-cpu-baseline --no-e2e --steps 60 > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
echo "variant $v"; python scratch/kernels_of.py gpurun_out/bench_v$v.json | head -3; tail -3 gpurun_out/bench_v$v.err
done
