mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or cfg2 or small4" 2>&1 | tail -4
for v in 0 1; do
VSB_REMAP_VARIANT=$v python bench.py --no-cpu-baseline --no-e2e --steps 60 > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
echo "variant $v"; python scratch/kernels_of.py gpurun_out/bench_v$v.json | head -3; tail -3 gpurun_out/bench_v$v.err
done
VSB_REMAP_VARIANT=1 VSB200_LIB=$PWD/scratch/variants/libvsb200_l4.so python bench.py --no-cpu-baseline --no-e2e --steps 60 > gpurun_out/bench_l4.json 2> gpurun_out/bench_l4.err
echo "variant l4 lanes"; python scratch/kernels_of.py gpurun_out/bench_l4.json | head -3
VSB_REMAP_VARIANT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "compose_matches or formats" 2>&1 | tail -4
