mkdir -p gpurun_out
for rep in 1 2; do
for v in 0 1; do
VSB_REMAP_VARIANT=$v python bench.py --no-cpu-baseline --no-e2e --steps 60 > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
echo "variant $v (minb4)"; python scratch/kernels_of.py gpurun_out/bench_v$v.json | head -2; tail -3 gpurun_out/bench_v$v.err
VSB_REMAP_VARIANT=$v VSB200_LIB=$PWD/scratch/variants/libvsb200_l5.so python bench.py --no-cpu-baseline --no-e2e --steps 60 > gpurun_out/bench_l5.json 2> gpurun_out/bench_l5.err
echo "variant $v (minb5)"; python scratch/kernels_of.py gpurun_out/bench_l5.json | head -2
done
done
