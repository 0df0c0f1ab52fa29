mkdir -p gpurun_out
for rep in 1 2; do
for v in prev cur; do
if [ $v = cur ]; then L=$PWD/video-stitcher_b200/libvsb200.so; else L=$PWD/scratch/variants/libvsb200_$v.so; fi
VSB200_LIB=$L python bench.py --no-cpu-baseline --no-e2e --steps 60 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
echo "variant $v"; python scratch/kernels_of.py gpurun_out/bench_$v.json; tail -3 gpurun_out/bench_$v.err
done
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_down_tail -s 4 -c 1 -f -o gpurun_out/r01c_prof_k_down_tail python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_dt.log 2>&1
tail -2 gpurun_out/ncu_dt.log
