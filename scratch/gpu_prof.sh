mkdir -p gpurun_out
for k in k_blend k_remap_stage1 k_down2; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof_d_$k python bench.py --steps 3 --warmup 3 --batch 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_d_$k.log 2>&1
tail -1 gpurun_out/ncu_d_$k.log
done
