TAG=r01d
mkdir -p gpurun_out
for k in k_blend k_remap_stage1_tab; do
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_$k.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_$k.log | cut -c1-150
done
