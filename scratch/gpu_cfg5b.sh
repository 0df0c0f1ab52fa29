for i in 1 2 3 4 5 6; do
python bench.py --workload cfg5 --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/c5_$i.json 2> gpurun_out/c5_$i.err; echo "run $i rc=$? bytes=$(stat -c %s gpurun_out/c5_$i.json)"; grep -v "Warning\|warn" gpurun_out/c5_$i.err | tail -4
done
