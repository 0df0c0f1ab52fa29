mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r01f_bench.json 2> gpurun_out/r01f_bench.err; python scratch/kernels_of.py gpurun_out/r01f_bench.json; tail -2 gpurun_out/r01f_bench.err
python bench.py --workload cfg4 --batch 2 --ring 2 --no-cpu-baseline --steps 20 > gpurun_out/r01f_bench_cfg4.json 2>/dev/null; python scratch/kernels_of.py gpurun_out/r01f_bench_cfg4.json | head -1
python -c "import __graft_entry__ as g; g.smoke()"
