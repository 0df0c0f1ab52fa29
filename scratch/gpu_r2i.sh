mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 60 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python scratch/kernels_of.py gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1]); print(d['cpu_baseline']); print(d['e2e']); print(d['e2e_wire']); print(d['clocks'])"
