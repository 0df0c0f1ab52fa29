set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; python scratch/kernels_of.py gpurun_out/bench_d.json; tail -5 gpurun_out/bench_d.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_d.json').read().strip().splitlines()[-1]); print(d['cpu_baseline']); print(d['e2e']); print(d['clocks'])"
python bench.py --impl reference --steps 20 > gpurun_out/bench_d_ref.json 2>gpurun_out/bench_d_ref.err; cat gpurun_out/bench_d_ref.json | cut -c1-900
nproc; nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv
