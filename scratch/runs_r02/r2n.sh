#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_shard.py -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2n_bench_n2.json 2> gpurun_out/r2n_bench_n2.err
echo "bench rc=$?"; tail -3 gpurun_out/r2n_bench_n2.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2n_bench_n2.json').read().strip().splitlines()[-1])
    print('N=2', d['config']['workload'], 'F', d['config']['frames_per_step'], 'value', round(d['value']), 'single', d.get('single_gpu_same_workload'), 'e2e', d['e2e'] and round(d['e2e']['value']), 'parity', d['parity_checked'], 'replicas', d['replicas'] and round(d['replicas']['value']), d['config'].get('control_plane'))
except Exception as e: print('ERR', e)
PY
