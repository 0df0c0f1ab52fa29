#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 2 --workload cfg4 --batch 2 --steps 5 --warmup 3 --no-replicas --no-e2e > gpurun_out/r2k_bench_n2_cfg4.json 2> gpurun_out/r2k_bench_n2_cfg4.err
grep -v "^$\|OMP_NUM\|\*\*\*" gpurun_out/r2k_bench_n2_cfg4.err | tail -12 | cut -c1-400
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2k_bench_n2_cfg4.json').read().strip().splitlines()[-1])
    print('N=2 cfg4', 'F', d['config']['frames_per_step'], 'value', round(d['value']), 'single', d['single_gpu_same_workload'], 'parity', d['parity_checked'], 'xbytes', d['exchange_bytes_per_frame'], [ (s['views'], s['strip']) for s in d['shards']])
except Exception as e: print('ERR', e)
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv | tail -2
