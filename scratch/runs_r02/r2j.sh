#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
N=$1
date +%s > gpurun_out/r2j_t0_$N
timeout 60 true > gpurun_out/r2j_pytest_n$N.log 2>&1; tail -3 gpurun_out/r2j_pytest_n$N.log
timeout 540 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2972$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2j_bench_n$N.json 2> gpurun_out/r2j_bench_n$N.err
tail -2 gpurun_out/r2j_bench_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2j_bench_n$N.json').read().strip().splitlines()[-1])
    print('N=$N', d['config']['workload'], 'F', d['config']['frames_per_step'], 'value', round(d['value']), 'single', d['single_gpu_same_workload'], 'e2e', d['e2e'] and round(d['e2e']['value']), 'parity', d['parity_checked'], 'replicas', d['replicas'] and round(d['replicas']['value']), 'xbytes', d['exchange_bytes_per_frame'], [ (s['views'], s['strip']) for s in d['shards']])
except Exception as e: print('ERR', e)
PY
echo "wall $(( $(date +%s) - $(cat gpurun_out/r2j_t0_$N) )) s"
