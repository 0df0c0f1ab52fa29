#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
date +%s > gpurun_out/r2o_t0
NCCL_DEBUG=WARN timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29738 bench.py --gpus 8 --steps 5 --warmup 3 --ring 2 --no-replicas --no-e2e --no-single --shard-deadline 90 > gpurun_out/r2o_bench_n8.json 2> gpurun_out/r2o_bench_n8.err
echo "bench rc=$? wall $(( $(date +%s) - $(cat gpurun_out/r2o_t0) )) s"
grep -v "^$\|OMP_NUM\|\*\*\*" gpurun_out/r2o_bench_n8.err | grep -i "warn\|error\|nccl" | head -20 | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2o_bench_n8.json').read().strip().splitlines()[-1])
    print('N=8', d['config']['workload'], 'F', d['config']['frames_per_step'], 'value', round(d['value']), 'parity', d['parity_checked'], 'xbytes', d.get('exchange_bytes_per_frame'), [ (s['views'], s['strip']) for s in d.get('shards', [])], d.get('sharded_error'))
except Exception as e: print('ERR', e)
PY
