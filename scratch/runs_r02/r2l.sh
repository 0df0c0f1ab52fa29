#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
t0=$(date +%s)
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2l_bench_n2.json 2> gpurun_out/r2l_bench_n2.err
echo "rc=$? wall $(( $(date +%s) - t0 )) s"
grep -v "^$\|OMP_NUM\|\*\*\*" gpurun_out/r2l_bench_n2.err | tail -5 | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2l_bench_n2.json').read().strip().splitlines()[-1])
    print('N=2', d['config']['workload'], 'F', d['config']['frames_per_step'], 'value', round(d['value']), 'single', d.get('single_gpu_same_workload'), 'e2e', d.get('e2e') and round(d['e2e']['value']), 'parity', d.get('parity_checked'), 'replicas', d.get('replicas') and round(d['replicas']['value']))
except Exception as e: print('ERR', e)
PY
t0=$(date +%s)
VSB_BENCH_INJECT_FAIL=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 20 --warmup 3 --shard-deadline 120 > gpurun_out/r2l_bench_n2_fail.json 2> gpurun_out/r2l_bench_n2_fail.err
echo "inject: rc=$? wall $(( $(date +%s) - t0 )) s"
tail -1 gpurun_out/r2l_bench_n2_fail.json | cut -c1-700
