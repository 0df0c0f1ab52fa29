#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shard.py tests/test_gpu_parity.py -m gpu -x -q -k "native or shard_compose or batched or small4" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -5 gpurun_out/r2h_pytest.log
for b in 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2971$b bench.py --gpus 2 --steps 20 --warmup 3 --batch $b --no-replicas > gpurun_out/r2h_bench_n2_b$b.json 2> gpurun_out/r2h_bench_n2_b$b.err
tail -1 gpurun_out/r2h_bench_n2_b$b.err | cut -c1-200
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2h_bench_n2_b$b.json').read().strip().splitlines()[-1])
    print('N2 batch $b', round(d['value']), 'single', d['single_gpu_same_workload'], 'e2e', d['e2e'] and round(d['e2e']['value']), 'parity', d['parity_checked'], [s['launches_per_step'] for s in d['shards']])
except Exception as e: print('ERR', e)
PY
done
timeout 200 python bench.py --no-cpu-baseline --no-e2e --batch 16 --ring 16 > gpurun_out/r2h_bench_f16.json 2> gpurun_out/r2h_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2h_bench_f16.json').read().strip().splitlines()[-1])
    print('F16', round(d['value']), {k:round(v['ms']*1000) for k,v in d['kernels'].items()})
except Exception as e: print('ERR', e)
PY
