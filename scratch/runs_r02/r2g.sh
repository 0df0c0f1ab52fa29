#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
VSB_STAGGER=1 timeout 200 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2g_bench_stagger2.json 2> gpurun_out/r2g_bench.err
VSB_STAGGER=1 VSB_SPLIT=4 timeout 200 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2g_bench_stagger4.json 2>> gpurun_out/r2g_bench.err
timeout 200 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2g_bench_base.json 2>> gpurun_out/r2g_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2g_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), {k:round(v['ms']*1000) for k,v in d['kernels'].items()}, 'f1', d.get('f1') and round(d['f1']['value_f1']))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 360 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
tail -2 gpurun_out/r2g_bench_n2.err | cut -c1-300; tail -1 gpurun_out/r2g_bench_n2.json | cut -c1-2500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 20 --warmup 3 --batch 8 --no-replicas --no-e2e > gpurun_out/r2g_bench_n2_b8.json 2>> gpurun_out/r2g_bench_n2.err
tail -1 gpurun_out/r2g_bench_n2_b8.json | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 20 --warmup 3 --batch 2 --no-replicas --no-e2e > gpurun_out/r2g_bench_n2_b2.json 2>> gpurun_out/r2g_bench_n2.err
tail -1 gpurun_out/r2g_bench_n2_b2.json | cut -c1-400
