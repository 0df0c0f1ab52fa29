#!/bin/bash
# round 2, call A: parity suite + bench on the new kernels, split variants
cd /root/repo; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
for sp in 1 4; do VSB_SPLIT=$sp timeout 200 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2a_bench_split$sp.json 2>> gpurun_out/r2a_bench.err; done
VSB_REMAP_VARIANT=0 timeout 200 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2a_bench_rv0.json 2>> gpurun_out/r2a_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), {k:round(v['ms']*1000) for k,v in d['kernels'].items()}, d.get('e2e') and round(d['e2e']['value']), d.get('e2e_wire') and round(d['e2e_wire']['value']))
    except Exception as e: print(f, 'ERR', e)
PY
