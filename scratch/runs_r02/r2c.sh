#!/bin/bash
# round 2, call C: full parity suite, new bench protocol, K1 2-px variant
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -6 gpurun_out/r2c_pytest.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -3 gpurun_out/r2c_bench.err
VSB_REMAP_VARIANT=2 timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2c_bench_rv2.json 2>> gpurun_out/r2c_bench.err
VSB_HOST_SUB=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c_bench_hs1.json 2>> gpurun_out/r2c_bench.err
VSB_HOST_SUB=4 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c_bench_hs4.json 2>> gpurun_out/r2c_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), {k:round(v['ms']*1000) for k,v in d['kernels'].items()}, 'e2e', d.get('e2e') and round(d['e2e']['value']), 'wire', d.get('e2e_wire') and round(d['e2e_wire']['value']), 'f1', d.get('f1'), d['timing']['runs_ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
