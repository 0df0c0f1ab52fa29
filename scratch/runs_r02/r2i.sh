#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -12 gpurun_out/r2i_pytest.log
timeout 400 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -3 gpurun_out/r2i_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
    print(round(d['value']), {k:round(v['ms']*1000) for k,v in d['kernels'].items()}, 'e2e', d['e2e'] and round(d['e2e']['value']), 'wire', d['e2e_wire'] and round(d['e2e_wire']['value']), 'f1', d['f1'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], d['roofline_path']['frac'])
except Exception as e: print('ERR', e)
PY
