#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 100 python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2t_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), {k: round(v['ms']*1000) for k,v in d['kernels'].items()}, 'e2e', round(d['e2e']['value']), 'wire', round(d['e2e_wire']['value']), 'f1', round(d['f1']['value_f1']), d['roofline_path']['frac'], d['cpu_baseline']['value'])
PY
