#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -6 gpurun_out/r2f_pytest.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -3 gpurun_out/r2f_bench.err
VSB_COARSE_TILE=64 timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench_ct64.json 2>> gpurun_out/r2f_bench.err
for v in co5 nostrips; do VSB200_LIB=scratch/variants/libvsb200_$v.so timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench_$v.json 2>> gpurun_out/r2f_bench.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), {k:round(v['ms']*1000) for k,v in d['kernels'].items()}, 'f1', d.get('f1') and round(d['f1']['value_f1']))
    except Exception as e: print(f, 'ERR', e)
PY
