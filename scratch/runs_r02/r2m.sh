#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -8 gpurun_out/r2m_pytest.log
timeout 400 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; tail -3 gpurun_out/r2m_bench.err
VSB200_LIB=scratch/variants/libvsb200_co64_4.so timeout 200 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2m_bench_co64_4.json 2>> gpurun_out/r2m_bench.err
VSB_SPLIT=4 timeout 200 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2m_bench_split4.json 2>> gpurun_out/r2m_bench.err
VSB_SPLIT=3 timeout 200 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2m_bench_split3.json 2>> gpurun_out/r2m_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), {k:round(v['ms']*1000) for k,v in d['kernels'].items()}, 'e2e', d.get('e2e') and round(d['e2e']['value']), 'wire', d.get('e2e_wire') and round(d['e2e_wire']['value']), 'f1', d.get('f1') and round(d['f1']['value_f1']))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_blend_seam|k_blend_int|k_remap_stage1_tab|k_remap_stage2_tab|k_coarse|k_down2|k_down_tail' --launch-skip 36 --launch-count 7 -o gpurun_out/r02_prof -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu.log 2>&1
tail -2 gpurun_out/r02_ncu.log
