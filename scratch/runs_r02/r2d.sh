#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or compose_matches or batched or nv12_input or shard_compose or recalibration" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -8 gpurun_out/r2d_pytest.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -3 gpurun_out/r2d_bench.err
VSB_REMAP_VARIANT=1 timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2d_bench_rv1.json 2>> gpurun_out/r2d_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2d_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), {k:round(v['ms']*1000) for k,v in d['kernels'].items()}, 'f1', d.get('f1') and round(d['f1']['value_f1']))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_remap_stage1_st' --launch-skip 3 --launch-count 1 -o gpurun_out/r2d_prof_k1s -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2d_ncu.log 2>&1
tail -2 gpurun_out/r2d_ncu.log
