#!/bin/bash
# round 2, call B: shard world=1 test, K1 prefetch variants, ncu full captures of the new kernels
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shard_compose or small4 or nv12" > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -4 gpurun_out/r2b_pytest.log
for v in pf1 pf2; do VSB200_LIB=scratch/variants/libvsb200_$v.so timeout 200 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench_$v.json 2>> gpurun_out/r2b_bench.err; done
timeout 200 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench_base.json 2>> gpurun_out/r2b_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), {k:round(v['ms']*1000) for k,v in d['kernels'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_blend_seam|k_blend_int|k_remap_stage1_tab|k_remap_stage2_tab|k_coarse|k_down2|k_down_tail' --launch-skip 36 --launch-count 7 -o gpurun_out/r2b_prof -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log
ls -la gpurun_out/r2b_prof*
