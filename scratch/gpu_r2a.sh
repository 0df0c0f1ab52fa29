set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for v in -1 0 1 2 3; do
  VSB_REMAP_VARIANT=$v python bench.py --no-cpu-baseline --no-e2e --steps 60 > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  echo "variant $v"; python scratch/kernels_of.py gpurun_out/bench_v$v.json; tail -3 gpurun_out/bench_v$v.err
done
