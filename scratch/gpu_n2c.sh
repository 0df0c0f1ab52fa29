mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shard.py -m gpu -x -q 2>&1 | tail -6
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode shard --workload cfg3 --batch 8 --steps 30 > gpurun_out/r01f_bench_n2_shard_cfg3_batched.json 2> gpurun_out/n2s.err
cut -c1-300 gpurun_out/r01f_bench_n2_shard_cfg3_batched.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/n2s.err | tail -4
