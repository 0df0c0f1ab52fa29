#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2e_gpus.txt
timeout 900 python -m pytest tests/test_gpu_shard.py -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -5 gpurun_out/r2e_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
tail -3 gpurun_out/r2e_bench_n2.err; cat gpurun_out/r2e_bench_n2.json | tail -1 | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 20 --warmup 3 --batch 8 --no-replicas --no-e2e > gpurun_out/r2e_bench_n2_b8.json 2>> gpurun_out/r2e_bench_n2.err
cat gpurun_out/r2e_bench_n2_b8.json | tail -1 | cut -c1-600
