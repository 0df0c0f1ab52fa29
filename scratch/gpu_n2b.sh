mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_shard.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --no-cpu-baseline > gpurun_out/r01f_bench_n2_replicas.json 2> gpurun_out/n2.err
python scratch/kernels_of.py gpurun_out/r01f_bench_n2_replicas.json | head -1; tail -3 gpurun_out/n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode shard --workload cfg3 --steps 40 > gpurun_out/r01f_bench_n2_shard_cfg3.json 2> gpurun_out/n2s.err
cut -c1-400 gpurun_out/r01f_bench_n2_shard_cfg3.json; tail -3 gpurun_out/n2s.err
