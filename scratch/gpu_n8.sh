mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --workload cfg4 --batch 2 --ring 2 --steps 30 --no-e2e --no-cpu-baseline > gpurun_out/r01f_bench_n8_cfg4.json 2> gpurun_out/n8.err
python scratch/kernels_of.py gpurun_out/r01f_bench_n8_cfg4.json | head -1; tail -3 gpurun_out/n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 60 --no-cpu-baseline > gpurun_out/r01f_bench_n8_replicas.json 2> gpurun_out/n8b.err
python scratch/kernels_of.py gpurun_out/r01f_bench_n8_replicas.json | head -1; tail -3 gpurun_out/n8b.err
