mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
$TR 29701 bench.py --gpus 2 --steps 60 --no-e2e > gpurun_out/bench_n2_replicas.json 2> gpurun_out/bench_n2_replicas.err; python scratch/kernels_of.py gpurun_out/bench_n2_replicas.json; tail -2 gpurun_out/bench_n2_replicas.err
$TR 29702 bench.py --gpus 2 --steps 200 --mode shard --workload cfg3 > gpurun_out/bench_n2_shard_cfg3.json 2> gpurun_out/bench_n2_shard_cfg3.err; cut -c1-700 gpurun_out/bench_n2_shard_cfg3.json; tail -2 gpurun_out/bench_n2_shard_cfg3.err
python bench.py --steps 200 --mode shard --workload cfg3 > gpurun_out/bench_n1_shard_cfg3.json 2> gpurun_out/bench_n1_shard_cfg3.err; cut -c1-300 gpurun_out/bench_n1_shard_cfg3.json; tail -2 gpurun_out/bench_n1_shard_cfg3.err
python bench.py --steps 40 --workload cfg3 --no-cpu-baseline --no-e2e > gpurun_out/bench_n1_cfg3.json 2> gpurun_out/bench_n1_cfg3.err; python scratch/kernels_of.py gpurun_out/bench_n1_cfg3.json
