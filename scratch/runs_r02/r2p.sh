#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 280 python -m pytest tests/test_gpu_shard.py -m gpu -x -q -k "any_world" > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2p_pytest.log | cut -c1-600
