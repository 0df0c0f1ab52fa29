#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config4" > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2q_pytest.log | cut -c1-800
