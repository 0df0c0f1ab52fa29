#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 12 python scratch/runs_r02/r2v_check.py > gpurun_out/r2v.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2v.log
