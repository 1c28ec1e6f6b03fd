#!/bin/bash
cd "$GRAFT_REPO_ROOT" || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ba.py -m gpu -q -x > gpurun_out/c18_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/c18_tests.log
PVIO_B200_TUNE_LIB=$PWD/tools/_variants/libpvio_stamps.so timeout 200 python tools/solve_stamps.py 2>&1 | grep "total\|cholesky +" | tee gpurun_out/c18_stamps.log
timeout 200 python tools/call_overhead.py 2>&1 | tail -4
