#!/bin/bash
cd "$GRAFT_REPO_ROOT" || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ba.py -m gpu -q -x -k "failed_linear or dogleg or batch_solve" > gpurun_out/c10_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/c10_tests.log
