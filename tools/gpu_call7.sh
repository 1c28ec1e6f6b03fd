#!/bin/bash
cd "$GRAFT_REPO_ROOT" || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ba.py -m gpu -q -x > gpurun_out/c7_tests.log 2>&1; echo "tests rc=$?"
tail -12 gpurun_out/c7_tests.log
PVIO_B200_TUNE_LIB=$PWD/tools/_variants/libpvio_stamps.so timeout 300 python tools/solve_stamps.py 2>&1 | grep -v "e+\|-5\|-6\|-4\|-3\|-7\|-8\|-9" | tee gpurun_out/c7_stamps.log
timeout 300 python tools/marg_probe.py 2>&1 | tail -4
timeout 300 python tools/cfg3_probe.py cfg4 2>&1 | tail -3
