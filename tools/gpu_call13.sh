#!/bin/bash
cd "$GRAFT_REPO_ROOT" || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ba.py -m gpu -q -x -k "marg or sequence or resident or chain" > gpurun_out/c13_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/c13_tests.log
PVIO_B200_TUNE_LIB=$PWD/tools/_variants/libpvio_margstamps.so timeout 120 python tools/marg_stamps.py 2>&1 | tee gpurun_out/c13_marg_stamps.log
