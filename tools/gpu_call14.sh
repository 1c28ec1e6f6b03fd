#!/bin/bash
cd "$GRAFT_REPO_ROOT" || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c14_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/c14_tests.log
timeout 900 python bench.py > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/c14_bench.err
PVIO_B200_TUNE_LIB=$PWD/tools/_variants/libpvio_stamps.so timeout 300 python tools/solve_stamps.py 2>&1 | grep "total\|outputs\|vision" | tee gpurun_out/c14_stamps.log
