#!/bin/bash
cd "$GRAFT_REPO_ROOT" || cd /root/repo
mkdir -p gpurun_out
PVIO_B200_TUNE_LIB=$PWD/tools/_variants/libpvio_stamps.so timeout 300 python tools/solve_stamps.py 2>&1 | tee gpurun_out/c6_stamps.log
