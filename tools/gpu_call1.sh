#!/bin/bash
# round-2 session-2 call 1: GPU tests (new F-RANSAC tests first), A/B of programmatic dependent launch, launch lists
cd "$GRAFT_REPO_ROOT" || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fm.py -x -q > gpurun_out/c1_fm.log 2>&1; echo "fm rc=$?" 
tail -15 gpurun_out/c1_fm.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fm.py > gpurun_out/c1_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/c1_tests.log
echo "== PDL build"; timeout 300 python tools/marg_probe.py 2>&1 | tail -4
echo "== no-PDL build"; PVIO_B200_TUNE_LIB=$PWD/tools/_variants/libpvio_nopdl.so timeout 300 python tools/marg_probe.py 2>&1 | tail -4
echo "== PDL again"; timeout 300 python tools/marg_probe.py 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/c1_marg_launches.csv python tools/marg_probe.py > gpurun_out/c1_marg_probe.log 2>&1
echo "ncu rc=$?"
