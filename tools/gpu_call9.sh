#!/bin/bash
cd "$GRAFT_REPO_ROOT" || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c9_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/c9_tests.log
timeout 900 python bench.py > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/c9_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^solve_kernel' -s 14 -c 1 -o gpurun_out/c9_solve -f python tools/cfg3_probe.py > gpurun_out/c9_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/ | tail -5
