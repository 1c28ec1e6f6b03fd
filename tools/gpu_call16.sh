#!/bin/bash
cd "$GRAFT_REPO_ROOT" || cd /root/repo
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/c16_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/c16_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_run.py > gpurun_out/c16_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/c16_racecheck.log
