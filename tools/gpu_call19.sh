#!/bin/bash
cd "$GRAFT_REPO_ROOT" || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^solve_kernel$' -s 14 -c 1 -o gpurun_out/c19_solve -f python tools/cfg3_probe.py > gpurun_out/c19_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'marg_eig_kernel' -s 2 -c 1 -o gpurun_out/c19_eig -f python tools/marg_probe.py > gpurun_out/c19_ncu2.log 2>&1; echo "ncu rc=$?"
