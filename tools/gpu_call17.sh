#!/bin/bash
cd "$GRAFT_REPO_ROOT" || cd /root/repo
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --kernel-regex-exclude kns=marg_eig_kernel python tools/sanitize_run.py > gpurun_out/c17_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/c17_racecheck.log
grep -E "Write Thread|Read Thread" gpurun_out/c17_racecheck.log | sed -E 's/0x[0-9a-f]+//g; s/Thread \([0-9,]+\)/Thread/' | sort | uniq -c | sort -rn | head -12
