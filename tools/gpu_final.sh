#!/bin/bash
# final validation of the round: all GPU tests, smoke, the bench line, the launch list of the bench under ncu
cd "$GRAFT_REPO_ROOT" || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/final_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/final_bench_under_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/call_overhead.py 2>&1 | tail -4 | tee gpurun_out/final_call_overhead.log
