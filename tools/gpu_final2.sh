#!/bin/bash
cd "$GRAFT_REPO_ROOT" || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final2_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/final2_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err; echo "bench rc=$?"
