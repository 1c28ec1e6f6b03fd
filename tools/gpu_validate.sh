#!/bin/bash
# Validation of a build on a B200 (under gpurun: `gpurun --timeout 1500 -- 'bash tools/gpu_validate.sh'`):
# every GPU test, smoke(), the bench line; with LAUNCHES=1 also the launch list of a short bench run (slow under ncu:
# keep -c small, ~0.1 s per profiled launch).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/validate_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/validate_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err; echo "bench rc=$?"
if [ -n "$LAUNCHES" ]; then
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/validate_launches.csv \
        python bench.py --steps 2 --warmup 1 > gpurun_out/validate_bench_under_ncu.log 2>&1; echo "ncu rc=$?"
fi
