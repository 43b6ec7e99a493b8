#!/bin/bash
# round-end validation: smoke(), the full GPU suite, the default bench line and the reference arm (no profiler)
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
bash scripts/gpu_bench_round.sh
