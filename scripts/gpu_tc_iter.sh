#!/bin/bash
# one iteration on k_update_tc: scale-shape parity, the c5s line, then the phase counters (scratch build, normal build restored)
timeout 600 python -m pytest tests/test_gpu_scale.py tests/test_gpu_r2.py -x -q -m gpu > gpurun_out/tc_scale_test.log 2>&1; tail -4 gpurun_out/tc_scale_test.log
timeout 200 python bench.py --workload c5s --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/tc_c5s.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('c5s ms/rollout %.3f' % d['ms_per_step'], d['roofline']['kernels'])"
tail -2 gpurun_out/tc_c5s.err
bash scripts/gpu_phase_tc.sh
