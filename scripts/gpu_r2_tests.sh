#!/bin/bash
# round-2 parity tests on the GPU + phase counters of k_update_tc
timeout 900 python -m pytest tests/test_gpu_r2.py tests/test_gpu_rollout.py -q -m gpu -x --deselect tests/test_gpu_rollout.py::test_rollout_grads_golden > gpurun_out/r2_tests.log 2>&1; tail -30 gpurun_out/r2_tests.log
bash scripts/gpu_phase_tc.sh
