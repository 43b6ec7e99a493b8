#!/bin/bash
# scale-shape parity tests + the c5s bench line with the tensor-core k_update and with the FFMA one (GNCA_NO_TC=1)
timeout 600 python -m pytest tests/test_gpu_scale.py -x -q -m gpu > gpurun_out/tc_scale_test.log 2>&1; tail -15 gpurun_out/tc_scale_test.log
timeout 200 python bench.py --workload c5s --steps 5 --warmup 3 > gpurun_out/tc_c5s.json 2> gpurun_out/tc_c5s.err; tail -3 gpurun_out/tc_c5s.err; cat gpurun_out/tc_c5s.json
GNCA_NO_TC=1 timeout 200 python bench.py --workload c5s --steps 5 --warmup 3 > gpurun_out/notc_c5s.json 2>gpurun_out/notc_c5s.err; cat gpurun_out/notc_c5s.json
