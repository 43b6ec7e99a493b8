#!/bin/bash
# ncu --set full capture of the tensor-core k_update (c5 slice), after the same command has run clean without ncu
CMD="python bench.py --workload c5s --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_tc_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_update_tc --launch-skip 30 -c 1 -f -o gpurun_out/r02_k_update_tc $CMD > gpurun_out/ncu_tc.log 2>&1
tail -3 gpurun_out/ncu_tc.log; ls -la gpurun_out/*.ncu-rep
