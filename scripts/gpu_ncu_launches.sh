#!/bin/bash
# ncu launch list of the default bench command (per-launch device times; cold-cache and serialised: compare SHARES) and
# a full capture of the headline kernel k_rep_fwd -- each after the same command has exited 0 without ncu
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_ll_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_default.csv $CMD > gpurun_out/ncu_ll.log 2>&1
tail -2 gpurun_out/ncu_ll.log
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train-extra"
$CMD2 > gpurun_out/ncu_rf_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_rep_fwd --launch-skip 3 -c 1 -f -o gpurun_out/r02_rep_fwd $CMD2 > gpurun_out/ncu_rf.log 2>&1
tail -2 gpurun_out/ncu_rf.log; ls -la gpurun_out/r02_*
