#!/bin/bash
# phase counters of k_update_tc (a GNCA_PHASE_COUNTERS build into a scratch lib dir, then the normal build restored)
cp graph_neural_cellular_automata_b200/lib/libgnca.so /tmp/libgnca_normal.so
GNCA_PHASE_COUNTERS=1 python -c "from graph_neural_cellular_automata_b200 import build; build.build(force=True, verbose=False)" > gpurun_out/phase_build.log 2>&1
timeout 200 python bench.py --workload c5s --steps 1 --warmup 1 2>&1 | grep "phases" | head -12 > gpurun_out/tc_phases.txt
cat gpurun_out/tc_phases.txt
cp /tmp/libgnca_normal.so graph_neural_cellular_automata_b200/lib/libgnca.so
