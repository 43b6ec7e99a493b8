#!/bin/bash
# timing experiment: phase counters of a scratch build with extra defines ($1), normal build restored afterwards
cp graph_neural_cellular_automata_b200/lib/libgnca.so /tmp/libgnca_normal.so
GNCA_EXTRA_DEFINES="$1" GNCA_PHASE_COUNTERS=1 python -c "from graph_neural_cellular_automata_b200 import build; build.build(force=True, verbose=False)" > gpurun_out/exp_build.log 2>&1
timeout 200 python bench.py --workload c5s --steps 3 --warmup 2 --no-cpu-baseline 2> gpurun_out/exp.err | grep -v "warpgroup [12]" | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('c5s ms/rollout %.3f' % d['ms_per_step'], d['roofline']['kernels'])
    elif 'phases' in l: print(l.strip())
" | tail -4
cp /tmp/libgnca_normal.so graph_neural_cellular_automata_b200/lib/libgnca.so
