#!/bin/bash
# Round-end evidence run on one B200 (everything lands in gpurun_out/): tests, smoke, bench lines, launch list, ncu captures.
set -x
O=gpurun_out
python -m pytest tests -q -m gpu > $O/r01_final_tests.log 2>&1; tail -2 $O/r01_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r01_smoke.log 2>&1; tail -2 $O/r01_smoke.log
python bench.py > $O/r01_bench_default.json 2> $O/r01_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r01_bench_reference.json 2>/dev/null
python bench.py --workload c3 --no-cpu-baseline > $O/r01_bench_c3.json 2>/dev/null
python bench.py --workload c1 --no-cpu-baseline --no-train-extra > $O/r01_bench_c1.json 2>/dev/null
python bench.py --workload c4 --no-cpu-baseline > $O/r01_bench_c4.json 2>/dev/null
python bench.py --workload c5s --steps 5 --warmup 3 --no-cpu-baseline > $O/r01_bench_c5s.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_update --launch-skip 30 -c 1 -f -o $O/r01_k_update python bench.py --workload c5s --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_ku.log 2>&1
ls -la $O/*.json | tail -8
ncu --set full --clock-control none --import-source on -k regex:k_rep_fwd --launch-skip 3 -c 1 -f -o $O/r01_rep_fwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train-extra > $O/ncu_repfwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rep_bwd --launch-skip 3 -c 1 -f -o $O/r01_rep_bwd python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_repbwd.log 2>&1
