#!/bin/bash
# full GPU test suite + default bench line + reference arm (no profiler)
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/gputests.log 2>&1; tail -8 gpurun_out/gputests.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; tail -3 gpurun_out/bench_reference.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("value %.4e  ms %.3f  e2e %.4e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
print("cpu", d["cpu_baseline"]); print("eager", d["gpu_eager"])
for k, v in (d["fwd_bwd"] or {}).items():
    if v and "value" in v: print(k, "%.4e" % v["value"], v.get("ms_per_step"), (v.get("e2e") or {}).get("value"), (v.get("roofline") or {}).get("frac"))
r = json.load(open("gpurun_out/bench_reference.json")); print("reference arm %.4e" % r["value"], r["cpu_baseline"])
PY
