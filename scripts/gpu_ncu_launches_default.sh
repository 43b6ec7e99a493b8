#!/bin/bash
# ncu launch list of the default bench command, library kernels only (per-launch device times are cold-cache and
# serialised: compare SHARES) -- after the same command has exited 0 without ncu.  One profiler use per call.
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_ll_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/r02_launches_default.csv $CMD > gpurun_out/ncu_ll.log 2>&1
tail -1 gpurun_out/ncu_ll.log | cut -c1-200
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_default.csv")) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(",", "")); v = v / 1000 if r[iu] in ("ns", "nsecond") else v
    a = agg.setdefault(r[ik][:64], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print("%-66s n=%4d total %9.1f us avg %8.2f us %5.1f%%" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
PY
