#!/bin/bash
# launch list (durations) of one zero-pad training iteration: which streaming-backward kernels cost what
python scripts/time_zeropad_train.py > gpurun_out/zp_plain.log 2>&1 && \
ZP_ONLY=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --launch-skip 1500 -c 1500 --csv --log-file gpurun_out/zp_launches.csv python scripts/time_zeropad_train.py > gpurun_out/zp_ncu.log 2>&1
tail -2 gpurun_out/zp_plain.log
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/zp_launches.csv")) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(",", "")); v = v / 1000 if r[iu] in ("ns", "nsecond") else v
    k = r[ik][:60]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print("%-62s n=%5d total %9.1f us  avg %7.2f us  %4.1f%%" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
PY
