#!/bin/bash
# launch list (ALL kernels, incl. torch's) of the torus training iteration: what the 0.6 ms outside the three big kernels is
ZP_ONLY=torus python scripts/time_zeropad_train.py > gpurun_out/tr_plain.log 2>&1 && \
ZP_ONLY=torus ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 5150 -c 330 --csv --log-file gpurun_out/train_launches.csv python scripts/time_zeropad_train.py > gpurun_out/tr_ncu.log 2>&1
tail -1 gpurun_out/tr_plain.log
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/train_launches.csv")) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
seq = []
for r in rows[1:]:
    v = float(r[iv].replace(",", "")); v = v / 1000 if r[iu] in ("ns", "nsecond") else v
    seq.append((r[ik][:58], v))
# one iteration = from one k_rep_fwd to the next
idx = [i for i, (k, v) in enumerate(seq) if "k_rep_fwd" in k]
print("launches between consecutive k_rep_fwd:", [b - a for a, b in zip(idx, idx[1:])])
if len(idx) >= 2:
    it = seq[idx[-2]:idx[-1]]
    tot = sum(v for k, v in it)
    small = [(k, v) for k, v in it if not any(s in k for s in ("k_rep_fwd", "k_rep_bwd", "k_rep_wgrad"))]
    print("iteration: %d launches, %.1f us total, %.1f us outside the three big kernels (%d launches)" % (len(it), tot, sum(v for k, v in small), len(small)))
    for k, v in it: print("  %-60s %8.2f us" % (k, v))
PY
