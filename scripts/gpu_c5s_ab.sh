#!/bin/bash
# c5s A/B of an environment switch ($1, e.g. GNCA_APPLY_NARROW=1) against the default, after the scale parity tests
timeout 600 python -m pytest tests/test_gpu_scale.py -x -q -m gpu 2>&1 | tail -2
for v in GNCA_NONE=1 "$1"; do
  env $v timeout 200 python bench.py --workload c5s --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/ab.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('$v', 'c5s ms/rollout %.3f' % d['ms_per_step'], {k: round(v['ms_per_step'], 3) for k, v in d['roofline']['kernels'].items()})"
done
