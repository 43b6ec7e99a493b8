#!/bin/bash
# N-GPU bench line (torchrun, NCCL) -- $1 = number of GPUs
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"; tail -3 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d = [json.loads(l) for l in open("gpurun_out/bench_${N}gpu.json") if l.startswith("{")][-1]
print("fwd value %.4e  ms %.3f  e2e %.4e n_gpus %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["n_gpus"]))
for k, v in (d["fwd_bwd"] or {}).items():
    if v and "value" in v: print(k, "%.4e" % v["value"], "ms %.3f" % v["ms_per_step"], "B_global", v["config"]["global_batch"], "e2e %.4e" % v["e2e"]["value"])
PY
