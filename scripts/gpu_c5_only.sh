#!/bin/bash
# the full BASELINE configs[4] on N GPUs (torchrun) + the one-GPU slice line
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload c5 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c5_${N}gpu.json 2> gpurun_out/bench_c5_${N}gpu.err; echo "c5 rc=$?"; tail -2 gpurun_out/bench_c5_${N}gpu.err
CUDA_VISIBLE_DEVICES=0 python bench.py --workload c5s --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5s.json 2> gpurun_out/bench_c5s.err
python - <<PY
import json
d = [json.loads(l) for l in open("gpurun_out/bench_c5_${N}gpu.json") if l.startswith("{")][-1]
print("c5 on ${N} GPUs: %.3f s per 1000-step rollout, %.4e cell-updates/s, e2e %.4e" % (d["ms_per_step"] / 1e3, d["value"], d["e2e"]["value"]))
print(d["roofline"]["kernels"])
d = [json.loads(l) for l in open("gpurun_out/bench_c5s.json") if l.startswith("{")][-1]
print("c5s: %.3f ms, %.4e" % (d["ms_per_step"], d["value"]), d["roofline"]["kernels"], d["roofline"]["ncu"])
PY
