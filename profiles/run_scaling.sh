#!/bin/bash
# usage: profiles/run_scaling.sh N [N ...]   -- default bench and config C under torchrun at each N, lines into gpurun_out/
for n in "$@"; do
  port=$((29500 + n))
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --no-single-thread 2>/dev/null | grep '^{' > gpurun_out/r02_bench_n1.json
    python bench.py --gpus 1 --config C 2>/dev/null | grep '^{' > gpurun_out/r02_configC_n1.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n 2>/dev/null | grep '^{' > gpurun_out/r02_bench_n$n.json
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((port + 50)) bench.py --gpus $n --config C 2>/dev/null | grep '^{' > gpurun_out/r02_configC_n$n.json
  fi
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_n$n.json")); print("N=$n default:", round(d["value"], 1), "fps, e2e", round(d["e2e"]["value"], 1), d["e2e"].get("host_gather"), "ms/step", round(d["ms_per_step"], 2))
c = json.load(open("gpurun_out/r02_configC_n$n.json")); print("N=$n config C:", {k: c.get(k) for k in ("value", "ms_per_step")}, {k: v for k, v in c.get("config", {}).items() if "ms" in k or "match" in k})
PY
done
