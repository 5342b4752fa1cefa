#!/bin/bash
# usage: profiles/run_scaling_c.sh N [N ...]   -- bench.py --config C under torchrun at each N, lines into gpurun_out/
for n in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --config C 2>/dev/null | grep '^{' > gpurun_out/r02_configC_n$n.json
  python - <<PY
import json
c = json.load(open("gpurun_out/r02_configC_n$n.json")); print("N=$n config C:", {k: c.get(k) for k in ("value", "ms_per_step")}, {k: v for k, v in c.get("config", {}).items() if "ms" in k or "match" in k})
PY
done
