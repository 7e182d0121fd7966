#!/bin/bash
# A/B of environment switches on one GPU box: tools/ab_env.sh <workload> "<ENV=..>" "<ENV=..>" ...  ("-" = no switch)
W=$1; shift
mkdir -p gpurun_out
i=0
for e in "$@"; do
  [ "$e" = "-" ] && e=""
  env $e python bench.py --workload $W --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-cli > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_$i.json"))
k = d["roofline"]["kernels"]
print("$e".ljust(28), "value %.2f  ms/step %.1f " % (d["value"], d["ms_per_step"]), {n: round(v["ms"] / d["steps"], 1) for n, v in k.items()})
PY
  i=$((i+1))
done
