#!/bin/bash
# A/B on the GPU box: device-resident bench under each of the given environment settings ("NAME=VALUE" or "-")
mkdir -p gpurun_out
i=0
for kv in "$@"; do
  i=$((i+1))
  if [ "$kv" = "-" ]; then python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/abenv_$i.json 2> gpurun_out/abenv_$i.err
  else env $kv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/abenv_$i.json 2> gpurun_out/abenv_$i.err; fi
  python - <<P
import json
d=json.loads(open("gpurun_out/abenv_$i.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print("$kv value", round(d["value"],2), "ms", round(d["ms_per_step"]), {n:round(v["ms"]) for n,v in k.items()}, d["stats"]["lookups_per_read"])
P
done
