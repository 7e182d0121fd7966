#!/bin/bash
# A/B of the k_count_part launch geometries (BFC_B200_PART_CFG) on the GPU box: parity tests + device-resident bench each
mkdir -p gpurun_out
for c in ${CFGS:-2 3 4}; do
  BFC_B200_PART_CFG=$c python -m pytest tests -m gpu -x -q -k "oracle_random or device_batches or sharded or golden_trim" > gpurun_out/ab_pytest_$c.log 2>&1; tail -1 gpurun_out/ab_pytest_$c.log
  BFC_B200_PART_CFG=$c python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ab_bench_$c.json 2> gpurun_out/ab_bench_$c.err
  python - <<P
import json
d=json.loads(open("gpurun_out/ab_bench_$c.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print("cfg $c value", round(d["value"],2), "ms", round(d["ms_per_step"]), {n:round(v["ms"]) for n,v in k.items()})
P
done
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ab_bench_e2e.json 2> gpurun_out/ab_bench_e2e.err
python - <<P
import json
d=json.loads(open("gpurun_out/ab_bench_e2e.json").read().strip().splitlines()[-1])
print("e2e", d["e2e"]["value"], d["e2e"].get("split"), "value", d["value"])
P
