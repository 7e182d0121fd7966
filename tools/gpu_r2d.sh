#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or oracle_random or scratch or device_batches or refine" > gpurun_out/r02j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
tail -4 gpurun_out/r02j_pytest.log
ONE="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-cli"
i=0
for kv in "-"; do
  if [ "$kv" = "-" ]; then timeout 600 python bench.py $ONE > gpurun_out/r02j_ab_$i.json 2> gpurun_out/r02j_ab_$i.err
  else env $kv timeout 600 python bench.py $ONE > gpurun_out/r02j_ab_$i.json 2> gpurun_out/r02j_ab_$i.err; fi
  echo "ab $i ($kv) rc=$?"; tail -2 gpurun_out/r02j_ab_$i.err
  i=$((i+1))
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02j_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernels"]
        print(f, d["metric"], round(d["value"],2), "search ms", round(k["correct"]["ms"]/d["steps"],1), "lookups/read", d["stats"]["search_lookups_per_read"], "redo", d["stats"]["redo"])
    except Exception as e:
        print(f, "unreadable", e)
PY
