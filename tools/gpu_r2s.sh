#!/bin/bash
set -u
mkdir -p gpurun_out
BFC_B200_FUSE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or oracle_random or sharded or grows or library_exchange" > gpurun_out/r02s_pytest.log 2>&1; tail -3 gpurun_out/r02s_pytest.log
ONE="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-cli"
i=0
for kv in "-" "BFC_B200_FUSE=1"; do
  if [ "$kv" = "-" ]; then timeout 600 python bench.py --workload count $ONE > gpurun_out/r02s_ab_$i.json 2> gpurun_out/r02s_ab_$i.err
  else env $kv timeout 600 python bench.py --workload count $ONE > gpurun_out/r02s_ab_$i.json 2> gpurun_out/r02s_ab_$i.err; fi
  echo "ab $i ($kv) rc=$?"; tail -2 gpurun_out/r02s_ab_$i.err
  i=$((i+1))
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02s_ab_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), {k:round(v["ms"]/d["steps"],1) for k,v in d["roofline"]["kernels"].items()}, d["stats"]["distinct_kmers_in_table"])
    except Exception as e:
        print(f, "unreadable", e)
PY
