#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -5 gpurun_out/r02b_pytest.log
ONE="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-cli"
i=0
for kv in "-" "BFC_B200_APPLY_PARK=1" "BFC_B200_NO_BULK=1" "BFC_B200_APPLY_PARK=1 BFC_B200_NO_BULK=1"; do
  if [ "$kv" = "-" ]; then timeout 600 python bench.py --workload count $ONE > gpurun_out/r02b_ab_$i.json 2> gpurun_out/r02b_ab_$i.err
  else env $kv timeout 600 python bench.py --workload count $ONE > gpurun_out/r02b_ab_$i.json 2> gpurun_out/r02b_ab_$i.err; fi
  echo "ab $i ($kv) rc=$?"; tail -2 gpurun_out/r02b_ab_$i.err
  i=$((i+1))
done
timeout 600 python bench.py $ONE > gpurun_out/r02b_v2.json 2> gpurun_out/r02b_v2.err; echo "v2 rc=$?"; tail -2 gpurun_out/r02b_v2.err
BFC_B200_EC_V1=1 timeout 600 python bench.py $ONE > gpurun_out/r02b_v1.json 2> gpurun_out/r02b_v1.err; echo "v1 rc=$?"; tail -2 gpurun_out/r02b_v1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02b_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["metric"], round(d["value"],2), d["stats"]["search_lookups_per_read"])
        for k,v in sorted(d["roofline"]["kernels"].items(), key=lambda kv:-kv[1]["ms"]):
            print("   %-14s %8.1f ms/step share %.3f %s" % (k, v["ms"]/d["steps"], v["share_of_step"], ("frac %.3f" % v["frac"]) if "frac" in v else ""))
    except Exception as e:
        print(f, "unreadable", e)
PY
