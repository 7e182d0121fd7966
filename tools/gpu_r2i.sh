#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
tail -5 gpurun_out/r02i_pytest.log
ONE="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-cli"
i=0
for kv in "-" "BFC_B200_NO_FUSE=1"; do
  if [ "$kv" = "-" ]; then timeout 600 python bench.py $ONE > gpurun_out/r02i_ab_$i.json 2> gpurun_out/r02i_ab_$i.err
  else env $kv timeout 600 python bench.py $ONE > gpurun_out/r02i_ab_$i.json 2> gpurun_out/r02i_ab_$i.err; fi
  echo "ab $i ($kv) rc=$?"; tail -2 gpurun_out/r02i_ab_$i.err
  i=$((i+1))
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02i_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["metric"], round(d["value"],2), "count phase", d["roofline"]["count_phase"]["ms"], "frac", round(d["roofline"]["count_phase"]["frac"],3))
        for k,v in sorted(d["roofline"]["kernels"].items(), key=lambda kv:-kv[1]["ms"]):
            print("   %-14s %8.1f ms/step share %.3f launches %d" % (k, v["ms"]/d["steps"], v["share_of_step"], v["launches"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
