#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -4 gpurun_out/r02c_pytest.log
ONE="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-cli"
timeout 600 python bench.py --workload count $ONE > gpurun_out/r02c_count.json 2> gpurun_out/r02c_count.err; echo "count rc=$?"; tail -2 gpurun_out/r02c_count.err
NCU="ncu --set full --import-source on --clock-control none"
timeout 900 $NCU -k regex:k_ec_search2 -s 3 -c 1 -f -o gpurun_out/r02c_search2 python bench.py --reads 40000000 $ONE > gpurun_out/r02c_ncu2.log 2>&1; echo "ncu v2 rc=$?"
BFC_B200_EC_V1=1 timeout 900 $NCU -k regex:k_ec_search -s 3 -c 1 -f -o gpurun_out/r02c_search1 python bench.py --reads 40000000 $ONE > gpurun_out/r02c_ncu1.log 2>&1; echo "ncu v1 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02c_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["metric"], round(d["value"],2))
        for k,v in sorted(d["roofline"]["kernels"].items(), key=lambda kv:-kv[1]["ms"]):
            print("   %-14s %8.1f ms/step share %.3f launches %d %s" % (k, v["ms"]/d["steps"], v["share_of_step"], v["launches"], ("frac %.3f" % v["frac"]) if "frac" in v else ""))
    except Exception as e:
        print(f, "unreadable", e)
PY
ls -la gpurun_out/*.ncu-rep | tail -3
