#!/bin/bash
# round 2, first GPU call: parity tests, then one bench line per workload
set -u
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
for w in count trim k55; do
  timeout 900 python bench.py --workload $w --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
  echo "bench $w rc=$?"; tail -3 gpurun_out/${TAG}_bench_$w.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02a_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["metric"], round(d["value"],2), "e2e", d["e2e"] and round(d["e2e"]["value"],2), "cli", d.get("e2e_cli") and d["e2e_cli"].get("value"), "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"])
        for k,v in sorted(d["roofline"]["kernels"].items(), key=lambda kv:-kv[1]["ms"]):
            print("   %-14s %8.1f ms/step share %.3f %s" % (k, v["ms"]/d["steps"], v["share_of_step"], ("frac %.3f" % v["frac"]) if "frac" in v else ""))
        print("   count_phase", d["roofline"]["count_phase"])
    except Exception as e:
        print(f, "unreadable", e)
PY
