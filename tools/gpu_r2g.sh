#!/bin/bash
# multi-GPU call: real-rank tests, then the bench under torchrun with the library's exchange and with the round-1 one
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_gpu.py -x -q > gpurun_out/r02g_pytest_n$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02g_pytest_n$N.log
tail -5 gpurun_out/r02g_pytest_n$N.log
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $N --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02g_${tag}_n$N.json 2> gpurun_out/r02g_${tag}_n$N.err
  echo "$tag rc=$?"; tail -3 gpurun_out/r02g_${tag}_n$N.err
}
run native BFC_DIST_PROFILE=1
run py BFC_DIST_PY=1 BFC_DIST_PROFILE=1
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r02g_*_n$N.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["metric"], round(d["value"],2), "e2e", d["e2e"] and round(d["e2e"]["value"],2), d["e2e"] and d["e2e"].get("equals_resident_result"), "exchange", d.get("exchange"))
    except Exception as e:
        print(f, "unreadable", e)
PY
grep "dist profile" gpurun_out/r02g_*_n$N.err | tail -4
